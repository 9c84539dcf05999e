"""CPU: the config-size goldens (tests/golden/digests/, made from the UNMODIFIED reference by make_config_digests.py) are
present, well formed and reproduced by the C restatement — the pin of oracle/pangenes_oracle.c at config size — and,
where the reference tree is present (the build container), by the reference's own netclu_ng.py for the `.clus`."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import cport, pangenes_java
from pandelos_b200 import digest, synth

HERE = os.path.dirname(os.path.abspath(__file__))
DIGESTS = os.path.join(HERE, "golden", "digests")
CONFIGS = ["salmonella7", "ecoli10", "xanthomonas14", "mycoplasma64"]
SCRIPT = "/root/reference/netclu_ng.py"


def load(name):
    with open(os.path.join(DIGESTS, name + ".json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", CONFIGS + ["scaleout1000_first80", "scaleout1000_sample"])
def test_golden_present_and_well_formed(name):
    d = load(name)
    assert d["genes"] > 0 and d["kmers"] >= d["entries"] > 0 and d["total_cost"] > 0 and len(d["workload_sha256"]) == 64
    assert len(d["per_genome"]) >= 6
    for g, e in d["per_genome"].items():
        assert 0 <= int(g) < d["genomes"]
        assert e["pairs"] >= e["cells"] > 0 and e["lookups"] >= e["pairs"] and e["rows"] > 0
        assert len(e["sum"]) == 16 and len(e["xor"]) == 16 and len(e["tables"]) == 64
    if name in CONFIGS:
        assert len(d["per_genome"]) == d["genomes"]
        assert d["net"]["lines"] > 0 and (d["clus"].get("lines", 0) > 0 or "unavailable" in d["clus"])
        assert sum(e["lookups"] for e in d["per_genome"].values()) == d["total_cost"]


def test_digest_is_order_independent_and_field_sensitive():
    w = synth.generate(4, 60, 100.0, 0.1, 77)
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, 4)
    s = o.compute_scores(1)
    o.close()
    assert s.scoresCount > 10
    d0 = digest.scores_digest(s)
    p = np.random.default_rng(1).permutation(s.scoresCount)
    for f in cport.Scores.FIELDS:
        setattr(s, f, getattr(s, f)[p])
    assert digest.scores_digest(s) == d0
    for f in cport.Scores.FIELDS:
        a = getattr(s, f).copy()
        old = a[3]
        a[3] = np.nextafter(old, np.float32(2)) if a.dtype == np.float32 else old + 1
        keep = getattr(s, f)
        setattr(s, f, a)
        assert digest.scores_digest(s) != d0, f
        setattr(s, f, keep)
    s.max_genome_score = s.max_genome_score.copy()
    s.max_genome_score[0, 0] += 1
    assert digest.scores_digest(s) != d0


def test_restatement_reproduces_the_reference_goldens_at_config_size():
    """salmonella7 at full size (PR1 config): C restatement == digests of the unmodified library; restated Java host
    gives the golden `.net`; the reference's netclu_ng.py (when present) gives the golden `.clus`."""
    name = "salmonella7"
    gold = load(name)
    w = synth.shape(name)
    k = synth.calculate_k(w)
    assert k == gold["k"]
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, k)
    try:
        assert o.total_lookups == gold["total_cost"] and o.num_entries == gold["entries"] and o.num_kmers == gold["kmers"]
        net = pangenes_java.PangeneNet()
        for g in range(w.G):
            s = o.compute_scores(g)
            e = gold["per_genome"][str(g)]
            assert digest.scores_digest(s) == {x: e[x] for x in ("cells", "sum", "xor", "tables")}, g
            for src, dst, sc in pangenes_java.genome_task(s, g, w.G):
                net.add_connection(src, dst, sc)
    finally:
        o.close()
    text = "".join(ln + "\n" for ln in net.lines())
    assert hashlib.sha256(text.encode()).hexdigest() == gold["net"]["sha256"]
    from pandelos_b200 import build
    build.build_host()
    with tempfile.TemporaryDirectory() as td:
        faa, netf = os.path.join(td, "in.faa"), os.path.join(td, "in.net")
        w.write_faa(faa)
        with open(netf, "w") as f:
            f.write(text)
        # the native netclu (connected components + Girvan-Newman, netclu_cc -g) gives the golden `.clus` everywhere;
        # the reference's own script re-derives that golden where the reference tree is present
        runs = [subprocess.run([build.NETCLU_BIN, faa, netf, "-g"], capture_output=True, text=True, check=True)]
        if os.path.exists(SCRIPT):
            runs.append(subprocess.run([sys.executable, SCRIPT, faa, netf], capture_output=True, text=True, check=True))
    for r in runs:
        fams = sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "") for ln in r.stdout.splitlines() if "F{ " in ln))
        ctext = "".join(f + "\n" for f in fams)
        assert len(fams) == gold["clus"]["lines"] and hashlib.sha256(ctext.encode()).hexdigest() == gold["clus"]["sha256"]


def test_first_girvan_newman_step_on_the_component_the_script_cannot_finish(tmp_path):
    """mycoplasma64: the `.net` (made here from the C restatement's scores, sha256 = the golden of the reference's) has
    one component of 11,325 genes that netclu_ng.py does not split in any reasonable time.  One networkx betweenness pass
    on it was run once (tests/golden/verify_netclu_splits.py --first-edge, ~40 minutes), and the script itself until its first `gn`
    line: the edge removed first and the first split are those of the native split (netclu_cc -g, PD_NETCLU_TRACE=2; stopped there)."""
    gold = load("mycoplasma64")
    first = json.load(open(os.path.join(DIGESTS, "mycoplasma64_first_removed_edge.json")))
    w = synth.shape("mycoplasma64")
    k = synth.calculate_k(w)
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, k)
    try:
        net = pangenes_java.PangeneNet()
        for g in range(w.G):
            for src, dst, sc in pangenes_java.genome_task(o.compute_scores(g), g, w.G):
                net.add_connection(src, dst, sc)
    finally:
        o.close()
    text = "".join(ln + "\n" for ln in net.lines())
    assert hashlib.sha256(text.encode()).hexdigest() == gold["net"]["sha256"]
    from pandelos_b200 import build
    build.build_host()
    faa, netf = str(tmp_path / "in.faa"), str(tmp_path / "in.net")
    w.write_faa(faa)
    open(netf, "w").write(text)
    p = subprocess.Popen([build.NETCLU_BIN, faa, netf, "-g"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True,
                         env=dict(os.environ, PD_NETCLU_TRACE="2"))
    lines = []
    try:
        while not lines or not lines[-1].startswith("gn ("):   # the large component is the first in the network's node order
            lines.append(p.stderr.readline().rstrip("\n"))
            assert lines[-1] and len(lines) < 10
    finally:
        p.kill()
        p.wait()
    assert lines[0].split() == ["rm", str(first["first_removed_edge"][0]), str(first["first_removed_edge"][1])]
    # ... and the first split is the one the unmodified script prints after 67 minutes: 10,635 + 690 genes, two edges removed
    assert len(lines) == 1 + first["first_split"]["removed_edges_before"]
    assert hashlib.sha256(lines[-1].encode()).hexdigest() == first["first_split"]["gn_line_sha256"]
