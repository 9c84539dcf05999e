"""GPU (-m gpu): the engine on the BASELINE.json configurations AT THE SIZES THE BENCH TIMES, against the unmodified
reference (reference ig/native/library.cpp:189-604).

* the four CPU-runnable configs at full size: every genome's Scores (digest of every field) against digests of the
  unmodified reference library's output (tests/golden/digests/<config>.json, made by make_config_digests.py in the build
  container), candidate pairs / lookups against the restatement, the `.net` the native `pangenes` CLI writes against the
  `.net` made from the reference's scores, and the `.clus` golden of the reference's own netclu_ng.py on that `.net`
  (the script cannot run on the GPU box; identical `.net` text is identical input to it);
* the scale-out config's large-index configuration (S > 2^18 genes, k = 7, family-ordered 65,536-row blocks): the first
  80 genomes against the unmodified reference library RUN LIVE through the fake JNIEnv (oracle/_ref travels to the box),
  all Scores fields bit for bit, `info.lookups` against the reference's printed "Total cost";
* the full 1,000-genome index: the genomes bench.py's reference arm samples, against digests from the C restatement
  (tests/golden/digests/scaleout1000_sample.json).
"""
import hashlib
import json
import os
import re
import subprocess

import numpy as np
import pytest

from parity import check_scores
from pandelos_b200 import digest, native, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
DIGESTS = os.path.join(HERE, "golden", "digests")


def golden(name):
    p = os.path.join(DIGESTS, name + ".json")
    if not os.path.exists(p):
        pytest.skip("%s not generated (tests/golden/make_config_digests.py)" % p)
    with open(p) as f:
        return json.load(f)


def workload_sha(w):
    h = hashlib.sha256()
    for a in (w.residues, w.offsets.astype(np.uint64), w.genome_of.astype(np.uint32)):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module", autouse=True)
def product_lib(engine_lib):
    native._lib = None
    native.load(native.ENGINE_LIB)
    yield


def check_genome(pn, g, gold, what):
    s = pn.generateScoresPart(g)
    d = digest.scores_digest(s)
    want = {k: gold[k] for k in ("cells", "sum", "xor", "tables")}
    assert d == want, "%s genome %d: Scores differ from the reference (%s vs %s)" % (what, g, d, want)
    assert pn.last_stats.pairs == gold["pairs"], "%s genome %d: candidate pairs" % (what, g)
    assert pn.last_stats.lookups == gold["lookups"], "%s genome %d: lookups" % (what, g)
    assert pn.last_stats.rows == gold["rows"]
    return s


@pytest.mark.parametrize("name", ["salmonella7", "ecoli10", "xanthomonas14", "mycoplasma64"])
def test_config_full_size_against_the_unmodified_reference(name, tmp_path):
    gold = golden(name)
    w = synth.shape(name)
    k = synth.calculate_k(w)
    assert workload_sha(w) == gold["workload_sha256"], "the synthetic generator no longer makes the input the goldens were made from"
    assert k == gold["k"] and w.S == gold["genes"] and w.G == gold["genomes"]
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        assert pn.info.N == gold["kmers"] and pn.info.U == gold["entries"]
        assert pn.info.lookups == gold["total_cost"], "Total cost (library.cpp:349)"
        cells = pairs = 0
        for g in range(w.G):
            s = check_genome(pn, g, gold["per_genome"][str(g)], name)
            cells += s.scoresCount
            pairs += pn.last_stats.pairs
        # the device-resident partition path (bench `value`) sees the same job
        st = pn.score_partition_device(0, w.S)
        assert st.cells == cells and st.pairs == pairs and st.lookups == gold["total_cost"]
    finally:
        pn.close()
    # end to end through the native CLI (pandelos.sh:73): .faa in, .net out
    from pandelos_b200 import build
    cli = build.build_host()
    faa, out = str(tmp_path / "in.faa"), str(tmp_path / "out.net")
    w.write_faa(faa)
    # --clus: the families too, clustered from the network in memory (mycoplasma64, whose Girvan-Newman split the
    # reference's script cannot finish: the golden is the native split verified level by level against networkx,
    # tests/golden/digests/mycoplasma64_clus_verified.json; its 11,325-gene component takes about a minute of host time)
    clus_file = str(tmp_path / "out.clus")
    r = subprocess.run([cli, "-i", faa, "-k", str(k), "-o", out] + (["--clus", clus_file] if "sha256" in gold["clus"] else []),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if "sha256" in gold["clus"]:
        assert hashlib.sha256(open(clus_file, "rb").read()).hexdigest() == gold["clus"]["sha256"], "pangenes --clus"
    m = re.search(r"Total cost: (\d+) lookups", r.stdout)
    assert m and int(m.group(1)) == gold["total_cost"]
    text = open(out, "rb").read()
    assert text.count(b"\n") == gold["net"]["lines"]
    assert hashlib.sha256(text).hexdigest() == gold["net"]["sha256"], ".net differs from the one made of the reference's scores"
    # connected components of that .net with the native tool: families + the components left to Girvan-Newman cover every gene once
    rest = str(tmp_path / "rest.net")
    r = subprocess.run([build.NETCLU_BIN, faa, out, "-r", rest], capture_output=True, text=True)
    assert r.returncode in (0, 3), r.stderr     # 3: some components are left to the script's Girvan-Newman split
    fam_lines = [ln for ln in r.stdout.splitlines() if ln.startswith("F{ ")]
    assert len(fam_lines) > 0
    if "lines" in gold["clus"]:   # mycoplasma64: the script's Girvan-Newman split does not end in reasonable time (golden says so)
        assert len(fam_lines) <= gold["clus"]["lines"]
    # .clus itself (pandelos.sh:78-79): the native netclu with its own Girvan-Newman split (netclu_cc -g) against the
    # golden of the unmodified netclu_ng.py on the identical .net — .faa in, families out, nothing of the reference run
    if "sha256" in gold["clus"] and name != "mycoplasma64":     # (mycoplasma64: pangenes --clus above ran the same code)
        r = subprocess.run([build.NETCLU_BIN, faa, out, "-g"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        fams = sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "") for ln in r.stdout.splitlines() if "F{ " in ln))
        ctext = "".join(f + "\n" for f in fams)
        assert len(fams) == gold["clus"]["lines"]
        assert hashlib.sha256(ctext.encode()).hexdigest() == gold["clus"]["sha256"], ".clus differs from netclu_ng.py's"
    if name == "salmonella7":
        # and the whole of pandelos.sh (:46-79) through the package's own pipeline script: dataset.faa in, out_prefix.clus out
        script = os.path.join(os.path.dirname(build.NETCLU_BIN), "pandelos.sh")
        r = subprocess.run(["bash", script, faa, str(tmp_path / "fam")], capture_output=True, text=True, cwd=str(tmp_path))
        assert r.returncode == 0, r.stdout + r.stderr
        assert "k = %d" % k in r.stdout
        assert hashlib.sha256(open(str(tmp_path / "fam.clus"), "rb").read()).hexdigest() == gold["clus"]["sha256"]


def _first80():
    w = synth.shape("scaleout1000").subset_genomes(80)
    return w, 7


def test_scaleout_prefix_against_the_unmodified_reference_live():
    """S = 305 K genes > 2^18, k = 7: the level set, seq_bits and family-ordered row blocks of the timed configuration."""
    from oracle import refjni
    if not refjni.available():
        pytest.skip("oracle/_ref not built")
    gold = golden("scaleout1000_first80")
    w, k = _first80()
    assert w.S > (1 << 18)
    assert workload_sha(w) == gold["workload_sha256"]
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_config_digests as mk
    ref = refjni.RefJni()
    _, text = mk.preprocess_capturing_stdout(ref, w, k)
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        assert pn.info.lookups == mk.total_cost(text) == gold["total_cost"]
        assert pn.info.N == gold["kmers"] and pn.info.U == gold["entries"]
        gb = np.searchsorted(w.genome_of, np.arange(w.G + 1))
        for g in sorted(int(x) for x in gold["per_genome"]):
            s = check_genome(pn, g, gold["per_genome"][str(g)], "scaleout1000[:80]")
            check_scores(s, ref.compute_scores(g), "scaleout1000[:80] genome %d vs library.cpp" % g)
            # the same rows through the device-resident partition path, in one 65,536-row family-ordered block
            st = pn.score_partition_device(int(gb[g]), int(gb[g + 1]))
            assert st.pairs == gold["per_genome"][str(g)]["pairs"] and st.cells == gold["per_genome"][str(g)]["cells"]
        st = pn.score_partition_device(0, w.S)   # all rows: five blocks of 65,536 rows
        assert st.lookups == gold["total_cost"] and st.rows == w.S
    finally:
        pn.close()


def test_scaleout_full_index_sample_genomes():
    """The index bench.py times (1,000 genomes, 3.8 M genes, 1.12 G k-mers): the reference arm's sample genomes."""
    gold = golden("scaleout1000_sample")
    w = synth.shape("scaleout1000")
    k = synth.calculate_k(w)
    assert k == gold["k"] and workload_sha(w) == gold["workload_sha256"]
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    try:
        assert pn.info.N == gold["kmers"] and pn.info.U == gold["entries"] and pn.info.lookups == gold["total_cost"]
        for g in sorted(int(x) for x in gold["per_genome"])[:12]:
            check_genome(pn, g, gold["per_genome"][str(g)], "scaleout1000")
    finally:
        pn.close()
