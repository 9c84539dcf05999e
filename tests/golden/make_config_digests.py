"""Generates tests/golden/digests/*.json in the build container (needs /root/reference): what the UNMODIFIED reference
produces on the BASELINE.json configurations at FULL size, digested (pandelos_b200/digest.py) so that the GPU box —
where neither /root/reference nor a JVM exists — can check the engine on the very inputs the bench times.

  <config>.json  (salmonella7, ecoli10, xanthomonas14, mycoplasma64; synthetic shapes of pandelos_b200/synth.py)
      per genome : digest of every Scores field returned by the unmodified library.cpp (oracle/_ref, fake JNIEnv),
                   candidate pairs and lookups (C restatement oracle/pangenes_oracle.c, whose own Scores digests are
                   asserted equal to the reference's here: the pin of the restatement at config size)
      total_cost : the "Total cost: N lookups" line the reference prints (library.cpp:349)
      net        : sha256 / line count of the .net text made from the reference's scores by the restated Java host
                   (oracle/pangenes_java.py: Pangenes.java:98-176, PangeneNet.java:159-179; no JVM here)
      clus       : sha256 / line count of the .clus the reference's own netclu_ng.py + pandelos.sh:79 make of that .net
                   (mycoplasma64: the script does not finish; the .clus of the native split, verified level by level
                   against networkx — <config>_clus_verified.json, tests/golden/verify_netclu_splits.py)
  scaleout1000_first80.json   the first 80 genomes of the scale-out config (S > 2^18 genes, k = 7: the engine's
                   large-index configuration), same content for a sample of genomes, no .net / .clus
  scaleout1000_sample.json    (--scaleout-full; ~35 GB of RAM, ~30 min) the FULL 1,000-genome index, genomes
                   0,125,..,875 and 24 more (32 in all) — the sample the reference arm of bench.py times: per genome rows, lookups, candidate
                   pairs, cells and Scores digest from the C restatement (the reference library needs > 62 GB there)

    python tests/golden/make_config_digests.py [config ...] [--scaleout-full]
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cport, pangenes_java, refjni  # noqa: E402
from pandelos_b200 import digest, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "digests")
NETCLU_TIMEOUT_S = 900
SMALL = ["salmonella7", "ecoli10", "xanthomonas14", "mycoplasma64"]
# 0,125,..,875 first (SURVEY.md §8d names them), then 24 more so that a host with up to 32 threads has a genome per thread
SAMPLE_GENOMES = list(range(0, 1000, 125)) + [g for g in ((i * 1000) // 32 for i in range(32)) if g % 125]


def workload_digest(w):
    h = hashlib.sha256()
    for a in (w.residues, w.offsets.astype(np.uint64), w.genome_of.astype(np.uint32)):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def preprocess_capturing_stdout(ref, w, k):
    """Runs preprocessSequences with the reference's own prints going to a file; returns (seconds, text)."""
    sys.stdout.flush()
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            t = ref.preprocess(w.residues, w.offsets, w.genome_of, k, quiet=False)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return t, tmp.read().decode("latin-1")


def total_cost(text):
    m = re.search(r"Total cost: (\d+) lookups", text)
    return int(m.group(1)) if m else None


def config_doc(name, w, k, genomes, with_net):
    t0 = time.time()
    ref = refjni.RefJni()
    t_pre, text = preprocess_capturing_stdout(ref, w, k)
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, k)
    kl, visited = o.gene_stats()
    doc = {"config": name, "generator": "tests/golden/make_config_digests.py", "workload_sha256": workload_digest(w), "k": int(k),
           "genes": int(w.S), "genomes": int(w.G), "kmers": int(o.num_kmers), "entries": int(o.num_entries),
           "total_cost": total_cost(text), "reference_preprocess_s": round(t_pre, 2), "per_genome": {}}
    assert doc["total_cost"] == o.total_lookups, (doc["total_cost"], o.total_lookups)
    net = pangenes_java.PangeneNet()
    for g in genomes:
        s = ref.compute_scores(g)
        d = digest.scores_digest(s)
        so = o.compute_scores(g)
        assert digest.scores_digest(so) == d, "restatement differs from the reference: %s genome %d" % (name, g)
        rows = np.flatnonzero(w.genome_of == g)
        d["rows"] = int(len(rows))
        d["lookups"] = int(visited[rows].sum())
        d["pairs"] = int(o.candidate_pairs(g))
        doc["per_genome"][str(g)] = d
        if with_net:
            for src, dst, sc in pangenes_java.genome_task(s, g, w.G):
                net.add_connection(src, dst, sc)
        print("  %s genome %d: %d cells, %d pairs (%.0fs)" % (name, g, d["cells"], d["pairs"], time.time() - t0), flush=True)
    # job totals over ALL genomes (candidate pairs = the bench metric's numerator), from the restatement
    tp = tc = 0
    for g in range(w.G):
        e = doc["per_genome"].get(str(g))
        if e is None:
            e = {"pairs": int(o.candidate_pairs(g)), "cells": int(o.compute_scores(g).scoresCount)}
        tp += e["pairs"]
        tc += e["cells"]
    doc["total_pairs"], doc["total_cells"] = tp, tc
    if with_net:
        lines = net.lines()
        text = "".join(ln + "\n" for ln in lines)
        doc["net"] = {"lines": len(lines), "sha256": hashlib.sha256(text.encode()).hexdigest()}
        with tempfile.TemporaryDirectory() as td:
            faa, netf = os.path.join(td, "in.faa"), os.path.join(td, "in.net")
            w.write_faa(faa)
            with open(netf, "w") as f:
                f.write(text)
            try:
                r = subprocess.run([sys.executable, "/root/reference/netclu_ng.py", faa, netf], capture_output=True, text=True, check=True,
                                   timeout=NETCLU_TIMEOUT_S)
            except subprocess.TimeoutExpired:
                # Girvan-Newman (networkx edge betweenness, recomputed per removed edge) on the large mixed components of
                # the many-genome config does not end in reasonable time; the .net golden above still pins the input to it
                why = "the reference's netclu_ng.py did not finish its Girvan-Newman split within %d s" % NETCLU_TIMEOUT_S
                # <config>_clus_verified.json: the native split's .clus, checked level by level against networkx
                # (tests/golden/verify_netclu_splits.py; hours of CPU, done once and committed with its evidence)
                ver = os.path.join(OUT, name + "_clus_verified.json")
                doc["clus"] = dict(json.load(open(ver)), script_unavailable=why) if os.path.exists(ver) else {"unavailable": why}
                o.close()
                return doc
        # pandelos.sh:79: grep "F{ " | sed s/F{\ //g | sed s/}//g | sed s/\ \;//g | sort | uniq   (byte order, LC_ALL=C)
        fams = sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "") for ln in r.stdout.splitlines() if "F{ " in ln))
        ctext = "".join(f + "\n" for f in fams)
        doc["clus"] = {"lines": len(fams), "sha256": hashlib.sha256(ctext.encode()).hexdigest()}
    o.close()
    return doc


def scaleout_full():
    w = synth.shape("scaleout1000")
    k = synth.calculate_k(w)
    t0 = time.time()
    print("scaleout1000: %d genes, k=%d; building the restatement's index (single thread, comparison sort) ..." % (w.S, k), flush=True)
    o = cport.OracleIndex(w.residues, w.offsets, w.genome_of, k)
    print("  index: %d entries, %d lookups (%.0fs)" % (o.num_entries, o.total_lookups, time.time() - t0), flush=True)
    kl, visited = o.gene_stats()
    doc = {"config": "scaleout1000", "generator": "tests/golden/make_config_digests.py --scaleout-full", "oracle": "port (oracle/pangenes_oracle.c)",
           "workload_sha256": workload_digest(w), "k": int(k), "genes": int(w.S), "genomes": int(w.G), "kmers": int(o.num_kmers),
           "entries": int(o.num_entries), "total_cost": int(o.total_lookups), "per_genome": {}}
    for g in SAMPLE_GENOMES:
        d = digest.scores_digest(o.compute_scores(g))
        rows = np.flatnonzero(w.genome_of == g)
        d["rows"] = int(len(rows))
        d["lookups"] = int(visited[rows].sum())
        d["pairs"] = int(o.candidate_pairs(g))
        doc["per_genome"][str(g)] = d
        print("  genome %d: %d cells, %d pairs (%.0fs)" % (g, d["cells"], d["pairs"], time.time() - t0), flush=True)
    # totals over all genes: what N GPUs score together in one strong-scaling step
    doc["total_rows"] = int(w.S)
    return doc


def write(name, doc):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote %s.json" % name, flush=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--scaleout-full" in sys.argv:
        write("scaleout1000_sample", scaleout_full())
        return
    for name in (args or SMALL + ["scaleout1000_first80"]):
        if name == "scaleout1000_first80":
            w = synth.shape("scaleout1000").subset_genomes(80)
            k = 7
            assert k == synth.calculate_k(synth.shape("scaleout1000"))
            write(name, config_doc(name, w, k, list(range(0, 80, 11)), with_net=False))
        else:
            w = synth.shape(name)
            write(name, config_doc(name, w, synth.calculate_k(w), list(range(w.G)), with_net=True))


if __name__ == "__main__":
    main()
