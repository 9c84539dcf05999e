"""e2e probe (GPU box): where does the host-buffer path spend its time?
    python tools/e2e_probe.py [NAME[:GENOMES]] [--genomes Q]
Prints the pinned D2H bandwidth of the box, then pd_build + pd_compute_scores over Q query genomes with 1..4 host threads,
and the same through pd_genome_edges."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pandelos_b200 import native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="scaleout1000")
    ap.add_argument("--genomes", type=int, default=125)
    a = ap.parse_args()
    name, _, g = a.workload.partition(":")
    w = synth.shape(name, genomes=int(g) if g else None)
    k = synth.calculate_k(w)

    dev = torch.device("cuda:0")
    n = 128 << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for direction in ("d2h", "h2d"):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(20):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        print(json.dumps({"pinned_" + direction + "_GBps": round(20 * n / (time.perf_counter() - t) / 1e9, 2)}), flush=True)

    res = torch.from_numpy(w.residues).pin_memory()
    data = native.PangeneIData(res.numpy(), w.offsets, w.genome_of)
    q = min(a.genomes, int(w.genome_of.max()) + 1)

    def scores(pn, g):
        stt, rel = pn.compute_scores_raw(g)
        ss = native.ScoreStats()
        pn._L.pd_last_score_stats(stt, ss)
        nb = 24 * stt.scoresCount + 4 * stt.rows * stt.G + 4 * stt.S
        rel()
        return nb, ss.total_ms, ss.kernel_ms

    def edges(pn, g):
        e, rel = pn.genome_edges_raw(g)
        nb = 12 * int(e.count)
        rel()
        return nb, 0.0, 0.0

    for label, fn in (("scores", scores), ("edges", edges)):
        for threads in [int(t) for t in os.environ.get("PROBE_THREADS", "2,4").split(",")]:
            walls, builds = [], []
            for rep in range(int(os.environ.get("PROBE_REPS", "6"))):
                t0 = time.perf_counter()
                pn = native.PangeneNative(k, data, device=0, contexts=threads)
                t1 = time.perf_counter()
                with ThreadPoolExecutor(max_workers=threads) as pool:
                    r = list(pool.map(lambda gg: fn(pn, gg), range(q)))
                t2 = time.perf_counter()
                pn.close()
                walls.append(round((t2 - t1) * 1e3, 1))
                builds.append(round((t1 - t0) * 1e3, 1))
            print(json.dumps({"call": label, "threads": threads, "score_ms": walls, "build_ms_min": min(builds),
                              "d2h_GB": round(sum(x[0] for x in r) / 1e9, 2), "sum_kernel_ms": round(sum(x[2] for x in r), 1)}), flush=True)


if __name__ == "__main__":
    main()
