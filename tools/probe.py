"""Perf probe (GPU box): index build + device-resident scoring of a workload; prints timings and roofline numbers.
    python tools/probe.py NAME[:GENOMES] ... [--rows N] [--hash-log2 H] [--repeat R]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pandelos_b200 import native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+")
    ap.add_argument("--rows", type=int, default=0, help="score only the first N rows")
    ap.add_argument("--hash-log2", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--e2e", action="store_true")
    a = ap.parse_args()
    for spec in a.workloads:
        name, _, g = spec.partition(":")
        t = time.time()
        w = synth.shape(name, genomes=int(g) if g else None)
        k = a.k or synth.calculate_k(w)
        tg = time.time() - t
        data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
        t = time.time()
        pn = native.PangeneNative(k, data, hash_log2=a.hash_log2)
        tb = time.time() - t
        i = pn.info
        print(json.dumps({"workload": spec, "gen_s": round(tg, 2), "S": i.S, "G": i.G, "k": k, "base": i.base, "N": i.N, "U": i.U, "R": i.R,
                          "groups": i.groups, "lookups": i.lookups, "build_wall_s": round(tb, 3),
                          "build_ms": {n: round(v, 3) for n, v in zip(["hist", "encode", "sort", "groups", "forward", "total", "h2d", "launches"], i.build_ms)}}), flush=True)
        rows = a.rows or i.S
        for r in range(a.repeat):
            t = time.time()
            st = pn.score_partition_device(0, min(rows, i.S))
            wall = time.time() - t
            alg = 8.0 * st.lookups + 12.0 * st.fwd_entries + 20.0 * st.cells + 4.0 * st.rows * i.G + 8.0 * i.S
            print(json.dumps({"workload": spec, "rows": st.rows, "lookups": st.lookups, "pairs": st.pairs, "cells": st.cells,
                              "fallback_rows": st.fallback_rows, "retry_rows": st.retry_rows, "launches": st.launches, "kernel_ms": round(st.kernel_ms, 3),
                              "total_ms": round(st.total_ms, 3), "wall_ms": round(wall * 1e3, 3),
                              "Glookups_per_s": round(st.lookups / st.kernel_ms / 1e6, 2), "Mpairs_per_s": round(st.pairs / st.kernel_ms / 1e3, 1),
                              "alg_GBps": round(alg / st.kernel_ms / 1e6, 1), "frac_of_6551": round(alg / st.kernel_ms / 1e6 / 6551, 4)}), flush=True)
        if a.e2e:
            t = time.time()
            cells = 0
            for g in range(i.G):
                stt, rel = pn.compute_scores_raw(g)
                cells += stt.scoresCount
                rel()
            print(json.dumps({"workload": spec, "e2e_scores_all_genomes_s": round(time.time() - t, 3), "cells": cells}), flush=True)
        pn.close()


if __name__ == "__main__":
    main()
