"""Tuning probe: one process, the bench's workload, several table configurations of score_rows_kernel (PD_LEVELS, read
by the engine at every scoring call).  Prints one JSON line per configuration: scoring kernel time of the whole job, rows
that overflowed their table and were re-run, and the job's cells / pairs (which must not depend on the configuration).

    python tools/levels_sweep.py [--config scaleout1000] [--steps 2] "t1:hbits:threads:fcap" ...      (level 2 only)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="scaleout1000")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("levels", nargs="*")
    args = ap.parse_args()
    import torch
    from pandelos_b200 import native, synth
    w = synth.shape(args.config)
    k = synth.calculate_k(w)
    dev = torch.device("cuda", 0)
    res_dev = torch.from_numpy(w.residues).to(dev)
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    base = None
    for lv2 in ["default"] + list(args.levels) + ["default"]:
        if lv2 == "default":
            os.environ.pop("PD_LEVELS", None)
        else:
            os.environ["PD_LEVELS"] = "512:9:128:256:256,2048:11:256:512:1536,%s:0,2048:14:1024:1024:0" % lv2
        kms, st = [], None
        t0 = time.time()
        for i in range(1 + args.steps):
            pn = native.PangeneNative(k, data, device=0, residues_device_ptr=res_dev.data_ptr())
            st = pn.score_partition_device(0, w.S)
            if i:
                kms.append(st.kernel_ms)
            pn.close()
        same = None
        if base is None:
            base = (st.cells, st.pairs, st.lookups)
        else:
            same = (st.cells, st.pairs, st.lookups) == base
        print(json.dumps({"level2": lv2, "score_kernel_ms": float(np.mean(kms)), "all": [round(x, 1) for x in kms], "retry_rows": int(st.retry_rows),
                          "fallback_rows": int(st.fallback_rows), "launches": int(st.launches), "same_job": same, "wall_s": round(time.time() - t0, 1)}),
              flush=True)


if __name__ == "__main__":
    main()
