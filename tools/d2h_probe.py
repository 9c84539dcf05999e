"""Aggregate device-to-host copy bandwidth of the box with N ranks copying at the same time (what bounds the e2e arm of
bench.py at N > 1: every rank returns its genomes' Scores arrays to pinned host memory).

    python -m torch.distributed.run --nproc-per-node N tools/d2h_probe.py [--mb 2048] [--reps 6]
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    out = {}
    for mode in ("alone", "together"):
        times = []
        for r in range(a.reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if mode == "alone" and world > 1:
                # one rank at a time
                t = 0.0
                for turn in range(world):
                    if turn == rank:
                        t0 = time.perf_counter()
                        h.copy_(d, non_blocking=True)
                        torch.cuda.synchronize()
                        t = time.perf_counter() - t0
                    dist.barrier()
            else:
                t0 = time.perf_counter()
                h.copy_(d, non_blocking=True)
                torch.cuda.synchronize()
                t = time.perf_counter() - t0
            if r:
                times.append(t)
        gbs = n / min(times) / 1e9
        v = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        if world > 1:
            s = v.clone()
            dist.all_reduce(s)
            m = v.clone()
            dist.all_reduce(m, op=dist.ReduceOp.MIN)
            out[mode] = {"sum_gbs": float(s.item()), "min_rank_gbs": float(m.item())}
        else:
            out[mode] = {"sum_gbs": gbs, "min_rank_gbs": gbs}
    if rank == 0:
        print(json.dumps({"ranks": world, "mb_per_copy": a.mb, "d2h": out, "host_cores": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
