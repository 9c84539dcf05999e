#!/usr/bin/env python
"""bench.py — the PanDelos Pangenes similarity hot path on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A STEP is one pass of the hot path over the rank's share of one synthetic pan-genome (SURVEY.md §8d shapes):
index build from residues already in HBM (k-mer encode, sort, dedup, groups, forward lists = preprocessSequences)
followed by scoring of the rank's query genes against the whole index (computeScores: accumulate along posting
lists, float32 Jaccard, best hits), cells left in HBM.  With N > 1 every rank holds the whole index (built
redundantly, "built once and replicated"); the query genomes of the job (N x 125 of the 1,000) are split into N
genome-aligned blocks of equal posting-list volume (pandelos_b200/multigpu.py), each rank scores its block, and the
best-hit slices are all-gathered over NCCL.  Per-rank work is fixed as N grows (scaling "weak"): N = 8 covers the
whole data set.

metric  = candidate gene pairs scored per second (distinct (row, col != row) cells evaluated, library.cpp:493)
e2e     = the same through the reference-facing C ABI with HOST buffers: pd_build from host residues, then one
          pd_compute_scores per query genome with every Scores array copied back to pinned host memory.
roofline= scoring kernels: algorithmic bytes / CUDA-event time of the kernels, vs measured HBM peak.
cpu_baseline / --impl reference = the UNMODIFIED reference library (oracle/_ref, fake JNIEnv driver, thread pool
          over genomes as Pangenes.java:54-66) on the box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "scaleout1000"
# genomes scored per rank (weak scaling); None = all genomes on every rank count once (N must be 1)
QUERY_GENOMES = {"scaleout1000": 125}
# genomes of the workload the CPU reference is timed on (bounded sample: ~10-30 s of host work)
CPU_SAMPLE_GENOMES = {"scaleout1000": 48, "mycoplasma64": 64}
# host threads calling pd_compute_scores in the e2e arm (the Java host uses a thread pool, Pangenes.java:54-66)
E2E_THREADS = int(os.environ.get("PD_E2E_THREADS", "0")) or max(
    1, min(4, (os.cpu_count() or 4) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed region (B200_PROFILING.md recipe): an `nvidia-smi -lms 100`
    loop beside the bench (default).  PD_CLOCKS=nvml reads the same counters through NVML in this process (falls back to
    nvidia-smi when the binding is missing); PD_CLOCKS=off disables sampling."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.mode = os.environ.get("PD_CLOCKS", "smi")
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.lines = []      # nvidia-smi csv lines
        self.samples = []    # (sm_mhz, reasons bitmask) from NVML
        self.max_mhz = None
        self.source = None

    def _nvml_loop(self, nv, handle):
        while not self.stop_flag.is_set():
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))))
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def start(self):
        if self.mode == "off":
            return
        if self.mode == "nvml":
            try:
                import pynvml as nv
                nv.nvmlInit()
                # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a list of indices
                vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
                phys = int(vis[self.index]) if self.index < len(vis) else self.index
                handle = nv.nvmlDeviceGetHandleByIndex(phys)
                self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
                self._nv = nv
                self.source = "nvml"
                self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
                self.thread.start()
                return
            except Exception:
                self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.source == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self._nv
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            reasons = sorted(n for n, b in bits.items() if any(r & b for _, r in self.samples))
            sm = [m for m, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable" if self.mode != "off" else "clock sampling off"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


def make_workload(name):
    from pandelos_b200 import synth
    t = time.time()
    w = synth.shape(name)
    k = synth.calculate_k(w)
    log("[bench] workload %s: %d genes, %d genomes, %d residues, k=%d (%.1fs)" % (name, w.S, w.G, len(w.residues), k, time.time() - t))
    return w, k


def genome_bounds(w):
    return np.searchsorted(w.genome_of, np.arange(w.G + 1), side="left").astype(np.int64)


def cpu_reference_sample(w, k, name, threads):
    """Times the unmodified reference on the first `n` genomes of the workload.  Returns dict or None."""
    from oracle import refjni
    if not refjni.available():
        return None
    n = min(CPU_SAMPLE_GENOMES.get(name, w.G), w.G)
    sub = w.subset_genomes(n) if n < w.G else w
    ref = refjni.RefJni()
    t_pre = ref.preprocess(sub.residues, sub.offsets, sub.genome_of, k)
    t_sc, cells = ref.compute_scores_pool(0, n, threads)
    return {"sample": sub, "genomes": n, "preprocess_s": t_pre, "scores_s": t_sc, "cells": cells}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pandelos_b200 import native
    name = args.workload
    w, k = make_workload(name)
    threads = os.cpu_count() or 1
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_reference_sample(w, k, name, threads)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
            return 0
        if i >= args.warmup:
            times.append(res["preprocess_s"] + res["scores_s"])
    # pair count of the sample (denominator only) from the engine when a GPU is here, else from the C port
    sub = res["sample"]
    try:
        pn = native.PangeneNative(k, native.PangeneIData(sub.residues, sub.offsets, sub.genome_of))
        pairs = pn.score_partition_device(0, sub.S).pairs
        lookups = pn.info.lookups
        pn.close()
    except Exception:
        from oracle import cport
        o = cport.OracleIndex(sub.residues, sub.offsets, sub.genome_of, k)
        pairs = sum(o.candidate_pairs(g) for g in range(sub.G))
        lookups = o.total_lookups
    t = float(np.mean(times))
    val = pairs / t
    sample = "first %d of %d genomes of %s (%d genes, k=%d, %d lookups); preprocess 1 thread %.2fs + computeScores %d threads %.2fs" % (
        res["genomes"], w.G, name, sub.S, k, lookups, res["preprocess_s"], threads, res["scores_s"])
    out = {"impl": "reference", "metric": "gene-pair Jaccard scores/sec", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
           "config": {"workload": name, "k": k, "genes": int(w.S), "genomes": int(w.G)},
           "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "reference", "sample": sample},
           "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--query-genomes", type=int, default=0, help="genomes scored per rank (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter) is sent to stderr, and the JSON line goes to the saved descriptor
    global print
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _print = print

    def print(*a, **kw):  # noqa: A001 — the result line
        kw.setdefault("file", real_stdout)
        kw.setdefault("flush", True)
        _print(*a, **kw)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pandelos_b200 import native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    native.load()

    from pandelos_b200 import multigpu
    name = args.workload
    w, k = make_workload(name)
    gb = multigpu.genome_bounds(w.genome_of, w.G)
    q = args.query_genomes or QUERY_GENOMES.get(name) or w.G
    q = min(q, w.G)
    if world * q > w.G:
        q = max(1, w.G // world)
    # the O(S) gene table (offsets, genome ids) in pinned host memory: it is re-sent with every index build
    off_pin = torch.from_numpy(w.offsets.astype(np.int64)).pin_memory()
    gid_pin = torch.from_numpy(w.genome_of.astype(np.int32)).pin_memory()
    data = native.PangeneIData(w.residues, off_pin.numpy().view(np.uint64), gid_pin.numpy().view(np.uint32))
    # the job's query genes = the first world x q genomes, split by posting-list volume at genome boundaries
    if world > 1:
        pn0 = native.PangeneNative(k, data, device=local)
        _, visited = pn0.gene_stats()
        pn0.close()
        bounds = multigpu.balanced_bounds(visited, 0, int(gb[world * q]), world, snap=gb)
    else:
        bounds = np.array([0, int(gb[q])], np.int64)
    row0, row1 = int(bounds[rank]), int(bounds[rank + 1])
    g0, g1 = int(np.searchsorted(gb, row0)), int(np.searchsorted(gb, row1))

    # residues resident in HBM before the timed region
    res_host = torch.from_numpy(w.residues).pin_memory()
    res_dev = res_host.to(dev, non_blocking=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    G = w.G
    rows_of_rank = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    gather = None
    if world > 1:
        # best-hit slices are all-gathered chunk by chunk behind the scoring of the next chunk (multigpu.py)
        gather = multigpu.ChunkedBestHitGather(dist, rows_of_rank, G, dev, chunks=max(1, -(-max(rows_of_rank) // 65536)))
        bh_local = gather.local
    else:
        bh_local = torch.zeros((max(rows_of_rank), G), dtype=torch.float32, device=dev)

    host_ms = [0.0, 0.0]  # wall clock of the last step's two calls (log only)

    class Stats:
        def __init__(self, d):
            self.__dict__.update(d)

        def as_dict(self):
            return dict(self.__dict__)

    def step():
        t0 = time.perf_counter()
        pn = native.PangeneNative(k, data, device=local, residues_device_ptr=res_dev.data_ptr())
        t1 = time.perf_counter()
        if gather is not None:
            st = Stats(multigpu.score_and_gather(pn, gather, rank, row0))
        else:
            st = pn.score_partition_device(row0, row1, best_hit_ptr=bh_local.data_ptr())
        host_ms[0], host_ms[1] = (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
        return pn, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        pn, st = step()
        pn.close()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = build_ms = 0.0
    launches = 0
    ms_steps = []
    info = None
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations
        barrier()
        ev0.record()
        t0 = time.perf_counter()
        pn, st = step()
        ev1.record()
        barrier()
        ms_steps.append(max(ev0.elapsed_time(ev1), 0.0))
        wall = (time.perf_counter() - t0) * 1e3
        kernel_ms += st.kernel_ms
        build_ms += pn.info.build_ms[5]
        launches += int(st.launches + pn.info.build_ms[7])
        info = pn.info
        stats = st.as_dict()
        if rank == 0:
            log("[bench] step %.1f ms (wall %.1f): build %.1f [%s], scoring kernels %.1f, scoring call %.1f; host wall: build call %.1f, scoring call %.1f" % (
                ms_steps[-1], wall, pn.info.build_ms[5], " ".join("%.1f" % v for v in pn.info.build_ms[:5]), st.kernel_ms, st.total_ms,
                host_ms[0], host_ms[1]))
        pn.close()
    clocks = sampler.stop() if rank == 0 else None
    ms = float(np.mean(ms_steps))
    t = torch.tensor([ms, float(stats["pairs"]), float(stats["lookups"]), float(stats["cells"]), kernel_ms / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0].item())
        pairs_all, lookups_all, cells_all = float(tsum[1].item()), float(tsum[2].item()), float(tsum[3].item())
    else:
        pairs_all, lookups_all, cells_all = float(stats["pairs"]), float(stats["lookups"]), float(stats["cells"])
    value = pairs_all / (ms * 1e-3)

    # ---- roofline of the scoring kernels on this rank (algorithmic bytes)
    rows_n = row1 - row0
    kms = kernel_ms / args.steps
    # 8 B per posting visited + 12 B per forward entry + 20 B per emitted cell + best-hit table
    # + gene metadata once (the kernel itself reads 4-byte postings: DESIGN.md "algorithmic bytes")
    alg_bytes = 8.0 * stats["lookups"] + 12.0 * stats["fwd_entries"] + 20.0 * stats["cells"] + 4.0 * rows_n * G + 8.0 * info.S
    peak, peak_kind = peaks()
    achieved = alg_bytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    # DRAM traffic of the dominant kernel: one `ncu --set full` capture of its largest launch, committed under profiles/
    traffic, traffic_launch = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_score_traffic.json")
    if args.workload == "scaleout1000" and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes"]
        traffic_launch = {"launch": tj["launch"], "algorithmic_bytes": tj["block_algorithmic_bytes"],
                          "traffic_over_algorithmic": tj["traffic_over_algorithmic"], "source": tj["source"]}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_launch": traffic_launch,
                "bytes_per_lookup_model": 8, "bytes_per_lookup_read": 4,
                "kernel": "score_rows_kernel", "kernel_ms_per_step": kms, "peak_kind": peak_kind,
                "lookups_per_s": stats["lookups"] / (kms * 1e-3) if kms > 0 else 0.0}

    # ---- e2e through the C ABI with host buffers (this rank's query genomes), max over ranks
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 2))
        e2e_ms = []
        d2h = h2d = 0
        pairs_e = 0
        # inputs in pinned host memory; computeScores from a pool of E2E_THREADS host threads, as Pangenes.java:54-66
        # calls it from its thread pool: one call's device->host copies overlap the other calls' kernels
        from concurrent.futures import ThreadPoolExecutor
        data_pinned = native.PangeneIData(res_host.numpy(), w.offsets, w.genome_of)

        def one_genome(pn, g):
            stt, rel = pn.compute_scores_raw(g)
            nbytes = 24 * stt.scoresCount + 4 * stt.rows * stt.G + 4 * stt.S  # six cell arrays (first_seq_genome is host-filled), best-hit table, colmax
            ss = native.ScoreStats()
            pn._L.pd_last_score_stats(stt, ss)
            rel()
            return nbytes, ss.pairs

        E2E_WARM = 2  # untimed passes: every score context has to have met a large genome before the timed ones
        for i in range(E2E_WARM + e2e_steps):
            barrier()
            t0 = time.perf_counter()
            pn = native.PangeneNative(k, data_pinned, device=local, contexts=E2E_THREADS)
            h2d = len(w.residues) + 8 * (w.S + 1) + 4 * w.S
            with ThreadPoolExecutor(max_workers=E2E_THREADS) as pool:
                res_g = list(pool.map(lambda g: one_genome(pn, g), range(g0, g1)))
            d2h = sum(r[0] for r in res_g)
            pairs_e = sum(r[1] for r in res_g)
            barrier()
            if i >= E2E_WARM:
                e2e_ms.append((time.perf_counter() - t0) * 1e3)
            if rank == 0:
                log("[bench] e2e pass %d: %.1f ms%s" % (i, (time.perf_counter() - t0) * 1e3, " (untimed)" if i < E2E_WARM else ""))
            pn.close()
        te = torch.tensor([float(np.mean(e2e_ms)), float(pairs_e)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = te.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = te.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            e_ms, e_pairs = float(tm[0].item()), float(ts[1].item())
        else:
            e_ms, e_pairs = float(te[0].item()), float(te[1].item())
        e2e = {"value": e_pairs / (e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_ms, "call": "pd_build + pd_compute_scores per genome (the JNI boundary), %d host threads" % E2E_THREADS}

        # the same step through the native-CLI call path: pd_genome_edges runs the Java host's BBH filter on the device,
        # so only network edges come back (reported next to e2e, not instead of it)
        def one_genome_edges(pn, g):
            e, rel = pn.genome_edges_raw(g)
            n = int(e.count)
            rel()
            return 12 * n, n

        n_ms_all = []
        for i in range(1 + e2e_steps):  # first pass untimed, like the arm above
            barrier()
            t0 = time.perf_counter()
            pn = native.PangeneNative(k, data_pinned, device=local, contexts=E2E_THREADS)
            with ThreadPoolExecutor(max_workers=E2E_THREADS) as pool:
                res_e = list(pool.map(lambda g: one_genome_edges(pn, g), range(g0, g1)))
            barrier()
            if i > 0:
                n_ms_all.append((time.perf_counter() - t0) * 1e3)
            pn.close()
        n_ms = float(np.mean(n_ms_all))
        n_pairs = float(pairs_e)
        if world > 1:
            tn = torch.tensor([n_ms, n_pairs], dtype=torch.float64, device=dev)
            tn_max = tn.clone()
            dist.all_reduce(tn_max, op=dist.ReduceOp.MAX)
            dist.all_reduce(tn, op=dist.ReduceOp.SUM)
            n_ms, n_pairs = float(tn_max[0].item()), float(tn[1].item())
        e2e["network_path"] = {"value": n_pairs / (n_ms * 1e-3), "unit": "pairs/s", "ms_per_step": n_ms,
                               "d2h_bytes_per_step": int(sum(r[0] for r in res_e)), "edges": int(sum(r[1] for r in res_e)),
                               "call": "pd_build + pd_genome_edges per genome (native pangenes CLI path), %d host threads" % E2E_THREADS}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        res = cpu_reference_sample(w, k, name, threads)
        if res is not None:
            sub = res["sample"]
            pn = native.PangeneNative(k, native.PangeneIData(sub.residues, sub.offsets, sub.genome_of), device=local)
            sp = pn.score_partition_device(0, sub.S)
            tt = res["preprocess_s"] + res["scores_s"]
            cpu = {"value": sp.pairs / tt, "unit": "pairs/s", "cores": threads, "kind": "reference",
                   "sample": "first %d of %d genomes of %s (%d genes, k=%d, %d lookups): preprocess (1 thread) %.2fs + computeScores (%d threads) %.2fs; "
                             "scoring only %.3g lookups/s" % (res["genomes"], w.G, name, sub.S, k, pn.info.lookups, res["preprocess_s"], threads,
                                                               res["scores_s"], pn.info.lookups / max(res["scores_s"], 1e-9))}
            pn.close()

    if rank == 0:
        out = {"metric": "gene-pair Jaccard scores/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "int32+f32", "data": "synthetic",
               "config": {"workload": name, "k": int(k), "genes": int(w.S), "genomes": int(G), "kmers": int(info.N),
                          "query_genomes_per_rank": int(q), "query_rows_per_rank": int(rows_n), "query_genomes_this_rank": [g0, g1],
                          "parallelism": "index replicated, query genomes split by posting-list volume" if world > 1 else "single GPU",
                          "l2": "flushed between timed steps (256 MiB fill)",
                          "step": "index build from HBM-resident residues + scoring of the rank's query rows" + (", NCCL allgather of the best-hit slices chunk by chunk behind the scoring" if world > 1 else "")},
               "lookups_per_s": lookups_all / (ms * 1e-3), "cells_per_step": cells_all, "pairs_per_step": pairs_all,
               "build_ms_per_step": build_ms / args.steps, "score_kernel_ms_per_step": kms,
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
