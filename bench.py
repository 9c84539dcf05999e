#!/usr/bin/env python
"""bench.py — the PanDelos Pangenes similarity hot path on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]
                    [--scaling strong|weak] [--exchange none|allgather]

THE JOB (default workload scaleout1000, BASELINE.json configs[4]): preprocessSequences + computeScores of EVERY genome of
the 1,000-genome synthetic pan-genome (3.8 M genes, 1.12 G k-mers, k = 7 from calculate_k.py's formula).

A STEP (--scaling strong, the default) is one pass of that whole job: index build from residues already in HBM (k-mer
encode, sort, dedup, rank groups, forward lists = preprocessSequences, reference library.cpp:189-371) followed by the
scoring of all query genes (computeScores, library.cpp:409-527: accumulate along posting lists, float32 Jaccard, best
hits), cells left in HBM.  With N > 1 the genomes are split into N genome-aligned blocks of equal posting-list volume
(the reference's own cost model, library.cpp:327); every computeScores(g) of the reference is served by exactly one rank,
so no best-hit exchange is needed (colmax_g[c] is a maximum over the rows of genome g, all on one rank).  `--exchange
allgather` adds the NCCL all-gather of the best-hit slices (the north star's design) as a cross-check mode.
--scaling weak keeps round 1's fixed 125 genomes per rank.

metric  = candidate gene pairs scored per second (distinct (row, col != row) cells evaluated, library.cpp:493), whole job
e2e     = the same job through the reference-facing C ABI with HOST buffers: pd_build from host residues, then one
          pd_compute_scores per genome from a pool of host threads (Pangenes.java:54-66), every Scores array copied
          back to pinned host memory.
roofline= scoring kernels: algorithmic bytes / CUDA-event time of the kernels, vs measured HBM peak.
parity  = after the timed steps the engine's Scores of the reference arm's sample genomes are digest-compared with the
          goldens of tests/golden/digests/ (same index, same genomes); a mismatch fails the bench.
cpu_baseline = the UNMODIFIED reference library (oracle/_ref, fake JNIEnv driver, thread pool over genomes) on a bounded
          sub-index (first 80 genomes), the engine digest-compared with it live.
--impl reference = the unmodified reference library on the box's host cores on THE SAME JOB: the full 1,000-genome index
          is built once (single thread, as the reference does; reported as preprocess_s), each step times computeScores of
          a fixed genome sample (0,125,..,875[, +8 more with 16 or more cores]) with one thread per genome, and
          value = sample pairs / (sample scoring time + preprocess_s x sample/1000): the whole job's rate, EXTRAPOLATED
          from the sample (a full CPU pass is hours) — SURVEY.md §8d, BASELINE.md §3.
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
WEAK_QUERY_GENOMES = {"scaleout1000": 125}   # --scaling weak: genomes scored per rank
SMALL_CONFIGS = ["salmonella7", "ecoli10", "xanthomonas14", "mycoplasma64"]
CPU_SUBINDEX_GENOMES = 80                    # cpu_baseline leg of the b200 arm: reference on the first 80 genomes
REF_FULL_INDEX_GB = 100                      # host RAM the unmodified reference needs for the 1,000-genome index (measured ~70)
DIGESTS = os.path.join(ROOT, "tests", "golden", "digests")
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
    """SM clock and throttle reasons sampled during the timed region (B200_PROFILING.md recipe): an `nvidia-smi -lms 250`
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
        self.first = 0
        self.first_nvml = 0

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
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark_timed_region(self):
        """Waits for the sampler's first sample (so that its start-up is over) and drops what was sampled so far."""
        t0 = time.time()
        while self.source == "nvidia-smi" and not self.lines and time.time() - t0 < 5.0:
            time.sleep(0.02)
        self.first = len(self.lines)
        self.first_nvml = len(self.samples)

    def stop(self):
        if self.source == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self._nv
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            samples = self.samples[self.first_nvml:]
            reasons = sorted(n for n, b in bits.items() if any(r & b for _, r in samples))
            sm = [m for m, _ in samples]
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
        for ln in self.lines[self.first:]:
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


def golden(name):
    p = os.path.join(DIGESTS, name + ".json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


def gold_total(gold, key):
    """Job total of `key` ("pairs" / "cells") from a golden file: stored, or summed when every genome is in the file."""
    if "total_" + key in gold:
        return gold["total_" + key]
    if len(gold["per_genome"]) == gold["genomes"]:
        return sum(e[key] for e in gold["per_genome"].values())
    return None


def host_kmers(w, k):
    ln = np.diff(w.offsets.astype(np.int64))
    return int(np.maximum(ln - k + 1, 0).sum())


def job_config(name, w, k):
    """The `config` object: identical in both arms (what is run on it differs by arm and is described beside it)."""
    return {"workload": name, "k": int(k), "genes": int(w.S), "genomes": int(w.G), "kmers": host_kmers(w, k),
            "job": "preprocessSequences + computeScores of every genome"}


def mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) / 1048576.0
    except Exception:
        pass
    return 0.0


def sample_order(gold):
    """Sample genomes in the order of the golden file's generator: 0,125,..,875 first, then the others ascending."""
    gs = sorted(int(g) for g in gold["per_genome"])
    first = [g for g in gs if g % 125 == 0]
    return first + [g for g in gs if g % 125]


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the same job on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refjni
    name = args.workload
    if not refjni.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return 0
    w, k = make_workload(name)
    cores = os.cpu_count() or 1
    cfg = job_config(name, w, k)
    big = name == "scaleout1000"
    gold = golden(name + "_sample") if big else golden(name)
    index_w, index_note, fallback = w, "full index (all %d genomes)" % w.G, False
    if big and (gold is None or mem_available_gb() < REF_FULL_INDEX_GB or os.environ.get("PD_REF_SUBINDEX")):
        # not enough host RAM for the reference's data layout at 1.12 G k-mers (16-B kmer_rank + radix temp + 24-B ranges)
        gold = golden("scaleout1000_first80")
        index_w = w.subset_genomes(CPU_SUBINDEX_GENOMES)
        index_note = "SUB-INDEX of the first %d genomes (host RAM %.0f GB < %d GB needed for the full reference index)" % (
            CPU_SUBINDEX_GENOMES, mem_available_gb(), REF_FULL_INDEX_GB)
        fallback = True
    if gold is None:
        print(json.dumps({"impl": "reference", "unavailable": "no golden pair counts for %s (tests/golden/digests)" % name}))
        return 0
    if big and not fallback:
        order = sample_order(gold)
        sample = order[:8] if cores < 16 else order[:16]
        pairs = sum(gold["per_genome"][str(g)]["pairs"] for g in sample)
        want_cells = np.array([gold["per_genome"][str(g)]["cells"] for g in sample], np.int64)
    else:
        sample = list(range(index_w.G))
        pairs = gold_total(gold, "pairs")
        want_cells = None
    threads = max(1, min(cores, len(sample)))
    log("[bench] reference: %s; preprocessSequences (1 thread) ..." % index_note)
    ref = refjni.RefJni()
    pre_s = ref.preprocess(index_w.residues, index_w.offsets, index_w.genome_of, k)
    share = pre_s * len(sample) / float(index_w.G)   # the sample's share of the one-off preprocess
    log("[bench] reference: preprocess %.1fs; each step = computeScores of %d genomes on %d threads" % (pre_s, len(sample), threads))
    times, cells_ok = [], True
    for i in range(args.warmup + args.steps):
        t, per = ref.compute_scores_list(sample, threads)
        if want_cells is not None:
            cells_ok = cells_ok and bool((per == want_cells).all())
        else:
            cells_ok = cells_ok and int(per.sum()) == gold_total(gold, "cells")
        if i >= args.warmup:
            times.append(t)
        log("[bench] reference step %d: %.2fs" % (i, t))
    t_sc = float(np.mean(times))
    t = t_sc + share
    val = pairs / t
    lookups = sum(gold["per_genome"][str(g)]["lookups"] for g in sample) if want_cells is not None else gold["total_cost"]
    desc = "%s; preprocessSequences once, 1 thread: %.1fs; per step computeScores of genomes %s on %d threads: %.2fs (%.3g lookups/s); " \
           "value = sample pairs / (scoring + preprocess x %d/%d)%s" % (
               index_note, pre_s, sample if len(sample) <= 16 else "0..%d" % (len(sample) - 1), threads, t_sc, lookups / t_sc,
               len(sample), index_w.G, ", extrapolated to the whole job" if len(sample) < w.G else "")
    out = {"impl": "reference", "metric": "gene-pair Jaccard scores/sec", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic", "config": cfg,
           "extrapolated": len(sample) < w.G, "same_index": not fallback, "preprocess_s": pre_s, "scores_s_per_step": t_sc,
           "sample_genomes": sample if len(sample) <= 32 else len(sample), "sample_pairs": int(pairs),
           "pairs_source": "tests/golden/digests (oracle restatement), not the engine",
           "reference_cells_match_golden": cells_ok, "host_max_rss_gb": __import__("resource").getrusage(__import__("resource").RUSAGE_SELF).ru_maxrss / 1048576.0,
           "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "reference", "sample": desc},
           "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ parity legs

def parity_against_golden(pn, gold, genomes, what):
    """Digest-compares the engine's computeScores of `genomes` with the golden file's; returns (ok, detail)."""
    from pandelos_b200 import digest
    bad = []
    for g in genomes:
        e = gold["per_genome"][str(g)]
        s = pn.generateScoresPart(g)
        d = digest.scores_digest(s)
        if d != {x: e[x] for x in ("cells", "sum", "xor", "tables")} or pn.last_stats.pairs != e["pairs"] or pn.last_stats.lookups != e["lookups"]:
            bad.append(int(g))
    ok = not bad and pn.info.lookups == gold["total_cost"] and pn.info.N == gold["kmers"] and pn.info.U == gold["entries"]
    return ok, {"against": what, "genomes": [int(g) for g in genomes], "mismatches": bad, "total_cost": int(pn.info.lookups),
                "fields": "every Scores field (cells as a multiset of bit patterns, the three tables), candidate pairs, lookups, N, U"}


def cpu_baseline_leg(w, k, name, local):
    """The unmodified reference on a bounded sub-index + the engine on the same sub-index, digest-compared live."""
    from oracle import refjni
    from pandelos_b200 import digest, native
    if not refjni.available():
        return None
    threads = os.cpu_count() or 1
    big = name == "scaleout1000"
    sub = w.subset_genomes(CPU_SUBINDEX_GENOMES) if big else w
    gold = golden("scaleout1000_first80") if big else golden(name)
    ref = refjni.RefJni()
    t_pre = ref.preprocess(sub.residues, sub.offsets, sub.genome_of, k)
    t_sc, per = ref.compute_scores_list(list(range(sub.G)), min(threads, sub.G))
    pn = native.PangeneNative(k, native.PangeneIData(sub.residues, sub.offsets, sub.genome_of), device=local)
    sp = pn.score_partition_device(0, sub.S)
    check = list(range(0, sub.G, max(1, sub.G // 8)))[:8]
    bad = [g for g in check if digest.scores_digest(pn.generateScoresPart(g)) != digest.scores_digest(ref.compute_scores(g))]
    ok = not bad and int(per.sum()) == int(sp.cells)
    if gold is not None:
        ok = ok and gold_total(gold, "pairs") == int(sp.pairs) and gold_total(gold, "cells") == int(sp.cells) and gold["total_cost"] == int(pn.info.lookups)
    tt = t_pre + t_sc
    out = {"value": sp.pairs / tt, "unit": "pairs/s", "cores": min(threads, sub.G), "kind": "reference",
           "sample": "%s of %s (%d genes, k=%d, %d lookups): preprocess (1 thread) %.2fs + computeScores of all its genomes (%d threads) %.2fs; "
                     "scoring only %.3g lookups/s" % ("sub-index of the first %d genomes" % sub.G if big else "the whole index", name, sub.S, k,
                                                       pn.info.lookups, t_pre, min(threads, sub.G), t_sc, pn.info.lookups / max(t_sc, 1e-9)),
           "parity_checked": True, "parity_ok": bool(ok),
           "parity": "engine vs the unmodified library.cpp on this sub-index: Scores digests of genomes %s, total cells%s" % (
               check, ", pairs / cells / Total cost vs tests/golden/digests" if gold is not None else "")}
    pn.close()
    return out, ok


def small_config_line(name, local, dev, steps=3, warmup=2):
    """One bench line for a CPU-runnable config (whole job on one GPU) + its parity against the reference goldens."""
    import torch
    from pandelos_b200 import native
    w, k = make_workload(name)
    gold = golden(name)
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    res_host = torch.from_numpy(w.residues).pin_memory()
    res_dev = res_host.to(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, kms, st, info = [], [], None, None
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        ev0.record()
        pn = native.PangeneNative(k, data, device=local, residues_device_ptr=res_dev.data_ptr())
        st = pn.score_partition_device(0, w.S)
        ev1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms.append(ev0.elapsed_time(ev1))
            kms.append(st.kernel_ms)
        info = pn.info
        pn.close()
    # e2e: host buffers in, every Scores array out, one call per genome
    data_pinned = native.PangeneIData(res_host.numpy(), w.offsets, w.genome_of)
    e_ms = []
    for i in range(1 + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pn = native.PangeneNative(k, data_pinned, device=local)
        for g in range(w.G):
            stt, rel = pn.compute_scores_raw(g)
            rel()
        if i:
            e_ms.append((time.perf_counter() - t0) * 1e3)
        pn.close()
    line = {"value": st.pairs / (np.mean(ms) * 1e-3), "unit": "pairs/s", "ms_per_step": float(np.mean(ms)), "k": int(k), "genes": int(w.S),
            "genomes": int(w.G), "kmers": int(info.N), "lookups": int(info.lookups), "pairs": int(st.pairs), "cells": int(st.cells),
            "build_ms": float(info.build_ms[5]), "score_kernel_ms": float(np.mean(kms)),
            "e2e": {"value": st.pairs / (np.mean(e_ms) * 1e-3), "unit": "pairs/s", "ms_per_step": float(np.mean(e_ms))}}
    ok = True
    if gold is not None:
        pn = native.PangeneNative(k, data, device=local)
        ok, det = parity_against_golden(pn, gold, list(range(w.G)), "tests/golden/digests/%s.json (unmodified library.cpp)" % name)
        ok = ok and int(st.pairs) == gold_total(gold, "pairs") and int(st.cells) == gold_total(gold, "cells")
        pn.close()
        line["parity_checked"], line["parity_ok"] = True, bool(ok)
    else:
        line["parity_checked"] = False
    return line, ok


# ------------------------------------------------------------------------------------------------ the b200 arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--exchange", default="none", choices=["none", "allgather"], help="best-hit slices between ranks (strong scaling)")
    ap.add_argument("--build", default="sharded", choices=["sharded", "replicated"],
                    help="N > 1, strong scaling: ranks build ONE index together (rank-range slices, postings all-gathered) or each its own copy")
    ap.add_argument("--query-genomes", type=int, default=0, help="--scaling weak: genomes scored per rank (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
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
    strong = args.scaling == "strong"
    w, k = make_workload(name)
    cfg = job_config(name, w, k)
    gb = multigpu.genome_bounds(w.genome_of, w.G)
    if strong:
        job_genomes = w.G
    else:
        q = args.query_genomes or WEAK_QUERY_GENOMES.get(name) or w.G
        q = min(q, w.G)
        if world * q > w.G:
            q = max(1, w.G // world)
        job_genomes = world * q
    # the O(S) gene table (offsets, genome ids) in pinned host memory: it is re-sent with every index build
    off_pin = torch.from_numpy(w.offsets.astype(np.int64)).pin_memory()
    gid_pin = torch.from_numpy(w.genome_of.astype(np.int32)).pin_memory()
    data = native.PangeneIData(w.residues, off_pin.numpy().view(np.uint64), gid_pin.numpy().view(np.uint32))
    # residues resident in HBM before the timed region
    res_host = torch.from_numpy(w.residues).pin_memory()
    res_dev = res_host.to(dev, non_blocking=False)
    off_dev = off_pin.to(dev)
    gid_dev = gid_pin.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    sharded = world > 1 and strong and args.build == "sharded"

    def make_index(host_residues=False, **kw):
        """The index as this run builds it: one per rank (single GPU, replicated) or one for all ranks (sharded)."""
        d = data_pinned if host_residues else data
        ptr = None if host_residues else res_dev.data_ptr()
        if not host_residues:   # the whole input is resident in HBM: residues and the gene table (offsets, genome ids)
            kw = dict(kw, table_device_ptrs=(off_dev.data_ptr(), gid_dev.data_ptr()))
        if sharded:
            return multigpu.build_sharded(dist, native, k, d, device=dev, device_index=local, residues_device_ptr=ptr, **kw)
        return native.PangeneNative(k, d, device=local, residues_device_ptr=ptr, **kw), None

    data_pinned = native.PangeneIData(res_host.numpy(), w.offsets, w.genome_of)
    # the job's query genes, split by posting-list volume at genome boundaries (the sharded build does it itself)
    if sharded:
        # (three untimed builds: NCCL sets up its peer-to-peer connections and maps the exchange buffers of the other ranks
        # lazily, per buffer address; until every address the block cache hands out has been seen — a few builds at N = 8 —
        # a build call can stall by 10 - 100 ms.  Setup cost of the communicator, not of the step; the warm-up steps follow.)
        for _ in range(3):
            pn0, bounds = make_index()
            pn0.close()
    elif world > 1:
        pn0 = native.PangeneNative(k, data, device=local)
        _, visited = pn0.gene_stats()
        pn0.close()
        bounds = multigpu.balanced_bounds(visited, 0, int(gb[job_genomes]), world, snap=gb)
    else:
        bounds = np.array([0, int(gb[job_genomes])], np.int64)
    row0, row1 = int(bounds[rank]), int(bounds[rank + 1])
    g0, g1 = int(np.searchsorted(gb, row0)), int(np.searchsorted(gb, row1))

    G = w.G
    rows_of_rank = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    exchange = world > 1 and (args.exchange == "allgather" or not strong)
    gather = None
    bh_local = None
    if exchange:
        # best-hit slices are all-gathered chunk by chunk behind the scoring of the next chunk (multigpu.py)
        gather = multigpu.ChunkedBestHitGather(dist, rows_of_rank, G, dev, chunks=max(1, -(-max(rows_of_rank) // 65536)))
        torch.cuda.synchronize()

    host_ms = [0.0, 0.0]  # wall clock of the last step's two calls (log only)

    class Stats:
        def __init__(self, d):
            self.__dict__.update(d)

        def as_dict(self):
            return dict(self.__dict__)

    def step():
        t0 = time.perf_counter()
        pn, _ = make_index()
        t1 = time.perf_counter()
        if gather is not None:
            st = Stats(multigpu.score_and_gather(pn, gather, rank, row0))
        else:
            # best hits stay in the engine's own table (served per genome by pd_compute_scores / pd_genome_edges)
            st = pn.score_partition_device(row0, row1)
        host_ms[0], host_ms[1] = (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
        return pn, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The clock sampler (an `nvidia-smi -lms` loop) is started BEFORE the warm-up steps: its start-up initialises NVML on every
    # GPU of the box, which holds a driver lock for ~0.1 s and showed up as a 10 - 120 ms host stall inside one of the first
    # timed steps (max over ranks: one stalled rank stalls the collectives of all).  Only samples taken during the timed
    # steps are used.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        pn, st = step()
        pn.close()
    if rank == 0:
        sampler.mark_timed_region()
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
    for tname in ("r02_score_traffic.json", "r01_score_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if args.workload == "scaleout1000" and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes"]
            traffic_launch = {"launch": tj["launch"], "algorithmic_bytes": tj["block_algorithmic_bytes"],
                              "traffic_over_algorithmic": tj["traffic_over_algorithmic"], "source": tj["source"]}
            break
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_launch": traffic_launch,
                "bytes_per_lookup_model": 8, "bytes_per_lookup_read": 4,
                "kernel": "score_rows_kernel", "kernel_ms_per_step": kms, "peak_kind": peak_kind,
                "launches_per_step": int(stats["launches"]),
                "lookups_per_s": stats["lookups"] / (kms * 1e-3) if kms > 0 else 0.0}

    # ---- parity on the timed index: the reference arm's sample genomes that this rank owns, against the goldens
    parity, parity_ok = None, True
    if not args.no_parity:
        gold = golden(name + "_sample") if name == "scaleout1000" else golden(name)
        if gold is not None:
            mine = [g for g in (sample_order(gold) if name == "scaleout1000" else sorted(int(x) for x in gold["per_genome"])) if g0 <= g < g1]
            mine = mine[:max(1, 8 // world)] if name == "scaleout1000" else mine
            pn, _ = make_index()
            parity_ok, parity = parity_against_golden(pn, gold, mine, "tests/golden/digests/%s.json (%s)" % (
                name + "_sample" if name == "scaleout1000" else name, gold.get("oracle", "unmodified library.cpp")))
            pn.close()
            if strong:
                tot_ok = int(lookups_all) == gold["total_cost"] and (gold_total(gold, "pairs") is None or int(pairs_all) == gold_total(gold, "pairs"))
                parity_ok = parity_ok and tot_ok
                parity["job_lookups_match_total_cost"] = bool(int(lookups_all) == gold["total_cost"])
            tp = torch.tensor([1.0 if parity_ok else 0.0, float(len(mine))], dtype=torch.float64, device=dev)
            if world > 1:
                tmin = tp.clone()
                dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
                dist.all_reduce(tp, op=dist.ReduceOp.SUM)
                parity_ok = bool(tmin[0].item() > 0.5)
                parity["genomes_checked_all_ranks"] = int(tp[1].item())
            parity["checked"], parity["ok"] = True, bool(parity_ok)
            if rank == 0:
                log("[bench] parity on the timed index: %s" % ("ok" if parity_ok else "MISMATCH"), parity)
        else:
            parity = {"checked": False, "why": "no golden file for %s" % name}

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

        def one_genome(pn, g):
            stt, rel = pn.compute_scores_raw(g)
            nbytes = 24 * stt.scoresCount + 4 * stt.rows * stt.G + 4 * stt.S  # six cell arrays (first_seq_genome is host-filled), best-hit table, colmax
            ss = native.ScoreStats()
            pn._L.pd_last_score_stats(stt, ss)
            rel()
            return nbytes, ss.pairs

        E2E_WARM = 1  # untimed: the first pass of a process grows the pinned / device result buffers (reported as cold_ms)
        cold_ms = None
        for i in range(E2E_WARM + e2e_steps):
            barrier()
            t0 = time.perf_counter()
            pn, _ = make_index(host_residues=True, contexts=E2E_THREADS)
            h2d = len(w.residues) + 8 * (w.S + 1) + 4 * w.S
            with ThreadPoolExecutor(max_workers=E2E_THREADS) as pool:
                res_g = list(pool.map(lambda g: one_genome(pn, g), range(g0, g1)))
            d2h = sum(r[0] for r in res_g)
            pairs_e = sum(r[1] for r in res_g)
            barrier()
            dt = (time.perf_counter() - t0) * 1e3
            if i >= E2E_WARM:
                e2e_ms.append(dt)
            elif i == 0:
                cold_ms = dt
            if rank == 0:
                log("[bench] e2e pass %d: %.1f ms%s" % (i, dt, " (untimed, cold)" if i < E2E_WARM else ""))
            pn.close()
        te = torch.tensor([float(np.mean(e2e_ms)), float(pairs_e), float(cold_ms), float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = te.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = te.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            e_ms, e_pairs, cold_ms, d2h_all = float(tm[0].item()), float(ts[1].item()), float(tm[2].item()), float(ts[3].item())
        else:
            e_ms, e_pairs, d2h_all = float(te[0].item()), float(te[1].item()), float(d2h)
        e2e = {"value": e_pairs / (e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "d2h_bytes_per_step_all_ranks": int(d2h_all), "ms_per_step": e_ms, "cold_ms": cold_ms,
               "call": "pd_build + pd_compute_scores per genome (the JNI boundary), %d host threads" % E2E_THREADS}

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
            pn, _ = make_index(host_residues=True, contexts=E2E_THREADS)
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
    others = None
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            r = cpu_baseline_leg(w, k, name, local)
            if r is not None:
                cpu, ok = r
                parity_ok = parity_ok and ok
        if not args.no_other_configs and name == "scaleout1000":
            others = {}
            for cname in SMALL_CONFIGS:
                line, ok = small_config_line(cname, local, dev)
                others[cname] = line
                parity_ok = parity_ok and ok

    if rank == 0:
        step_desc = "index build from the HBM-resident input (residues, gene offsets, genome ids) + scoring of the rank's query rows (cells left in HBM)"
        if gather is not None:
            step_desc += ", NCCL allgather of the best-hit slices chunk by chunk behind the scoring"
        out = {"metric": "gene-pair Jaccard scores/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
               "dtype": "int32+f32", "data": "synthetic", "config": cfg,
               "run": {"query_genomes": int(job_genomes), "query_rows_this_rank": int(rows_n), "query_genomes_this_rank": [g0, g1],
                       "parallelism": ("%s; query genomes split by posting-list volume at genome boundaries; best-hit exchange: %s" % (
                                       "ONE index built by all ranks (k-mer rank slices sorted per rank; NCCL: all-gather of 4 B per posting + group bits, "
                                       "all-reduce of 16 B per gene), forward lists per rank for its own rows" if sharded else "index replicated",
                                       "NCCL all-gather" if gather is not None else "none needed (every genome on one rank)")) if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB fill)", "step": step_desc},
               "lookups_per_s": lookups_all / (ms * 1e-3), "cells_per_step": cells_all, "pairs_per_step": pairs_all,
               "build_ms_per_step": build_ms / args.steps, "score_kernel_ms_per_step": kms,
               "roofline": roofline, "parity": parity, "parity_checked": bool(parity and parity.get("checked")), "parity_ok": bool(parity_ok),
               "cpu_baseline": cpu, "e2e": e2e, "other_configs": others, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        log("[bench] PARITY MISMATCH: the numbers above are void")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
