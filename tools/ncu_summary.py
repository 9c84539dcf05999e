"""Summarises an .ncu-rep (raw + source pages) into text: key metrics per kernel and the hottest SASS lines.
    python tools/ncu_summary.py REPORT.ncu-rep [--top N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = page(rep, "raw")
    h, u = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(h, row))
        un = dict(zip(h, u))
        print("== %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                print("  %-90s %s %s" % (k, d[k], un.get(k, "")))
    src = page(rep, "source")
    # one header line per kernel: "Kernel Name", name
    hdr = None
    rows = []
    for r in src:
        if r and r[0] == "Kernel Name":
            continue
        if r and r[0] == "Address":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr and len(r) >= len(hdr):
            rows.append(r)
    if not rows:
        return
    tot = sum(float(r[hdr["Instructions Executed"]] or 0) for r in rows)
    smp = sum(float(r[hdr["# Samples"]] or 0) for r in rows)
    print("\n== hottest SASS by stall samples (total inst %.3g, samples %d)" % (tot, smp))
    rows.sort(key=lambda r: -float(r[hdr["# Samples"]] or 0))
    for r in rows[:top]:
        print("  smp %5.2f%%  inst %5.2f%%  thr %4s  %s" % (100 * float(r[hdr["# Samples"]] or 0) / max(smp, 1),
                                                          100 * float(r[hdr["Instructions Executed"]] or 0) / max(tot, 1),
                                                          r[hdr["Avg. Threads Executed"]], r[hdr["Source"]][:100]))


if __name__ == "__main__":
    main()
