"""Summarise ncu outputs into profiles/ (tracked): launch-list shares and selected `--set full` metrics per kernel.
Usage: python tools/ncu_summary.py <tag> [launches.csv] [prof.ncu-rep]"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "sm__icc_request_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) > vi:
            v = float(r[vi].replace(",", ""))
            v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
            agg[r[ki].split("(")[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    out.append("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n")
    out.append("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append("| %s | %d | %.3f | %.4f | %.1f%% |" % (k[:70], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    out.append("")


def full(path, out):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    out.append("## ncu --set full (per captured launch)\n")
    for r in rows[2:]:
        out.append("### %s\n" % r[H.index("Kernel Name")].split("(")[0])
        out.append("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in H:
                out.append("| %s | %s | %s |" % (m, r[H.index(m)], units[H.index(m)]))
        out.append("")


if __name__ == "__main__":
    tag = sys.argv[1]
    out = ["# ncu summary %s\n" % tag]
    if len(sys.argv) > 2 and sys.argv[2] != "-":
        launches(sys.argv[2], out)
    if len(sys.argv) > 3:
        full(sys.argv[3], out)
    open("profiles/%s.md" % tag, "w").write("\n".join(out) + "\n")
    print("wrote profiles/%s.md" % tag)
