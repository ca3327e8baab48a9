"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

usage: python profiles/summarize.py launches <launches.csv> > profiles/<name>.txt
       python profiles/summarize.py full <report.ncu-rep>    > profiles/<name>.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(x["Metric Unit"], v)
        name = x["Kernel Name"].split("(")[0][:70]
        tot[name] += v
        cnt[name] += 1
    all_us = sum(tot.values())
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)")
    print("%-72s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for n in sorted(tot, key=lambda k: -tot[k]):
        print("%-72s %6d %12.1f %10.1f %6.1f%%" % (n, cnt[n], tot[n], tot[n] / cnt[n], 100 * tot[n] / all_us))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# selected metrics from `ncu --set full --clock-control none` (%s)" % path.split("/")[-1])
    for r in rows[2:]:
        print("\n== %s  (id %s)" % (r[idx["Kernel Name"]][:90], r[idx["ID"]]))
        for m in METRICS:
            if m in idx:
                print("  %-88s %16s %s" % (m, r[idx[m]], units[idx[m]]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
