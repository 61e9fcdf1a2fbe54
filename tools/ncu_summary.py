"""Turns gpurun_out/*.ncu-rep and the launch-list CSV into the small text summaries kept under profiles/."""
import csv
import subprocess
import sys
from collections import defaultdict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_active.avg",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, source report: {rep}\n")
        for r in rows[2:]:
            f.write(f"\n## {r[idx['Kernel Name']][:100]}\n")
            for k in KEEP:
                if k in idx:
                    f.write(f"{k:80s} {r[idx[k]]:>16s} {units[idx[k]]}\n")
    print("wrote", out)


def launches(csv_path, out):
    rows = [r for r in csv.reader(open(csv_path)) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, cnt = defaultdict(float), defaultdict(int)
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0][:70]
        agg[name] += float(r[mv].replace(",", ""))
        cnt[name] += 1
    tot = sum(agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {csv_path}\n")
        f.write("# per-kernel totals over the captured launches (cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write(f"{k:72s} {cnt[k]:8d} {v / 1e3:12.1f} {v / 1e3 / cnt[k]:10.1f} {100 * v / tot:6.1f}%\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2], sys.argv[3])
