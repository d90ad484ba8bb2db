"""Turns an ncu report into the compact per-launch CSV kept under profiles/.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/rN_ncu_full_<cfg>_summary.csv
  python tools/ncu_summary.py --launches gpurun_out/x_launches.csv profiles/rN_launches_<cfg>.csv
"""
import csv
import subprocess
import sys

METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def full(rep, out):
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr = rows[0]
    cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[c] for c in cols])
    print(f"{out}: {len(rows) - 2} launches")


def launches(src, out):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    body = [(r[ki], float(r[vi].replace(",", "")) * scale[r[ui]]) for r in rows[1:]]
    tot = sum(v for _, v in body)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "gpu__time_duration_us", "share_of_step_pct"])
        for i, (k, v) in enumerate(body):
            w.writerow([i, k, f"{v:.3f}", f"{100 * v / tot:.2f}"])
        agg = {}
        for k, v in body:
            agg[k] = agg.get(k, 0.0) + v
        w.writerow([])
        w.writerow(["total", "kernel", "sum_us", "share_pct"])
        for k, v in agg.items():
            w.writerow(["", k, f"{v:.3f}", f"{100 * v / tot:.2f}"])
    print(f"{out}: {len(body)} launches, {tot:.1f} us")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
