"""Summarise an .ncu-rep into the CSV kept under profiles/: one row per kernel launch with the
metrics the roofline discussion uses.  usage: ncu_summary.py report.ncu-rep out.csv [git-sha] [src-hash]
The first line of the CSV is a comment with the git SHA and the hash of the CUDA sources
(bench.source_hash) of the binary the capture was taken from: bench.py recomputes the hash where it
runs and marks the traffic figure stale when the sources have changed since."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]

rep, out = sys.argv[1], sys.argv[2]
sha = sys.argv[3] if len(sys.argv) > 3 else "unknown"
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import bench
src_hash = sys.argv[4] if len(sys.argv) > 4 else bench.source_hash()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
with open(out, "w", newline="") as f:
    f.write("# git_sha=%s src_hash=%s source=%s\n" % (sha, src_hash, rep.split("/")[-1]))
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + [c for c, _ in cols])
    w.writerow([""] + [units[i] for _, i in cols])
    for r in data:
        w.writerow([r[ki]] + [r[i] for _, i in cols])
print("wrote", out, len(data), "launches")
