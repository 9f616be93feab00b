"""pull the handful of metrics the roofline entries quote out of an .ncu-rep (ncu -i rep --page raw --csv) and write
them as `name unit value` lines — small enough to commit under profiles/; bench.py reads dram__bytes_* from there.
usage: ncu_extract.py in.ncu-rep out.txt [note-key note-value ...]   (notes are appended as `key - value` lines)"""
import csv
import io
import subprocess
import sys

WANT = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        f.write(f"# kernel: {name}\n")
        for i, h in enumerate(hdr):
            if h in WANT or h.startswith("dram__bytes"):
                f.write(f"{h} {units[i] or '-'} {r[i]}\n")
    for k, v in zip(sys.argv[3::2], sys.argv[4::2]):
        f.write(f"{k} - {v}\n")
