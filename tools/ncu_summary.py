"""Key counters of an ncu --set full capture: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt (runs without a GPU)."""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "pipe_fp64", "pipe_tensor", "dmma",
        "bank_conflict", "issue_stalled_barrier", "issue_stalled_long_scoreboard", "issue_stalled_short_scoreboard", "issue_stalled_membar",
        "issue_stalled_math_pipe", "issue_stalled_wait", "issue_stalled_mio", "issue_stalled_lg_throttle", "warps_eligible", "one_or_more_eligible"]
print(f"# {sys.argv[1]}")
for r in rows[2:]:
    print(f"== {r[hdr.index('Kernel Name')][:100]}  (id {r[0]})")
    for i, h in enumerate(hdr):
        if any(w in h for w in want) and r[i] not in ("", "n/a"):
            print(f"  {h} = {r[i]} {units[i]}")
