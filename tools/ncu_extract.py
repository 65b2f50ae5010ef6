"""Pull the judged metrics out of an `ncu --page raw --csv` dump (one kernel) into a small metric,unit,value csv."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import csv
import sys

WANT = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "sm__cycles_elapsed.max.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
print("metric,unit,value")
for name in WANT:
    if name in hdr:
        i = hdr.index(name)
        print(f'{name},{units[i]},"{vals[i]}"')
