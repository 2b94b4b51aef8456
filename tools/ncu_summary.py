#!/usr/bin/env python
"""Key metrics of `ncu --set full` captures (run here, on the CPU box, on .ncu-rep files brought back in gpurun_out/):
    python tools/ncu_summary.py gpurun_out/prof_match_tc3.ncu-rep gpurun_out/prof_score.ncu-rep > profiles/r1_ncu_summary.txt"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg"]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"== {path.split('/')[-1]}")
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(w, r[i], units[i])
        print("---")
