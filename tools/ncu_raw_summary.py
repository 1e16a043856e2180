"""Print the key raw metrics of every kernel in an .ncu-rep (via `ncu --page raw --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "smsp__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_tensor.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=====", name[:110])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:72s} {r[i]:>18s} {units[i]}")
