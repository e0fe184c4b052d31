"""Key metrics of one kernel from an .ncu-rep (ncu --set full) as a small text file for profiles/ (development aid).
usage: python tools/ncu_summary.py report.ncu-rep "header comment" > profiles/ncu_xxx.txt"""
import csv, io, subprocess, sys
KEYS = """dram__bytes_read.sum dram__bytes_read.sum.per_second dram__bytes_write.sum gpu__time_duration.sum launch__block_size
launch__grid_size launch__registers_per_thread launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__cycles_elapsed.max sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.per_cycle_active smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__average_warp_latency_per_inst_issued.ratio""".split()
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u, v = rows[0], rows[1], rows[2]
print("#", sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("Kernel Name,,", v[h.index("Kernel Name")] if "Kernel Name" in h else "")
for i, n in enumerate(h):
    if n in KEYS or ("issue_stalled" in n and n.endswith("per_issue_active.ratio") and float(v[i] or 0) > 0.05):
        print(f"{n},{u[i]},{v[i]}")
