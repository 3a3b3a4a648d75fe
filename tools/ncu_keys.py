"""Print the key metrics of an ncu report: usage ncu_keys.py report.ncu-rep [launch-index]"""
import csv, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h, u, r = rows[0], rows[1], rows[2 + idx]
want = """gpu__time_duration.sum launch__registers_per_thread launch__grid_size launch__block_size launch__occupancy_limit_registers
launch__shared_mem_per_block_dynamic sm__warps_active.avg.pct_of_peak_sustained_active dram__bytes_read.sum dram__bytes_write.sum
dram__throughput.avg.pct_of_peak_sustained_elapsed lts__t_bytes.sum lts__t_sectors.sum lts__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__t_sector_hit_rate.pct l1tex__throughput.avg.pct_of_peak_sustained_elapsed l1tex__t_bytes.sum
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio sm__cycles_elapsed.max sm__throughput.avg.pct_of_peak_sustained_elapsed
sass__inst_executed_local_loads sass__inst_executed_local_stores l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct
l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum smsp__sass_thread_inst_executed_op_dadd_pred_on.sum smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2 smsp__inst_executed_pipe_fp64.sum""".split()
for k, uu, v in zip(h, u, r):
    if k in want or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") or k.startswith("smsp__average_warp_latency"):
        print("%-95s %-16s %s" % (k, uu, v))
