"""Print the key raw metrics of an .ncu-rep (first kernel)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__shared_mem_per_block_dynamic','launch__block_size','launch__grid_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg','sm__cycles_elapsed.avg.per_second','lts__t_sectors_op_read.sum','lts__t_bytes.sum','l1tex__t_sector_hit_rate.pct',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__maximum_warps_per_active_cycle_pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want: print(f"{h:80s} {vals[i]:>18s} {units[i]}")
