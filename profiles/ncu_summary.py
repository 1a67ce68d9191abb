import csv, sys, subprocess
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; units = rows[1]
want = ['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.per_cycle_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','sm__cycles_elapsed.max','lts__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w, '|', units[i], '|', ' | '.join(r[i][:48] for r in rows[2:]))
    else:
        print(w, 'MISSING')
