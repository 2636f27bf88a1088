import csv,sys,subprocess
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','launch__occupancy_limit_blocks','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__warps_eligible.avg.per_cycle_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__waves_per_multiprocessor','launch__grid_size','launch__block_size','sm__cycles_elapsed.max','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_dispatch_stall','smsp__pcsamp_warps_issue_stalled_imc_miss','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','dram__bytes_read.sum.per_second']
for rep in sys.argv[1:]:
    out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr,units,vals=rows[0],rows[1],rows[2]
    print('==',rep.split('/')[-1])
    for h,u,v in zip(hdr,units,vals):
        if h in want: print(f'  {h} = {v} {u}')
