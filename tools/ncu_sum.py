import csv, sys, collections
raw=sys.argv[1]
rows=list(csv.reader(open(raw)))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__block_size','l1tex__t_sector_hit_rate.pct','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','sm__inst_executed_pipe_fp64.sum','smsp__inst_executed_pipe_fp64.sum']
for w in hdr:
    if 'pcsamp_warps_issue_stalled' in w and 'not_issued' not in w: want.append(w)
for w in want:
    if w in idx: print('%-80s'%w, ' | '.join('%14s'%r[idx[w]] for r in rows[2:]))
