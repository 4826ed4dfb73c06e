"""Print the metrics we track from an `ncu --set full` report exported with --page raw --csv."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_global_st.sum', 'smsp__inst_executed_op_global_ld.sum', 'lts__t_bytes.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'sm__cycles_elapsed.max', 'launch__grid_size']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]: {[r[i] for r in data]}")
st = [(h, [r[i] for r in data]) for i, h in enumerate(hdr)
      if re.search(r'smsp__average_warps_issue_stalled_.*_per_issue_active', h)]
st.sort(key=lambda x: -float(x[1][0].replace(',', '')))
print("-- warp stall reasons (avg warps stalled per issue-active cycle) --")
for h, v in st[:10]:
    print(" ", h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
