"""Key metrics of every kernel in an `ncu --page raw --csv` export -> text (profiles/*.txt)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_requests_srcunit_tex_op_red.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_shared_mem']
ki = hdr.index('Kernel Name')
units = rows[1]
for r in rows[2:]:
    if len(r) <= ki:
        continue
    print(r[ki].split('(')[0])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print('    %-76s %s %s' % (w, r[i], units[i]))
