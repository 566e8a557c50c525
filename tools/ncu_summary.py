"""Print the metrics that matter from `ncu -i X.ncu-rep --page raw --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.avg', 'smsp__average_warp_latency_issue_stalled',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu', 'smsp__inst_executed_pipe_lsu',
        'l1tex__lsu_writeback_active', 'l1tex__m_xbar2l1tex_read_bytes', 'lts__t_bytes.sum', 'smsp__pcsamp_warps_issue_stalled',
        'smsp__warps_issue_stalled', 'sm__inst_executed_pipe']
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print('==', name[:120])
    for i, h in enumerate(hdr):
        if any(h == w or h.startswith(w) for w in WANT) and 'pct_of_peak_sustained_elapsed' not in h.replace('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '').replace('sm__throughput.avg.pct_of_peak_sustained_elapsed','').replace('l1tex__throughput.avg.pct_of_peak_sustained_elapsed','').replace('lts__throughput.avg.pct_of_peak_sustained_elapsed',''):
            try:
                if float(vals[i].replace(',', '')) == 0: continue
            except ValueError: pass
            print(f"  {h:90s} {units[i]:16s} {vals[i]}")
