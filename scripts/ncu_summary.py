"""Prints the metrics we care about from an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__maximum_warps_per_active_cycle_pct',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'smsp__inst_executed_op_global_ld.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for d in data:
    print('---', d[hdr.index('Kernel Name')][:70])
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"  {w} = {d[i]} {units[i]}")
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            try:
                v = float(d[i])
                if v > 2:
                    print(f"  {h.replace('smsp__average_warp_latency_issue_stalled_','').replace('smsp__average_warps_issue_stalled_','')} = {v:.1f}")
            except ValueError:
                pass
