"""Prints the metrics that matter from an .ncu-rep (run here, no GPU needed): ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_bytes.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_sleeping_per_warp_active.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_ltcfabric.sum', 'lts__t_sectors_lookup_miss.sum', 'lts__t_sectors_lookup_hit.sum',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'lts__d_sectors_fill_device.sum', 'lts__t_sectors_srcnode_gpc.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            print(f'   {w:85s} {r[hdr.index(w)]:>18s} {rows[1][hdr.index(w)]}')
    if len(sys.argv) > 2:
        for i, h in enumerate(hdr):
            if sys.argv[2] in h:
                print(f'   {h:85s} {r[i]:>18s} {rows[1][i]}')
