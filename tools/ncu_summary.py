#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline needs."""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_sleeping_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct']
def main(path, extra=()):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:90], ' grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
        for k in list(KEYS) + list(extra):
            ms = [h for h in hdr if h == k] or [h for h in hdr if k in h][:3]
            for m in ms:
                print('   %-75s %s %s' % (m, r[hdr.index(m)], units[hdr.index(m)]))
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
