#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.
usage: scripts/ncu_summary.py <file.ncu-rep> [metric-substring ...]"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
    'lts__t_bytes.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_global_op_ld.sum',
    'l1tex__data_pipe_lsu_wavefronts.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
    'smsp__cycles_active.avg', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'l1tex__m_xbar2l1tex_read_sectors.sum', 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
]


def main():
    path = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('--- kernel', d.get('Kernel Name'), 'launch id', d.get('ID'))
        for k in hdr:
            if k in KEYS or any(e in k for e in extra):
                print(f'{k:90s} {d[k]:>22s} {units[hdr.index(k)]}')


if __name__ == '__main__':
    main()
