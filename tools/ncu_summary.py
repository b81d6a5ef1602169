"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers DESIGN.md quotes."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__cluster_dim_x', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print(f'## {name[:110]}')
        for k in KEYS:
            if k in hdr:
                print(f'  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')


if __name__ == '__main__':
    main(sys.argv[1])
