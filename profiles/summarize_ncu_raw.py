#!/usr/bin/env python
"""Compact table out of `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch).
usage: ncu -i rep.ncu-rep --page raw --csv | python profiles/summarize_ncu_raw.py"""
import csv
import sys

WANT = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('launch__registers_per_thread', 'regs'), ('launch__waves_per_multiprocessor', 'waves')]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f'{"kernel":46s} {"grid":14s} ' + ' '.join(f'{n:>9s}' for _, n in WANT))
    for r in data:
        name = r[idx['Kernel Name']].replace('void ', '').split('(')[0][:46]
        grid = r[idx['Grid Size']].replace(' ', '')
        vals = []
        for k, _ in WANT:
            v = r[idx[k]] if k in idx else ''
            u = units[idx[k]] if k in idx else ''
            try:
                f = float(v.replace(',', ''))
                if u == 'Mbyte':
                    v = f'{f:.1f}MB'
                elif u == 'Kbyte':
                    v = f'{f / 1e3:.1f}MB'
                elif u == 'byte':
                    v = f'{f / 1e6:.1f}MB'
                elif u in ('ns', 'nsecond'):
                    v = f'{f / 1e3:.1f}'
                elif u in ('us', 'usecond'):
                    v = f'{f:.1f}'
                else:
                    v = f'{f:.1f}'
            except ValueError:
                pass
            vals.append(f'{v:>9s}')
        print(f'{name:46s} {grid:14s} ' + ' '.join(vals))


if __name__ == '__main__':
    main()
