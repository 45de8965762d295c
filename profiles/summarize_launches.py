#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv [n_steps_in_capture] > summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    nsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'^void ', '', name)[:100]
        agg[name][0] += 1
        agg[name][1] += float(r['Metric Value'].replace(',', ''))
    tot = sum(v for _, v in agg.values())
    print(f'# {path}: {len(rows)} launches, {tot / 1e3:.1f} us total, {nsteps:g} steps in capture')
    print(f'# {"us/step":>10s} {"launches/step":>14s} {"share":>7s}  kernel')
    ours = 0.0
    for name, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if name.startswith('cpgb::'):
            ours += v
        print(f'  {v / 1e3 / nsteps:10.1f} {c / nsteps:14.1f} {100 * v / tot:6.1f}%  {name}')
    print(f'# cpgb:: kernels: {100 * ours / tot:.1f}% of device time')


if __name__ == '__main__':
    main()
