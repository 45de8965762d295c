#!/usr/bin/env python
"""Turn the files tools/r2_final.sh / tools/r2_profile.sh left in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/finalize_profiles.py"""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def launches():
    src = os.path.join(G, 'r2_launches.csv')
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if 'stage_weights_batched' in r['Kernel Name']]
    a, b, nsteps = idx[1], idx[-1], len(idx) - 2
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[a:b]:
        name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r['Kernel Name']))[:100]
        agg[name][0] += 1
        agg[name][1] += float(r['Metric Value'].replace(',', ''))
    tot = sum(v for _, v in agg.values())
    out = ['# profiles/r2_launches.csv, rows between the 2nd and the last stage_weights_batched launch: %d launches, '
           '%.1f us per step (ncu: serialised, cold caches), %d whole steps' % (b - a, tot / 1e3 / nsteps, nsteps),
           '# %10s %14s %7s  kernel' % ('us/step', 'launches/step', 'share')]
    ours = 0.0
    for name, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        ours += v if name.startswith('cpgb::') else 0.0
        out.append('  %10.1f %14.1f %6.1f%%  %s' % (v / 1e3 / nsteps, c / nsteps, 100 * v / tot, name))
    out.append('# cpgb:: kernels: %.1f%% of device time' % (100 * ours / tot))
    open(os.path.join(P, 'r2_launches_summary.txt'), 'w').write('\n'.join(out) + '\n')
    shutil.copy(src, os.path.join(P, 'r2_launches.csv'))
    print('\n'.join(out[:14]))


def bench_lines():
    for src, dst in (('r2_final_bench.json', 'r2_bench_line.json'), ('r2_final_bench_reference.json', 'r2_bench_line_reference.json'),
                     ('r2_final_vgg16_prune_cycle.json', 'r2_bench_line_prune_cycle.json'),
                     ('r2_final_spherenet20.json', 'r2_bench_line_spherenet20.json'),
                     ('r2_final_resnet50.json', 'r2_bench_line_resnet50.json')):
        p = os.path.join(G, src)
        if not os.path.exists(p):
            continue
        last = [l for l in open(p).read().strip().splitlines() if l.startswith('{')][-1]
        d = json.loads(last)
        open(os.path.join(P, dst), 'w').write(json.dumps(d) + '\n')
        print(dst, {k: d.get(k) for k in ('value', 'ms_per_step', 'loss')}, 'e2e', (d.get('e2e') or {}).get('value'))


if __name__ == '__main__':
    launches()
    bench_lines()
