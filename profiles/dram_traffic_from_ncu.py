#!/usr/bin/env python
"""DRAM traffic per pass out of
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file traffic.csv python -m tools.prof_ops 1 vgg
(the 15 sharable layers of VGG16-BN-cifar at batch 128, each pass launched once, no piggymask).
usage: python profiles/dram_traffic_from_ncu.py traffic.csv > profiles/r1_dram_traffic.json"""
import collections
import csv
import json
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1,
        'ms': 1e3, 'msecond': 1e3}
kern = collections.OrderedDict()
for r in rows:
    k = kern.setdefault(r['ID'], {'name': r['Kernel Name']})
    k[r['Metric Name']] = float(r['Metric Value'].replace(',', '')) * UNIT.get(r['Metric Unit'], 1)

out = {p: {'dram_bytes': 0.0, 'us_under_ncu': 0.0, 'launches': 0} for p in ('stage', 'fprop', 'dgrad', 'wgrad')}
last = 'fprop'
for k in kern.values():
    name = k['name']
    if 'stage_weights' in name:
        pas = 'stage'
    elif 'stem_fprop' in name:
        pas = last = 'fprop'
    elif 'conv_gemm_kernel' in name:
        m = re.search(r'conv_gemm_kernel<\s*\d+,\s*(\w+)', name)      # second template argument: B_MN (dgrad)
        pas = 'dgrad' if (m and m.group(1) in ('1', 'true')) else 'fprop'
        last = pas
    elif 'splitk_reduce' in name or 'im2col' in name or 'col2im' in name:
        pas = last
    elif 'wgrad' in name:
        pas = last = 'wgrad'
    else:
        continue
    out[pas]['dram_bytes'] += k.get('dram__bytes_read.sum', 0.0) + k.get('dram__bytes_write.sum', 0.0)
    out[pas]['us_under_ncu'] += k.get('gpu__time_duration.sum', 0.0)
    out[pas]['launches'] += 1
out['note'] = ('sum over the 15 sharable layers of VGG16-BN-cifar, batch 128, one launch of each pass; '
               'dram__bytes_read.sum + dram__bytes_write.sum per kernel from ncu (cold caches per replay)')
json.dump(out, sys.stdout, indent=1)
print()
