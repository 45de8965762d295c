#!/usr/bin/env python
"""SASS opcode histogram per kernel of cpg_b200/libcpgb200.so (cuobjdump -sass): which kernels really carry tcgen05
(UTCHMMA / UTCBAR / LDTM), TMA (UTMALDG / UBLKCP), cluster (UCGABAR / MAPA-class) and packed fp32 (FFMA2) instructions.
usage: python tools/sass_histogram.py [lib] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMAPF', 'UBLKCP', 'SYNCS', 'UCGABAR', 'MAPA', 'FFMA2', 'FFMA', 'HMMA',
       'ATOMS', 'ATOMG', 'RED', 'REDUX', 'SHFL', 'LDG', 'STG', 'LDS', 'STS', 'BAR', 'ACQBULK', 'CCTL', 'DFMA', 'MUFU']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'cpg_b200', 'libcpgb200.so')
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    demangle = {}
    hist = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    names = list(hist)
    try:
        dm = subprocess.run(['cu++filt'] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    print('# %s: %d kernels; columns = instruction counts in the SASS of each kernel (static, not executed counts)' %
          (os.path.basename(lib), len(names)))
    tot = collections.Counter()
    rows = []
    for n in names:
        h = hist[n]
        d = demangle.get(n, n)
        if d.endswith(')'):                      # drop the argument list: the parenthesis that matches the last one
            depth = 0
            for i in range(len(d) - 1, -1, -1):
                depth += d[i] == ')'
                depth -= d[i] == '('
                if depth == 0:
                    d = d[:i]
                    break
        d = re.sub(r'\((?:int|bool)\)', '', d)
        d = re.sub(r'^void ', '', d)
        d = d.replace('cpgb::', '').replace('(anonymous namespace)::', '')
        cols = [(k, sum(v for op, v in h.items() if op == k or op.startswith(k + '.') or (k in ('UTMALDG', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'BAR', 'UTCHMMA', 'UTCBAR', 'LDTM', 'UBLKCP', 'ATOMS', 'ATOMG', 'RED', 'MUFU', 'SYNCS') and op.startswith(k)))) for k in KEY]
        rows.append((d, sum(h.values()), cols))
        for k, v in cols:
            tot[k] += v
    rows.sort(key=lambda r: -dict(r[2])['UTCHMMA'] * 100000 - r[1])
    for d, n_ins, cols in rows:
        nz = ' '.join('%s=%d' % (k, v) for k, v in cols if v)
        print('%-78s %6d instr | %s' % (d[:78], n_ins, nz))
    print('# totals: ' + ' '.join('%s=%d' % (k, tot[k]) for k in KEY if tot[k]))


if __name__ == '__main__':
    main()
