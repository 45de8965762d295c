"""Per-layer roofline floors of the 15 sharable VGG16 layers ("report per kernel max(t_flop, t_hbm)") from
the per-layer device times of a committed bench line.

    python profiles/per_layer_roofline.py profiles/r2_bench_line.json > profiles/r2_per_layer_roofline.txt

t_flop = algorithmic FLOPs of the pass / TF32 peak (bf16_tflops / 2 of MEASURED_PEAKS.json: tcgen05 kind::tf32 issues at
half the kind::f16 rate); t_hbm = algorithmic bytes of the pass / measured HBM copy bandwidth; floor = max of the two;
frac = floor / measured time.  Algorithmic bytes (task-1 regime, batch 128, fp32): fprop reads X and W and
writes Y; dgrad reads dY and W and writes dX; wgrad + epilogue reads X, dY, W (4 B) and T (1 B) and writes dW.
The measured times are the bench's cold-L2, event-timed launches (kernels.per_layer of the JSON line)."""
import json
import re
import sys

BATCH = 128


def shape(name):
    m = re.match(r'conv(\d+)x(\d+)@(\d+)$', name)
    if m:
        c, k, hw = map(int, m.groups())
        return c, k, hw, 9
    m = re.match(r'fc(\d+)x(\d+)$', name)
    c, k = map(int, m.groups())
    return c, k, 1, 1


def main(path):
    line = json.load(open(path))
    peaks = line.get('peaks', {})
    tf32 = float(line['roofline']['peak']) * 1e12
    hbm = float(peaks.get('hbm_gbs', 6549.1)) * 1e9
    rows = []
    tot = {p: [0.0, 0.0] for p in ('fprop', 'dgrad', 'wgrad')}
    for lay in line['kernels']['per_layer']:
        c, k, hw, rs = shape(lay['layer'])
        n_w = lay['n_weights']
        x = BATCH * c * hw * hw * 4
        y = BATCH * k * hw * hw * 4
        byts = {'fprop': x + y + 4 * n_w, 'dgrad': x + y + 4 * n_w, 'wgrad': x + y + 9 * n_w}
        for p in ('fprop', 'dgrad', 'wgrad'):
            ms = lay[p + '_ms']
            if ms <= 0:
                continue
            t_flop = lay['flop'] / tf32 * 1e3
            t_hbm = byts[p] / hbm * 1e3
            floor = max(t_flop, t_hbm)
            rows.append((lay['layer'], p, lay['flop'] / 1e9, byts[p] / 1e6, t_flop * 1e3, t_hbm * 1e3,
                         'tensor' if t_flop >= t_hbm else 'hbm', ms * 1e3, floor / ms))
            tot[p][0] += floor
            tot[p][1] += ms
    print('peaks: TF32 %.2f TFLOP/s (bf16_tflops / 2), HBM %.1f GB/s; batch %d, task-1 regime; source %s' %
          (tf32 / 1e12, hbm / 1e9, BATCH, path))
    print('%-16s %-6s %8s %8s %9s %9s %-6s %9s %6s' % ('layer', 'pass', 'GFLOP', 'MB', 't_flop us', 't_hbm us', 'bound',
                                                      'meas us', 'frac'))
    for r in rows:
        print('%-16s %-6s %8.2f %8.1f %9.1f %9.1f %-6s %9.1f %6.2f' % r)
    print()
    for p in ('fprop', 'dgrad', 'wgrad'):
        print('%-6s sum of floors %.3f ms, measured %.3f ms, floor / measured = %.2f' % (p, tot[p][0], tot[p][1],
                                                                                         tot[p][0] / tot[p][1]))
    allf, allm = sum(v[0] for v in tot.values()), sum(v[1] for v in tot.values())
    print('all    sum of floors %.3f ms, measured %.3f ms, floor / measured = %.2f' % (allf, allm, allf / allm))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'profiles/r2_bench_line.json')
