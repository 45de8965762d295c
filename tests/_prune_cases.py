"""Seeded adversarial inputs of the a7 prune step (utils/prune.py:30-53), shared by the processes of
tests/test_prune_differential_cpu.py: heavy ties, zeros left by apply_mask, quantised values, +-inf, NaN, empty and tiny
pools, and ratios whose k = round(ratio * |pool|) sits on a half (banker's rounding) or outside 1..|pool| (exit 2)."""
import numpy as np


def cases(n_cases=240):
    rng = np.random.RandomState(20261017)
    out = []
    for i in range(n_cases):
        n = int(rng.choice([1, 2, 3, 7, 16, 33, 100, 216, 301]))
        kind = i % 8
        if kind == 0:
            w = rng.standard_normal(n)
        elif kind == 1:
            w = np.round(rng.standard_normal(n) * 4) / 4                 # quantised: many ties
        elif kind == 2:
            w = rng.standard_normal(n) * (rng.rand(n) < 0.5)             # zeros as apply_mask leaves them
        elif kind == 3:
            w = np.full(n, float(rng.choice([0.0, 0.5, -0.5])))          # constant
        elif kind == 4:
            w = rng.standard_normal(n)
            w[rng.rand(n) < 0.2] = np.nan
        elif kind == 5:
            w = rng.standard_normal(n)
            w[rng.rand(n) < 0.1] = np.inf
            w[rng.rand(n) < 0.1] = -np.inf
        elif kind == 6:
            w = np.exp(rng.uniform(-60, 60, n)) * rng.choice([-1.0, 1.0], n)   # denormals .. huge
        else:
            w = rng.standard_normal(n) * 1e-3
            w[::2] = -w[::2][::-1] if n > 1 else w[::2]                  # +-pairs: equal magnitudes
        w = w.astype(np.float32)
        cur = int(rng.randint(1, 4))
        t = rng.randint(0, 4, n).astype(np.uint8)
        if i % 11 == 0:
            t[:] = (cur % 3) + 1 if cur != (cur % 3) + 1 else 0          # pool may be empty
        m = int(((t == cur) | (t == 0)).sum())
        pick = i % 6
        if pick == 0 and m > 0:
            ratio = (int(rng.randint(0, m)) + 0.5) / m                   # k on a half
        elif pick == 1:
            ratio = 0.0
        elif pick == 2:
            ratio = 1.0
        elif pick == 3 and m > 0:
            ratio = 0.4375 if m == 216 else 1.0 / (2 * m)                # 94.5 / 0.5: rounds to even
        else:
            ratio = float(rng.uniform(0, 1))
        out.append((w, t, cur, float(ratio)))
    return out
