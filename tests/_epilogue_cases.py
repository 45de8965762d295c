"""Seeded adversarial inputs of the a6 gradient epilogue (utils/prune.py:195-211) and the a1 Binarizer
(models/layers.py:15-19), shared by the processes of tests/test_epilogue_differential_cpu.py: NaN / +-inf gradients
under masked positions (the reference ASSIGNS zero there, it does not multiply), huge and denormal weights, piggymask
values at, next to and far from the threshold."""
import numpy as np

THR = np.float32(5e-3)


def _special(rng, n, scale):
    v = (rng.standard_normal(n) * scale).astype(np.float32)
    r = rng.rand(n)
    v[r < 0.06] = np.nan
    v[(r >= 0.06) & (r < 0.10)] = np.inf
    v[(r >= 0.10) & (r < 0.14)] = -np.inf
    v[(r >= 0.14) & (r < 0.18)] = 0.0
    v[(r >= 0.18) & (r < 0.20)] = -0.0
    return v


def epilogue_cases(n_cases=96):
    rng = np.random.RandomState(60)
    out = []
    for i in range(n_cases):
        shape = [(3, 2, 3, 3), (5, 7), (1, 1, 1, 1), (4, 3, 1, 1)][i % 4]
        n = int(np.prod(shape))
        w = (rng.standard_normal(n) * 10.0 ** rng.randint(-20, 20)).astype(np.float32).reshape(shape)
        g = _special(rng, n, 1.0).reshape(shape) if i % 3 else rng.standard_normal(n).astype(np.float32).reshape(shape)
        gp = _special(rng, n, 1.0).reshape(shape)
        t = rng.randint(0, 5, n).astype(np.uint8).reshape(shape)
        cur = int(rng.randint(1, 5))
        mode = ('finetune', 'prune')[i % 2]
        wd = float([4e-5, 0.0, 0.5][i % 3])
        has_piggy = i % 5 != 0
        out.append((w, g, gp if has_piggy else None, t, cur, mode, wd))
    return out


def binarizer_cases():
    rng = np.random.RandomState(61)
    near = np.array([THR, np.nextafter(THR, np.float32(1)), np.nextafter(THR, np.float32(-1)), 0.0, -0.0, np.nan, np.inf,
                     -np.inf, 1e-45, -1e-45, 0.01, 0.005, 0.0049999998], dtype=np.float32)
    return [near, rng.uniform(0, 0.01, 257).astype(np.float32), _special(rng, 64, 0.01)]
