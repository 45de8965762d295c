"""tcgen05.mma kind::tf32 issue-rate microbenchmark (GPU box only)."""
import ctypes
import torch
from cpg_b200 import _lib

lib = _lib.load()
lib.cpgb_debug_mma_rate.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p]
lib.cpgb_debug_mma_rate.restype = ctypes.c_int
out = torch.zeros(1, dtype=torch.int64, device='cuda:0')
iters = 2000
for grid in (1, 148):
    for bn in (64, 128, 256):
        for a_mn, b_mn in ((0, 0), (0, 1), (1, 1)):
            rc = lib.cpgb_debug_mma_rate(bn, a_mn, b_mn, iters, grid, out.data_ptr(), _lib.stream_ptr())
            torch.cuda.synchronize()
            cyc = out.item() / (iters * 4)
            print(f'grid {grid:3d} M128 N{bn:3d} K8  A_{"MN" if a_mn else "K "} B_{"MN" if b_mn else "K "}: {cyc:7.1f} clk/MMA  '
                  f'({128 * bn * 8 / cyc:7.0f} MAC/clk/SM, ideal {128 * bn / 256:.0f} clk)', flush=True)
