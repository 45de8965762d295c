"""First-light driver of cpg_b200/csrc/experimental/gemm2cta_probe.cu (cta_group::2 bring-up probe; not part
of the product).  Build the probe library first (command in the .cu header) -> cpg_b200/libcpgb_probe.so.
Compares C = A @ B^T against torch fp32 and times the single-CTA and CTA-pair kernels."""
import ctypes
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, 'cpg_b200', 'libcpgb_probe.so'))
vp, i32 = ctypes.c_void_p, ctypes.c_int
lib.cpgb_probe_gemm.restype = i32
lib.cpgb_probe_gemm.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
DEV = 'cuda:0'
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False


def run(M, N, K, bn, pair, nstage):
    A = torch.randn(M, K, device=DEV)
    B = torch.randn(N, K, device=DEV) * 0.05
    C = torch.full((M, N), float('nan'), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.cpgb_probe_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, bn, pair, nstage, st)
    torch.cuda.synchronize()
    ref = A @ B.t()
    err = ((C - ref).abs().max() / ref.abs().max()).item()
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.cpgb_probe_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, bn, pair, nstage, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    print(f'M{M} N{N} K{K} bn{bn} pair{pair} stages{nstage}: rc {rc} rel err {err:.2e}  {us:7.1f} us  '
          f'{2.0 * M * N * K / us / 1e6:6.1f} TF/s', flush=True)


if __name__ == '__main__':
    cases = [(8192, 256, 2304, 128, 0, 4), (8192, 256, 2304, 128, 1, 4), (8192, 256, 2304, 256, 0, 4),
             (8192, 256, 2304, 256, 1, 4), (32768, 128, 1152, 128, 0, 4), (32768, 128, 1152, 128, 1, 4),
             # three stages: two single CTAs (96 KB) or three pair halves (72 KB) fit on one SM
             (8192, 256, 2304, 128, 0, 3), (8192, 256, 2304, 128, 1, 3), (32768, 128, 1152, 128, 0, 3),
             (32768, 128, 1152, 128, 1, 3)]
    if len(sys.argv) > 1 and sys.argv[1] == 'stages3':
        cases = cases[6:]
    for c in cases:
        try:
            run(*c)
        except Exception as ex:      # a trapped launch poisons the context: stop at the first failure
            print('FAILED', c, type(ex).__name__, str(ex)[:200], flush=True)
            sys.exit(1)
