import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpg_b200 import _lib
lib = _lib.load()
lib.cpgb_debug_cluster_probe.restype = ctypes.c_int
lib.cpgb_debug_cluster_probe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
for z, pdl, smem in ((2, 0, 1024), (4, 1, 1024), (4, 1, 100 * 1024), (8, 1, 199680)):
    out = torch.full((z * 6,), -7, dtype=torch.int32, device='cuda')
    rc = lib.cpgb_debug_cluster_probe(z, pdl, smem, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print('z', z, 'pdl', pdl, 'smem', smem, 'rc', rc, lib.cpgb_last_error() if rc else '')
    for row in out.view(z, 6).tolist():
        print('   rank %d of %d peer-token %d  &static 0x%x  &dyn 0x%x  mapa(own) 0x%x' % tuple(v & 0xffffffff if i >= 3 else v for i, v in enumerate(row)), flush=True)
