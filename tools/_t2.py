import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import libstb_b200 as stb, bench
from tests import harness
L = stb.lib()
R = C.CDLL(harness.REF_SO)
d, u32p = C.c_double, C.POINTER(C.c_uint32)
R.samplea.restype = d
R.samplea.argtypes = [d, C.c_int, C.POINTER(C.c_int), u32p, C.POINTER(u32p), C.POINTER(C.POINTER(C.c_uint16)), C.c_void_p, C.POINTER(d), C.c_void_p, C.c_int, C.c_int]
libc = C.CDLL(None); libc.srand.argtypes = [C.c_uint]; libc.rand.restype = C.c_int
cts = bench.config4_counts()
bpar = np.full(cts.I, 10.0)
dp = C.POINTER(C.c_double)
idx = [0, 1, 2, 500, 913, 1000]
a0 = np.array([0.05 + 0.9 * (c + 0.5) / 1024 for c in idx])
seeds = [777 + c for c in idx]
a1, rnd, st = stb.samplea_batch_ars(a0, cts, bpar, stb.rand31_states(seeds), trace_cap=16)
for j, c in enumerate(idx):
    libc.srand(seeds[j])
    a_ref = R.samplea(float(a0[j]), *cts.args(), None, bpar.ctypes.data_as(dp), None, 1, 0)
    r_ref = libc.rand()
    r_our = L.stb_rand31_next(rnd.ctypes.data + j * stb.RAND31_DTYPE.itemsize)
    print(c, "a0 %.6f ours %.12f ref %.12f rel diff %.2e  stream in step: %s" % (a0[j], a1[j], a_ref, abs(a1[j]-a_ref)/a_ref, r_ref == r_our))
tr = st["trace"]
print({k: (v[:2] if hasattr(v, '__len__') else v) for k, v in tr.items()} if isinstance(tr, dict) else type(tr))
