"""Where the wall time of a bench step goes (development aid): S_remake wall vs kernel, stb_S_batch wall."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libstb_b200 as stb
L = stb.lib()
N, M, a = 200000, 20000, 0.7
t = stb.Table(N, M, N, M, a, stb.S_STABLE | stb.S_NOMIRROR)
rng = np.random.default_rng(1)
n_np = rng.integers(3, N + 1, size=100000).astype(np.uint32)
m_np = np.minimum(rng.integers(2, M + 1, size=100000), n_np - 1).astype(np.uint32)
n_h = torch.from_numpy(n_np.view(np.int32)).pin_memory(); m_h = torch.from_numpy(m_np.view(np.int32)).pin_memory()
out_h = torch.empty(100000, dtype=torch.float64).pin_memory()
u32p, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
n_p, m_p, out_p = C.cast(n_h.data_ptr(), u32p), C.cast(m_h.data_ptr(), u32p), C.cast(out_h.data_ptr(), dp)
for _ in range(3):
    L.S_remake(t.sp, a); L.stb_S_batch(t.sp, n_p, m_p, out_p, 100000)
tr, tb, tk = [], [], []
for _ in range(10):
    t0 = time.perf_counter(); L.S_remake(t.sp, a); t1 = time.perf_counter()
    L.stb_S_batch(t.sp, n_p, m_p, out_p, 100000); t2 = time.perf_counter()
    tr.append(t1 - t0); tb.append(t2 - t1); tk.append(L.stb_last_fill_ms(t.sp))
print("S_remake wall %.3f ms (kernel %.3f ms) ; stb_S_batch wall %.3f ms" % (1e3*np.median(tr), np.median(tk), 1e3*np.median(tb)))
