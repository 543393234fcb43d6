import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, ctypes as C
import libstb_b200 as stb
import bench
L = stb.lib()
cts = bench.config4_counts()
Cn = 1024
bpar = np.full(cts.I, 10.0)
a0 = 0.05 + 0.9 * (np.arange(Cn) + 0.5) / Cn
r0 = np.array([L.stb_rng48_state(12345 + c) for c in range(Cn)], dtype=np.uint64)
for rep in range(3):
    t0 = time.perf_counter()
    a1, r1, sa = stb.samplea_batch(a0, cts, bpar, r0, loops=1)
    t1 = time.perf_counter()
    b1, r2, sb = stb.sampleb_batch(np.full(Cn, 10.0), cts, 1.1, 20.0, a1, r1, loops=1)
    t2 = time.perf_counter()
    print("samplea %.1f ms (device %.1f, rounds %d, evals %d)  sampleb %.1f ms" % ((t1-t0)*1e3, sa["eval_ms"], sa["rounds"], sa["evals"], (t2-t1)*1e3), flush=True)
