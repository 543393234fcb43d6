"""Device timing of a discount sweep (development aid): python tools/quick_sweep.py N M na"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libstb_b200 as stb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
na = int(sys.argv[3]) if len(sys.argv) > 3 else 64
cells = (M - 1) * (M - 2) // 2 + (N - M) * (M - 1)
rng = np.random.default_rng(1)
n = rng.integers(3, N + 1, size=100000).astype(np.uint32)
m = np.minimum(rng.integers(2, M + 1, size=100000), n - 1).astype(np.uint32)
w = stb.Sweep(N, M)
w.set_pairs(n, m)
a = (np.arange(na) + 0.5) / na
for rep in range(3):
    t0 = time.time()
    g, s, _ = w.run(a, gather=False, sums=True)
    wall = time.time() - t0
    ms = w.last_fill_ms
    print(f"N={N} M={M} na={na} tables/launch={w.tables_in_flight}: fill {ms:.2f} ms device, {wall*1e3:.1f} ms wall -> "
          f"{cells*na/ms/1e-3:.3e} cells/s ({cells*na*8/ms/1e-3/1e9:.0f} GB/s)", flush=True)
w.free()
