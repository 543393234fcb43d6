"""config 4's statistics through the seat-partition kernel, for ncu captures:
python tools/prof_partition.py [exact(0|1)] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import libstb_b200 as stb  # noqa: E402

exact = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cts = bench.config4_counts()
n = np.concatenate(cts.n_rows)
t = np.concatenate(cts.t_rows)
keep = (t > 1) & (t < n)
n, t = n[keep], t[keep]
maxn, maxt = int(n.max()) + 1, int(t.max()) + 1
tab = stb.Table(maxn, maxt, maxn, maxt, 0.5, stb.S_STABLE)
rng = np.random.default_rng(4)
logu = np.log(rng.random(int((t.astype(np.int64) - 1).sum()) if exact else n.shape[0]))
for _ in range(reps):
    tab.partition_sample(0.5, n, t, logu, exact=exact)
print("nodes", n.shape[0], "kernel ms", tab.last_partition_ms)
tab.free()
