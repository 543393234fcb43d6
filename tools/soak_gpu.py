"""Soak test (development aid): the same tables filled again and again must come out bit-identical
every time -- a rare race in the producer/consumer/boundary protocol would show as a differing
digest.  usage: python tools/soak_gpu.py [minutes]"""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import libstb_b200 as stb  # noqa: E402

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
t_end = time.time() + 60 * minutes
S, V = stb.S_STABLE, stb.S_UVTABLE
rng = np.random.default_rng(3)
# (1) config-2 size, S+V: one million sampled cells per fill
N, M, a = 200000, 20000, 0.7
big = stb.Table(N, M, N, M, a, S | V | stb.S_NOMIRROR)
n = rng.integers(3, N + 1, size=1_000_000).astype(np.uint32)
m = np.minimum(rng.integers(2, M + 1, size=1_000_000), n - 1).astype(np.uint32)
ref = None
fills = 0
while time.time() < t_end - 60 * minutes * 0.5:
    big.remake(a)
    d = hashlib.sha256(big.S_batch(n, m).tobytes() + big.V_batch(n, m).tobytes()).hexdigest()
    ref = ref or d
    assert d == ref, f"config-2 fill {fills} differs"
    fills += 1
big.free()
print(f"config 2 S+V: {fills} fills identical", flush=True)
# (2) medium tables of several geometries, whole table hashed
shapes = [(3000, 700, 0.3, None), (5000, 1000, 0.9, "7"), (2500, 2500, 0.5, "1"), (4000, 333, 0.0, "3")]
tabs, refs = [], []
for (N, M, a, k) in shapes:
    if k:
        os.environ["STB_STRIP_K"] = k
    else:
        os.environ.pop("STB_STRIP_K", None)
    tabs.append(stb.Table(N, M, N, M, a, S | V))
    refs.append(None)
rounds = 0
while time.time() < t_end:
    for i, (N, M, a, k) in enumerate(shapes):
        if k:
            os.environ["STB_STRIP_K"] = k
        else:
            os.environ.pop("STB_STRIP_K", None)
        tabs[i].remake(a)
        d = hashlib.sha256(tabs[i].rows(0, 1, N).tobytes() + tabs[i].rows(1, 1, N).tobytes()).hexdigest()
        refs[i] = refs[i] or d
        assert d == refs[i], f"shape {shapes[i]} round {rounds} differs"
    rounds += 1
print(f"medium tables: {rounds} rounds x {len(shapes)} shapes identical")
