"""One table fill for ncu captures: python tools/prof_fill.py N M a flags reps"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libstb_b200 as stb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
a = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
flags = int(sys.argv[4]) if len(sys.argv) > 4 else stb.S_STABLE
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
t = stb.Table(N, M, N, M, a, flags | stb.S_NOMIRROR)
for _ in range(reps):
    t.remake(a)
print("fill ms", t.last_fill_ms)
t.free()
