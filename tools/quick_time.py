"""Quick device timings of the table fill for a few shapes (development aid; bench.py is the
measured benchmark)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libstb_b200 as stb  # noqa: E402


def cells(N, M, v=False):
    return (M * (M - 1) // 2 if v else (M - 1) * (M - 2) // 2) + (N - M) * (M - 1)


def run(N, M, a, flags, reps=3, label=""):
    t0 = time.time()
    t = stb.Table(N, M, N, M, a, flags | stb.S_NOMIRROR)
    t_make = time.time() - t0
    ms = [t.last_fill_ms]
    for _ in range(reps):
        t.remake(a)
        ms.append(t.last_fill_ms)
    c = (cells(N, M) if flags & stb.S_STABLE else 0) + (cells(N, M, True) if flags & stb.S_UVTABLE else 0)
    best = min(ms)
    bpc = 4 if flags & stb.S_FLOAT else 8
    print(f"{label:28s} N={N} M={M} a={a}: make {t_make*1e3:.1f} ms wall; fill ms {['%.3f' % x for x in ms]}"
          f" -> {c/best/1e-3:.3e} cells/s, {c*bpc/best/1e-3/1e9:.1f} GB/s written", flush=True)
    t.free()


if __name__ == "__main__":
    S, V, F = stb.S_STABLE, stb.S_UVTABLE, stb.S_FLOAT
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "shape":  # python tools/quick_time.py shape N M [a] [flags]
        N, M = int(sys.argv[2]), int(sys.argv[3])
        a = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
        fl = int(sys.argv[5]) if len(sys.argv) > 5 else S
        run(N, M, a, fl, label=f"K={os.environ.get('STB_STRIP_K', 'auto')} L={os.environ.get('STB_STRIP_L', 'auto')}")
        sys.exit(0)
    run(10000, 1000, 0.5, S | V, label="C1 S+V")
    run(10000, 1000, 0.5, S, label="C1 S")
    run(50000, 5000, 0.7, S, label="C3-shape S")
    run(50000, 5000, 0.7, S | V, label="C3-shape S+V")
    for k in ("1", "3", "5", "7"):
        os.environ["STB_STRIP_K"] = k
        run(50000, 5000, 0.7, S, label=f"C3-shape S K={k}")
    os.environ.pop("STB_STRIP_K")
    if which != "small":
        run(200000, 20000, 0.7, S, label="C2 S")
        run(200000, 20000, 0.7, S | F, label="C2 S float")
        run(200000, 20000, 0.7, S | V, reps=1, label="C2 S+V")
    if which == "all":
        run(2000, 300, 0.7, S | V | stb.S_MIRROR_ORDER, reps=1, label="mirror small")
        run(10000, 1000, 0.5, S | V | stb.S_MIRROR_ORDER, reps=1, label="mirror C1")
