"""Generates tests/golden/tables.json from the UNMODIFIED reference (oracle/_ref/libstb_ref.so,
built by oracle/build_ref.sh from /root/reference/lib).  Run in the build container:

    python tests/golden/make_golden.py

The reference ships no golden files, so these vectors are the outputs of
the reference itself: spot cells printed with 17 significant digits (repr round-trips doubles
exactly) and SHA-256 digests of whole tables read cell by cell through S_S/S_V.
"""
import hashlib
import json
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import harness  # noqa: E402

L = harness.ref()
FLAGS = harness.S_STABLE | harness.S_UVTABLE
out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libstb_ref.so (wbuntine/libstb, lib/Makefile flags)",
       "tables": []}

CONFIGS = [
    # (N, M, a, spot cells)
    (10, 10, 0.0, [(10, 3), (10, 2), (5, 4), (3, 2), (10, 9), (7, 1)]),
    (10, 10, 0.3, [(5, 5), (5, 4), (3, 2), (10, 9), (10, 2)]),
    (200, 60, 0.5, [(200, 60), (200, 2), (100, 50), (61, 60), (60, 59), (150, 33)]),
    (200, 60, 0.01, [(200, 60), (200, 2), (100, 50), (61, 60)]),
    (200, 60, 0.98, [(200, 60), (200, 2), (100, 50), (61, 60)]),
    (2000, 300, 0.7, [(2000, 300), (2000, 2), (1000, 150), (301, 300), (1999, 299), (1234, 77)]),
    (10000, 1000, 0.5, [(10000, 1000), (10000, 2), (9998, 998), (5000, 500), (1001, 1000), (7777, 3)]),
]

for N, M, a, spots in CONFIGS:
    sp = L.S_make(N, M, N, M, a, FLAGS)
    hS, hV = hashlib.sha256(), hashlib.sha256()
    for n in range(2, N + 1):
        top = min(n, M)
        hS.update(struct.pack("<%dd" % top, *[L.S_S(sp, n, m) for m in range(1, top + 1)]))
        if top >= 2:
            hV.update(struct.pack("<%dd" % (top - 1), *[L.S_V(sp, n, m) for m in range(2, top + 1)]))
    ent = {"N": N, "M": M, "a": a, "sha256_S": hS.hexdigest(), "sha256_V": hV.hexdigest(), "spots": []}
    for n, m in spots:
        ent["spots"].append({"n": n, "m": m, "S": repr(L.S_S(sp, n, m)),
                             "V": repr(L.S_V(sp, n, m)) if m >= 2 else None,  # S_V(n,1) reads out of bounds in the reference
                             "U": repr(L.S_U(sp, n, m)), "UV": repr(L.S_UV(sp, n, m))})
    out["tables"].append(ent)
    L.S_free(sp)

# larger shapes: spot cells only (the full tables are too slow to hash cell by cell from python)
out["big"] = []
for N, M, a, spots in [(50000, 5000, 0.7, [(50000, 5000), (50000, 2), (25000, 2500), (49999, 4999), (5001, 5000)])]:
    sp = L.S_make(N, M, N, M, a, FLAGS)
    ent = {"N": N, "M": M, "a": a, "spots": []}
    for n, m in spots:
        ent["spots"].append({"n": n, "m": m, "S": repr(L.S_S(sp, n, m)), "V": repr(L.S_V(sp, n, m))})
    out["big"].append(ent)
    L.S_free(sp)

# asymptote (S_ASYMPT past maxN) and look-up conventions
asy = []
for a in (0.0, 0.3, 0.7):
    sp = L.S_make(20, 10, 20, 10, a, FLAGS | harness.S_ASYMPT)
    for n, m in [(25, 3), (100, 5), (1000, 10), (100000, 7), (10000000, 4)]:
        asy.append({"a": a, "n": n, "m": m, "S": repr(L.S_S(sp, n, m)), "V": repr(L.S_V(sp, n, m)),
                    "asympt": repr(L.S_asympt(sp, n, m))})
    L.S_free(sp)
out["asympt"] = asy

edge = []
sp = L.S_make(30, 12, 30, 12, 0.3, FLAGS)
for n, m in [(2, 2), (2, 1), (0, 0), (5, 0), (3, 5), (5, 1), (5, 6), (5, 5), (31, 3), (20, 13), (30, 12), (12, 12), (13, 12)]:
    edge.append({"n": n, "m": m, "S": repr(L.S_S(sp, n, m)), "V": repr(L.S_V(sp, n, m)) if m >= 2 else None,
                 # S_UV(n,0) / S_U(n,0) read out of bounds / exit in the reference: not sampled
                 "UV": repr(L.S_UV(sp, n, m)) if m >= 1 else None, "U": repr(L.S_U(sp, n, m)) if m >= 1 else None})
L.S_free(sp)
out["edge"] = {"N": 30, "M": 12, "a": 0.3, "cells": edge}

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
