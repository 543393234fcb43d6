"""Drop-in proof with the reference's OWN programs: test/list.c, test/demo.c and test/check.c of
wbuntine/libstb compile and link UNMODIFIED against this repo's include/ and libstb_b200.so
(oracle/build_ref.sh builds them into oracle/_ref/dropin_*; binaries travel to the GPU box, the
sources do not).  `list` is deterministic: its stdout (S_S, S_V, S_U, S_UV, S_asympt over a table
that grows on demand, plus the asymptote-difference mode) is compared with the stdout of the same
source linked with the reference library (tests/golden/list/*.txt, made by
tests/golden/make_golden_list.py).  Values are printed with 6 significant digits; the bar is
2e-6 relative on every number, identical text otherwise."""
import math
import os
import re
import subprocess

import pytest

from tests import harness
from tests.golden.make_golden_list import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = {p: os.path.join(ROOT, "oracle", "_ref", "dropin_" + p) for p in ("list", "demo", "check")}
NUM = re.compile(r"[-+]?(?:inf|nan|\d+\.?\d*(?:[eE][-+]?\d+)?)")


def test_reference_programs_link_unchanged():
    """(no GPU needed) the three programs were compiled from the reference's sources against our
    headers and linked with our library; every symbol they import from it is exported"""
    if not all(os.path.exists(p) for p in DROPIN.values()):
        pytest.skip("oracle/_ref/dropin_* not built (no /root/reference here)")
    lib = os.path.join(ROOT, "libstb_b200", "lib", "libstb_b200.so")
    exported = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True,
                                                      text=True, check=True).stdout.splitlines() if l.strip()}
    api = re.compile(r"^(S_|yaps_|sample[ab]$|SliceSimple$|arms|gammadiff$|psidiff$|[gpq]cache_|digamma|ML|gsl_rng_)")
    for prog, path in DROPIN.items():
        und = [l.split()[-1] for l in subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True,
                                                     text=True, check=True).stdout.splitlines() if l.strip()]
        ours = [s for s in und if api.match(s.split("@")[0])]
        assert ours, prog
        missing = [s for s in ours if s.split("@")[0] not in exported]
        assert not missing, (prog, missing)


def _numbers_match(a, b):
    if a == b:
        return True
    x, y = float(a), float(b)
    if math.isnan(x) or math.isnan(y) or math.isinf(x) or math.isinf(y):
        return (math.isnan(x) and math.isnan(y)) or x == y
    return abs(x - y) <= 2e-6 * max(abs(x), abs(y)) + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_list_program_output_matches_reference(case):
    if not os.path.exists(DROPIN["list"]):
        pytest.skip("oracle/_ref/dropin_list not built")
    want = open(os.path.join(ROOT, "tests", "golden", "list", case + ".txt")).read().splitlines()
    got = subprocess.run([DROPIN["list"]] + CASES[case], capture_output=True, text=True, timeout=300)
    assert got.returncode == 0, got.stderr[-2000:]
    keep = lambda lines: [l for l in lines if l.strip() and not l.startswith("S-table")]  # the report line differs by design
    w, g = keep(want), keep(got.stdout.splitlines())
    assert len(w) == len(g), (len(w), len(g))
    bad = []
    for lw, lg in zip(w, g):
        if NUM.sub("#", lw) != NUM.sub("#", lg):
            bad.append((lw, lg))
            continue
        nw, ng = NUM.findall(lw), NUM.findall(lg)
        if len(nw) != len(ng) or not all(_numbers_match(a, b) for a, b in zip(nw, ng)):
            bad.append((lw, lg))
    assert not bad, bad[:5]
