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
from tests.golden.make_golden_programs import CASES as PROGRAM_CASES
from tests.golden.make_golden_programs import run as run_fixed_clock

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


@pytest.mark.gpu
def test_demo_and_check_programs_run_end_to_end():
    """test/demo.c (simulated PYP hierarchy: Gibbs on table counts with S_V / S_U, samplea and
    sampleb every few cycles) and test/check.c (table-indicator sampler, a/b sampling) run to
    completion on the GPU engine and report estimates inside the samplers' bounds.  Their random
    decisions depend on the last bits of the table values, so only the summaries are checked."""
    if not (os.path.exists(DROPIN["demo"]) and os.path.exists(DROPIN["check"])):
        pytest.skip("oracle/_ref/dropin_* not built")
    demo = subprocess.run([DROPIN["demo"], "-a", "0.3", "-b", "5", "-N", "200", "-C", "50", "-I", "10", "-H", "10",
                           "-s", "7"], capture_output=True, text=True, timeout=300)
    assert demo.returncode == 0, demo.stderr[-2000:]
    out = demo.stdout + demo.stderr
    a = float(re.search(r"^a=([0-9.]+)", out, re.M).group(1))
    b = float(re.search(r"^b=([0-9.]+)", out, re.M).group(1))
    assert 0.01 <= a <= 0.98 and 0.01 <= b <= 2000
    assert len(re.findall(r"^T\[\d\]=", out, re.M)) == 3
    chk = subprocess.run([DROPIN["check"], "-a", "0.3", "-b", "5", "-N", "200", "-C", "20", "-I", "5", "-H", "5",
                          "-s", "7", "-STI"], capture_output=True, text=True, timeout=300)
    assert chk.returncode == 0, chk.stderr[-2000:]
    out = chk.stdout + chk.stderr
    a = float(re.search(r"Run ave a: = ([0-9.]+)", out).group(1))
    b = float(re.search(r"Run ave b: = ([0-9.]+)", out).group(1))
    T = float(re.search(r"Run ave T: = ([0-9.]+)", out).group(1))
    assert 0.01 <= a <= 0.98 and 0.01 <= b <= 2000 and 5 <= T <= 200


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(PROGRAM_CASES))
def test_demo_and_check_output_matches_reference_under_a_fixed_clock(case):
    """The reference's simulation programs -- Gibbs sweeps over table counts through S_V / S_S / S_U,
    samplea / sampleb (ARS, the reference's default build) or the programs' own slice / ARS code every
    few cycles, S_remake after every new discount -- print what they print when linked with the
    reference library: same source, same seeds, same clock (oracle/_ref/shim_time.so; goldens from
    tests/golden/make_golden_programs.py).  Every accept/reject of thousands of draws has to agree
    for the summaries to agree; numbers are compared at 2e-6 relative (6 digits are printed)."""
    prog, args = PROGRAM_CASES[case]
    shim = os.path.join(ROOT, "oracle", "_ref", "shim_time.so")
    if not (os.path.exists(DROPIN[prog]) and os.path.exists(shim)):
        pytest.skip("oracle/_ref/dropin_* not built")
    want = open(os.path.join(ROOT, "tests", "golden", "programs", case + ".txt")).read().splitlines()
    got = run_fixed_clock(DROPIN[prog], args, {"STB_SAMPLER": "ars"})
    assert got.returncode == 0, got.stdout[-2000:]
    keep = lambda lines: [l for l in lines if l.strip() and not l.startswith(("S-table", "Time"))]
    w, g = keep(want), keep(got.stdout.splitlines())
    assert len(w) == len(g), (w, g)
    bad = []
    for lw, lg in zip(w, g):
        nw, ng = NUM.findall(lw), NUM.findall(lg)
        if NUM.sub("#", lw) != NUM.sub("#", lg) or len(nw) != len(ng) or not all(
                _numbers_match(a, b) for a, b in zip(nw, ng)):
            bad.append((lw, lg))
    assert not bad, bad[:6]
