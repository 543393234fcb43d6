"""Generates tests/golden/programs/*.txt: the output of the reference's own simulation programs
test/demo.c and test/check.c (UNMODIFIED, linked with the reference library in its default
configuration: oracle/_ref/ref_demo, ref_check, built by oracle/build_ref.sh) under a fixed clock
(oracle/_ref/shim_time.so: the programs reseed from time(NULL) before their Gibbs loops).
tests/test_dropin_gpu.py runs the SAME sources linked against this repo's library, under the same
clock, and compares.  Run in the build container:  python tests/golden/make_golden_programs.py"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
CASES = {
    # program, arguments
    "demo_ab": ("demo", ["-a", "0.3", "-b", "5", "-N", "200", "-C", "50", "-I", "10", "-H", "10", "-s", "7"]),
    "demo_fixed": ("demo", ["-a", "0.6", "-b", "2", "-N", "300", "-C", "40", "-s", "11"]),
    "demo_a0": ("demo", ["-a", "0", "-b", "3", "-N", "150", "-C", "30", "-H", "5", "-s", "3"]),
    "check_TI": ("check", ["-a", "0.3", "-b", "5", "-N", "200", "-C", "20", "-I", "5", "-H", "5", "-s", "7", "-STI"]),
    "check_CT": ("check", ["-a", "0.5", "-b", "2", "-N", "150", "-C", "20", "-I", "4", "-H", "4", "-s", "5", "-SCT"]),
    "check_CTW": ("check", ["-a", "0.2", "-b", "8", "-N", "250", "-C", "15", "-s", "9", "-SCTW"]),
    "check_TI_ars": ("check", ["-A", "-a", "0.4", "-b", "4", "-N", "120", "-C", "20", "-I", "5", "-H", "5", "-s", "2",
                               "-STI"]),
    "demo_long": ("demo", ["-a", "0.4", "-b", "10", "-N", "3000", "-C", "200", "-I", "5", "-H", "5", "-s", "21"]),
    "check_TI_long": ("check", ["-a", "0.25", "-b", "20", "-N", "2000", "-C", "100", "-I", "5", "-H", "5", "-s", "13",
                                "-STI"]),
    "check_CT_long": ("check", ["-a", "0.7", "-b", "1", "-N", "1500", "-C", "60", "-I", "6", "-H", "3", "-s", "17",
                                "-SCT"]),
    "check_SA": ("check", ["-a", "0.3", "-b", "5", "-N", "100", "-C", "10", "-I", "5", "-s", "4", "-SSA"]),
}


def run(exe, args, extra_env=None):
    """merged stdout + stderr of one run under the fixed clock"""
    env = dict(os.environ, LD_PRELOAD=os.path.join(REF, "shim_time.so"))
    env.update(extra_env or {})
    return subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                          timeout=600)


if __name__ == "__main__":
    for name, (prog, args) in CASES.items():
        r = run(os.path.join(REF, "ref_" + prog), args)
        assert r.returncode == 0, (name, r.stdout[-2000:])
        with open(os.path.join(HERE, "programs", name + ".txt"), "w") as f:
            f.write(r.stdout)
        print(name, len(r.stdout.splitlines()), "lines")
