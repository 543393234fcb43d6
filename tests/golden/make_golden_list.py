"""Generates tests/golden/list/*.txt: stdout of the reference's own test program test/list.c
(UNMODIFIED, linked with the reference library: oracle/_ref/ref_list, built by oracle/build_ref.sh)
for several argument sets.  tests/test_dropin_gpu.py runs the SAME source linked against this
repo's library and compares.  Run in the build container:  python tests/golden/make_golden_list.py"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES = {
    "default": [],
    "a0.1_N100_T20": ["-a", "0.1", "-N", "100", "-T", "20"],
    "a0.9_N60_T60_n12": ["-a", "0.9", "-N", "60", "-T", "60", "-n", "12"],
    "a0_N80_T25": ["-a", "0", "-N", "80", "-T", "25"],
    "asympt_a0.3_N50_T20": ["-A", "-a", "0.3", "-N", "50", "-T", "20"],
}

if __name__ == "__main__":
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_list")
    for name, args in CASES.items():
        out = subprocess.run([exe] + args, capture_output=True, text=True, check=True).stdout
        with open(os.path.join(HERE, "list", name + ".txt"), "w") as f:
            f.write(out)
        print(name, len(out.splitlines()), "lines")
