"""Builds and runs tests/cpp/ut_oracle.cpp: the reference's own unit-test formulas ("ref code vs
code to test", /root/reference/unit-tests) applied to the CPU oracle with the
reference's fixtures, seed and tolerance.  This is what pins the oracle's per-formula pieces."""
import os
import subprocess

import cases

ROOT = cases.ROOT


def test_reference_unit_test_formulas(oracle, tmp_path):
    exe = str(tmp_path / "ut_oracle")
    so = os.path.join(ROOT, "oracle", "libfg_oracle.so")
    assert os.path.exists(so)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "cpp", "ut_oracle.cpp"), "-o", exe, so,
                           "-Wl,-rpath," + os.path.dirname(so), "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout
