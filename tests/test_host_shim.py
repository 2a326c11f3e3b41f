"""The C++17 drop-in (feellgood_b200/host/feellgood_b200.hpp): compiled and linked against the
C-ABI library on every box; run end to end (reference-style time loop + ports of the reference's
algebra unit tests) where there is a GPU."""
import os
import subprocess

import pytest

import cases
from feellgood_b200 import capi


def _build(tmp_path):
    exe = str(tmp_path / "host_shim_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror",
                           os.path.join(cases.ROOT, "tests", "cpp", "host_shim_test.cpp"), "-o", exe,
                           capi.LIB_PATH, "-Wl,-rpath," + os.path.dirname(capi.LIB_PATH)])
    return exe


def test_shim_compiles_and_links(gpu_lib, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "--link-check"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "fg_version" in r.stdout


def test_shim_timestepper_matches_reference_bits(gpu_lib, tmp_path):
    """The C++ TimeStepper / LogStats of the drop-in header against outputs of the reference's own
    classes (tests/golden/ref_timestepper.npz): bit-exact."""
    import numpy as np
    G = np.load(os.path.join(cases.GOLDEN, "ref_timestepper.npz"))
    exe = _build(tmp_path)
    for k in range(3):
        init, mn, mx = G["ts%d_prm" % k]
        lines = ["%.17g %.17g %.17g" % (init, mn, mx)]
        lines += ["%d %.17g" % (int(op), val) for op, val in G["ts%d_ops" % k]]
        if k == 0:
            lines += ["2 %.17g" % x for x in G["ls_x"]]
        r = subprocess.run([exe, "--timestepper"], input="\n".join(lines) + "\n", capture_output=True,
                           text=True, timeout=120)
        assert r.returncode == 0
        out = r.stdout.split("\n")
        ref = G["ts%d_out" % k]
        for i, v in enumerate(ref):
            if not np.isnan(v):
                assert float(out[i]) == v
        if k == 0:
            for i, refrow in enumerate(G["ls_out"]):
                n, m, sd = out[len(ref) + i].split()
                assert (float(n), float(m), float(sd)) == tuple(refrow)


@pytest.mark.gpu
def test_shim_time_loop_and_algebra(gpu_lib, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ok (0 failures)" in r.stdout
