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


@pytest.mark.gpu
def test_shim_time_loop_and_algebra(gpu_lib, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ok (0 failures)" in r.stdout
