"""Host-side run of the __host__ __device__ math helpers of the CUDA library (no GPU needed: nvcc
compiles the host path): the unit-quaternion encoding of the tangent-plane basis."""
import os
import subprocess

import cases


def test_quaternion_basis_round_trip(tmp_path):
    exe = str(tmp_path / "device_math_test")
    src = os.path.join(cases.ROOT, "tests", "cpp", "device_math_test.cu")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, src])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DEVICE_MATH_OK" in r.stdout, r.stdout + r.stderr
