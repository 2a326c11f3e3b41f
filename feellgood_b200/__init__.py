"""feellgood_b200 — B200-native (sm_100a, FP64) per-time-step LLG hot path of FeeLLGood.

csrc/               CUDA kernels + the C ABI (include/feellgood_b200.h) -> libfeellgood_b200.so
host/               C++17 drop-in keeping the reference's LinAlgebra / timing / algebra:: surface
linear_algebra.py   Python mirror of LinAlgebra / timing (what tests/ and bench.py drive)
algebra.py          Python mirror of algebra::SparseMatrix, bicg, bicg_dir, cg, cg_dir
meshgen.py          mesh inputs: gmsh ASCII reader, Cuboid/disk/tube generators, dMs, node sort
"""
from . import capi, meshgen  # noqa: F401
from .linear_algebra import LinAlgebra, Settings, timing  # noqa: F401
