"""ctypes binding of libfeellgood_b200.so (include/feellgood_b200.h).

Thin by design: argument marshalling only.  There is no CPU fallback anywhere in this package —
if the shared library is missing, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfeellgood_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

FG_OK = 0
FG_UNDEFINED, FG_CONVERGED, FG_ITER_OVERFLOW, FG_CANNOT_CONVERGE = -1, 0, 1, 2
FG_IDX_UNDEF, FG_IDX_X, FG_IDX_Y, FG_IDX_Z = -1, 0, 1, 2

# every symbol include/feellgood_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "fg_last_error", "fg_version", "fg_create", "fg_destroy", "fg_get_sizes", "fg_set_state",
    "fg_set_next_v", "fg_set_potentials", "fg_get_state", "fg_commit", "fg_set_ext_space_field",
    "fg_base_projection", "fg_prepare_elements", "fg_prepare_elements_space", "fg_solve",
    "fg_step", "fg_get_basis", "fg_get_elements", "fg_get_tri_elements", "fg_get_csr_pattern",
    "fg_get_system", "fg_apply_operator", "fg_get_solution", "fg_get_tet_tables",
    "fg_matrix_create", "fg_matrix_destroy", "fg_matrix_set_values", "fg_matrix_mult", "fg_bicg",
    "fg_bicg_dir", "fg_cg", "fg_cg_dir", "fg_kernel_launches", "fg_stream", "fg_set_profiling",
    "fg_get_phase_times", "fg_get_krylov_state", "fg_get_krylov_history", "fg_set_operator", "fg_get_spmv_times", "fg_get_kernel_times", "fg_energy", "fg_energy_space", "fg_avg", "fg_max_angle", "fg_calc_charges", "fg_demag_direct", "fg_bench_spmv", "fg_host_plan", "fg_dist_create", "fg_dist_export", "fg_dist_connect", "fg_get_layout",
    "fg_set_solver", "fg_get_solve_times", "fg_get_precond", "fg_get_records", "fg_solver_launch_shape",
]


class TetPrm(C.Structure):
    """fg_tet_prm — Tetra::prm fields read by the hot path (reference src/tetra.h:86-112)."""
    _fields_ = [("alpha_LLG", C.c_double), ("A", C.c_double), ("Ms", C.c_double),
                ("K", C.c_double), ("uk", C.c_double * 3), ("K3", C.c_double),
                ("ex", C.c_double * 3), ("ey", C.c_double * 3), ("ez", C.c_double * 3)]


class TriPrm(C.Structure):
    """fg_tri_prm — Triangle::prm fields (reference src/triangle.h:69-82)."""
    _fields_ = [("Ks", C.c_double), ("uk", C.c_double * 3), ("suppress_charges", C.c_int),
                ("pad_", C.c_int)]


class CMesh(C.Structure):
    _fields_ = [("NOD", C.c_int), ("node_p", c_double_p), ("NT", C.c_int), ("tet_ind", c_int_p),
                ("tet_reg", c_int_p), ("NF", C.c_int), ("tri_ind", c_int_p), ("tri_reg", c_int_p),
                ("tri_dMs", c_double_p)]


class CParams(C.Structure):
    _fields_ = [("nreg_tet", C.c_int), ("prm_tet", C.POINTER(TetPrm)), ("nreg_tri", C.c_int),
                ("prm_tri", C.POINTER(TriPrm)), ("npi_tet", C.c_int), ("npi_tri", C.c_int),
                ("tol", C.c_double), ("maxiter", C.c_int)]


class StepResult(C.Structure):
    _fields_ = [("failed", C.c_int), ("status", C.c_int), ("iters", C.c_int), ("pad_", C.c_int),
                ("res", C.c_double), ("rhsnorm", C.c_double), ("v_max", C.c_double)]


class IterResult(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("res", C.c_double),
                ("rhsnorm", C.c_double)]


def tet_prm(alpha=0.5, A=1e-11, Ms=795774.7, K=0.0, uk=(0, 0, 1), K3=0.0, ex=(1, 0, 0),
            ey=(0, 1, 0), ez=(0, 0, 1)):
    """Defaults = default-settings.yml:104-125 of the reference."""
    p = TetPrm()
    p.alpha_LLG, p.A, p.Ms, p.K, p.K3 = alpha, A, Ms, K, K3
    p.uk[:] = uk
    p.ex[:] = ex
    p.ey[:] = ey
    p.ez[:] = ez
    return p


def tri_prm(Ks=0.0, uk=(0, 0, 1), suppress_charges=False):
    p = TriPrm()
    p.Ks = Ks
    p.uk[:] = uk
    p.suppress_charges = int(suppress_charges)
    return p


class FgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("feellgood_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load the shared library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                "%s not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C feellgood_b200/csrc`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.fg_last_error.restype = C.c_char_p
        L.fg_kernel_launches.restype = C.c_longlong
        L.fg_kernel_launches.argtypes = [C.c_void_p]
        L.fg_stream.restype = C.c_void_p
        L.fg_stream.argtypes = [C.c_void_p]
        L.fg_destroy.restype = None
        L.fg_destroy.argtypes = [C.c_void_p]
        L.fg_matrix_destroy.restype = None
        L.fg_matrix_destroy.argtypes = [C.c_void_p]
        # the per-step calls: with argtypes ctypes converts Python numbers itself, which costs a fraction of
        # building c_double objects per call (the host path of a step is GPU idle time on small meshes)
        L.fg_step.argtypes = [C.c_void_p, C.c_double, c_double_p, C.c_double, C.c_double, C.c_int, C.c_double,
                              C.POINTER(StepResult)]
        L.fg_commit.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != FG_OK:
        raise FgError(rc, lib().fg_last_error().decode())


def solver_launch_shape(nslice, n_sm=0):
    """Launch shape of the persistent solve kernel for `nslice` SELL slices on `n_sm` SMs (0: 148), computed
    on the host: dict(block, warps, grid, head)  (fg_solver_launch_shape)."""
    out = (C.c_int * 4)()
    check(lib().fg_solver_launch_shape(int(nslice), int(n_sm), out))
    return dict(block=out[0], warps=out[1], grid=out[2], head=bool(out[3]))


def dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
