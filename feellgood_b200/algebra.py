"""Mirror of the reference's `algebra::` namespace (src/algebra/*.h) on top of the C ABI:
SparseMatrix (MatrixShape, clear/set/add/mult/build_diag_precond), iteration, bicg, bicg_dir,
cg, cg_dir.  Same argument order as the reference: solver(iter, A, x, rhs[, xd], ld).

Matrix construction (set/add) is host-side bookkeeping exactly like the reference's; the values
are shipped to the device once per solve and every SpMV / BLAS-1 / reduction runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import capi
from .capi import check, dp, f64, i32, ip

UNDEFINED, CONVERGED, ITER_OVERFLOW, CANNOT_CONVERGE = -1, 0, 1, 2   # iter.h:24-30


class iteration:
    """algebra::iteration<T>, src/algebra/iter.h:37-179."""

    def __init__(self, name="", resmax=1e-8, verbose=False, maxiter=-1):
        self.solver_name, self.resmax, self.verbose, self.maxiter = name, resmax, verbose, maxiter
        self.reset()

    def reset(self):
        self.status, self.nit = UNDEFINED, 0
        self.res, self.rhsn = sys.float_info.max, 1.0

    def get_res(self):
        return self.res

    def get_iteration(self):
        return self.nit

    def get_rhsnorm(self):
        return self.rhsn

    def infos(self):
        names = {UNDEFINED: "undefined", CONVERGED: "converged", ITER_OVERFLOW: "iter overflow",
                 CANNOT_CONVERGE: "cannot converge"}
        return "%s %s after %d iterations, residu= %g" % (self.solver_name, names[self.status],
                                                          self.nit, self.res)

    def _take(self, r):
        self.status, self.nit, self.res, self.rhsn = r.status, r.iters, r.res, r.rhsnorm


class SparseMatrix:
    """algebra::SparseMatrix, src/algebra/sparseMat.h:45-191.  `shape` is a MatrixShape: one
    iterable of column indices per row (sparseMat.h:38)."""

    def __init__(self, shape, device=0):
        rows = [np.unique(np.asarray(sorted(r), dtype=np.int32)) for r in shape]
        self.N = len(rows)
        self.rowptr = np.zeros(self.N + 1, dtype=np.int32)
        self.rowptr[1:] = np.cumsum([r.size for r in rows])
        self.col = (np.concatenate(rows) if self.N and self.rowptr[-1] else
                    np.zeros(0, dtype=np.int32)).astype(np.int32)
        self.val = np.zeros(self.col.size)
        self._dirty = True
        h = C.c_void_p()
        check(capi.lib().fg_matrix_create(C.c_int(self.N), ip(self.rowptr), ip(self.col),
                                          C.c_int(device), C.byref(h)))
        self._h = h

    @classmethod
    def from_csr(cls, rowptr, col, val=None, device=0):
        m = cls.__new__(cls)
        m.rowptr, m.col = i32(rowptr).copy(), i32(col).copy()
        m.N = m.rowptr.size - 1
        m.val = f64(val).copy() if val is not None else np.zeros(m.col.size)
        m._dirty = True
        h = C.c_void_p()
        check(capi.lib().fg_matrix_create(C.c_int(m.N), ip(m.rowptr), ip(m.col), C.c_int(device),
                                          C.byref(h)))
        m._h = h
        return m

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().fg_matrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _slot(self, i, j):
        lo, hi = self.rowptr[i], self.rowptr[i + 1]
        k = lo + int(np.searchsorted(self.col[lo:hi], j))
        if k >= hi or self.col[k] != j:
            raise IndexError("(%d,%d) is outside the matrix shape" % (i, j))   # assert in the ref
        return k

    def clear(self):
        self.val[:] = 0.0
        self._dirty = True

    def set(self, i, j, v):
        self.val[self._slot(i, j)] = v
        self._dirty = True

    def add(self, i, j, v):
        self.val[self._slot(i, j)] += v
        self._dirty = True

    def __call__(self, i, j):
        lo, hi = self.rowptr[i], self.rowptr[i + 1]
        k = lo + int(np.searchsorted(self.col[lo:hi], j))
        return float(self.val[k]) if k < hi and self.col[k] == j else 0.0

    def _sync(self):
        if self._dirty:
            check(capi.lib().fg_matrix_set_values(self._h, dp(self.val)))
            self._dirty = False

    def mult(self, x):
        self._sync()
        x = f64(x)
        y = np.empty(self.N)
        check(capi.lib().fg_matrix_mult(self._h, dp(x), dp(y)))
        return y

    def build_diag_precond(self):
        """sparseMat.h:174-183 (host mirror, used by tests only; the solvers build D on the GPU)."""
        with np.errstate(divide="ignore"):
            return np.array([1.0 / self(i, i) for i in range(self.N)])


def mult(A, x):
    return A.mult(x)


def _run(fn, it, A, x, rhs, xd, ld):
    A._sync()
    x = f64(x)
    rhs = f64(rhs)
    r = capi.IterResult()
    args = [A._h, dp(x), dp(rhs)]
    keep = []
    if fn in ("fg_bicg_dir", "fg_cg_dir"):
        xd_a = f64(xd) if xd is not None else None
        ld_a = i32(ld)
        keep += [xd_a, ld_a]
        args += [dp(xd_a), ip(ld_a), C.c_int(ld_a.size)]
    args += [C.c_double(it.resmax), C.c_int(it.maxiter), C.byref(r)]
    it.reset()
    check(getattr(capi.lib(), fn)(*args))
    it._take(r)
    return x


def bicg(it, A, x, rhs):
    """algebra::bicg, src/algebra/bicg.h:14-72.  x is updated in place when it is a float64
    contiguous ndarray; the solution is also returned."""
    out = _run("fg_bicg", it, A, x, rhs, None, None)
    return _writeback(x, out)


def bicg_dir(it, A, x, rhs, *args):
    """algebra::bicg_dir: (iter, A, x, rhs, xd, ld) src/algebra/bicg.h:83-154, or
    (iter, A, x, rhs, ld) with zero Dirichlet values :163-234.  Returns (x, res/rhsn-like)."""
    xd, ld = (args if len(args) == 2 else (None, args[0]))
    out = _run("fg_bicg_dir", it, A, x, rhs, xd, ld)
    return _writeback(x, out)


def cg(it, A, x, rhs):
    """algebra::cg, src/algebra/cg.h:15-58."""
    out = _run("fg_cg", it, A, x, rhs, None, None)
    return _writeback(x, out)


def cg_dir(it, A, x, rhs, xd, ld):
    """algebra::cg_dir, src/algebra/cg.h:68-121."""
    out = _run("fg_cg_dir", it, A, x, rhs, xd, ld)
    return _writeback(x, out)


def _writeback(x, out):
    if isinstance(x, np.ndarray) and x.dtype == np.float64 and x is not out:
        x[...] = out
    return out
