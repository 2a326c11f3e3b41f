"""ctypes binding of the CPU oracle (oracle/libfg_oracle.so) and of the compiled reference algebra
(oracle/_ref/libfgref_algebra.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libfg_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libfgref_algebra.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class TetPrm(C.Structure):
    _fields_ = [("alpha_LLG", C.c_double), ("A", C.c_double), ("Ms", C.c_double),
                ("K", C.c_double), ("uk", C.c_double * 3), ("K3", C.c_double),
                ("ex", C.c_double * 3), ("ey", C.c_double * 3), ("ez", C.c_double * 3)]


class TriPrm(C.Structure):
    _fields_ = [("Ks", C.c_double), ("uk", C.c_double * 3), ("suppress_charges", C.c_int),
                ("pad_", C.c_int)]


class Iter(C.Structure):
    _fields_ = [("resmax", C.c_double), ("maxiter", C.c_int), ("status", C.c_int),
                ("nit", C.c_int), ("res", C.c_double), ("rhsn", C.c_double)]


def tet_prm(alpha=0.5, A=1e-11, Ms=795774.7, K=0.0, uk=(0, 0, 1), K3=0.0, ex=(1, 0, 0),
            ey=(0, 1, 0), ez=(0, 0, 1)):
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


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(ORACLE_SO) or (
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "fg_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "libfg_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/algebra") and (force or not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        L.fgo_timing_dt0.restype = C.c_double
        L.fgo_timing_dt0.argtypes = [C.c_double, C.c_double]
        L.fgo_timing_prefactor.restype = C.c_double
        L.fgo_timing_prefactor.argtypes = [C.c_double, C.c_double]
        for name in ("fgo_tet_a", "fgo_tet_pds", "fgo_tri_a", "fgo_tri_pds"):
            getattr(L, name).restype = c_double_p
            getattr(L, name).argtypes = [C.c_int]
        L.fgo_tet_setup.restype = C.c_double
        L.fgo_create.restype = C.c_void_p
        L.fgo_get_v_max.restype = C.c_double
        L.fgo_get_v_max.argtypes = [C.c_void_p]
        L.fgo_total_mag_vol.restype = C.c_double
        L.fgo_total_mag_vol.argtypes = [C.c_void_p]
        L.fgo_max_angle.restype = C.c_double
        L.fgo_max_angle.argtypes = [C.c_void_p]
        L.fgo_bicg_dir.restype = C.c_double
        L.fgo_n_sources.restype = C.c_longlong
        L.fgo_n_sources.argtypes = [C.c_void_p]
        L.fgo_tri_potential.restype = C.c_double
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
# single-element functions
# ---------------------------------------------------------------------------------------------
def timing_prefactor(dt, dtmax):
    return lib().fgo_timing_prefactor(dt, dtmax)


def node_set_basis(u, r):
    u = _f64(u)
    ep, eq = np.zeros(3), np.zeros(3)
    lib().fgo_node_set_basis(_dp(u), C.c_double(r), _dp(ep), _dp(eq))
    return ep, eq


def node_make_evol(u0, ep, eq, vp, vq, dt):
    u1, v1 = np.zeros(3), np.zeros(3)
    lib().fgo_node_make_evol(_dp(_f64(u0)), _dp(_f64(ep)), _dp(_f64(eq)), C.c_double(vp),
                             C.c_double(vq), C.c_double(dt), _dp(u1), _dp(v1))
    return u1, v1


def tet_setup(node_p, ind, npi=5):
    node_p = _f64(node_p)
    ind = _i32(ind).copy()
    sw = lib().fgo_tet_orientate(_dp(node_p), _ip(ind))
    da, w = np.zeros(12), np.zeros(npi)
    detJ = lib().fgo_tet_setup(_dp(node_p), _ip(ind), C.c_int(npi), _dp(da), _dp(w))
    return ind, da.reshape(4, 3), w, detJ, sw


def calc_alpha_eff(dt, alpha, uHeff):
    uHeff = _f64(uHeff)
    out = np.zeros_like(uHeff)
    lib().fgo_calc_alpha_eff(C.c_int(uHeff.size), C.c_double(dt), C.c_double(alpha), _dp(uHeff),
                             _dp(out))
    return out


def tet_integrales(prm, dt, prefactor, da, weight, u, v, phi, phiv, ep, eq, Hext, idx_dir=-1,
                   Vdrift=0.0):
    """All node arrays (4,3)/(4,); Hext (3,npi). Returns Kp (8,8), Lp (8,)."""
    npi = len(weight)
    Kp, Lp = np.zeros(64), np.zeros(8)
    Hext = _f64(Hext)
    if Hext.shape == (3,):
        Hext = np.repeat(Hext[:, None], npi, axis=1)
    lib().fgo_tet_integrales(C.c_int(npi), C.byref(prm), C.c_double(dt), C.c_double(prefactor),
                             _dp(_f64(da)), _dp(_f64(weight)), _dp(_f64(u)), _dp(_f64(v)),
                             _dp(_f64(phi)), _dp(_f64(phiv)), _dp(_f64(ep)), _dp(_f64(eq)),
                             _dp(_f64(Hext)), C.c_int(idx_dir), C.c_double(Vdrift), _dp(Kp), _dp(Lp))
    return Kp.reshape(8, 8), Lp


def tri_integrales(prm, dMs, weight, u, ep, eq):
    Lp = np.zeros(6)
    lib().fgo_tri_integrales(C.c_int(len(weight)), C.byref(prm), C.c_double(dMs), _dp(_f64(weight)),
                             _dp(_f64(u)), _dp(_f64(ep)), _dp(_f64(eq)), _dp(Lp))
    return Lp


# ---------------------------------------------------------------------------------------------
# sparse algebra on CSR arrays
# ---------------------------------------------------------------------------------------------
def spmv(rowptr, col, val, x):
    rowptr, col, val, x = _i32(rowptr), _i32(col), _f64(val), _f64(x)
    y = np.zeros_like(x)
    lib().fgo_spmv(C.c_int(x.size), _ip(rowptr), _ip(col), _dp(val), _dp(x), _dp(y))
    return y


def _solve(fn, rowptr, col, val, x0, rhs, tol, maxiter, xd=None, ld=None):
    rowptr, col, val = _i32(rowptr), _i32(col), _f64(val)
    x = _f64(x0).copy()
    rhs = _f64(rhs)
    it = Iter(resmax=tol, maxiter=maxiter)
    n = C.c_int(x.size)
    args = [C.byref(it), n, _ip(rowptr), _ip(col), _dp(val), _dp(x), _dp(rhs)]
    if xd is not None:
        xd = _f64(xd)
        args.append(_dp(xd))
    if ld is not None:
        ld = _i32(ld)
        args += [_ip(ld), C.c_int(ld.size)]
    getattr(lib(), fn)(*args)
    return x, dict(status=it.status, nit=it.nit, res=it.res, rhsn=it.rhsn)


def bicg(rowptr, col, val, x0, rhs, tol=1e-6, maxiter=700):
    return _solve("fgo_bicg", rowptr, col, val, x0, rhs, tol, maxiter)


def bicg_dir(rowptr, col, val, x0, rhs, ld, tol=1e-6, maxiter=700, xd=None):
    fn = "fgo_bicg_dir" if xd is None else "fgo_bicg_dir_xd"
    return _solve(fn, rowptr, col, val, x0, rhs, tol, maxiter, xd=xd, ld=np.asarray(ld))


def cg(rowptr, col, val, x0, rhs, tol=1e-6, maxiter=700):
    return _solve("fgo_cg", rowptr, col, val, x0, rhs, tol, maxiter)


def cg_dir(rowptr, col, val, x0, rhs, xd, ld, tol=1e-6, maxiter=700):
    return _solve("fgo_cg_dir", rowptr, col, val, x0, rhs, tol, maxiter, xd=xd, ld=np.asarray(ld))


# ---------------------------------------------------------------------------------------------
# mesh-level context (the LinAlgebra call surface)
# ---------------------------------------------------------------------------------------------
class OracleCtx:
    def __init__(self, mesh, prm_tet, prm_tri, npi=5, npi_tri=4, tol=1e-6, maxiter=700):
        L = lib()
        self.L = L
        self.npi, self.npi_tri = npi, npi_tri
        self.NOD, self.NT, self.NF = mesh.NOD, mesh.NT, mesh.NF
        pt = (TetPrm * len(prm_tet))(*prm_tet)
        pf = (TriPrm * max(1, len(prm_tri)))(*prm_tri)
        p, ti, tr = _f64(mesh.node_p), _i32(mesh.tet_ind), _i32(mesh.tet_reg)
        fi, fr, fd = _i32(mesh.tri_ind), _i32(mesh.tri_reg), _f64(mesh.tri_dMs)
        self.h = C.c_void_p(L.fgo_create(
            C.c_int(mesh.NOD), _dp(p), C.c_int(mesh.NT), _ip(ti), _ip(tr), C.c_int(mesh.NF),
            _ip(fi), _ip(fr), _dp(fd), C.c_int(len(prm_tet)), pt, C.c_int(len(prm_tri)), pf,
            C.c_int(npi), C.c_int(npi_tri), C.c_double(tol), C.c_int(maxiter)))
        out = (C.c_longlong * 10)()
        L.fgo_sizes(self.h, out)
        (_, _, _, self.n_magTet, self.n_magTri, self.E, self.E_mag, self.n, self.nnz,
         self.nlvd) = list(out)

    def close(self):
        if self.h:
            self.L.fgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_num_threads(self, n):
        self.L.fgo_set_num_threads(self.h, C.c_int(n))

    def use_reference_algebra(self, on=True):
        if on and not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        r = self.L.fgo_use_reference_algebra(self.h, REF_SO.encode() if on else None)
        if r != 0:
            raise RuntimeError("fgo_use_reference_algebra failed: %d" % r)

    def set_state(self, u, v=None, phi=None, phiv=None):
        u = _f64(u)
        v = _f64(v) if v is not None else None
        phi = _f64(phi) if phi is not None else None
        phiv = _f64(phiv) if phiv is not None else None
        self.L.fgo_set_state(self.h, _dp(u), _dp(v) if v is not None else None,
                             _dp(phi) if phi is not None else None,
                             _dp(phiv) if phiv is not None else None)

    def set_next_v(self, v):
        v = _f64(v)
        self.L.fgo_set_next_v(self.h, _dp(v))

    def set_potentials_next(self, phi, phiv):
        phi, phiv = _f64(phi), _f64(phiv)
        self.L.fgo_set_potentials_next(self.h, _dp(phi), _dp(phiv))

    def get_state(self, step=1):
        u, v = np.zeros((self.NOD, 3)), np.zeros((self.NOD, 3))
        phi, phiv = np.zeros(self.NOD), np.zeros(self.NOD)
        self.L.fgo_get_state(self.h, C.c_int(step), _dp(u), _dp(v), _dp(phi), _dp(phiv))
        return u, v, phi, phiv

    def get_basis(self):
        ep, eq = np.zeros((self.NOD, 3)), np.zeros((self.NOD, 3))
        self.L.fgo_get_basis(self.h, _dp(ep), _dp(eq))
        return ep, eq

    def evolution(self):
        self.L.fgo_evolution(self.h)

    def set_ext_space_field(self, field):
        field = _f64(field)
        assert field.shape == (self.NT, 3, self.npi)
        self.L.fgo_set_ext_space_field(self.h, _dp(field))

    def base_projection(self, r):
        self.L.fgo_base_projection(self.h, C.c_double(r))

    def prepare_elements(self, Hext, dt, prefactor, idx_dir=-1, Vdrift=0.0):
        H = _f64(Hext)
        self.L.fgo_prepare_elements(self.h, _dp(H), C.c_double(dt), C.c_double(prefactor),
                                    C.c_int(idx_dir), C.c_double(Vdrift))

    def prepare_elements_space(self, A_Hext, dt, prefactor, idx_dir=-1, Vdrift=0.0):
        self.L.fgo_prepare_elements_space(self.h, C.c_double(A_Hext), C.c_double(dt),
                                          C.c_double(prefactor), C.c_int(idx_dir),
                                          C.c_double(Vdrift))

    def assemble(self):
        self.L.fgo_assemble(self.h)

    def solve(self, dt):
        return bool(self.L.fgo_solve(self.h, C.c_double(dt)))

    def v_max(self):
        return self.L.fgo_get_v_max(self.h)

    def iter_info(self):
        it = Iter()
        self.L.fgo_get_iter(self.h, C.byref(it))
        return dict(status=it.status, nit=it.nit, res=it.res, rhsn=it.rhsn)

    def tet_ind(self):
        a = np.zeros((self.NT, 4), dtype=np.int32)
        self.L.fgo_get_tet_ind(self.h, _ip(a))
        return a

    def tet_geom(self):
        da, w = np.zeros((self.NT, 4, 3)), np.zeros((self.NT, self.npi))
        self.L.fgo_get_tet_geom(self.h, _dp(da), _dp(w))
        return da, w

    def element(self, t):
        Kp, Lp = np.zeros(64), np.zeros(8)
        self.L.fgo_get_element(self.h, C.c_int(t), _dp(Kp), _dp(Lp))
        return Kp.reshape(8, 8), Lp

    def tri_element(self, f):
        Lp = np.zeros(6)
        self.L.fgo_get_tri_element(self.h, C.c_int(f), _dp(Lp))
        return Lp

    def csr(self):
        rowptr, col = np.zeros(self.n + 1, dtype=np.int32), np.zeros(self.nnz, dtype=np.int32)
        self.L.fgo_get_csr(self.h, _ip(rowptr), _ip(col))
        return rowptr, col

    def system(self):
        val, rhs, x = np.zeros(self.nnz), np.zeros(self.n), np.zeros(self.n)
        self.L.fgo_get_system(self.h, _dp(val), _dp(rhs), _dp(x))
        return val, rhs, x

    def masks(self):
        mag = np.zeros(self.NOD, dtype=np.uint8)
        lvd = np.zeros(max(1, self.nlvd), dtype=np.int32)
        self.L.fgo_get_masks(self.h, mag.ctypes.data_as(C.POINTER(C.c_ubyte)), _ip(lvd))
        return mag.astype(bool), lvd[:self.nlvd]

    def edges(self):
        e = np.zeros((max(1, self.E), 2), dtype=np.int32)
        self.L.fgo_get_edges(self.h, _ip(e))
        return e[:self.E]

    def energy(self, Hext):
        E = np.zeros(4)
        H = _f64(Hext)
        self.L.fgo_energy(self.h, _dp(H), _dp(E))
        return E

    def energy_space(self, fieldAmp):
        E = np.zeros(4)
        self.L.fgo_energy_space(self.h, C.c_double(fieldAmp), _dp(E))
        return E

    def avg(self, what=0, region=-1):
        out = np.zeros(3)
        self.L.fgo_avg_region(self.h, C.c_int(what), C.c_int(region), _dp(out))
        return out

    def n_sources(self):
        return int(self.L.fgo_n_sources(self.h))

    def source_positions(self):
        pos = np.zeros((max(1, self.n_sources()), 3))
        self.L.fgo_source_positions(self.h, _dp(pos))
        return pos[:self.n_sources()]

    def calc_charges(self, which=0):
        src, corr = np.zeros(max(1, self.n_sources())), np.zeros(self.NOD)
        self.L.fgo_calc_charges(self.h, C.c_int(which), _dp(src), _dp(corr))
        return src[:self.n_sources()], corr

    def demag_direct(self, second_order=True):
        self.L.fgo_calc_demag_direct(self.h, C.c_int(int(second_order)))

    def total_mag_vol(self):
        return self.L.fgo_total_mag_vol(self.h)

    def max_angle(self):
        return self.L.fgo_max_angle(self.h)


# ---------------------------------------------------------------------------------------------
# compiled reference algebra (oracle/_ref)
# ---------------------------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_SO)
        R.fgref_matrix_create.restype = C.c_void_p
        R.fgref_bicg_dir.restype = C.c_double
        R.fgref_dot.restype = C.c_double
        R.fgref_norm.restype = C.c_double
        if hasattr(R, "fgref_ts_new"):
            R.fgref_ts_new.restype = C.c_void_p
            R.fgref_ts_new.argtypes = [C.c_double, C.c_double, C.c_double]
            R.fgref_ts_free.argtypes = [C.c_void_p]
            R.fgref_ts_set_soft_limit.argtypes = [C.c_void_p, C.c_double]
            R.fgref_ts_step.restype = C.c_double
            R.fgref_ts_step.argtypes = [C.c_void_p, C.c_double]
            R.fgref_ls_new.restype = C.c_void_p
            R.fgref_ls_free.argtypes = [C.c_void_p]
            R.fgref_ls_add.argtypes = [C.c_void_p, C.c_double]
            R.fgref_ls_get.argtypes = [C.c_void_p, c_double_p]
        _ref = R
    return _ref


class RefTimeStepper:
    """The reference's own TimeStepper (src/time_integration.cpp:11-38) from oracle/_ref."""

    def __init__(self, initial, dtmin, dtmax):
        self.R = ref_lib()
        self.h = C.c_void_p(self.R.fgref_ts_new(initial, dtmin, dtmax))

    def set_soft_limit(self, mx):
        self.R.fgref_ts_set_soft_limit(self.h, mx)

    def __call__(self, stride):
        return self.R.fgref_ts_step(self.h, stride)

    def __del__(self):
        if getattr(self, "h", None):
            self.R.fgref_ts_free(self.h)
            self.h = None


class RefLogStats:
    """The reference's own LogStats (src/log-stats.h) from oracle/_ref."""

    def __init__(self):
        self.R = ref_lib()
        self.h = C.c_void_p(self.R.fgref_ls_new())

    def add(self, x):
        self.R.fgref_ls_add(self.h, x)

    def get(self):
        out = np.zeros(3)
        self.R.fgref_ls_get(self.h, _dp(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.R.fgref_ls_free(self.h)
            self.h = None


class RefMatrix:
    """The reference's algebra::SparseMatrix built from a CSR pattern."""

    def __init__(self, rowptr, col, val=None):
        self.R = ref_lib()
        self.rowptr, self.col = _i32(rowptr), _i32(col)
        self.n = self.rowptr.size - 1
        self.h = C.c_void_p(self.R.fgref_matrix_create(C.c_int(self.n), _ip(self.rowptr),
                                                       _ip(self.col)))
        if val is not None:
            val = _f64(val)
            self.R.fgref_matrix_set_values(self.h, _dp(val))

    def __del__(self):
        try:
            if self.h:
                self.R.fgref_matrix_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add(self, i, j, v):
        self.R.fgref_matrix_add(self.h, C.c_int(i), C.c_int(j), C.c_double(v))

    def clear(self):
        self.R.fgref_matrix_clear(self.h)

    def values(self):
        v = np.zeros(self.col.size)
        self.R.fgref_matrix_get_values(self.h, _dp(v))
        return v

    def mult(self, x):
        x = _f64(x)
        y = np.zeros_like(x)
        self.R.fgref_matrix_mult(self.h, _dp(x), _dp(y))
        return y

    def _run(self, fn, x0, rhs, tol, maxiter, xd=None, ld=None):
        x, rhs = _f64(x0).copy(), _f64(rhs)
        st, nit = C.c_int(), C.c_int()
        res, rhsn = C.c_double(), C.c_double()
        args = [self.h, _dp(x), _dp(rhs)]
        if xd is not None:
            xd = _f64(xd)
            args.append(_dp(xd))
        args.append(C.c_int(self.n))
        if ld is not None:
            ld = _i32(ld)
            args += [_ip(ld), C.c_int(ld.size)]
        args += [C.c_double(tol), C.c_int(maxiter), C.byref(st), C.byref(nit), C.byref(res),
                 C.byref(rhsn)]
        getattr(self.R, fn)(*args)
        return x, dict(status=st.value, nit=nit.value, res=res.value, rhsn=rhsn.value)

    def bicg(self, x0, rhs, tol=1e-6, maxiter=700):
        return self._run("fgref_bicg", x0, rhs, tol, maxiter)

    def bicg_dir(self, x0, rhs, ld, tol=1e-6, maxiter=700, xd=None):
        fn = "fgref_bicg_dir" if xd is None else "fgref_bicg_dir_xd"
        return self._run(fn, x0, rhs, tol, maxiter, xd=xd, ld=np.asarray(ld))

    def cg(self, x0, rhs, tol=1e-6, maxiter=700):
        return self._run("fgref_cg", x0, rhs, tol, maxiter)

    def cg_dir(self, x0, rhs, xd, ld, tol=1e-6, maxiter=700):
        return self._run("fgref_cg_dir", x0, rhs, tol, maxiter, xd=xd, ld=np.asarray(ld))


def ref_timing(tf, dtmin, dtmax, dt):
    out = np.zeros(3)
    ref_lib().fgref_timing(C.c_double(tf), C.c_double(dtmin), C.c_double(dtmax), C.c_double(dt),
                           _dp(out))
    return dict(dt0=out[0], prefactor0=out[1], prefactor=out[2])
