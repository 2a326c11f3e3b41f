"""Host-side mirror of the reference's call surface for the LLG hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference classes the time-integration
loop drives (reference src/time_integration.cpp:193-206):

    timing       src/time_integration.h:6-59
    LinAlgebra   src/linear_algebra.h:40-118  (base_projection, prepareElements x2, solve,
                                               get_v_max, set_DW_vz, buildInitGuess)

The C++17 drop-in (feellgood_b200/host/feellgood_b200.hpp) exposes the same surface to C++
callers; this module is what tests/ and bench.py drive.  All compute runs in
libfeellgood_b200.so on the GPU; nothing here has a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi
from .capi import check, dp, f64, i32, ip

MU0 = 1.25663706127e-6                 # src/config.h.in:31
GAMMA0 = 1.76085962784e11 * MU0        # src/config.h.in:32
M_2_PI = 0.63661977236758134308        # <cmath>: 2/pi (sic — what the reference multiplies by)

_libc = None


def _c_rand():
    """glibc rand(): the stream `--seed N` / srand(N) controls (reference src/main.cpp:211)."""
    global _libc
    if _libc is None:
        _libc = C.CDLL("libc.so.6")
        _libc.rand.restype = C.c_int
    return _libc.rand()


def c_srand(seed):
    global _libc
    if _libc is None:
        _c_rand()
    _libc.srand(C.c_uint(seed))


def mt19937_uniform01(seed):
    """First draw of std::uniform_real_distribution<>(0,1) from std::mt19937(seed)
    (libstdc++ generate_canonical<double,53>: two 32-bit outputs)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
    x0, x1 = (int(v) for v in bg.random_raw(2))
    r = (x0 + x1 * 4294967296.0) / 18446744073709551616.0
    if r >= 1.0:
        r = math.nextafter(1.0, 0.0)
    return r


class timing:
    """reference src/time_integration.h:6-59 (prefactor kept in sync with dt)."""

    def __init__(self, tf, dtmin, dtmax):
        self.tf, self.DTMIN, self.DTMAX = tf, dtmin, dtmax
        self.TAUR = 100. * dtmax
        self._t = 0.0
        self.set_dt(math.sqrt(dtmin * dtmax))

    def get_dt(self):
        return self._dt

    def set_dt(self, dt):
        self._dt = dt
        t_tilde = dt / self.TAUR
        self.prefactor = 1. + t_tilde * abs(math.log(t_tilde))

    def is_dt_TooSmall(self):
        return self._dt < self.DTMIN

    def inc_t(self):
        self._t += self._dt

    def get_t(self):
        return self._t

    def set_t(self, t):
        self._t = t


class Settings:
    """The Settings fields LinAlgebra reads (reference src/settings.h): paramTetra, paramTriangle,
    TOL, MAXITER, verbose, recenter, recentering_direction, plus the NPI compile-time switch."""

    def __init__(self, paramTetra, paramTriangle=(), TOL=1e-6, MAXITER=700, verbose=False,
                 recenter=False, recentering_direction=capi.FG_IDX_Z, npi_tet=5, npi_tri=4,
                 time_step=1e-11, DUMAX=0.02, evol_columns=None, field=None, field_time=None):
        self.paramTetra = list(paramTetra)
        self.paramTriangle = list(paramTriangle)
        self.TOL, self.MAXITER, self.verbose = TOL, MAXITER, verbose
        self.recenter, self.recentering_direction = recenter, recentering_direction
        self.npi_tet, self.npi_tri = npi_tet, npi_tri
        # what Fem::time_integration reads (feellgood_b200.fem): outputs.evol_time_step,
        # time_integration.max(du) (default-settings.yml), outputs.evol_columns, and the applied
        # field: `field` t -> 3-vector in A/m (RtoR3) or `field_time` t -> amplitude (R4toR3)
        self.time_step, self.DUMAX, self.evol_columns = time_step, DUMAX, evol_columns
        self.field, self.field_time = field, field_time


class LinAlgebra:
    """reference src/linear_algebra.h:40-118 on one B200.

    `mesh` is a feellgood_b200.meshgen.Mesh (what Mesh::mesh hands over: sorted, scaled nodes,
    zero-based connectivity, region ids, dMs).  The node state the reference keeps inside
    Mesh::mesh (u, v, phi, phiv CURRENT/NEXT) lives in the device context; `set_state`,
    `set_potentials`, `evolution`, `get_state` are the mesh-side accessors the loop needs.
    """

    def __init__(self, settings, mesh, device=0):
        self._init_ctx(settings, mesh, device,
                       lambda L, cm, cp, dev, h: L.fg_create(cm, cp, dev, h))

    def _init_ctx(self, settings, mesh, device, create):
        L = capi.lib()
        self._L = L
        self.settings = settings
        self.verbose = settings.verbose
        self.NOD = mesh.NOD
        self._keep = (f64(mesh.node_p), i32(mesh.tet_ind), i32(mesh.tet_reg), i32(mesh.tri_ind),
                      i32(mesh.tri_reg), f64(mesh.tri_dMs))
        p, ti, tr, fi, fr, fd = self._keep
        cm = capi.CMesh(mesh.NOD, dp(p), mesh.NT, ip(ti), ip(tr), mesh.NF, ip(fi), ip(fr), dp(fd))
        pt = (capi.TetPrm * len(settings.paramTetra))(*settings.paramTetra)
        ntri = len(settings.paramTriangle)
        pf = (capi.TriPrm * max(1, ntri))(*settings.paramTriangle)
        cp = capi.CParams(len(settings.paramTetra), pt, ntri, pf, settings.npi_tet,
                          settings.npi_tri, settings.TOL, settings.MAXITER)
        h = C.c_void_p()
        check(create(L, C.byref(cm), C.byref(cp), C.c_int(device), C.byref(h)))
        self._h = h
        self._step_args = None
        out = (C.c_longlong * 10)()
        check(L.fg_get_sizes(h, out))
        (_, self.NT, self.NF, self.n_magTet, self.n_magTri, self.E, self.E_mag, self.n, self.nnz,
         self.nlvd) = list(out)
        self.npi = settings.npi_tet
        lay = (C.c_longlong * 4)()
        check(L.fg_get_layout(h, lay))
        # device layout: index bytes per stored node pair, stored pairs (with padding), fast-path flag
        self.col_bytes, self.stored_pairs, self.iso_fast_path = int(lay[0]), int(lay[1]), bool(lay[2])
        # linear_algebra.h:50-53
        self.idx_dir = settings.recentering_direction if settings.recenter else capi.FG_IDX_UNDEF
        self.DW_vz = 0.0   # never initialised in the reference: explicit here
        self.v_max = 0.0
        self.iter = dict(status=capi.FG_UNDEFINED, nit=0, res=0.0, rhsn=0.0)
        self.last_angle = None

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None):
            self._L.fg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- mesh-side state accessors (Mesh::mesh in the reference) ----
    def set_state(self, u, v=None, phi=None, phiv=None):
        """mesh::init_distrib + first evolution: CURRENT = NEXT = (u, v, phi, phiv)."""
        u = f64(u)
        v = f64(v) if v is not None else None
        phi = f64(phi) if phi is not None else None
        phiv = f64(phiv) if phiv is not None else None
        check(self._L.fg_set_state(self._h, dp(u), dp(v), dp(phi), dp(phiv)))

    def set_next_v(self, v):
        v = f64(v)
        check(self._L.fg_set_next_v(self._h, dp(v)))

    def set_potentials(self, phi, phiv):
        """Nodes::set_phi / set_phiv on NEXT (what the demag solver writes each step)."""
        phi, phiv = f64(phi), f64(phiv)
        check(self._L.fg_set_potentials(self._h, dp(phi), dp(phiv)))

    def get_state(self, step=1, what="uvpq"):
        """(u, v, phi, phiv) of step 0 = CURRENT | 1 = NEXT; `what` selects the arrays."""
        N = self.NOD
        u = np.empty((N, 3)) if "u" in what else None
        v = np.empty((N, 3)) if "v" in what else None
        phi = np.empty(N) if "p" in what else None
        phiv = np.empty(N) if "q" in what else None
        check(self._L.fg_get_state(self._h, C.c_int(step), dp(u), dp(v), dp(phi), dp(phiv)))
        return u, v, phi, phiv

    def get_state_into(self, step, u=None, v=None, phi=None, phiv=None):
        """Same, into caller-owned (e.g. pinned) float64 buffers."""
        check(self._L.fg_get_state(self._h, C.c_int(step), dp(u), dp(v), dp(phi), dp(phiv)))

    def evolution(self):
        """mesh::evolution (src/mesh.h:189-193): NEXT -> CURRENT."""
        check(self._L.fg_commit(self._h))

    def set_ext_space_field(self, field):
        field = f64(field)
        assert field.shape == (self.NT, 3, self.npi)
        check(self._L.fg_set_ext_space_field(self._h, dp(field)))

    # ---- observables of the resident NEXT state (Fem::energy, mesh::avg, mesh::max_angle) ----
    def energy(self, Hext):
        """Fem::energy (src/energy.cpp:5-68): E[exchange, anisotropy, demag, zeeman].  A 3-vector
        is the uniform field (RtoR3), a scalar the amplitude of mesh.extSpaceField (R4toR3)."""
        E = np.empty(4)
        if np.ndim(Hext) == 0:
            check(self._L.fg_energy_space(self._h, C.c_double(float(Hext)), dp(E)))
        else:
            H = f64(Hext)
            assert H.shape == (3,)
            check(self._L.fg_energy(self._h, dp(H), dp(E)))
        return E

    def avg(self, what="u", region=-1):
        """mesh::avg (src/mesh.cpp:89-106) of all three components of u or v (NEXT)."""
        out = np.empty(3)
        check(self._L.fg_avg(self._h, C.c_int({"u": 0, "v": 1}[what]), C.c_int(region), dp(out)))
        return out

    def max_angle(self):
        """mesh::max_angle (src/mesh.h:295-306), radians."""
        a = C.c_double()
        check(self._L.fg_max_angle(self._h, C.byref(a)))
        return a.value

    # ---- charges and the all-pairs demag stand-in (scal_fmm::fmm, src/fmm_demag.h) ----
    def calc_charges(self, which=0):
        """fmm::calc_charges (src/fmm_demag.h:155-185): (srcDen, corr) from u (0) or v (1), NEXT."""
        nsrc = self.n_magTet * self.npi + self.n_magTri * self.settings.npi_tri
        src, corr = np.empty(max(1, nsrc)), np.empty(self.NOD)
        check(self._L.fg_calc_charges(self._h, C.c_int(which), dp(src), dp(corr)))
        return src[:nsrc], corr

    def demag_direct(self, second_order=True):
        """All-pairs stand-in for myFMM.calc_demag: writes phi (and phiv) of NEXT on the device."""
        check(self._L.fg_demag_direct(self._h, C.c_int(int(second_order))))

    # ---- the reference's LinAlgebra surface ----
    def base_projection(self, angle=None):
        """src/linear_algebra.cpp:3-11.  With angle=None the angle is drawn exactly like the
        reference does: mt19937 seeded by rand(), one uniform draw, times M_2_PI."""
        if angle is None:
            angle = M_2_PI * mt19937_uniform01(_c_rand())
        self.last_angle = angle
        check(self._L.fg_base_projection(self._h, C.c_double(angle)))

    def prepareElements(self, Hext, t_prm):
        """Both overloads (src/linear_algebra.cpp:26-52 and :54-81): a 3-vector is the uniform
        applied field, a scalar is the amplitude multiplying mesh.extSpaceField."""
        if np.ndim(Hext) == 0:
            check(self._L.fg_prepare_elements_space(
                self._h, C.c_double(float(Hext)), C.c_double(t_prm.get_dt()),
                C.c_double(t_prm.prefactor), C.c_int(self.idx_dir), C.c_double(self.DW_vz)))
        else:
            H = f64(Hext)
            assert H.shape == (3,)
            check(self._L.fg_prepare_elements(
                self._h, dp(H), C.c_double(t_prm.get_dt()), C.c_double(t_prm.prefactor),
                C.c_int(self.idx_dir), C.c_double(self.DW_vz)))

    def _store(self, r):
        self.iter = dict(status=r.status, nit=r.iters, res=r.res, rhsn=r.rhsnorm)
        self.v_max = r.v_max
        return bool(r.failed)

    def solve(self, t_prm):
        """src/solver.cpp:6-90.  Returns True on FAILURE, like the reference."""
        r = capi.StepResult()
        check(self._L.fg_solve(self._h, C.c_double(t_prm.get_dt()), C.byref(r)))
        return self._store(r)

    def step(self, Hext, t_prm, angle=None):
        """base_projection + prepareElements(Hext) + solve in one enqueue (one host sync)."""
        if angle is None:
            angle = M_2_PI * mt19937_uniform01(_c_rand())
        self.last_angle = angle
        sc = self._step_args
        if sc is None:  # argument buffers of the per-step call, marshalled once (5 us per step otherwise)
            r = capi.StepResult()
            sc = self._step_args = ((C.c_double * 3)(), r, C.pointer(r))
        H = sc[0]
        H[0], H[1], H[2] = Hext
        check(self._L.fg_step(self._h, angle, H, t_prm.get_dt(), t_prm.prefactor, self.idx_dir, self.DW_vz, sc[2]))
        return self._store(sc[1])

    def get_v_max(self):
        return self.v_max

    def set_DW_vz(self, vz):
        self.DW_vz = vz

    def buildInitGuess(self, t_prm):
        """src/linear_algebra.cpp:13-24 (as assembled for the system of the last prepareElements)."""
        return self.system(t_prm)[2]

    # ---- taps for the parity tests ----
    def basis(self):
        ep, eq = np.empty((self.NOD, 3)), np.empty((self.NOD, 3))
        check(self._L.fg_get_basis(self._h, dp(ep), dp(eq)))
        return ep, eq

    def elements(self, first=0, count=None):
        count = self.NT - first if count is None else count
        Kp, Lp = np.empty((count, 8, 8)), np.empty((count, 8))
        check(self._L.fg_get_elements(self._h, C.c_int(first), C.c_int(count), dp(Kp), dp(Lp)))
        return Kp, Lp

    def records(self, first=0, count=None):
        """(count, 4, 4) production element records {contrib, BE(3)} per local node (fg_get_records)."""
        count = self.NT - first if count is None else count
        out = np.zeros((count, 4, 4))
        check(self._L.fg_get_records(self._h, C.c_int(first), C.c_int(count), dp(out)))
        return out

    def tri_elements(self, first=0, count=None):
        count = self.NF - first if count is None else count
        Lp = np.empty((max(count, 1), 6))
        check(self._L.fg_get_tri_elements(self._h, C.c_int(first), C.c_int(count), dp(Lp)))
        return Lp[:count]

    def csr(self):
        rowptr, col = np.empty(self.n + 1, dtype=np.int32), np.empty(self.nnz, dtype=np.int32)
        check(self._L.fg_get_csr_pattern(self._h, ip(rowptr), ip(col)))
        return rowptr, col

    def system(self, t_prm):
        val, rhs, x0 = np.empty(self.nnz), np.empty(self.n), np.empty(self.n)
        check(self._L.fg_get_system(self._h, C.c_double(t_prm.get_dt()), dp(val), dp(rhs), dp(x0)))
        return val, rhs, x0

    def jacobi_diagonal(self, t_prm):
        """D = 1/diag(K), 0 on masked dofs (src/algebra/sparseMat.h:174-183), as the solver uses it."""
        D = np.empty(self.n)
        check(self._L.fg_get_precond(self._h, C.c_double(t_prm.get_dt()), dp(D)))
        return D

    def apply_operator(self, x):
        x = f64(x)
        y = np.empty(self.n)
        check(self._L.fg_apply_operator(self._h, dp(x), dp(y)))
        return y

    def solution(self):
        x = np.empty(self.n)
        check(self._L.fg_get_solution(self._h, dp(x)))
        return x

    def tet_tables(self):
        ind = np.empty((self.NT, 4), dtype=np.int32)
        da, w = np.empty((self.NT, 4, 3)), np.empty((self.NT, self.npi))
        check(self._L.fg_get_tet_tables(self._h, ip(ind), dp(da), dp(w)))
        return ind, da, w

    # ---- instrumentation ----
    def kernel_launches(self):
        return int(self._L.fg_kernel_launches(self._h))

    def stream(self):
        return self._L.fg_stream(self._h)

    def set_profiling(self, on=1):
        """0 off, 1 phase timers (adds syncs), 2 CUDA-event pairs around every SpMV launch."""
        check(self._L.fg_set_profiling(self._h, C.c_int(int(on))))

    def spmv_times(self):
        """(total device ms, launches) of the SpMV kernels since profiling mode 2 was set."""
        ms, cnt = C.c_double(), C.c_int()
        check(self._L.fg_get_spmv_times(self._h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    KERNEL_CLASSES = ("basis", "tet", "tri", "assemble", "spmv_setup", "bicg_p", "spmv_v", "bicg_s",
                      "spmv_t", "bicg_xr", "halo", "update", "solve", "other", "gaps")

    def kernel_times(self):
        """{class: (device ms, launches)} since set_profiling(3); call before spmv_times()."""
        ms = (C.c_double * 15)()
        cnt = (C.c_int * 15)()
        check(self._L.fg_get_kernel_times(self._h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    def set_solver(self, kind):
        """"persistent" (default): one cooperative kernel per solve; "multi": one kernel per phase."""
        check(self._L.fg_set_solver(self._h, C.c_int({"persistent": 0, "multi": 1}[kind])))
        lay = (C.c_longlong * 4)()
        check(self._L.fg_get_layout(self._h, lay))
        self.col_bytes = int(lay[0])   # the persistent kernel may address staged images with 2-byte indices

    SOLVE_PHASES = ("kernel", "setup", "A_p", "B_spmv_v", "C_s", "D_spmv_t", "E_xr", "halo_x", "update")

    def solve_times(self):
        """{phase: (ms, count)} of the persistent solve kernel since set_profiling(2|3): "kernel" from
        CUDA events around its launches, the phases from its in-kernel time stamps."""
        ms = (C.c_double * 27)()
        cnt = (C.c_longlong * 9)()
        check(self._L.fg_get_solve_times(self._h, ms, cnt))
        self.solve_breakdown = {k: dict(work_ms=ms[9 + i], cross_gpu_ms=ms[18 + i], count=int(cnt[i]))
                                for i, k in enumerate(self.SOLVE_PHASES) if i > 0}
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.SOLVE_PHASES)}

    def phase_times(self):
        out = (C.c_double * 8)()
        check(self._L.fg_get_phase_times(self._h, out))
        return dict(basis=out[0], elements=out[1], assemble=out[2], solve=out[3])

    def set_operator(self, kind):
        """"node3" (default): matrix-free K; "blocks": assembled 2x2 blocks (A/B checks)."""
        check(self._L.fg_set_operator(self._h, C.c_int({"node3": 0, "blocks": 1}[kind])))

    def krylov_history(self, rows=0):
        """First call: start recording; later: (rows, 8) array rho_1,(v,rt),alpha,|s|^2,(t,s),(t,t),omega,|r|^2."""
        out = np.zeros((max(rows, 1), 8))
        check(self._L.fg_get_krylov_history(self._h, C.c_int(rows), dp(out)))
        return out[:rows]

    def krylov_state(self):
        out = (C.c_double * 8)()
        check(self._L.fg_get_krylov_state(self._h, out))
        return dict(zip(("rho1", "rho2", "alpha", "omega", "res", "rhsn", "nit", "status"), out))

    def bench_spmv(self, reps=20):
        ms = C.c_double()
        check(self._L.fg_bench_spmv(self._h, C.c_int(reps), C.byref(ms)))
        return ms.value
