"""Seeded test cases shared by the CPU (oracle) and GPU (parity) tests.

A Case bundles what the reference's Mesh::mesh + Settings + timing hand to LinAlgebra for one
step: mesh, region parameters, node state (u, v, phi, phiv), field, dt/prefactor and the basis
angle.  `oracle_ctx` builds the CPU oracle for it, `gpu_linalg` the CUDA LinAlgebra.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from feellgood_b200 import capi, meshgen  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MU0 = 1.25663706127e-6


@dataclass
class Case:
    name: str
    mesh: meshgen.Mesh
    tet_regions: list          # dicts of fg_tet_prm fields, index 0 = __default__
    tri_regions: list          # dicts of fg_tri_prm fields, index 0 = __default__
    u: np.ndarray
    v: np.ndarray
    phi: np.ndarray
    phiv: np.ndarray
    Hext: np.ndarray
    dt: float
    dtmax: float
    angle: float
    npi: int = 5
    npi_tri: int = 4
    tol: float = 1e-6
    maxiter: int = 700
    extra: dict = field(default_factory=dict)

    @property
    def prefactor(self):
        t = self.dt / (100. * self.dtmax)
        return 1. + t * abs(np.log(t))


def unit_rows(a):
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def smooth_state(p, rng, Ms=8e5, rough=0.15):
    """Smooth magnetisation + small random roughness, a velocity field, and analytic surrogate
    potentials (demag is outside the path, both sides get identical phi/phiv)."""
    L = p.max(axis=0) - p.min(axis=0)
    L[L == 0] = 1.0
    x = (p - p.min(axis=0)) / L
    u = np.stack([np.cos(2.1 * x[:, 0] + 0.3) * np.cos(1.3 * x[:, 1]),
                  np.sin(2.1 * x[:, 0] + 0.3) * np.cos(0.7 * x[:, 2] + 0.2),
                  0.35 + 0.5 * np.sin(1.7 * x[:, 1] + x[:, 2])], axis=1)
    u = unit_rows(u + rough * rng.standard_normal(u.shape))
    w = rng.standard_normal(u.shape)
    v = np.cross(u, w) * 2.0e9                      # tangent velocities, ~rad/ns
    c = np.array([0.3, -0.2, 0.5])
    phi = Ms * (p @ c)
    phiv = 0.1 * phi / 7e-15
    return u, v, phi, phiv


def small_cuboid(npi=5, seed=5489, nx=6, ny=5, nz=4, with_nonmag=True, aniso=True):
    """Two magnetic regions (uniaxial / cubic anisotropy) + a non-magnetic one, surface triangles
    with Neel anisotropy on part of the boundary, one suppress_charges surface."""
    rng = np.random.default_rng(seed)
    m = meshgen.cuboid([0, 0, 0], [12.0 * nx / 6, 10.0 * ny / 5, 8.0 * nz / 4], nx, ny, nz, scale=1e-9)
    # jitter interior-and-all nodes a little so that tets are not all congruent
    m.node_p = m.node_p + 0.12e-9 * rng.uniform(-1, 1, size=m.node_p.shape)
    cen = m.node_p[m.tet_ind].mean(axis=1)
    xr = (cen[:, 0] - m.node_p[:, 0].min()) / np.ptp(m.node_p[:, 0])
    reg = np.where(xr < 0.45, 1, 2).astype(np.int32)
    if with_nonmag:
        reg = np.where(xr > 0.8, 3, reg).astype(np.int32)
    m.tet_reg = reg
    tcen = m.node_p[m.tri_ind].mean(axis=1)
    zr = (tcen[:, 2] - m.node_p[:, 2].min()) / np.ptp(m.node_p[:, 2])
    m.tri_reg = np.where(zr > 0.99, 1, np.where(zr < 0.01, 2, 3)).astype(np.int32)
    meshgen.sort_nodes(m)
    tet_regions = [dict(),
                   dict(alpha=0.05, A=1.3e-11, Ms=8e5, K=3e5, uk=(0, 1, 0)),
                   dict(alpha=0.5, A=1e-11, Ms=795774.7, K3=-1.2e4,
                        ex=(1 / np.sqrt(2), 1 / np.sqrt(2), 0), ey=(-1 / np.sqrt(2), 1 / np.sqrt(2), 0),
                        ez=(0, 0, 1)),
                   dict(Ms=0.0)]
    if not with_nonmag:
        tet_regions = tet_regions[:3]
    if not aniso:   # K = K3 = 0 everywhere: the library's element fast path (k_tet_iso)
        for r in tet_regions:
            r.pop("K", None)
            r.pop("K3", None)
    tri_regions = [dict(), dict(Ks=2.5e-4, uk=(0, 0, 1)), dict(Ks=1.0e-4, uk=(1, 0, 0), suppress_charges=True),
                   dict(Ks=0.0)]
    Ms = [r.get("Ms", 795774.7) for r in tet_regions]
    m.tri_dMs = meshgen.compute_dMs(m, Ms)
    u, v, phi, phiv = smooth_state(m.node_p, rng)
    return Case("small_cuboid" if aniso else "small_cuboid_iso", m, tet_regions, tri_regions, u, v, phi, phiv,
                Hext=np.array([-24.6e-3, 4.3e-3, 1e-3]) / MU0, dt=2.0e-14, dtmax=5e-13,
                angle=0.35580211334117789, npi=npi, npi_tri=4 if npi == 5 else 1)


def ellipsoid(npi=5, K=3e5, seed=5489):
    """BASELINE.json config 1 mesh (reference examples/ellipsoid.msh, committed as a fixture) with
    the ci-tests/full_test.py material (K = 3e5 along y) and field Bext = (1, 0, -1) T."""
    z = np.load(os.path.join(GOLDEN, "ellipsoid_mesh.npz"))
    m = meshgen.Mesh(node_p=z["node_p"], tet_ind=z["tet_ind"], tet_reg=z["tet_reg"],
                     tri_ind=z["tri_ind"], tri_reg=z["tri_reg"], tri_dMs=z["tri_dMs"])
    rng = np.random.default_rng(seed)
    u, v, phi, phiv = smooth_state(m.node_p, rng, Ms=795774.7, rough=0.05)
    tet_regions = [dict(), dict(K=K, uk=(0, 1, 0))]
    tri_regions = [dict(), dict()]
    return Case("ellipsoid", m, tet_regions, tri_regions, u, v, phi, phiv,
                Hext=np.array([1.0, 0.0, -1.0]) / MU0, dt=5e-14, dtmax=1e-12,
                angle=0.54963651201209474, npi=npi, npi_tri=4 if npi == 5 else 1)


def film(nx, ny, nz, cell=2.0, seed=5489, npi=5):
    """Cuboid film of nx x ny x nz cells (BASELINE configs 2 and 5 family), permalloy."""
    rng = np.random.default_rng(seed)
    m = meshgen.cuboid([0, 0, 0], [cell * nx, cell * ny, cell * nz], nx, ny, nz, scale=1e-9,
                       with_surface=False)
    meshgen.sort_nodes(m)
    p = m.node_p
    kx = 2 * np.pi / (p[:, 0].max() - p[:, 0].min() + 1e-30) * 3
    ky = 2 * np.pi / (p[:, 1].max() - p[:, 1].min() + 1e-30) * 2
    u = unit_rows(np.stack([np.cos(kx * p[:, 0]), np.sin(kx * p[:, 0]) * np.cos(ky * p[:, 1]),
                            0.1 + 0 * p[:, 0]], axis=1))
    v = np.zeros_like(u)
    tet_regions = [dict(), dict(alpha=0.02, A=1.3e-11, Ms=8e5)]
    return Case("film_%dx%dx%d" % (nx, ny, nz), m, tet_regions, [dict()], u, v,
                np.zeros(m.NOD), np.zeros(m.NOD), Hext=np.array([0.0, 10e-3, 0.0]) / MU0,
                dt=1e-13, dtmax=5e-13, angle=0.23551244070763139, npi=npi,
                npi_tri=4 if npi == 5 else 1)


# ------------------------------------------------------------------------------------------
def oracle_ctx(case):
    from oracle import fg_oracle_py as fo
    pt = [fo.tet_prm(**r) for r in case.tet_regions]
    pf = [fo.tri_prm(**r) for r in case.tri_regions]
    return fo.OracleCtx(case.mesh, pt, pf, npi=case.npi, npi_tri=case.npi_tri, tol=case.tol,
                        maxiter=case.maxiter)


def gpu_linalg(case, device=0):
    from feellgood_b200 import LinAlgebra, Settings
    pt = [capi.tet_prm(**r) for r in case.tet_regions]
    pf = [capi.tri_prm(**r) for r in case.tri_regions]
    s = Settings(pt, pf, TOL=case.tol, MAXITER=case.maxiter, npi_tet=case.npi, npi_tri=case.npi_tri)
    return LinAlgebra(s, case.mesh, device=device)


class FixedTiming:
    """timing with dt fixed by the case (prefactor from the case's dtmax)."""

    def __init__(self, case):
        self._dt, self.prefactor = case.dt, case.prefactor

    def get_dt(self):
        return self._dt


def rel_max(a, b):
    """max |a-b| / max |b| (norm-wise relative error, 0/0 = 0)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = np.max(np.abs(b)) if b.size else 0.0
    return 0.0 if d == 0.0 else d / s


# ------------------------------------------------------------------------------------------
class OracleLinAlgebra:
    """The CPU oracle behind the LinAlgebra surface that feellgood_b200.fem.Fem drives, so that the
    very same loop code can be run on the checker and on the GPU (test infrastructure)."""

    def __init__(self, oc):
        self.oc = oc
        self.v_max = 0.0
        self.iter = {}

    def base_projection(self, angle=None):
        from feellgood_b200.linear_algebra import M_2_PI, _c_rand, mt19937_uniform01
        if angle is None:
            angle = M_2_PI * mt19937_uniform01(_c_rand())
        self.oc.base_projection(angle)

    def prepareElements(self, Hext, t_prm):
        if np.ndim(Hext) == 0:
            self.oc.prepare_elements_space(float(Hext), t_prm.get_dt(), t_prm.prefactor)
        else:
            self.oc.prepare_elements(np.asarray(Hext, dtype=np.float64), t_prm.get_dt(), t_prm.prefactor)

    def solve(self, t_prm):
        failed = self.oc.solve(t_prm.get_dt())
        self.iter = self.oc.iter_info()
        self.v_max = self.oc.v_max()
        return failed

    def get_v_max(self):
        return self.v_max

    def evolution(self):
        self.oc.evolution()

    def energy(self, Hext):
        return self.oc.energy_space(float(Hext)) if np.ndim(Hext) == 0 else self.oc.energy(Hext)

    def avg(self, what="u", region=-1):
        return self.oc.avg({"u": 0, "v": 1}[what], region)

    def max_angle(self):
        return self.oc.max_angle()

    def get_state(self, step=1, what="uvpq"):
        return self.oc.get_state(step)

    def set_potentials(self, phi, phiv):
        self.oc.set_potentials_next(phi, phiv)


def local_demag_surrogate(mesh, Ms=795774.7):
    """A cheap state-dependent stand-in for the demag solver (ScalFMM is outside the path): the
    callback Fem.compute_all runs where the reference calls myFMM.calc_demag."""
    p = mesh.node_p / np.max(np.abs(mesh.node_p))

    def demag(la):
        u, v = la.get_state(1, "uv")[:2]
        phi = 2e-9 * Ms * np.sum(u * p, axis=1)
        phiv = 2e-9 * Ms * np.sum(v * p, axis=1)
        la.set_potentials(phi, phiv)
    return demag
