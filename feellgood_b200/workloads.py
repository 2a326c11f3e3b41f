"""The five BASELINE.json configurations as synthetic inputs.

Host-side only: each builder returns a ``Workload`` = what the reference's Mesh::mesh + Settings +
timing hand to LinAlgebra (mesh, region parameters, initial magnetisation, applied field, dt).
gmsh is not available, so configs 2-5 are generated with the gmsh-free ``Cuboid`` scheme of the
reference's python-modules/meshMaker.py:440-519 (6 tets per hexahedron); config 1 is the reference's
own examples/ellipsoid.msh, committed as the fixture tests/golden/ellipsoid_mesh.npz.

Demag potentials are outside the timed path (BASELINE.json north_star): throughput runs feed
phi = phiv = 0, parity runs feed an analytic surrogate identically to both sides.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import capi, meshgen

MU0 = 1.25663706127e-6
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@dataclass
class Workload:
    name: str
    config: str                 # which BASELINE.json config this is
    mesh: meshgen.Mesh
    tet_regions: list           # dicts of fg_tet_prm fields, index 0 = __default__
    tri_regions: list
    u: np.ndarray               # (NOD, 3) unit vectors
    Hext: np.ndarray            # A/m
    dt: float
    dtmin: float = 1e-16
    dtmax: float = 5e-13
    tol: float = 1e-6           # default-settings.yml:232
    maxiter: int = 700          # default-settings.yml:229
    npi: int = 5
    extra: dict = field(default_factory=dict)

    def settings(self):
        from .linear_algebra import Settings
        return Settings([capi.tet_prm(**r) for r in self.tet_regions],
                        [capi.tri_prm(**r) for r in self.tri_regions], TOL=self.tol,
                        MAXITER=self.maxiter, npi_tet=self.npi, npi_tri=4 if self.npi == 5 else 1)

    def timing(self):
        from .linear_algebra import timing
        t = timing(1.0, self.dtmin, self.dtmax)
        t.set_dt(self.dt)
        return t


def _unit(a):
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def ellipsoid(B=0.02):
    """Config 1: examples/ellipsoid.msh with examples/example1.py settings (default region
    parameters of default-settings.yml:104-125, M0 = (0,0,1), Bext = (0,0,B))."""
    z = np.load(os.path.join(_ROOT, "tests", "golden", "ellipsoid_mesh.npz"))
    m = meshgen.Mesh(node_p=z["node_p"], tet_ind=z["tet_ind"], tet_reg=z["tet_reg"],
                     tri_ind=z["tri_ind"], tri_reg=z["tri_reg"], tri_dMs=z["tri_dMs"])
    u = np.zeros((m.NOD, 3))
    u[:, 2] = 1.0
    # a uniform state along the field is a fixed point; tilt it so that the steps do real work
    u = _unit(u + 0.3 * np.stack([np.sin(3e7 * m.node_p[:, 0]), np.cos(2e7 * m.node_p[:, 1]),
                                  0 * m.node_p[:, 2]], axis=1))
    return Workload("ellipsoid", "configs[0] examples/ellipsoid.msh relaxation", m,
                    [dict(), dict()], [dict(), dict()], u, np.array([0.0, 0.0, B]) / MU0,
                    dt=2e-13)


def film(nx, ny, nz, cell=2.0, name=None, config="film"):
    """Permalloy film of nx x ny x nz cells (the Cuboid scheme), spin-wave initial state
    normalize(cos kx, sin kx cos ky, 0.1), Bext = (0, 10 mT, 0)  (SURVEY.md §8d C5)."""
    m = meshgen.cuboid([0, 0, 0], [cell * nx, cell * ny, cell * nz], nx, ny, nz, scale=1e-9,
                       with_surface=False)
    meshgen.sort_nodes(m)
    p = m.node_p
    kx = 2 * np.pi / (np.ptp(p[:, 0]) + 1e-30) * 3
    ky = 2 * np.pi / (np.ptp(p[:, 1]) + 1e-30) * 2
    u = _unit(np.stack([np.cos(kx * p[:, 0]), np.sin(kx * p[:, 0]) * np.cos(ky * p[:, 1]),
                        0.1 + 0 * p[:, 0]], axis=1))
    return Workload(name or "film_%dx%dx%d" % (nx, ny, nz), config, m,
                    [dict(), dict(alpha=0.02, A=1.3e-11, Ms=8e5)], [dict()], u,
                    np.array([0.0, 10e-3, 0.0]) / MU0, dt=1e-13)


def film20m(scale=1.0):
    """Config 5: synthetic 20M-tet extended film, 1290 x 1290 x 2 cells of 2 nm (19.97M tets,
    5.0M nodes).  `scale` < 1 shrinks the lateral extent at the same cell size and thickness (the
    bounded CPU-baseline sample: same local matrix structure and conditioning)."""
    n = max(4, int(round(1290 * scale)))
    return film(n, n, 2, 2.0, name="film20m" if scale == 1.0 else "film20m_x%.3g" % scale,
                config="configs[4] synthetic 20M-tet extended film 1290x1290x2 cells (2 nm)")


def film20m_k(scale=1.0):
    """Config 5's mesh with a uniaxial anisotropy (K = 3e5 J/m^3 along y, the material of the reference's
    ci-tests/full_test.py:33-36): exercises the GENERAL element kernel (k_tet) instead of the isotropic
    fast path, so that its throughput is measured by the same bench."""
    w = film20m(scale)
    w.name = "film20m_k" if scale == 1.0 else "film20m_k_x%.3g" % scale
    w.config = "configs[4] mesh with uniaxial anisotropy K = 3e5 J/m^3 along y (general element kernel)"
    w.tet_regions = [dict(), dict(alpha=0.02, A=1.3e-11, Ms=8e5, K=3e5, uk=(0, 1, 0))]
    return w


def sp4(scale=1.0):
    """Config 2: muMAG standard problem 4, 500 x 125 x 3 nm permalloy, 250 x 62 x 2 cells
    (186k tets), s-state, reversal field 1 = (-24.6, 4.3, 0) mT."""
    nx, ny = max(4, int(round(250 * scale))), max(2, int(round(62 * scale)))
    m = meshgen.cuboid([0, 0, 0], [2.0 * nx, 125.0 / 62 * ny, 3.0], nx, ny, 2, scale=1e-9,
                       with_surface=False)
    meshgen.sort_nodes(m)
    p = m.node_p
    u = _unit(np.stack([np.ones(m.NOD), 0.1 * np.sin(np.pi * p[:, 0] / np.ptp(p[:, 0])),
                        np.zeros(m.NOD)], axis=1))
    return Workload("sp4", "configs[1] muMAG SP4 film 500x125x3 nm, 186k tets", m,
                    [dict(), dict(alpha=0.02, A=1.3e-11, Ms=8e5)], [dict()], u,
                    np.array([-24.6e-3, 4.3e-3, 0.0]) / MU0, dt=5e-14, dtmax=1e-13)


def disk1m(scale=1.0):
    """Config 3: permalloy nanodisk r = 250 nm, t = 20 nm, vortex state; 2.5 nm cells, 5 layers
    (0.94M tets)."""
    n = max(8, int(round(200 * scale)))
    m = meshgen.disk(250.0, 20.0, n, 5, scale=1e-9, with_surface=False)
    meshgen.sort_nodes(m)
    x, y = m.node_p[:, 0] * 1e9, m.node_p[:, 1] * 1e9
    u = _unit(np.stack([-y, x, 20.0 * np.exp(-(x * x + y * y) / 100.0) + 1e-3], axis=1))
    return Workload("disk1m", "configs[2] permalloy nanodisk vortex, ~1M tets", m,
                    [dict(), dict(alpha=0.01, A=1.3e-11, Ms=8e5)], [dict()], u,
                    np.array([5e-3, 0.0, 0.0]) / MU0, dt=1e-13)


def tube5m(scale=1.0):
    """Config 4: nanotube r1 = 50, r2 = 70 nm (examples/tube.py geometry) lengthened to 1725 nm,
    structured annulus of 8 x 152 x 690 cells of ~2.5 nm (5.03M tets); two opposite azimuthal
    domains (a domain wall at z = 0), Bext = (0, 0, 10 mT)."""
    nz = max(4, int(round(690 * scale)))
    m = meshgen.tube(50.0, 70.0, 2.5 * nz, 8, 152, nz, scale=1e-9, with_surface=False)
    meshgen.sort_nodes(m)
    x, y, z = (m.node_p[:, k] for k in range(3))
    sg = np.tanh(z / 20e-9)
    r = np.sqrt(x * x + y * y)
    u = _unit(np.stack([-y / r * sg, x / r * sg, 0.05 + (1 - sg * sg)], axis=1))
    return Workload("tube5m", "configs[3] nanotube domain wall, ~5M tets", m,
                    [dict(), dict(alpha=0.05, A=1.3e-11, Ms=8e5)], [dict()], u,
                    np.array([0.0, 0.0, 10e-3]) / MU0, dt=1e-13)


BUILDERS = dict(ellipsoid=lambda scale=1.0: ellipsoid(), sp4=sp4, disk1m=disk1m, tube5m=tube5m,
                film20m=film20m, film20m_k=film20m_k)


def build(name, scale=1.0):
    if name not in BUILDERS:
        raise KeyError("unknown workload %r (have %s)" % (name, sorted(BUILDERS)))
    return BUILDERS[name](scale=scale)


# ---- algorithmic bytes (the figures bench.py's roofline is computed from) --------
def spmv_bytes(n, nnz, col_bytes=4, reads_x=False):
    """One launch of the matrix-free operator K = cS P^T (S x I3) P + Dg: per
    stored node pair (= nnz/4) 8 B of S + `col_bytes` of column index (2 when the 16-bit row offsets
    fit the mesh, else 4); per node (= n/2) the gathered 3-vector image of x (32 B, read once from
    HBM), its basis as a unit quaternion (32 B), the node-diagonal pair (Ma, a_w) (16 B), the
    identity-row flag (1 B), the fused second operand (16 B), y written (16 B), one slice pointer per
    32 nodes; x itself (16 B) only in the setup stage and the taps -- inside the BiCGStab loop the
    node's unknowns are recovered from its image."""
    return ((8 + col_bytes) * (nnz // 4) + (32 + 32 + 16 + 1 + 16 + 16 + (16 if reads_x else 0)) * (n // 2)
            + 4 * (n // 64))


def spmv_bytes_blocks(n, nnz):
    """The same product with the assembled 2x2 blocks (round-1 layout: 36 B per block)."""
    return 8 * nnz + nnz + 4 * (n // 2) + 8 * n + 8 * n


def spmv_bytes_csr(n, nnz):
    """SURVEY.md §8d figure for the reference's CSR (12 B per nnz, 20 B per row)."""
    return 12 * nnz + 20 * n


def iter_bytes(n, nnz, col_bytes=4):
    """One BiCGStab iteration as the library runs it (5 kernels): 2 SpMV + per node
    p-update  R r,p,v,D (64) + basis quaternion (32), W p (16) + image of D p (32)   = 144 B
    s-update  R r,v,D (48) + basis quaternion (32),   W s (16) + image of D s (32)   = 128 B
    x/r       R x,p,D,s,t,rt (96),                    W x,r (32)                     = 128 B
    (D p and D s are never stored: 13 vector reads + 5 writes + 2 images, against SURVEY.md §8d's
    minimal 18 passes for a formulation that stores them)."""
    return 2 * spmv_bytes(n, nnz, col_bytes) + (144 + 128 + 128) * (n // 2)


def solve_bytes(n, nnz, iters, col_bytes=4):
    """One launch of the persistent solve kernel (k_llg_solve): the setup product r = b - K x0 (reads x0,
    writes r and rt), `iters` BiCGStab iterations and the node update (reads x, u, the basis: 16 + 32 + 48 B;
    writes u, v of NEXT: 48 B per node)."""
    return (spmv_bytes(n, nnz, col_bytes, reads_x=True) + 16 * (n // 2) + iters * iter_bytes(n, nnz, col_bytes)
            + (16 + 32 + 48 + 48) * (n // 2))


def step_bytes(NOD, NT, n, nnz, iters, col_bytes=4):
    """B_step of SURVEY.md §8d with this library's matrix layout."""
    b_basis = (72 + 32) * NOD   # + the quaternion copy of the basis the Krylov kernels read
    # elements (124 B/tet tables, 112 B/node gathered) + records written and read back (2 x 128 B/tet)
    # + per node rhs, guess and its image, Dg, D (K itself is never written)
    b_asm = 124 * NT + 112 * NOD + 256 * NT + (16 + 16 + 32 + 32 + 16) * NOD
    b_guess = 88 * NOD
    b_setup = spmv_bytes(n, nnz, col_bytes, reads_x=True) + 10 * 8 * n
    b_update = 136 * NOD
    return b_basis + b_asm + b_guess + b_setup + iters * iter_bytes(n, nnz, col_bytes) + b_update


def step_bytes_survey(NOD, NT, n, nnz, iters):
    """B_step exactly as SURVEY.md §8d defines it for the reference's data structures (CSR with 12 B
    per nnz and 20 B per row, assembled K written once per step): the layout-independent yardstick.
    This library moves fewer bytes than that (K is never materialised), so a rate computed from this
    figure is an EQUIVALENT bandwidth and may exceed the HBM peak."""
    b_spmv = spmv_bytes_csr(n, nnz)
    b_iter = 2 * b_spmv + 18 * 8 * n
    b_asm = 124 * NT + 112 * NOD + 8 * nnz + 8 * n
    return 72 * NOD + b_asm + 88 * NOD + (b_spmv + 10 * 8 * n) + iters * b_iter + 136 * NOD
