"""GPU parity at BASELINE.json's FULL sizes (configs 2-5: sp4 186k, disk1m 0.94M, tube5m 5.0M,
film20m 20M tets), where the C oracle is too slow to run whole: size-independent properties plus
sampled comparisons against the independent dense numpy restatement (tests/np_restatement.py).

  * sizes: n = 2 NOD, nnz = 4 (NOD + 2 E) (reference src/solver.h:80-101)
  * basis orthonormal and tangent at every node (unit-tests/ut_node.cpp:58-121, 5e-15)
  * sampled tets: Kp (8x8), Lp (8) vs the dense Eigen-style restatement, 1e-12 (north-star tolerance)
  * sampled node rows of the global system: K x and L_rhs rebuilt from the elements of every tet
    around the node (solver::buildMat / buildVect, src/solver.h:110-143) vs the device operator
    (matrix-free, K never materialised) and the device right-hand side, 1e-12
  * linearity of the operator
  * solve: converged, TRUE residual |b - K x| <= TOL |b| recomputed through the operator tap,
    |u| = 1 after the node update, v tangent to u, v_max consistent (src/solver.cpp:74-88)
  * the same step repeated from the same state is bit-identical
"""
import ctypes as C

import numpy as np
import pytest

import np_restatement as npr
from cases import rel_max

pytestmark = pytest.mark.gpu

WORKLOADS = ["sp4", "disk1m", "tube5m", "film20m"]
ANGLE = 0.41887902047863906
TOL_ELEM = 1e-12


def smooth_potentials(p, Ms, dt):
    c = np.array([0.3, -0.2, 0.5])
    phi = Ms * (p @ c)
    return phi, 0.1 * phi / (70.0 * dt)


@pytest.fixture(scope="module", params=WORKLOADS)
def full(request, gpu_lib):
    from feellgood_b200 import LinAlgebra, workloads
    w = workloads.build(request.param)
    la = LinAlgebra(w.settings(), w.mesh)
    rng = np.random.default_rng(5489)
    p = w.mesh.node_p
    # tangent velocities of a few rad/ns and smooth surrogate potentials, so that every term of
    # Tet::integrales is exercised (demag is outside the path: SURVEY §8d)
    v = np.cross(w.u, rng.standard_normal(w.u.shape)) * 2.0e9
    phi, phiv = smooth_potentials(p, 8e5, w.dt)
    la.set_state(w.u, v, phi, phiv)
    t = w.timing()
    la.base_projection(ANGLE)
    la.prepareElements(w.Hext, t)
    # oriented connectivity as the library keeps it (Tet::orientate, src/tetra.cpp:410-424); only
    # the indices are fetched (da and weights of 20M tets would be 2.7 GB of host copies)
    from feellgood_b200.capi import check, ip
    tet_ind = np.empty((w.mesh.NT, 4), dtype=np.int32)
    check(la._L.fg_get_tet_tables(la._h, ip(tet_ind), None, None))
    state = dict(w=w, la=la, t=t, v=v, phi=phi, phiv=phiv, tet_ind=tet_ind)
    yield state
    la.close()


def system_rhs_x0(la, t):
    """fg_get_system without the 8-byte-per-nnz value copy (val = NULL)."""
    from feellgood_b200.capi import check, dp
    rhs, x0 = np.empty(la.n), np.empty(la.n)
    check(la._L.fg_get_system(la._h, C.c_double(t.get_dt()), None, dp(rhs), dp(x0)))
    return rhs, x0


def region_prm(w, reg):
    full = dict(alpha=0.5, A=1e-11, Ms=795774.7, K=0.0, uk=(0, 0, 1), K3=0.0, ex=(1, 0, 0),
                ey=(0, 1, 0), ez=(0, 0, 1))
    full.update(w.tet_regions[reg] if reg < len(w.tet_regions) else {})
    return full


def test_sizes(full):
    la, m = full["la"], full["w"].mesh
    assert la.n == 2 * m.NOD and la.NT == m.NT
    assert la.nnz == 4 * (m.NOD + 2 * la.E_mag)
    assert la.nlvd == 0 and la.n_magTet == m.NT


def test_basis_orthonormal(full):
    la, u = full["la"], full["w"].u
    ep, eq = la.basis()
    dot = lambda a, b: np.einsum("ij,ij->i", a, b)
    assert np.max(np.abs(dot(ep, u))) < 5e-15 and np.max(np.abs(dot(eq, u))) < 5e-15
    assert np.max(np.abs(dot(ep, eq))) < 5e-15
    assert np.max(np.abs(np.linalg.norm(ep, axis=1) - 1)) < 5e-15
    assert np.max(np.abs(np.linalg.norm(eq, axis=1) - 1)) < 5e-15
    # eq = u x ep rotated with ep: right-handed triad
    assert np.max(np.abs(np.cross(u, ep) - eq)) < 5e-15
    full["ep"], full["eq"] = ep, eq


def dense_element(full, tet, ind):
    """Kp, Lp of one tet from the dense restatement; geometry recomputed from the node positions."""
    w, m = full["w"], full["w"].mesh
    npi = w.npi
    da, wt = npr.tet_geometry(m.node_p[ind], npi)
    Hext = np.repeat(np.asarray(w.Hext, dtype=float)[:, None], npi, axis=1)
    t = full["t"]
    return npr.tet_integrales(region_prm(w, int(m.tet_reg[tet])), t.get_dt(), t.prefactor, da, wt,
                              w.u[ind], full["v"][ind], full["phi"][ind], full["phiv"][ind],
                              full["ep"][ind], full["eq"][ind], Hext)


def test_sampled_elements(full):
    la, m = full["la"], full["w"].mesh
    if "ep" not in full:
        full["ep"], full["eq"] = la.basis()
    rng = np.random.default_rng(7)
    starts = np.unique(np.concatenate([[0, m.NT - 8], rng.integers(0, m.NT - 8, size=10)]))
    worstK = worstL = 0.0
    for s in starts:
        Kp, Lp = la.elements(int(s), 8)
        for k in range(8):
            tet = int(s) + k
            ind = full["tet_ind"][tet]
            Kn, Ln = dense_element(full, tet, ind)
            worstK = max(worstK, rel_max(Kp[k], Kn))
            worstL = max(worstL, rel_max(Lp[k], Ln))
    assert worstK < TOL_ELEM, worstK
    assert worstL < TOL_ELEM, worstL


def test_sampled_system_rows(full):
    """Rows 2a, 2a+1 of K x and of L_rhs rebuilt on the host from the element blocks of every tet
    around node a (the reference's scatter), against the device's matrix-free product and rhs."""
    la, w, t = full["la"], full["w"], full["t"]
    m = w.mesh
    if "ep" not in full:
        full["ep"], full["eq"] = la.basis()
    tet_ind = full["tet_ind"]
    rhs, _ = system_rhs_x0(la, t)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(la.n)
    y = la.apply_operator(x)
    nodes = np.unique(np.concatenate([[0, m.NOD - 1, m.NOD // 2], rng.integers(0, m.NOD, size=9)]))
    worst_y = worst_b = 0.0
    scale_y, scale_b = np.max(np.abs(y)), np.max(np.abs(rhs))
    for a in nodes:
        tets = np.nonzero(np.any(tet_ind == a, axis=1))[0]
        assert tets.size > 0
        ya, ba = np.zeros(2), np.zeros(2)
        for tet in tets:
            ind = tet_ind[tet]
            Kn, Ln = dense_element(full, int(tet), ind)
            ie = int(np.nonzero(ind == a)[0][0])
            for di in range(2):
                ba[di] += Ln[di * 4 + ie]
                for dj in range(2):
                    for je in range(4):
                        ya[di] += Kn[di * 4 + ie, dj * 4 + je] * x[2 * ind[je] + dj]
        worst_y = max(worst_y, np.max(np.abs(ya - y[2 * a:2 * a + 2])) / scale_y)
        worst_b = max(worst_b, np.max(np.abs(ba - rhs[2 * a:2 * a + 2])) / scale_b)
    assert worst_y < TOL_ELEM, worst_y
    assert worst_b < TOL_ELEM, worst_b


def test_operator_linearity(full):
    la = full["la"]
    rng = np.random.default_rng(13)
    x, z = rng.standard_normal(la.n), rng.standard_normal(la.n)
    a, b = 0.7, -1.9
    lhs = la.apply_operator(a * x + b * z)
    rhs = a * la.apply_operator(x) + b * la.apply_operator(z)
    assert rel_max(lhs, rhs) < 1e-13


def test_solve_properties_and_reproducibility(full):
    from feellgood_b200.linear_algebra import GAMMA0
    la, w, t = full["la"], full["w"], full["t"]
    rhs, _ = system_rhs_x0(la, t)
    failed = la.solve(t)
    assert failed is False and la.iter["status"] == 0
    assert 0 < la.iter["nit"] < 100
    assert la.iter["res"] <= w.tol * la.iter["rhsn"]
    assert abs(la.iter["rhsn"] - np.linalg.norm(rhs)) <= 1e-12 * la.iter["rhsn"]
    x = la.solution()
    r = rhs - la.apply_operator(x)
    assert np.linalg.norm(r) <= 1.01 * w.tol * np.linalg.norm(rhs)
    u1, v1, phi1, phiv1 = la.get_state(1)
    assert np.max(np.abs(np.linalg.norm(u1, axis=1) - 1)) < 1e-15
    # v = gamma0 (vp ep + vq eq) lies in the tangent plane of u CURRENT (src/solver.cpp:80-81)
    ep, eq = la.basis()
    vp, vq = x[0::2], x[1::2]
    v_expect = GAMMA0 * (vp[:, None] * ep + vq[:, None] * eq)
    assert rel_max(v1, v_expect) < 1e-14
    assert abs(la.get_v_max() - GAMMA0 * np.sqrt(np.max(vp * vp + vq * vq))) <= 1e-13 * la.get_v_max()
    un = w.u + t.get_dt() * v_expect
    un /= np.linalg.norm(un, axis=1, keepdims=True)
    assert np.max(np.abs(u1 - un)) < 1e-15
    assert np.array_equal(phi1, full["phi"]) and np.array_equal(phiv1, full["phiv"])
    nit = la.iter["nit"]
    # same state, same angle, same field again: bit-identical trajectory
    la.set_state(w.u, full["v"], full["phi"], full["phiv"])
    la.base_projection(ANGLE)
    la.prepareElements(w.Hext, t)
    assert la.solve(t) is False and la.iter["nit"] == nit
    u2, v2, _, _ = la.get_state(1)
    assert np.array_equal(u1, u2) and np.array_equal(v1, v2)


@pytest.mark.parametrize("name,nsteps", [("sp4", 20), ("disk1m", 8), ("tube5m", 5), ("film20m", 3)])
def test_trajectory_vs_oracle_at_full_size(oracle, gpu_lib, name, nsteps):
    """BASELINE configs 2 to 5 at their FULL size against the CPU oracle run whole, all host threads (the
    north-star criterion: average magnetisation and energies after a fixed number of steps within 1e-6
    relative; per-step solutions within the solver tolerance).  The 5 M- and 20 M-tet meshes take the
    oracle ~1 and ~4 s per step on 16 threads."""
    import cases
    from feellgood_b200 import workloads
    from feellgood_b200.linear_algebra import M_2_PI, mt19937_uniform01
    w = workloads.build(name)
    z = np.zeros(w.mesh.NOD)
    case = cases.Case(w.name, w.mesh, w.tet_regions, w.tri_regions, w.u, np.zeros_like(w.u), z, z, w.Hext,
                      w.dt, w.dtmax, ANGLE, npi=w.npi, npi_tri=4 if w.npi == 5 else 1, tol=w.tol,
                      maxiter=w.maxiter)
    oc, la = cases.oracle_ctx(case), cases.gpu_linalg(case)
    import os
    oc.set_num_threads(os.cpu_count() or 1)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    la.set_state(case.u, case.v, case.phi, case.phiv)
    t = cases.FixedTiming(case)
    for step in range(nsteps):
        ang = M_2_PI * mt19937_uniform01(2000 + step)
        oc.base_projection(ang)
        oc.prepare_elements(case.Hext, case.dt, case.prefactor)
        failed_o = oc.solve(case.dt)
        failed_g = la.step(case.Hext, t, angle=ang)
        assert failed_g == failed_o is False, step
        assert abs(la.iter["nit"] - oc.iter_info()["nit"]) <= 3, step
        oc.evolution()
        la.evolution()
    u_o, u_g = oc.get_state(0)[0], la.get_state(0, "u")[0]
    # 1e-6 relative to the unit magnetisation, component-wise (the vortex average is ~0 by symmetry,
    # where a norm-wise relative bound is ill-posed: two converged solves differ by ~cond*TOL)
    assert np.max(np.abs(u_g - u_o)) < 1e-6
    assert np.max(np.abs(u_g.mean(axis=0) - u_o.mean(axis=0))) < 1e-6 * max(1e-2, np.max(np.abs(u_o.mean(axis=0))))
    assert np.max(np.abs(la.avg("u") - oc.avg(0))) < 1e-6 * max(1e-2, np.max(np.abs(oc.avg(0))))
    E_o, E_g = oc.energy(case.Hext), la.energy(case.Hext)
    assert np.max(np.abs(E_g - E_o)) <= 1e-6 * np.max(np.abs(E_o)), (E_g, E_o)
    assert abs(la.max_angle() - oc.max_angle()) < 1e-6
    la.close()
    oc.close()
