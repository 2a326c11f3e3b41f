"""Cross-checks the oracle's composition of Tet::integrales / Tri::integrales / the scatter /
the node update against an independent dense numpy restatement (tests/np_restatement.py) and
against the structured closed form of SURVEY.md §8a — the part no reference unit test pins."""
import numpy as np
import pytest

import cases
import np_restatement as npr
from cases import rel_max


def _rand_unit(rng, n):
    a = rng.standard_normal((n, 3))
    return a / np.linalg.norm(a, axis=1, keepdims=True)


PRMS = [dict(alpha=0.5, A=1e-11, Ms=795774.7),
        dict(alpha=0.05, A=1.3e-11, Ms=8e5, K=3e5, uk=(0, 1, 0)),
        dict(alpha=0.1, A=1e-11, Ms=6e5, K3=-1.2e4, ex=(2 ** -0.5, 2 ** -0.5, 0),
             ey=(-2 ** -0.5, 2 ** -0.5, 0), ez=(0, 0, 1)),
        dict(alpha=0.02, A=1.3e-11, Ms=8e5, K=-2e5, uk=(0.6, 0, 0.8), K3=5e3)]


@pytest.mark.parametrize("npi", [5, 1])
@pytest.mark.parametrize("ip", range(len(PRMS)))
@pytest.mark.parametrize("idx_dir", [-1, 0, 2])
def test_tet_integrales_vs_dense_numpy(oracle, npi, ip, idx_dir):
    rng = np.random.default_rng(100 * ip + npi + 7 * (idx_dir + 1))
    prm = PRMS[ip]
    full = dict(K=0.0, uk=(0, 0, 1), K3=0.0, ex=(1, 0, 0), ey=(0, 1, 0), ez=(0, 0, 1))
    full.update(prm)
    for trial in range(8):
        p4 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]]) * 2e-9 \
            + 0.3e-9 * rng.uniform(-1, 1, (4, 3))
        ind, da_o, w_o, detJ, _ = oracle.tet_setup(p4, [0, 1, 2, 3], npi)
        p4 = p4[ind]
        da, w = npr.tet_geometry(p4, npi)
        assert rel_max(da, da_o) < 1e-13 and rel_max(w, w_o) < 1e-13
        u, v = _rand_unit(rng, 4), 1e9 * rng.standard_normal((4, 3))
        phi, phiv = 1e-3 * rng.standard_normal(4), 1e10 * rng.standard_normal(4)
        ep, eq = np.zeros((4, 3)), np.zeros((4, 3))
        for i in range(4):
            ep[i], eq[i] = npr.set_basis(u[i], 0.4)
            e1, e2 = oracle.node_set_basis(u[i], 0.4)
            assert np.max(np.abs(e1 - ep[i])) < 5e-16 and np.max(np.abs(e2 - eq[i])) < 5e-16
        Hext = 1e4 * rng.standard_normal((3, npi))
        dt = 10 ** rng.uniform(-15, -12.5)
        pref = 1.0 + 0.01 * trial
        Ko, Lo = oracle.tet_integrales(oracle.tet_prm(**prm), dt, pref, da_o, w_o, u, v, phi, phiv,
                                       ep, eq, Hext, idx_dir=idx_dir, Vdrift=12.5)
        Kn, Ln = npr.tet_integrales(full, dt, pref, da_o, w_o, u, v, phi, phiv, ep, eq, Hext,
                                    idx_dir=idx_dir, Vdrift=12.5)
        assert rel_max(Ko, Kn) < 1e-13, (ip, npi, trial)
        assert rel_max(Lo, Ln) < 1e-13, (ip, npi, trial)


def test_structured_closed_form(oracle):
    """K[2a+r, 2b+c] = E_ab (e_r,a . e_c,b) + delta_ab a_w e_r,a . (m_a x e_c,a)
    with row 0 tested by eq and row 1 by ep — what the CUDA kernels implement."""
    rng = np.random.default_rng(3)
    npi = 5
    p4 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]]) * 1e-9
    ind, da, w, detJ, _ = oracle.tet_setup(p4, [0, 1, 2, 3], npi)
    u = _rand_unit(rng, 4)
    ep, eq = zip(*[oracle.node_set_basis(u[i], 0.2) for i in range(4)])
    ep, eq = np.array(ep), np.array(eq)
    prm = dict(alpha=0.3, A=1e-11, Ms=8e5)
    dt, pref = 1e-13, 1.02
    Hext = np.zeros((3, npi))
    Ko, _ = oracle.tet_integrales(oracle.tet_prm(**prm), dt, pref, da, w, u, np.zeros((4, 3)),
                                  np.zeros(4), np.zeros(4), ep, eq, Hext)
    a, _ = npr.tet_tables(npi)
    Abis = 2 * prm["A"] / (npr.MU0 * prm["Ms"])
    s_dt = npr.THETA * dt * npr.GAMMA0
    dU = u.T @ da
    U = u.T @ a
    uHeff = np.full(npi, -Abis * np.sum(dU * dU))
    aeff = npr.calc_alpha_eff(dt, prm["alpha"], uHeff)
    E = pref * s_dt * Abis * w.sum() * (da @ da.T) + np.diag(a @ (w * aeff))
    a_w = a @ w
    e = {0: eq, 1: ep}      # row/test function r: 0 -> eq, 1 -> ep ; column/unknown c: 0 -> ep, 1 -> eq
    f = {0: ep, 1: eq}
    for A_ in range(4):
        for B_ in range(4):
            for r in range(2):
                for c in range(2):
                    k = E[A_, B_] * e[r][A_].dot(f[c][B_])
                    if A_ == B_:
                        k += a_w[A_] * e[r][A_].dot(np.cross(u[A_], f[c][A_]))
                    assert abs(Ko[r * 4 + A_, c * 4 + B_] - k) <= 1e-13 * np.max(np.abs(Ko))


def test_tri_integrales_vs_numpy(oracle):
    rng = np.random.default_rng(11)
    for npi in (4, 1):
        a, pds = npr.tri_tables(npi)
        p3 = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0.]]) * 1e-9 + 1e-10 * rng.uniform(-1, 1, (3, 3))
        surf = 0.5 * np.linalg.norm(np.cross(p3[1] - p3[0], p3[2] - p3[0]))
        w = 2.0 * surf * pds
        u = _rand_unit(rng, 3)
        ep, eq = zip(*[oracle.node_set_basis(u[i], 0.1) for i in range(3)])
        ep, eq = np.array(ep), np.array(eq)
        uk = np.array([0.6, 0.0, 0.8])
        Lo = oracle.tri_integrales(oracle.tri_prm(Ks=2.5e-4, uk=uk), 8e5, w, u, ep, eq)
        Ln = npr.tri_integrales(2.5e-4, uk, 8e5, w, u, ep, eq)
        assert rel_max(Lo, Ln) < 1e-13


def test_scatter_and_solve_vs_dense(oracle):
    """solver<2>::buildMat / buildVect / mask / init guess / node update restated densely from the
    oracle's element blocks (src/solver.h:110-143, src/solver.cpp:25-88)."""
    case = cases.small_cuboid(nx=4, ny=3, nz=2)
    oc = cases.oracle_ctx(case)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    oc.base_projection(case.angle)
    oc.prepare_elements(case.Hext, case.dt, case.prefactor)
    n = oc.n
    K = np.zeros((n, n))
    L = np.zeros(n)
    ind = oc.tet_ind()
    mag, lvd = oc.masks()
    for t in range(oc.NT):
        Kp, Lp = oc.element(t)
        if not np.any(Kp):
            continue
        for ie in range(4):
            for di in range(2):
                L[2 * ind[t, ie] + di] += Lp[di * 4 + ie]
                for je in range(4):
                    for dj in range(2):
                        K[2 * ind[t, ie] + di, 2 * ind[t, je] + dj] += Kp[di * 4 + ie, dj * 4 + je]
    tri = case.mesh.tri_ind
    for f in range(oc.NF):
        Lp = oc.tri_element(f)
        supp = case.tri_regions[case.mesh.tri_reg[f]].get("suppress_charges", False)
        if supp or not mag[tri[f]].all():
            continue                      # magTri excludes them (src/mesh.h:127-131)
        for ie in range(3):
            for di in range(2):
                L[2 * tri[f, ie] + di] += Lp[di * 3 + ie]
    L[lvd] = 0.0
    for i in lvd:
        K[i, i] = 1.0
    oc.assemble()
    val, rhs, x0 = oc.system()
    rp, col = oc.csr()
    Kd = np.zeros((n, n))
    for i in range(n):
        Kd[i, col[rp[i]:rp[i + 1]]] = val[rp[i]:rp[i + 1]]
    assert rel_max(Kd, K) < 1e-14 and rel_max(rhs, L) < 1e-14
    ep, eq = oc.get_basis()
    g0 = npr.GAMMA0
    x0n = np.stack([np.einsum("ij,ij->i", case.v, ep), np.einsum("ij,ij->i", case.v, eq)], 1) / g0
    x0n[~mag] = 0.0
    assert rel_max(x0, x0n.reshape(-1)) < 1e-14
    failed = oc.solve(case.dt)
    assert not failed
    x = oc.system()[2]
    xd = np.linalg.solve(K, L)
    assert np.linalg.norm(x - xd) <= 1e-5 * np.linalg.norm(xd)
    # node update (src/node.h:116-122) from the oracle's own solution
    u1, v1, _, _ = oc.get_state(1)
    vp, vq = g0 * x[0::2], g0 * x[1::2]
    vn = vp[:, None] * ep + vq[:, None] * eq
    un = case.u + case.dt * vn
    un /= np.linalg.norm(un, axis=1, keepdims=True)
    assert np.max(np.abs(u1[mag] - un[mag])) < 1e-15 and rel_max(v1[mag], vn[mag]) < 1e-15
    assert np.array_equal(u1[~mag], case.u[~mag])
    assert abs(oc.v_max() - g0 * np.sqrt(np.max(x[0::2][mag] ** 2 + x[1::2][mag] ** 2))) <= 1e-12 * oc.v_max()
    oc.close()
