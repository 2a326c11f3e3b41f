"""A second, independent restatement of Tet::integrales / Tri::integrales in dense numpy, written
the way the reference writes it with Eigen (dense 12x12 AE, dense 8x12 P, Kp = Perm P AE P^T;
reference src/tetra.cpp:108-148,171-307, src/element.h:81-96, src/triangle.cpp:6-36).

Used only by the CPU tests to cross-check the C oracle's composition of the pieces that the
reference's unit tests pin one by one (SURVEY.md §4 "gaps": no reference test covers integrales
end to end).  Deliberately shares no code with oracle/fg_oracle.c.
"""
import numpy as np

MU0 = 1.25663706127e-6
GAMMA0 = 1.76085962784e11 * MU0
THETA = 0.5


def tet_tables(npi):
    if npi == 1:
        return np.array([[0.25], [0.25], [0.25], [0.25]]), np.array([1. / 6.])
    A, B, C, D, E = 1. / 4., 1. / 6., 1. / 2., -2. / 15., 3. / 40.
    u = np.array([A, B, B, B, C])
    v = np.array([A, B, B, C, B])
    w = np.array([A, B, C, B, B])
    return np.stack([1 - u - v - w, u, v, w]), np.array([D, E, E, E, E])


def tri_tables(npi):
    if npi == 1:
        return np.array([[1. / 3.], [1. / 3.], [1. / 3.]]), np.array([0.5])
    u = np.array([1 / 3., 1 / 5., 3 / 5., 1 / 5.])
    v = np.array([1 / 3., 1 / 5., 1 / 5., 3 / 5.])
    return np.stack([1 - u - v, u, v]), np.array([-27 / 96., 25 / 96., 25 / 96., 25 / 96.])


def tet_geometry(p4, npi):
    """da (4,3), weight (npi,) of an already oriented tet (src/tetra.h:140-163)."""
    J = np.stack([p4[1] - p4[0], p4[2] - p4[0], p4[3] - p4[0]], axis=1)
    detJ = np.linalg.det(J)
    dadu = np.array([[-1., -1., -1.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    da = dadu @ np.linalg.inv(J)
    _, pds = tet_tables(npi)
    return da, detJ * pds


def calc_alpha_eff(dt, alpha, h):
    rdt = GAMMA0 * dt
    M = 2. * alpha * 0.1 / rdt
    out = np.empty_like(h)
    for g, x in enumerate(h):
        if x > 0:
            out[g] = alpha + rdt / 2. * (M if x > M else x)
        else:
            out[g] = alpha / (1. + rdt / (2. * alpha) * M) if x < -M else alpha / (1. - rdt / (2. * alpha) * x)
    return out


def mat_P(ep, eq):
    N = ep.shape[0]
    P = np.zeros((2 * N, 3 * N))
    for d in range(3):
        P[:N, d * N:(d + 1) * N] = np.diag(ep[:, d])
        P[N:, d * N:(d + 1) * N] = np.diag(eq[:, d])
    return P


def tet_integrales(prm, dt, prefactor, da, weight, u, v, phi, phiv, ep, eq, Hext, idx_dir=-1,
                   Vdrift=0.0):
    """prm: dict(alpha, A, Ms, K, uk, K3, ex, ey, ez); node arrays (4,3)/(4,); Hext (3,npi)."""
    npi = weight.size
    a, _ = tet_tables(npi)
    alpha, Ms = prm["alpha"], prm["Ms"]
    Abis = 2.0 * prm["A"] / (MU0 * Ms)
    s_dt = THETA * dt * GAMMA0
    U = u.T @ a                                   # (3, npi)
    V = v.T @ a
    dU = u.T @ da                                 # (3, 3): column k = dU/dx_k
    Hd = np.repeat((-(phi @ da))[:, None], npi, axis=1)
    Hv = np.repeat((-(phiv @ da))[:, None], npi, axis=1)
    uHeff = np.full(npi, -Abis * np.sum(dU * dU))
    Han = np.zeros((3, npi))
    if prm.get("K", 0.0) != 0:
        Kbis = 2.0 * prm["K"] / (MU0 * Ms)
        uk = np.asarray(prm["uk"], dtype=float)
        sd = s_dt / GAMMA0
        for g in range(npi):
            Han[:, g] += (Kbis * uk.dot(U[:, g] + sd * V[:, g])) * uk
        uHeff += Kbis * (U.T @ uk) ** 2
    if prm.get("K3", 0.0) != 0:
        K3bis = 2.0 * prm["K3"] / (MU0 * Ms)
        ex, ey, ez = (np.asarray(prm[k], dtype=float) for k in ("ex", "ey", "ez"))
        sd = s_dt / GAMMA0
        res = np.zeros(npi)
        for g in range(npi):
            uk_u = np.array([ex.dot(U[:, g]), ey.dot(U[:, g]), ez.dot(U[:, g])])
            uk_v = np.array([ex.dot(V[:, g]), ey.dot(V[:, g]), ez.dot(V[:, g])])
            uuu = uk_u * (1.0 - uk_u * uk_u)
            tmp = uk_v * ex
            Han[:, g] += -K3bis * (uuu[0] * ex + uuu[1] * ey + uuu[2] * ez
                                   + sd * tmp * (1 - 3 * uk_u * uk_u))
            res[g] = uk_u.dot(uuu)
        uHeff += -K3bis * res
    Heff = Hd + Hext
    H = Heff.copy()
    uHeff = uHeff + np.sum(U * Heff, axis=0)
    a_eff = calc_alpha_eff(dt, alpha, uHeff)
    # lumping
    AE = np.zeros((12, 12))
    contrib = a @ (weight * a_eff)
    blk = (da @ da.T) * (prefactor * s_dt * Abis * weight.sum()) + np.diag(contrib)
    for d in range(3):
        AE[4 * d:4 * d + 4, 4 * d:4 * d + 4] += blk
    a_w = a @ weight
    AE[4:8, 8:12] -= np.diag(a_w * u[:, 0])
    AE[8:12, 4:8] += np.diag(a_w * u[:, 0])
    AE[0:4, 8:12] += np.diag(a_w * u[:, 1])
    AE[8:12, 0:4] -= np.diag(a_w * u[:, 1])
    AE[0:4, 4:8] -= np.diag(a_w * u[:, 2])
    AE[4:8, 0:4] += np.diag(a_w * u[:, 2])
    P = mat_P(ep, eq)
    perm = [4, 5, 6, 7, 0, 1, 2, 3]      # Eigen: (Perm * M).row(indices[i]) = M.row(i)
    Kp = np.zeros((8, 8))
    Kp[perm, :] = P @ AE @ P.T
    BE = np.zeros((3, 4))
    if idx_dir >= 0:
        dUd = np.repeat(dU[:, idx_dir][:, None], npi, axis=1)
        dVd = np.repeat((v.T @ da[:, idx_dir])[:, None], npi, axis=1)
        for g in range(npi):
            interim = np.zeros((3, 4))
            for i in range(4):
                interim[:, i] = a[i, g] * (alpha * dUd[:, g] + np.cross(U[:, g], dUd[:, g])
                                           + s_dt * (alpha * dVd[:, g] + np.cross(U[:, g], dVd[:, g])
                                                     + np.cross(V[:, g], dUd[:, g])))
            BE += Vdrift * weight[g] * interim
    H = H + Han + (s_dt / GAMMA0) * Hv
    for g in range(npi):
        w = weight[g]
        for i in range(4):
            BE[:, i] -= w * Abis * (da[i, 0] * dU[:, 0] + da[i, 1] * dU[:, 1] + da[i, 2] * dU[:, 2])
            BE[:, i] += w * a[i, g] * H[:, g]
    Lp = np.zeros(8)
    Lp[perm] = P @ BE.reshape(-1)
    return Kp, Lp


def tri_integrales(Ks, uk, dMs, weight, u, ep, eq):
    npi = weight.size
    a, _ = tri_tables(npi)
    uk = np.asarray(uk, dtype=float)
    Kbis = 2.0 * Ks / dMs
    ug = u.T @ a
    BE = np.zeros((3, 3))
    for g in range(npi):
        pf = weight[g] * Kbis * uk.dot(ug[:, g])
        for i in range(3):
            BE[:, i] += pf * a[i, g] * uk
    P = mat_P(ep, eq)
    Lp = np.zeros(6)
    Lp[[3, 4, 5, 0, 1, 2]] = P @ BE.reshape(-1)
    return Lp


def set_basis(u, r):
    """Node::setBasis (src/node.h:73-102)."""
    k = int(np.argmin(np.abs(u)))       # first minimum, like the reference's strict comparisons
    e = np.zeros(3)
    e[k] = 1.0
    e = e - e.dot(u) * u
    e /= np.linalg.norm(e)
    f = np.cross(u, e)
    return np.cos(r) * e - np.sin(r) * f, np.sin(r) * e + np.cos(r) * f
