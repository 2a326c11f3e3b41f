"""Oracle pinning of SURVEY §8(f) rank 1: energies, averages, maximum angle.

The reference's unit-tests/ut_energy.cpp pins the energy formulas against "ref code"; the two
numeric cases (uniaxialAnisotropyEnergy :18-76, demagEnergy :78-168) are ported here with the
same fixture (unit tetrahedron (0,0,0),(1,0,0),(0,1,0),(0,0,1), mt19937(5489)-style random unit
vectors).  The composition Fem::energy / mesh::avg / mesh::max_angle is additionally checked
against an independent numpy restatement on a multi-region mesh.
"""
import numpy as np
import pytest

import cases
from cases import MU0
from feellgood_b200 import meshgen
from oracle import fg_oracle_py as fo

UT_TOL = 5e-16


def unit_tet_mesh(with_faces):
    p = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    tet = np.array([[0, 1, 2, 3]], dtype=np.int32)
    if with_faces:   # ut_energy.cpp:148-153: {ia,ic,ib}, {ib,ic,id}, {ia,id,ic}, {ia,ib,id}
        tri = np.array([[0, 2, 1], [1, 2, 3], [0, 3, 2], [0, 1, 3]], dtype=np.int32)
    else:
        tri = np.zeros((0, 3), dtype=np.int32)
    return meshgen.Mesh(node_p=p, tet_ind=tet, tet_reg=np.zeros(1, dtype=np.int32), tri_ind=tri,
                        tri_reg=np.zeros(len(tri), dtype=np.int32), tri_dMs=np.zeros(len(tri)))


def rand_unit(rng, n):
    th, ph = np.pi * rng.random(n), 2 * np.pi * rng.random(n)
    return np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)


def tet_tables():
    a = np.ctypeslib.as_array(fo.lib().fgo_tet_a(5), shape=(20,)).reshape(4, 5).copy()
    w = np.ctypeslib.as_array(fo.lib().fgo_tet_pds(5), shape=(5,)).copy()
    return a, w


def test_uniaxial_anisotropy_energy(oracle):
    """ut_energy.cpp:18-76."""
    rng = np.random.default_rng(5489)
    u = rand_unit(rng, 4)
    K, uk = rng.random(), rand_unit(rng, 1)[0]
    m = unit_tet_mesh(False)
    oc = fo.OracleCtx(m, [fo.tet_prm(K=K, uk=uk, Ms=1.0)], [fo.tri_prm()], npi=5, npi_tri=4)
    oc.set_state(u)
    E = oc.energy(np.zeros(3))
    a, pds = tet_tables()
    ug = u.T @ a                                   # 3 x NPI (tiny::mult of the ref code)
    weight = pds * 1.0                              # detJ of the unit tet = 1
    ref = sum(weight[g] * (-K * (uk @ ug[:, g]) ** 2) for g in range(5))
    assert abs(E[1] - ref) <= UT_TOL * max(1.0, abs(ref))
    oc.close()


def test_demag_energy_volume_plus_surface(oracle):
    """ut_energy.cpp:78-168: -1/2 mu0 Ms int(u.Hd) over the tet equals the volume-charge term plus
    the surface-charge terms of its four faces computed from the scalar potential."""
    rng = np.random.default_rng(5489)
    u = rand_unit(rng, 4)
    phi = rng.random(4)
    Ms = rng.random()
    m = unit_tet_mesh(True)
    m.tri_dMs = np.full(4, Ms)
    oc = fo.OracleCtx(m, [fo.tet_prm(Ms=Ms)], [fo.tri_prm()], npi=5, npi_tri=4)
    oc.set_state(u, None, phi, None)
    E = oc.energy(np.zeros(3))
    a, pds = tet_tables()
    da = oc.tet_geom()[0][0].reshape(4, 3)
    ug = u.T @ a
    Hd = -(phi @ da)                                # constant over the element
    ref = sum(pds[g] * (-0.5 * MU0 * Ms * (ug[:, g] @ Hd)) for g in range(5))
    assert (E[2] - ref) ** 2 <= UT_TOL
    assert abs(E[2] - ref) <= 1e-13 * abs(ref)
    oc.close()


def numpy_observables(case, oc, Hext=None, space=None, amp=0.0, region=-1):
    """Independent dense restatement of Fem::energy, mesh::avg, mesh::max_angle (NEXT state)."""
    m = case.mesh
    a, pds = tet_tables()
    lib = fo.lib()
    at = np.ctypeslib.as_array(lib.fgo_tri_a(case.npi_tri), shape=(3 * case.npi_tri,)).reshape(3, case.npi_tri)
    pt = np.ctypeslib.as_array(lib.fgo_tri_pds(case.npi_tri), shape=(case.npi_tri,))
    if case.npi == 1:
        a = np.ctypeslib.as_array(lib.fgo_tet_a(1), shape=(4,)).reshape(4, 1)
        pds = np.ctypeslib.as_array(lib.fgo_tet_pds(1), shape=(1,))
    ind = oc.tet_ind()
    da, w = oc.tet_geom()
    u, v, phi, _ = oc.get_state(1)
    E = np.zeros(4)
    num = np.zeros((2, 3))
    vol = 0.0
    regvol = {}
    for t in range(m.NT):
        prm = case.tet_regions[m.tet_reg[t]]
        Ms = prm.get("Ms", 795774.7)
        regvol[m.tet_reg[t]] = regvol.get(m.tet_reg[t], 0.0) + w[t].sum()
        if not Ms > 0:
            continue
        un, pn = u[ind[t]], phi[ind[t]]
        G = un.T @ da[t].reshape(4, 3)             # G[d, k] = d u_d / d x_k
        ug = un.T @ a
        pg = pn @ a
        E[0] += prm.get("A", 1e-11) * np.sum(w[t] * np.sum(G * G))
        E[2] += -0.5 * MU0 * Ms * np.sum(w[t] * np.trace(G) * pg)
        if prm.get("K", 0.0) != 0.0:
            E[1] += -prm["K"] * np.sum(w[t] * (np.asarray(prm.get("uk", (0, 0, 1))) @ ug) ** 2)
        if prm.get("K3", 0.0) != 0.0:
            al = [np.asarray(prm.get(k, d), dtype=float) @ ug for k, d in
                  (("ex", (1, 0, 0)), ("ey", (0, 1, 0)), ("ez", (0, 0, 1)))]
            E[1] += prm["K3"] * np.sum(w[t] * ((al[0] * al[1]) ** 2 + (al[1] * al[2]) ** 2 + (al[2] * al[0]) ** 2))
        if space is None:
            E[3] += -MU0 * Ms * np.sum(w[t] * (Hext @ ug))
        else:
            E[3] += -MU0 * Ms * amp * np.sum(w[t] * np.sum(ug * space[t], axis=0))
        if region in (-1, m.tet_reg[t]):
            num[0] += (w[t][None, :] * ug).sum(axis=1)
            num[1] += (w[t][None, :] * (v[ind[t]].T @ a)).sum(axis=1)
        vol += w[t].sum()
    mag = oc.masks()[0]
    for f in range(m.NF):
        tp = case.tri_regions[m.tri_reg[f]]
        i3 = m.tri_ind[f]
        if not mag[i3].all() or tp.get("suppress_charges", False):
            continue
        p = m.node_p[i3]
        nv = np.cross(p[1] - p[0], p[2] - p[0])
        surf = 0.5 * np.linalg.norm(nv)
        n = nv / np.linalg.norm(nv)
        wg = 2.0 * surf * pt
        ug = u[i3].T @ at
        pg = phi[i3] @ at
        if tp.get("Ks", 0.0) != 0.0:
            E[1] += -tp["Ks"] * np.sum(wg * (np.asarray(tp.get("uk", (0, 0, 1)), dtype=float) @ ug) ** 2)
        E[2] += 0.5 * MU0 * m.tri_dMs[f] * np.sum((n @ ug) * pg * wg)
    e = oc.edges()
    min_dot = min(1.0, np.nanmin(np.sum(u[e[:, 0]] * u[e[:, 1]], axis=1)))
    if region != -1:
        vol = regvol[region]
    return E, num / vol, np.arccos(min_dot)


@pytest.mark.parametrize("npi", [5, 1])
def test_fem_energy_avg_max_angle_vs_numpy(oracle, npi):
    case = cases.small_cuboid(npi=npi)
    oc = cases.oracle_ctx(case)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    E_np, avg_np, ang_np = numpy_observables(case, oc, Hext=case.Hext)
    E = oc.energy(case.Hext)
    assert cases.rel_max(E, E_np) < 1e-12
    assert np.all(E != 0.0)                         # every term is exercised by this case
    assert cases.rel_max(oc.avg(0), avg_np[0]) < 1e-12
    assert cases.rel_max(oc.avg(1), avg_np[1]) < 1e-12
    assert abs(oc.max_angle() - ang_np) < 1e-13
    # per-region average (save.cpp:78-98) divides by the region volume of mesh.h:81-90
    for region in (1, 2):
        avg_r = numpy_observables(case, oc, Hext=case.Hext, region=region)[1]
        assert cases.rel_max(oc.avg(0, region), avg_r[0]) < 1e-12
        assert cases.rel_max(oc.avg(1, region), avg_r[1]) < 1e-12
    oc.close()


def test_zeeman_energy_space_field(oracle):
    """Tet::zeemanEnergy with mesh.extSpaceField (src/tetra.cpp:382-391; ut_energy.cpp:170-193 pins
    the column-wise dot product it is built on)."""
    case = cases.small_cuboid()
    oc = cases.oracle_ctx(case)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    rng = np.random.default_rng(7)
    field = rng.standard_normal((case.mesh.NT, 3, 5)) * 1e4
    oc.set_ext_space_field(field)
    amp = 0.37
    E_np = numpy_observables(case, oc, space=field, amp=amp)[0]
    E = oc.energy_space(amp)
    assert cases.rel_max(E, E_np) < 1e-12
    # a space field that is uniform reproduces the uniform-field energy
    uni = np.broadcast_to(case.Hext[None, :, None], field.shape).copy()
    oc.set_ext_space_field(uni)
    assert cases.rel_max(oc.energy_space(1.0), oc.energy(case.Hext)) < 1e-14
    oc.close()
