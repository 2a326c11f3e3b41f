"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libfeellgood_b200.so),
against the CPU oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star):
element matrices / vectors and assembled system 1e-12 relative (norm-wise per element / per
system); per-step solution within the solver tolerance; averages after N steps 1e-6 relative.
"""
import numpy as np
import pytest

import cases
from cases import FixedTiming, rel_max

pytestmark = pytest.mark.gpu

TOL_ELEM = 1e-12


def _pair(case):
    oc = cases.oracle_ctx(case)
    la = cases.gpu_linalg(case)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    la.set_state(case.u, case.v, case.phi, case.phiv)
    return oc, la


def _prepare(case, oc, la):
    t = FixedTiming(case)
    oc.base_projection(case.angle)
    la.base_projection(case.angle)
    oc.prepare_elements(case.Hext, case.dt, case.prefactor)
    la.prepareElements(case.Hext, t)
    return t


@pytest.fixture(scope="module", params=["small_cuboid", "small_cuboid_npi1", "small_cuboid_iso",
                                        "small_cuboid_iso_npi1", "ellipsoid"])
def prepared(request, oracle, gpu_lib):
    case = {"small_cuboid": lambda: cases.small_cuboid(),
            "small_cuboid_npi1": lambda: cases.small_cuboid(npi=1),
            "small_cuboid_iso": lambda: cases.small_cuboid(aniso=False),
            "small_cuboid_iso_npi1": lambda: cases.small_cuboid(npi=1, aniso=False),
            "ellipsoid": lambda: cases.ellipsoid()}[request.param]()
    oc, la = _pair(case)
    t = _prepare(case, oc, la)
    yield case, oc, la, t
    la.close()
    oc.close()


def test_sizes_and_pattern(prepared):
    case, oc, la, _ = prepared
    assert (la.NT, la.NF, la.n_magTet, la.n_magTri, la.E, la.E_mag, la.n, la.nnz, la.nlvd) == \
        (oc.NT, oc.NF, oc.n_magTet, oc.n_magTri, oc.E, oc.E_mag, oc.n, oc.nnz, oc.nlvd)
    rp_o, col_o = oc.csr()
    rp_g, col_g = la.csr()
    assert np.array_equal(rp_o, rp_g) and np.array_equal(col_o, col_g)   # bit-exact index work
    # the element fast path is offered exactly when no magnetic region has K or K3 (fg_get_layout)
    assert la.iso_fast_path == case.name.startswith("small_cuboid_iso")
    assert la.col_bytes in (2, 4) and la.stored_pairs * 4 >= la.nnz


def test_tet_tables(prepared):
    case, oc, la, _ = prepared
    ind, da, w = la.tet_tables()
    assert np.array_equal(ind, oc.tet_ind())                             # Tet::orientate
    da_o, w_o = oc.tet_geom()
    assert rel_max(da, da_o) < 1e-15 and rel_max(w, w_o) < 1e-15


def test_basis(prepared):
    case, oc, la, _ = prepared
    ep, eq = la.basis()
    ep_o, eq_o = oc.get_basis()
    assert np.max(np.abs(ep - ep_o)) < 5e-15 and np.max(np.abs(eq - eq_o)) < 5e-15
    # node_setBasis properties (unit-tests/ut_node.cpp:58-121): orthonormal to 5e-15
    u = case.u
    assert np.max(np.abs(np.einsum("ij,ij->i", ep, u))) < 5e-15
    assert np.max(np.abs(np.einsum("ij,ij->i", ep, eq))) < 5e-15
    assert np.max(np.abs(np.linalg.norm(eq, axis=1) - 1)) < 5e-15


def test_element_matrices(prepared):
    case, oc, la, _ = prepared
    Kp, Lp = la.elements()
    worstK = worstL = 0.0
    for t in range(oc.NT):
        Ko, Lo = oc.element(t)
        worstK = max(worstK, rel_max(Kp[t], Ko))
        worstL = max(worstL, rel_max(Lp[t], Lo))
    assert worstK < TOL_ELEM, worstK
    assert worstL < TOL_ELEM, worstL


def test_production_records(prepared):
    """The records the PRODUCTION element kernels wrote (k_tet_iso / k_tet_lean / k_tet, not the tap kernel
    of test_element_matrices), against the oracle's element: Lp = Perm P vec(BE) (src/tetra.cpp:306) from the
    stored BE and the node bases, and the state-dependent diagonal of E, E_aa = c (da_a . da_a) + contrib_a
    (src/tetra.cpp:108-131), read off Kp[(eq_a),(eq_a)]."""
    case, oc, la, t = prepared
    from feellgood_b200.linear_algebra import GAMMA0, MU0
    rec = la.records()
    ep, eq = la.basis()
    ind, da, w = la.tet_tables()
    n_mag = 0
    for k in range(oc.NT):
        Ko, Lo = oc.element(k)
        if not np.any(Ko):
            assert not np.any(rec[k])
            continue
        n_mag += 1
        nd = ind[k]
        be = rec[k, :, 1:4]
        Lp = np.concatenate([np.einsum("id,id->i", eq[nd], be), np.einsum("id,id->i", ep[nd], be)])
        assert rel_max(Lp, Lo) < TOL_ELEM, k
        prm = case.tet_regions[case.mesh.tet_reg[k]]
        Abis = 2.0 * prm.get("A", 1e-11) / (MU0 * prm.get("Ms", 795774.7))
        c = t.prefactor * (0.5 * t.get_dt() * GAMMA0) * Abis * w[k].sum()
        Ko = Ko.reshape(8, 8)
        for a in range(4):
            # row 2a of the node pair = the eq-tested equation = row a of Kp (Perm), column 4 + a = unknown vq of a
            E_aa = Ko[a, 4 + a]
            contrib = E_aa - c * np.dot(da[k, a], da[k, a])
            assert abs(rec[k, a, 0] - contrib) <= 1e-10 * abs(E_aa), (k, a)
    assert n_mag > 0


def test_tri_vectors(prepared):
    case, oc, la, _ = prepared
    Lp = la.tri_elements()
    nz = 0
    for f in range(oc.NF):
        Lo = oc.tri_element(f)
        nz += bool(np.any(Lo != 0))
        assert rel_max(Lp[f], Lo) < TOL_ELEM
    if case.name == "small_cuboid":
        assert nz > 0


def test_assembled_system(prepared):
    case, oc, la, t = prepared
    oc.assemble()
    val_o, rhs_o, x0_o = oc.system()
    val, rhs, x0 = la.system(t)
    assert rel_max(val, val_o) < TOL_ELEM
    assert rel_max(rhs, rhs_o) < TOL_ELEM
    assert rel_max(x0, x0_o) < TOL_ELEM
    # masked dofs: identity rows and zero rhs, bit-exact
    _, lvd = oc.masks()
    rp, col = la.csr()
    for i in lvd:
        row = val[rp[i]:rp[i + 1]]
        assert rhs[i] == 0.0 and np.all(row[col[rp[i]:rp[i + 1]] != i] == 0.0)
        assert row[col[rp[i]:rp[i + 1]] == i][0] == 1.0
    # the device SpMV on that K (SparseMatrix::mult)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(la.n)
    y = la.apply_operator(x)
    assert rel_max(y, oracle_spmv(rp, col, val_o, x)) < 1e-13


def oracle_spmv(rp, col, val, x):
    from oracle import fg_oracle_py as fo
    return fo.spmv(rp, col, val, x)


def test_solve_and_update(prepared):
    case, oc, la, t = prepared
    failed_o = oc.solve(case.dt)
    failed_g = la.solve(t)
    io = oc.iter_info()
    assert failed_g == failed_o and la.iter["status"] == io["status"] == 0
    assert abs(la.iter["nit"] - io["nit"]) <= 3
    assert abs(la.iter["rhsn"] - io["rhsn"]) <= 1e-12 * io["rhsn"]
    # per-step solution within the solver tolerance: the GPU solution satisfies the ORACLE's system
    val_o, rhs_o, x_o = oc.system()
    rp, col = oc.csr()
    x_g = la.solution()
    r = rhs_o - oracle_spmv(rp, col, val_o, x_g)
    assert np.linalg.norm(r) <= 1.01 * case.tol * np.linalg.norm(rhs_o)
    # observed ~1e-9 (equal iteration counts); 5*TOL leaves room for a one-iteration difference only
    assert np.linalg.norm(x_g - x_o) <= 5.0 * case.tol * np.linalg.norm(x_o)
    u_o, v_o, _, _ = oc.get_state(1)
    u_g, v_g, phi_g, phiv_g = la.get_state(1)
    # the node update is exact given x (src/node.h:116-122): |du| <= dt*gamma0*|dx| per node, and x
    # itself agrees to the solver tolerance (two converged BiCGStab runs differ by ~cond*TOL)
    from feellgood_b200.linear_algebra import GAMMA0
    dx = np.max(np.hypot(x_g[0::2] - x_o[0::2], x_g[1::2] - x_o[1::2]))
    assert np.max(np.abs(u_g - u_o)) <= 1.01 * case.dt * GAMMA0 * dx + 1e-14
    assert np.max(np.abs(u_g - u_o)) < 1e-6
    assert rel_max(v_g, v_o) < 2e-5
    assert abs(la.get_v_max() - oc.v_max()) <= 2e-5 * oc.v_max()
    assert np.array_equal(phi_g, case.phi) and np.array_equal(phiv_g, case.phiv)
    mag, _ = oc.masks()
    assert np.max(np.abs(np.linalg.norm(u_g[mag], axis=1) - 1)) < 1e-15
    assert np.array_equal(u_g[~mag], case.u[~mag])            # non-magnetic nodes untouched


def test_golden_llg_system(oracle, gpu_lib):
    """Committed fixture produced with the REFERENCE's own SparseMatrix::add + bicg_dir."""
    z = np.load(cases.GOLDEN + "/llg_system.npz")
    case = cases.small_cuboid()
    la = cases.gpu_linalg(case)
    la.set_state(case.u, case.v, case.phi, case.phiv)
    t = FixedTiming(case)
    la.base_projection(case.angle)
    la.prepareElements(case.Hext, t)
    Kp, Lp = la.elements(0, 1)
    assert rel_max(Kp[0], z["Kp0"]) < TOL_ELEM and rel_max(Lp[0], z["Lp0"]) < TOL_ELEM
    val, rhs, _ = la.system(t)
    assert rel_max(val, z["val"]) < TOL_ELEM and rel_max(rhs, z["rhs"]) < TOL_ELEM
    assert la.solve(t) == bool(z["failed"])
    assert abs(la.iter["nit"] - int(z["info"][1])) <= 3
    u1, v1, _, _ = la.get_state(1)
    assert np.max(np.abs(u1 - z["u1"])) < 1e-7 and rel_max(v1, z["v1"]) < 1e-4
    assert abs(la.get_v_max() - float(z["v_max"])) <= 1e-4 * float(z["v_max"])
    la.close()


def test_space_field_and_drift(oracle, gpu_lib):
    """prepareElements(A_Hext) overload (linear_algebra.cpp:54-81) and the recentring drift term
    (tetra.cpp:150-169, 279-288)."""
    case = cases.small_cuboid(seed=7)
    oc, la = _pair(case)
    rng = np.random.default_rng(3)
    fieldv = rng.standard_normal((case.mesh.NT, 3, case.npi)) * 2e4
    oc.set_ext_space_field(fieldv)
    la.set_ext_space_field(fieldv)
    t = FixedTiming(case)
    oc.base_projection(case.angle)
    la.base_projection(case.angle)
    la.idx_dir, la.DW_vz = 2, 37.5
    oc.prepare_elements_space(1.7, case.dt, case.prefactor, idx_dir=2, Vdrift=37.5)
    la.prepareElements(1.7, t)
    Kp, Lp = la.elements()
    for tt in range(oc.NT):
        Ko, Lo = oc.element(tt)
        assert rel_max(Kp[tt], Ko) < TOL_ELEM and rel_max(Lp[tt], Lo) < TOL_ELEM
    la.close()
    oc.close()


def test_trajectory_averages(oracle, gpu_lib):
    """N accepted steps with NEXT->CURRENT commits: <m> and v_max agree to 1e-6 relative."""
    case = cases.ellipsoid()
    oc, la = _pair(case)
    t = FixedTiming(case)
    from feellgood_b200.linear_algebra import M_2_PI, mt19937_uniform01
    for step in range(12):
        ang = M_2_PI * mt19937_uniform01(1000 + step)
        oc.base_projection(ang)
        oc.prepare_elements(case.Hext, case.dt, case.prefactor)
        fo_ = oc.solve(case.dt)
        fg_ = la.step(case.Hext, t, angle=ang)
        assert fg_ == fo_ is False
        oc.evolution()
        la.evolution()
    u_o = oc.get_state(0)[0]
    u_g = la.get_state(0, "u")[0]
    assert rel_max(u_g.mean(axis=0), u_o.mean(axis=0)) < 1e-6
    assert np.max(np.abs(u_g - u_o)) < 1e-6
    assert abs(la.get_v_max() - oc.v_max()) <= 1e-6 * oc.v_max() * 100
    la.close()
    oc.close()


@pytest.mark.parametrize("name", ["small_cuboid", "small_cuboid_npi1", "ellipsoid"])
def test_observables(oracle, gpu_lib, name):
    """Fem::energy, mesh::avg (all regions and per region), mesh::max_angle on the resident state
     against the oracle on the SAME state: sums of ~1e3 element terms in a
    different order, so 1e-12 relative per energy term."""
    case = {"small_cuboid": lambda: cases.small_cuboid(),
            "small_cuboid_npi1": lambda: cases.small_cuboid(npi=1),
            "ellipsoid": lambda: cases.ellipsoid()}[name]()
    oc, la = _pair(case)
    E_o, E_g = oc.energy(case.Hext), la.energy(case.Hext)
    for k in range(4):
        assert abs(E_g[k] - E_o[k]) <= 1e-12 * np.max(np.abs(E_o)), (k, E_g, E_o)
    for what, w in (("u", 0), ("v", 1)):
        assert rel_max(la.avg(what), oc.avg(w)) < 1e-12
        for region in range(1, len(case.tet_regions)):
            a_o, a_g = oc.avg(w, region), la.avg(what, region)
            assert rel_max(a_g, a_o) < 1e-12, (what, region)   # non-magnetic region: 0 / vol = 0
    assert np.all(np.isnan(la.avg("u", 0))) and np.all(np.isnan(oc.avg(0, 0)))  # no tet: 0 / 0
    assert abs(la.max_angle() - oc.max_angle()) < 1e-14
    la.close()
    oc.close()


def test_energy_space_field(oracle, gpu_lib):
    """Zeeman energy with mesh.extSpaceField x amplitude (src/energy.cpp:41-43)."""
    case = cases.small_cuboid()
    oc, la = _pair(case)
    rng = np.random.default_rng(11)
    field = rng.standard_normal((case.mesh.NT, 3, case.npi)) * 1e4
    oc.set_ext_space_field(field)
    la.set_ext_space_field(field)
    E_o, E_g = oc.energy_space(0.37), la.energy(0.37)
    assert np.max(np.abs(E_g - E_o)) <= 1e-12 * np.max(np.abs(E_o))
    la.close()
    oc.close()


@pytest.mark.parametrize("name", ["small_cuboid", "small_cuboid_npi1", "ellipsoid"])
def test_charges_and_direct_demag(oracle, gpu_lib, name):
    """fmm::calc_charges (Tet::charges, Tri::charges, Tri::correctionCharges / potential) and the
    all-pairs potential that stands in for ScalFMM against the oracle."""
    case = {"small_cuboid": lambda: cases.small_cuboid(),
            "small_cuboid_npi1": lambda: cases.small_cuboid(npi=1),
            "ellipsoid": lambda: cases.ellipsoid()}[name]()
    oc, la = _pair(case)
    for which in (0, 1):
        src_o, corr_o = oc.calc_charges(which)
        src_g, corr_g = la.calc_charges(which)
        assert src_g.shape == src_o.shape
        assert rel_max(src_g, src_o) < 1e-13
        assert rel_max(corr_g, corr_o) < 1e-10      # cancellation between 1/r terms and the analytic potential
    oc.demag_direct(True)
    la.demag_direct(True)
    _, _, phi_o, phiv_o = oc.get_state(1)
    _, _, phi_g, phiv_g = la.get_state(1)
    assert rel_max(phi_g, phi_o) < 1e-12 and rel_max(phiv_g, phiv_o) < 1e-12
    mag = oc.masks()[0]
    assert np.array_equal(phi_g[~mag], case.phi[~mag])                    # only magnetic nodes are targets
    la.demag_direct(False)                                                 # FIRST_ORDER: phiv untouched
    assert np.array_equal(la.get_state(1)[3], phiv_g)
    la.close()
    oc.close()


def test_failure_semantics(oracle, gpu_lib):
    """solve() returns True (failure) on ITER_OVERFLOW and leaves NEXT and v_max untouched
    (src/solver.cpp:62-69)."""
    case = cases.small_cuboid()
    case.maxiter = 2
    oc, la = _pair(case)
    t = _prepare(case, oc, la)
    fo_ = oc.solve(case.dt)
    fg_ = la.solve(t)
    assert fo_ and fg_
    assert la.iter["status"] == oc.iter_info()["status"] == 1
    assert la.iter["nit"] == oc.iter_info()["nit"] == 2
    u_g, v_g, _, _ = la.get_state(1)
    assert np.array_equal(u_g, case.u) and np.array_equal(v_g, case.v)
    assert la.get_v_max() == 0.0
    la.close()
    oc.close()


def test_error_paths(gpu_lib):
    """C-ABI error behaviour: bad meshes and call-sequence errors return codes, never exit()."""
    import ctypes as C
    from feellgood_b200 import capi, meshgen, LinAlgebra, Settings
    case = cases.small_cuboid()
    s = Settings([capi.tet_prm(**r) for r in case.tet_regions],
                 [capi.tri_prm(**r) for r in case.tri_regions])
    bad = meshgen.Mesh(case.mesh.node_p, case.mesh.tet_ind.copy(), case.mesh.tet_reg,
                       case.mesh.tri_ind, case.mesh.tri_reg, case.mesh.tri_dMs)
    bad.tet_ind[3] = [0, 0, 1, 2]                           # degenerate (Tet::orientate exits in the ref)
    with pytest.raises(capi.FgError) as e:
        LinAlgebra(s, bad)
    assert e.value.code == -3
    bad.tet_ind[3] = [0, 1, 2, case.mesh.NOD + 5]
    with pytest.raises(capi.FgError):
        LinAlgebra(s, bad)
    la = LinAlgebra(s, case.mesh)
    la.set_state(case.u)
    with pytest.raises(capi.FgError) as e:
        la.solve(FixedTiming(case))                         # solve before prepareElements
    assert e.value.code == -4
    # |u| = 1 on the magnetic nodes is a documented precondition of fg_set_state: violations are rejected
    import numpy as np
    bad_u = case.u.copy()
    mag = np.unique(case.mesh.tet_ind[np.array([case.tet_regions[r].get("Ms", 795774.7) > 0 for r in case.mesh.tet_reg])])
    bad_u[mag[3]] *= 1.01
    with pytest.raises(capi.FgError) as e:
        la.set_state(bad_u)
    assert e.value.code == -1 and "unit vector" in str(e.value)
    bad_u[mag[3]] = np.nan
    with pytest.raises(capi.FgError):
        la.set_state(bad_u)
    la.set_state(case.u)                                    # and the context is still usable
    la.close()
