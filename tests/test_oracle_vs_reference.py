"""Pins the oracle's sparse algebra and timing against outputs of the REFERENCE's own unmodified
src/algebra headers and src/time_integration.h (compiled into oracle/_ref, run in the build
container, committed as tests/golden/ref_algebra.npz by tests/golden/make_golden.py), and the
oracle's whole LLG step against the fixture produced with the reference's SparseMatrix::add +
bicg_dir driving the solve (tests/golden/llg_system.npz)."""
import numpy as np
import pytest

import cases
from cases import rel_max


@pytest.fixture(scope="module")
def gold():
    return np.load(cases.GOLDEN + "/ref_algebra.npz")


def _info(d):
    return np.array([d["status"], d["nit"], d["res"], d["rhsn"]])


def test_spmv_bit_exact(oracle, gold):
    y = oracle.spmv(gold["u_rowptr"], gold["u_col"], gold["u_val"], gold["u_x"])
    assert np.array_equal(y, gold["u_mult"])          # same left-to-right row fold, no FMA


def test_bicg_variants(oracle, gold):
    n = gold["u_x"].size
    args = (gold["u_rowptr"], gold["u_col"], gold["u_val"], np.zeros(n), gold["u_rhs"])
    x, info = oracle.bicg(*args, tol=1e-10, maxiter=200)
    assert np.array_equal(x, gold["u_bicg_x"]) and np.array_equal(_info(info), gold["u_bicg_info"])
    x, info = oracle.bicg_dir(*args, gold["u_ld"], tol=1e-10, maxiter=200)
    assert np.array_equal(x, gold["u_bicg_dir_x"])
    assert np.array_equal(_info(info), gold["u_bicg_dir_info"])
    x, info = oracle.bicg_dir(*args, gold["u_ld"], tol=1e-10, maxiter=200, xd=gold["u_xd"])
    assert np.array_equal(x, gold["u_bicg_dir_xd_x"])
    assert np.array_equal(_info(info), gold["u_bicg_dir_xd_info"])


def test_cg_variants(oracle, gold):
    n = gold["s_rhs"].size
    x, info = oracle.cg_dir(gold["s_rowptr"], gold["s_col"], gold["s_val"], np.zeros(n),
                            gold["s_rhs"], gold["s_xd"], gold["s_ld"], tol=1e-10, maxiter=2000)
    assert np.array_equal(x, gold["s_cg_dir_x"]) and np.array_equal(_info(info), gold["s_cg_dir_info"])
    # ut_algebra.cpp:192-243 property: the Dirichlet 1-D Laplacian solution is the linear ramp
    assert np.max(np.abs(x - np.linspace(0, 1, n))) < 1e-7
    x, info = oracle.cg(gold["s_rowptr"], gold["s_col"], gold["s_val2"], np.zeros(n),
                        gold["s_rhs2"], tol=1e-10, maxiter=2000)
    assert np.array_equal(x, gold["s_cg_x"]) and np.array_equal(_info(info), gold["s_cg_info"])


def test_timing(oracle, gold):
    for row, dt in zip(gold["timing"], (1e-16, 7.07e-15, 5e-13, 1e-13)):
        assert oracle.lib().fgo_timing_dt0(1e-16, 5e-13) == row[0]
        assert oracle.timing_prefactor(dt, 5e-13) == row[2]
    # the python mirror of `timing` used by the product host code
    from feellgood_b200 import timing
    t = timing(2e-11, 1e-16, 5e-13)
    assert t.get_dt() == gold["timing"][0, 0] and t.prefactor == gold["timing"][0, 1]
    for row, dt in zip(gold["timing"], (1e-16, 7.07e-15, 5e-13, 1e-13)):
        t.set_dt(dt)
        assert t.prefactor == row[2]


def test_llg_step_fixture(oracle):
    """Oracle with its OWN algebra reproduces the fixture generated with the reference's."""
    z = np.load(cases.GOLDEN + "/llg_system.npz")
    case = cases.small_cuboid()
    oc = cases.oracle_ctx(case)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    oc.base_projection(case.angle)
    oc.prepare_elements(case.Hext, case.dt, case.prefactor)
    Kp0, Lp0 = oc.element(0)
    assert np.array_equal(Kp0, z["Kp0"]) and np.array_equal(Lp0, z["Lp0"])
    failed = oc.solve(case.dt)
    val, rhs, _ = oc.system()
    # the matrix is summed by `omp atomic` / row mutexes in element order that may differ:
    # last-bit differences allowed
    assert rel_max(val, z["val"]) < 1e-14 and rel_max(rhs, z["rhs"]) < 1e-14
    info = oc.iter_info()
    assert failed == bool(z["failed"]) and info["status"] == int(z["info"][0])
    assert abs(info["nit"] - int(z["info"][1])) <= 1
    u1, v1, _, _ = oc.get_state(1)
    assert np.max(np.abs(u1 - z["u1"])) < 1e-9
    assert abs(oc.v_max() - float(z["v_max"])) < 1e-5 * float(z["v_max"])
    oc.close()


def test_live_reference_algebra_if_built(oracle):
    """In the build container oracle/_ref exists: drive the oracle's LLG solve with the
    reference's own SparseMatrix + bicg_dir and compare with the restated Krylov loop."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built here (no /root/reference)")
    case = cases.ellipsoid()
    res = []
    for use_ref in (False, True):
        oc = cases.oracle_ctx(case)
        oc.use_reference_algebra(use_ref)
        oc.set_state(case.u, case.v, case.phi, case.phiv)
        oc.base_projection(case.angle)
        oc.prepare_elements(case.Hext, case.dt, case.prefactor)
        failed = oc.solve(case.dt)
        res.append((failed, oc.iter_info(), oc.get_state(1)[0], oc.system()[2]))
        oc.close()
    assert res[0][0] == res[1][0] is False
    # K is summed in a different element order by the two paths (row mutexes vs omp atomic), so
    # the Krylov trajectories agree to rounding amplified by ~50 iterations: solver tolerance
    assert abs(res[0][1]["nit"] - res[1][1]["nit"]) <= 2
    assert np.max(np.abs(res[0][2] - res[1][2])) < 1e-6
    assert rel_max(res[0][3], res[1][3]) < 1e-4
