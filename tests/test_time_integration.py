"""Fem::time_integration (src/time_integration.cpp:133-245) restated in feellgood_b200.fem: the
CPU part runs the loop on the oracle backend (control flow, target times, statistics); the GPU
part runs the same loop on the CUDA LinAlgebra and compares the whole .evol trajectory."""
import numpy as np
import pytest

import cases
from cases import MU0
from feellgood_b200 import Settings, capi, timing
from feellgood_b200.fem import Fem
from feellgood_b200.linear_algebra import c_srand


def full_test_settings(case, tf=4e-12):
    """ci-tests/full_test.py:20-42 (ellipsoid, K = 3e5 along y, Bext = (1, 0, -1) T) with a
    shorter final time; the demag solver is replaced by a local surrogate on both sides."""
    B = np.array([1.0, 0.0, -1.0]) / MU0
    s = Settings([capi.tet_prm(**r) for r in case.tet_regions],
                 [capi.tri_prm(**r) for r in case.tri_regions], TOL=case.tol, MAXITER=case.maxiter,
                 npi_tet=case.npi, npi_tri=case.npi_tri, time_step=1e-12, DUMAX=0.1,
                 evol_columns=["iter", "t", "dt", "max_dm", "max_angle", "<Mx>", "<My>", "<Mz>",
                               "<dMx/dt>", "E_ex", "E_aniso", "E_demag", "E_zeeman", "E_tot", "Hx"],
                 field=lambda t: B)
    return s, timing(tf, 5e-14, 1e-12)


def initial_state(case):
    u = np.zeros((case.mesh.NOD, 3))
    u[:, 2] = 1.0                                    # initial_magnetization: [0, 0, 1]
    return u


def run_oracle(case, tf=4e-12):
    s, t_prm = full_test_settings(case, tf)
    oc = cases.oracle_ctx(case)
    oc.set_state(initial_state(case))
    c_srand(2)                                       # feellgood --seed 2
    fem = Fem(s, cases.OracleLinAlgebra(oc), demag=cases.local_demag_surrogate(case.mesh))
    status, nt = fem.time_integration(t_prm)
    return fem, t_prm, status, nt, oc


def test_loop_on_oracle_backend(oracle):
    case = cases.ellipsoid()
    fem, t_prm, status, nt, oc = run_oracle(case)
    assert status == 0 and nt >= 4
    rows = np.array(fem.evol, dtype=float)
    assert rows.shape == (5, 15)
    # visible steps land exactly on the targets (src/time_integration.cpp:226-230)
    assert np.array_equal(rows[:, 1], [0.0, 1e-12, 2e-12, 3e-12, 4e-12])
    assert t_prm.get_t() == 4e-12
    # accepted steps respect DUMAX; the statistics saw every attempt
    assert np.all(rows[1:, 3] <= 0.1)
    st = fem.stats
    assert st.good_dt.count() >= nt and st.good_dt.count() + st.bad_dt.count() >= nt
    # relaxation in a fixed field with damping: the total energy decreases
    assert np.all(np.diff(rows[:, 13]) < 0)
    assert np.allclose(rows[:, 13], rows[:, 9:13].sum(axis=1), rtol=1e-14, atol=0)
    assert np.allclose(np.linalg.norm(oc.get_state(1)[0], axis=1), 1.0, atol=1e-14)
    oc.close()


def test_step_count_rounds_half_away_from_zero(oracle):
    """step_count = std::round((tf - t0) / time_step) (src/time_integration.cpp:170): tf = 2.5 time steps
    gives 3 visible steps after the initial row (the reference and the C++ shim), not Python's round(2.5) = 2."""
    from feellgood_b200.fem import round_half_away
    assert [round_half_away(x) for x in (0.5, 1.5, 2.5, 3.5, 2.4999, 2.0, 0.0)] == [1, 2, 3, 4, 2, 2, 0]
    case = cases.ellipsoid()
    fem, t_prm, status, nt, oc = run_oracle(case, tf=2.5e-12)
    rows = np.array(fem.evol, dtype=float)
    assert status == 0 and rows.shape[0] == 4          # t = 0, 1, 2, 3 ps
    assert np.array_equal(rows[:, 1], [0.0, 1e-12, 2e-12, 3e-12])
    oc.close()


def test_dt_too_small_aborts(oracle):
    """**ABORTED**: dt < DTMIN returns status 1 (src/time_integration.cpp:185-190)."""
    case = cases.ellipsoid()
    case.maxiter = 1                                 # solves fail until dt is tiny -> dt < DTMIN
    s, t_prm = full_test_settings(case)
    oc = cases.oracle_ctx(case)
    oc.set_state(initial_state(case))
    c_srand(2)
    fem = Fem(s, cases.OracleLinAlgebra(oc))
    status, nt = fem.time_integration(t_prm)
    assert status == 1 and nt < 20
    assert fem.stats.bad_dt.count() >= 2
    assert t_prm.is_dt_TooSmall() and t_prm.get_t() < t_prm.tf
    oc.close()


@pytest.mark.gpu
def test_loop_gpu_matches_oracle(oracle, gpu_lib):
    """Same loop, same seed, GPU backend: identical accept/reject trajectory, .evol columns (averages
    and energies) within 1e-6 relative — the north-star criterion after a fixed number of steps."""
    case = cases.ellipsoid()
    fem_o, t_o, status_o, nt_o, oc = run_oracle(case)
    s, t_prm = full_test_settings(case)
    la = cases.gpu_linalg(case)
    la.set_state(initial_state(case))
    c_srand(2)
    fem_g = Fem(s, la, demag=cases.local_demag_surrogate(case.mesh))
    status_g, nt_g = fem_g.time_integration(t_prm)
    assert (status_g, nt_g) == (status_o, nt_o)
    ro, rg = np.array(fem_o.evol, dtype=float), np.array(fem_g.evol, dtype=float)
    assert ro.shape == rg.shape
    assert np.array_equal(ro[:, 0], rg[:, 0]) and np.array_equal(ro[:, 1], rg[:, 1])
    scale = np.max(np.abs(ro), axis=0)
    scale[scale == 0] = 1.0
    scale[9:14] = np.max(scale[9:14])      # energies: relative to the largest term (E_ex ~ 1e-32 J
                                           # of a uniform state is rounding noise on both sides)
    scale[5:8] = 1.0                       # <M> components: absolute on the unit sphere
    scale[4] = 1.0                         # max_angle in radians: acos near 1 amplifies 1e-13 -> 1e-7
    err = np.max(np.abs(rg - ro) / scale, axis=0)
    # state-like columns (t, dt, <M>, energies): 1e-6; velocity-like columns (max_dm, <dMx/dt>) are
    # linear in the solution of one solve, which is only determined to the solver tolerance
    tol = np.full(err.size, 1e-6)
    tol[[2, 3, 8]] = 1e-4                 # dt follows DUMAX / vmax (time_integration.cpp:213)
    assert np.all(err < tol), dict(zip(s.evol_columns, err))
    assert fem_g.stats.good_dt.count() == fem_o.stats.good_dt.count()
    assert fem_g.stats.bad_dt.count() == fem_o.stats.bad_dt.count()
    assert abs(fem_g.stats.max_angle - fem_o.stats.max_angle) < 1e-6
    la.close()
    oc.close()
