"""The reference's own end-to-end known-answer test, ci-tests/full_test.py:20-50: ellipsoid.msh,
K = 3e5 J/m3 along y, M0 = z, Bext = (1, 0, -1) T, dt in [5e-14, 1e-12], max(du) = 0.1, 20 ps,
`--seed 2`.  Expected final <m> and E_tot are the numbers stored in that script (for NPI = 5 and for
ONE_GAUSS_POINT), with its own acceptance threshold 1e-5.

This pins the COMPOSITION that no reference unit test covers — Tet::integrales end to end, the
scatter, the masked BiCGStab solve, the node update, the time-step controller, the energies, the
charges and the seeded basis-angle stream — against output of the reference program itself.  The
demag potential comes from the all-pairs sum that ScalFMM (third party, absent) approximates.
"""
import numpy as np
import pytest

import cases
from cases import MU0
from feellgood_b200 import Settings, capi, timing
from feellgood_b200.fem import Fem
from feellgood_b200.linear_algebra import c_srand

EXPECTED = {5: dict(m=[0.307542, 0.475877, -0.823989], E_tot=-4.445605e-19),     # full_test.py:49
            1: dict(m=[0.307432, 0.476202, -0.823843], E_tot=-4.446980e-19)}     # full_test.py:47
THRESHOLD = 1e-5                                                                  # full_test.py:70


def settings_and_timing(case):
    B = np.array([1.0, 0.0, -1.0]) / MU0
    s = Settings([capi.tet_prm(**r) for r in case.tet_regions],
                 [capi.tri_prm(**r) for r in case.tri_regions], TOL=1e-6, MAXITER=700,
                 npi_tet=case.npi, npi_tri=case.npi_tri, time_step=1e-12, DUMAX=0.1,
                 evol_columns=["t", "<Mx>", "<My>", "<Mz>", "E_ex", "E_demag", "E_zeeman", "E_tot"],
                 field=lambda t: B)
    return s, timing(2e-11, 5e-14, 1e-12)


def check_last_row(row, npi):
    exp = EXPECTED[npi]
    m_error = np.linalg.norm(np.array(row[1:4]) - exp["m"])
    e_error = abs((row[7] - exp["E_tot"]) / exp["E_tot"])
    assert row[0] == 2e-11
    assert m_error < THRESHOLD and e_error < THRESHOLD, (m_error, e_error)


@pytest.mark.parametrize("npi", [5, 1])
def test_full_test_golden_oracle(oracle, npi):
    case = cases.ellipsoid(npi=npi)
    s, t_prm = settings_and_timing(case)
    oc = cases.oracle_ctx(case)
    u = np.zeros((case.mesh.NOD, 3))
    u[:, 2] = 1.0
    oc.set_state(u)
    c_srand(2)
    fem = Fem(s, cases.OracleLinAlgebra(oc), demag=lambda la: la.oc.demag_direct(True))
    status, nt = fem.time_integration(t_prm)
    assert status == 0 and nt > 20
    check_last_row(fem.evol[-1], npi)
    oc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("npi", [5, 1])
def test_full_test_golden_gpu(gpu_lib, npi):
    """Same run entirely on the GPU: hot path, energies, averages, charges and the all-pairs demag."""
    case = cases.ellipsoid(npi=npi)
    s, t_prm = settings_and_timing(case)
    la = cases.gpu_linalg(case)
    u = np.zeros((case.mesh.NOD, 3))
    u[:, 2] = 1.0
    la.set_state(u)
    c_srand(2)
    fem = Fem(s, la, demag=lambda la: la.demag_direct(True))
    status, nt = fem.time_integration(t_prm)
    assert status == 0 and nt > 20
    check_last_row(fem.evol[-1], npi)
    la.close()
