"""How sensitive is the reference's full_test golden to the basis-angle seed and to the rounding of
the Krylov solver?  Runs the 20 ps ellipsoid scenario for several seeds on the oracle and on the GPU
with both forms of the operator; prints the distance to the golden and the number of failed solves."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import cases
from test_full_test_golden import settings_and_timing, EXPECTED
from feellgood_b200.fem import Fem
from feellgood_b200.linear_algebra import c_srand

def run(kind, npi, seed):
    case = cases.ellipsoid(npi=npi)
    u = np.zeros((case.mesh.NOD, 3)); u[:, 2] = 1.0
    s, t_prm = settings_and_timing(case)
    if kind == "oracle":
        oc = cases.oracle_ctx(case); oc.set_state(u)
        la = cases.OracleLinAlgebra(oc); dem = lambda l: l.oc.demag_direct(True)
    else:
        la = cases.gpu_linalg(case); la.set_state(u); la.set_operator(kind)
        dem = lambda l: l.demag_direct(True)
    c_srand(seed)
    fem = Fem(s, la, demag=dem)
    status, nt = fem.time_integration(t_prm)
    row = fem.evol[-1]
    exp = EXPECTED[npi]
    return (np.linalg.norm(np.array(row[1:4]) - exp["m"]), abs(row[7] / exp["E_tot"] - 1), nt,
            fem.stats.bad_dt.count(), fem.stats.good_dt.count())

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "stats":
    for kind in ("blocks", "node3"):
        for npi in (5, 1):
            bad_runs, total_bad, total_good = 0, 0, 0
            for seed in range(100, 100 + int(sys.argv[2])):
                m_err, e_err, nt, bad, good = run(kind, npi, seed)
                bad_runs += bad > 0
                total_bad += bad
                total_good += good
            print("%s npi=%d: %d runs, %d with a failed solve, %d failed / %d good solves"
                  % (kind, npi, int(sys.argv[2]), bad_runs, total_bad, total_good), flush=True)
    sys.exit(0)

for npi in (5, 1):
    for seed in (2, 1, 3, 4, 5, 6, 7, 8):
        out = []
        for kind in ("oracle", "blocks", "node3"):
            m_err, e_err, nt, bad, good = run(kind, npi, seed)
            out.append("%s: dm=%.1e dE=%.1e nt=%d bad=%d" % (kind, m_err, e_err, nt, bad))
        print("npi=%d seed=%d | %s" % (npi, seed, " | ".join(out)), flush=True)
