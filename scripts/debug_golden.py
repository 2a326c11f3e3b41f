import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import cases
from test_full_test_golden import settings_and_timing
from feellgood_b200.fem import Fem
from feellgood_b200.linear_algebra import c_srand

npi = int(sys.argv[1]) if len(sys.argv) > 1 else 1
KIND = sys.argv[2] if len(sys.argv) > 2 else "node3"
WATCH = int(sys.argv[3]) if len(sys.argv) > 3 else -1
case = cases.ellipsoid(npi=npi)
u = np.zeros((case.mesh.NOD, 3)); u[:, 2] = 1.0

class Spy:
    def __init__(self, la): self.la, self.log = la, []
    def __getattr__(self, k): return getattr(self.la, k)
    def solve(self, t):
        f = self.la.solve(t)
        it = self.la.iter
        if hasattr(self.la, "krylov_history") and (f or len(self.log) == WATCH):
            ks = self.la.krylov_state()
            print("SOLVE", len(self.log), "failed" if f else "ok", ks)
            h = self.la.krylov_history(int(ks["nit"]) + 2)
            np.set_printoptions(linewidth=250, precision=4)
            print("  cols: rho_1, (v,rt), alpha, |s|^2, (t,s), (t,t), omega, |r|^2")
            for k in list(range(3)) + list(range(max(3, len(h) - 12), len(h))):
                print("  it", k, h[k])
        self.log.append((t.get_dt(), f, it.get("status"), it.get("nit"), it.get("res"), it.get("rhsn", it.get("rhsnorm")), self.la.get_v_max()))
        return f

s, t_prm = settings_and_timing(case)
oc = cases.oracle_ctx(case); oc.set_state(u); c_srand(2)
so = Spy(cases.OracleLinAlgebra(oc))
fo_ = Fem(s, so, demag=lambda la: la.oc.demag_direct(True)); print("oracle", fo_.time_integration(t_prm))
s, t_prm = settings_and_timing(case)
la = cases.gpu_linalg(case); la.set_state(u); c_srand(2); la.set_operator(KIND); la.krylov_history(0)
sg = Spy(la)
fg_ = Fem(s, sg, demag=lambda l: l.demag_direct(True)); print("gpu", fg_.time_integration(t_prm))
print(len(so.log), len(sg.log))
for k, (a, b) in enumerate(zip(so.log, sg.log)):
    flag = "" if (a[1] == b[1] and a[3] == b[3]) else "   <<<<"
    if (a[1] != b[1]) or k < 2:
        print(k, a, b, flag)
print(fo_.evol[-1]); print(fg_.evol[-1])
