"""Host-side runs of the __host__ __device__ math of the CUDA library (no GPU needed: nvcc compiles
the host path of the very functions the kernels call):

  * the unit-quaternion encoding of the tangent-plane basis read by the Krylov kernels;
  * the element math of Tet::integrales — the general core (tet_core) and the fast path
    (tet_iso_front / tet_iso_be, checked bit for bit against the general core by the C++ program) —
    against the independent dense numpy restatement (tests/np_restatement.py), 1e-12 (north star).
"""
import json
import os
import subprocess

import numpy as np

import cases
import np_restatement as npr

CSRC = os.path.join(cases.ROOT, "feellgood_b200", "csrc")


def _nvcc(tmp_path, name, *extra):
    exe = str(tmp_path / name)
    src = os.path.join(cases.ROOT, "tests", "cpp", name + ".cu")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fopenmp", "-o", exe, src, *extra])
    return exe


def test_quaternion_basis_round_trip(tmp_path):
    exe = _nvcc(tmp_path, "device_math_test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DEVICE_MATH_OK" in r.stdout, r.stdout + r.stderr


def test_element_math_against_dense_restatement(tmp_path):
    exe = _nvcc(tmp_path, "element_math_test", os.path.join(CSRC, "fg_setup.cpp"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ELEMENT_MATH_OK" in r.stderr, r.stderr[-2000:]   # fast path == general core
    allrows = [json.loads(ln) for ln in r.stdout.splitlines() if ln.strip()]
    rows = [c for c in allrows if "tri" not in c]
    tris = [c for c in allrows if "tri" in c]
    assert len(rows) == 36 and {c["npi"] for c in rows} == {1, 5}
    assert len(tris) == 12 and {c["npi"] for c in tris} == {1, 4}
    for c in tris:   # Tri::integrales: with ep = eq = axis d, Lp[i] = BE[d][i]
        a, pds = npr.tri_tables(c["npi"])
        weight = 2.0 * c["surf"] * pds          # triangle.h:121-122
        u = np.array(c["u"]).reshape(3, 3)
        BE_dev = np.array(c["BE"]).reshape(3, 3)
        for d in range(3):
            e = np.zeros((3, 3))
            e[:, d] = 1.0
            Lp = npr.tri_integrales(c["Ks"], c["uk"], c["dMs"], weight, u, e, e)
            assert np.max(np.abs(Lp[:3] - BE_dev[d])) <= 1e-13 * np.max(np.abs(BE_dev))
    assert any(c["drift"] for c in rows) and any(c["K"] == 0 for c in rows) and any(c["K"] != 0 for c in rows)
    worst_be = worst_c = 0.0
    for c in rows:
        npi = c["npi"]
        a, pds = npr.tet_tables(npi)
        da, u, v = (np.array(c[k]).reshape(4, 3) for k in ("da", "u", "v"))
        phi, phiv = np.array(c["phi"]), np.array(c["phiv"])
        weight = c["detJ"] * pds
        prm = dict(alpha=c["alpha"], A=c["A"], Ms=c["Ms"], K=c["K"], uk=c["uk"], K3=c["K3"], ex=c["ex"],
                   ey=c["ey"], ez=c["ez"])
        Hext = np.repeat(np.array(c["Hext"])[:, None], npi, axis=1)
        BE_dev = np.array(c["BE"]).reshape(3, 4)
        BE_np = np.empty((3, 4))
        Kd = None
        for d in range(3):   # ep = eq = axis d: Lp[i] = BE[d][i], Kp[i, i] = E_ii (np formulas are linear in P)
            e = np.zeros((4, 3))
            e[:, d] = 1.0
            Kp, Lp = npr.tet_integrales(prm, c["dt"], c["prefactor"], da, weight, u, v, phi, phiv, e, e, Hext,
                                        idx_dir=c["idx_dir"], Vdrift=c["Vdrift"])
            BE_np[d] = Lp[:4]
            assert np.array_equal(Lp[:4], Lp[4:])
            Kd = np.diag(Kp)[:4] if Kd is None else Kd
        worst_be = max(worst_be, np.max(np.abs(BE_dev - BE_np)) / np.max(np.abs(BE_np)))
        # E_ii = prefactor s_dt Abis sum(w) |grad a_i|^2 + contrib_i   (src/tetra.cpp:108-131,261)
        s_dt = npr.THETA * c["dt"] * npr.GAMMA0
        cw = c["prefactor"] * s_dt * (2.0 * c["A"] / (npr.MU0 * c["Ms"])) * weight.sum()
        contrib_np = Kd - cw * np.einsum("ij,ij->i", da, da)
        contrib_dev = np.array(c["contrib"])
        # contrib is recovered as a difference of E_ii and the exchange term: compare at the scale of E_ii
        worst_c = max(worst_c, np.max(np.abs(contrib_dev - contrib_np)) / np.max(np.abs(Kd)))
    assert worst_be < 1e-12, worst_be
    assert worst_c < 1e-12, worst_c


def _emulate(exe, rowptr, col, val, rhs, x0, ld, tol, maxiter, mode="bicg"):
    txt = ["%d %d %d %.17g %d" % (rhs.size, col.size, len(ld), tol, maxiter)]
    for a, fmt in ((rowptr, "%d"), (col, "%d"), (val, "%.17g"), (rhs, "%.17g"), (x0, "%.17g"), (ld, "%d")):
        txt.append(" ".join(fmt % v for v in a))
    r = subprocess.run([exe, mode], input="\n".join(txt) + "\n", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.split("\n")
    st, nit, res, rhsn, failed = lines[0].split()
    return np.array([float(v) for v in lines[1:1 + rhs.size]]), dict(status=int(st), nit=int(nit), res=float(res),
                                                                     rhsn=float(rhsn), failed=bool(int(failed)))


def test_device_bicgstab_state_machine_against_reference(oracle, tmp_path):
    """The library's Krylov scalar code (iteration monitor, alpha/omega/rho, breakdown, overflow and
    mid-iteration exit: fg_krylov_state.cuh) driven on the CPU in the kernels' order, against the
    reference's own bicg_dir: the committed golden of the real reference and the oracle (which
    reproduces the reference bit for bit, tests/test_oracle_vs_reference.py) on special cases."""
    exe = _nvcc(tmp_path, "krylov_state_test")
    gold = np.load(os.path.join(cases.GOLDEN, "ref_algebra.npz"))
    rp, col, val, rhs, ld = (gold[k] for k in ("u_rowptr", "u_col", "u_val", "u_rhs", "u_ld"))
    n = rhs.size
    # (1) the reference's own run: same status, same iteration count, same solution to the tolerance
    x, info = _emulate(exe, rp, col, val, rhs, np.zeros(n), ld, 1e-10, 200)
    g_status, g_nit = int(gold["u_bicg_dir_info"][0]), int(gold["u_bicg_dir_info"][1])
    assert info["status"] == g_status == 0 and abs(info["nit"] - g_nit) <= 1
    assert np.linalg.norm(x - gold["u_bicg_dir_x"]) <= 1e-8 * np.linalg.norm(gold["u_bicg_dir_x"])
    assert np.all(x[ld] == 0.0)                                   # masked dofs never move
    rhs_m = rhs.copy()
    rhs_m[ld] = 0.0
    assert abs(info["rhsn"] - np.linalg.norm(rhs_m)) <= 1e-14 * info["rhsn"]
    # (2) iteration overflow: maxiter full iterations + the half iteration up to the ||s|| test
    for maxiter in (1, 3, 7):
        x, info = _emulate(exe, rp, col, val, rhs, np.zeros(n), ld, 1e-10, maxiter)
        xo, io = oracle.bicg_dir(rp, col, val, np.zeros(n), rhs, ld, tol=1e-10, maxiter=maxiter)
        assert info["status"] == io["status"] == 1 and info["nit"] == io["nit"] == maxiter
        assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
        assert abs(info["res"] - io["res"]) <= 1e-6 * io["res"]
    # (3) loose tolerances: exits on ||s|| in the middle of an iteration (x += alpha phat only) or at the top
    for tol in (1e-1, 1e-2, 1e-3, 1e-5):
        x, info = _emulate(exe, rp, col, val, rhs, np.zeros(n), ld, tol, 200)
        xo, io = oracle.bicg_dir(rp, col, val, np.zeros(n), rhs, ld, tol=tol, maxiter=200)
        assert info["status"] == io["status"] == 0 and info["nit"] == io["nit"], (tol, info, io)
        assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
        assert abs(info["res"] - io["res"]) <= 1e-6 * io["res"]
    # (4) zero right-hand side and zero guess: converged before the first iteration, x untouched
    x, info = _emulate(exe, rp, col, val, np.zeros(n), np.zeros(n), ld, 1e-10, 200)
    xo, io = oracle.bicg_dir(rp, col, val, np.zeros(n), np.zeros(n), ld, tol=1e-10, maxiter=200)
    assert info["status"] == io["status"] == 0 and info["nit"] == io["nit"] == 0 and np.all(x == 0.0)
    # (5) a zero on the diagonal: 1/0 in the Jacobi preconditioner, NaN residual => CANNOT_CONVERGE
    val_bad = val.copy()
    i = int(np.setdiff1d(np.arange(n), ld)[5])
    j = rp[i] + int(np.searchsorted(col[rp[i]:rp[i + 1]], i))
    val_bad[j] = 0.0
    x, info = _emulate(exe, rp, col, val_bad, rhs, np.zeros(n), ld, 1e-10, 200)
    xo, io = oracle.bicg_dir(rp, col, val_bad, np.zeros(n), rhs, ld, tol=1e-10, maxiter=200)
    assert info["status"] == io["status"] == 2, (info, io)
    # (7) LinAlgebra::solve's failure predicate (src/solver.cpp:62-69): overflow and breakdown fail; a
    #     relatively converged solve still FAILS when the absolute residual exceeds TOL (|b| > 1)
    x, info = _emulate(exe, rp, col, val, rhs, np.zeros(n), ld, 1e-10, 3)
    assert info["status"] == 1 and info["failed"]
    big = 1e6 * rhs                                   # |b| ~ 1e7: res <= TOL |b| but res > TOL
    x, info = _emulate(exe, rp, col, val, big, np.zeros(n), ld, 1e-6, 200)
    assert info["status"] == 0 and info["res"] <= 1e-6 * info["rhsn"] and info["res"] > 1e-6 and info["failed"]
    small = rhs / np.linalg.norm(rhs) * 0.5           # |b| < 1: relative convergence implies res <= TOL
    x, info = _emulate(exe, rp, col, val, small, np.zeros(n), ld, 1e-6, 200)
    assert info["status"] == 0 and info["res"] <= 1e-6 and not info["failed"]
    # (6) a non-zero initial guess
    x0 = np.random.default_rng(3).standard_normal(n)
    x0[ld] = 0.0
    x, info = _emulate(exe, rp, col, val, rhs, x0, ld, 1e-10, 200)
    xo, io = oracle.bicg_dir(rp, col, val, x0, rhs, ld, tol=1e-10, maxiter=200)
    assert info["status"] == io["status"] == 0 and abs(info["nit"] - io["nit"]) <= 1
    assert np.linalg.norm(x - xo) <= 1e-8 * np.linalg.norm(xo)


def test_device_cg_state_machine_against_reference(oracle, tmp_path):
    """Same for the Jacobi-preconditioned CG (src/algebra/cg.h:15-58): the reference's golden run on
    the SPD problem of make_golden.py, overflow, and the (q, p) = 0 breakdown."""
    exe = _nvcc(tmp_path, "krylov_state_test")
    gold = np.load(os.path.join(cases.GOLDEN, "ref_algebra.npz"))
    rp, col, val, rhs = (gold[k] for k in ("s_rowptr", "s_col", "s_val2", "s_rhs2"))
    n = rhs.size
    none = np.zeros(0, dtype=np.int32)
    x, info = _emulate(exe, rp, col, val, rhs, np.zeros(n), none, 1e-10, 2000, mode="cg")
    g = gold["s_cg_info"]
    assert info["status"] == int(g[0]) == 0 and abs(info["nit"] - int(g[1])) <= 1
    assert np.linalg.norm(x - gold["s_cg_x"]) <= 1e-8 * np.linalg.norm(gold["s_cg_x"])
    assert abs(info["rhsn"] - g[3]) <= 1e-14 * g[3]
    for maxiter in (1, 4):
        x, info = _emulate(exe, rp, col, val, rhs, np.zeros(n), none, 1e-10, maxiter, mode="cg")
        xo, io = oracle.cg(rp, col, val, np.zeros(n), rhs, tol=1e-10, maxiter=maxiter)
        assert info["status"] == io["status"] == 1 and info["nit"] == io["nit"] == maxiter
        assert np.linalg.norm(x - xo) <= 1e-10 * np.linalg.norm(xo)
    for tol in (1e-1, 1e-3):
        x, info = _emulate(exe, rp, col, val, rhs, np.zeros(n), none, tol, 2000, mode="cg")
        xo, io = oracle.cg(rp, col, val, np.zeros(n), rhs, tol=tol, maxiter=2000)
        assert info["status"] == io["status"] == 0 and info["nit"] == io["nit"]
        assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
    # zero right-hand side and zero guess: finished before the first iteration
    x, info = _emulate(exe, rp, col, val, np.zeros(n), np.zeros(n), none, 1e-10, 50, mode="cg")
    xo, io = oracle.cg(rp, col, val, np.zeros(n), np.zeros(n), tol=1e-10, maxiter=50)
    assert info["status"] == io["status"] and info["nit"] == io["nit"] == 0


def test_slice_ownership_of_the_persistent_kernel(tmp_path):
    """SliceIter / slices_balanced_at (csrc/fg_slice_iter.cuh, the code the kernels call) enumerated on the
    CPU: every slice owned once, fronts of W slices, the last round dealt out per CTA to within one slice."""
    exe = str(tmp_path / "slice_iter_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", CSRC, "-o", exe,
                           os.path.join(cases.ROOT, "tests", "cpp", "slice_iter_test.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SLICE_ITER_OK" in r.stdout, r.stdout + r.stderr[-2000:]
