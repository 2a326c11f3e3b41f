"""Generates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference and oracle/_ref exist; the GPU box only reads the .npz files).

  ellipsoid_mesh.npz   the reference's own example mesh examples/ellipsoid.msh (BASELINE config 1)
                       parsed with feellgood_b200.meshgen.read_msh, nodes sorted like
                       mesh::sortNodes, dMs like mesh::controlTriangles.
  ref_algebra.npz      outputs of the REFERENCE's unmodified src/algebra headers (compiled into
                       oracle/_ref/libfgref_algebra.so) on seeded inputs: SparseMatrix::mult,
                       bicg, bicg_dir (both overloads), cg, cg_dir, timing.  These pin the oracle
                       (tests/test_oracle_vs_reference.py) and, through it, the CUDA path.
  ref_timestepper.npz  outputs of the REFERENCE's own TimeStepper (cut out of
                       src/time_integration.cpp at build time) and LogStats (src/log-stats.h) on
                       seeded call sequences: pins feellgood_b200.fem and the C++ host mirror.
  bad_cuboids.npz      the four defective meshes of the reference's controlTriangles unit test.
  llg_system.npz       the oracle's K, L_rhs, x0, solution and next state for one LLG step on a
                       small two-region cuboid with the reference's own SparseMatrix::add +
                       bicg_dir driving the solve (fgo_use_reference_algebra).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from feellgood_b200 import meshgen  # noqa: E402
from oracle import fg_oracle_py as fo  # noqa: E402
import cases  # noqa: E402


def ellipsoid():
    m = meshgen.read_msh("/root/reference/examples/ellipsoid.msh", ["ellipsoid_volume"],
                         ["ellipsoid_surface"], scale=1e-9)
    meshgen.sort_nodes(m)
    m.tri_dMs = meshgen.compute_dMs(m, [0.0, 795774.7])
    np.savez_compressed(os.path.join(HERE, "ellipsoid_mesh.npz"), node_p=m.node_p,
                        tet_ind=m.tet_ind, tet_reg=m.tet_reg, tri_ind=m.tri_ind, tri_reg=m.tri_reg,
                        tri_dMs=m.tri_dMs)
    print("ellipsoid: NOD=%d NT=%d NF=%d" % (m.NOD, m.NT, m.NF))


def ref_algebra():
    rng = np.random.default_rng(5489)
    out = {}
    # (1) random diagonally dominant unsymmetric matrix, n=400, ~9 nnz/row
    n = 400
    rows = []
    for i in range(n):
        c = set(rng.integers(0, n, size=8).tolist())
        c.add(i)
        rows.append(np.array(sorted(c), dtype=np.int32))
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum([r.size for r in rows])
    col = np.concatenate(rows).astype(np.int32)
    val = rng.uniform(-1, 1, size=col.size)
    for i in range(n):
        k = rowptr[i] + int(np.searchsorted(col[rowptr[i]:rowptr[i + 1]], i))
        val[k] = 6.0 + rng.uniform(0, 1)
    x = rng.uniform(-1, 1, size=n)
    rhs = rng.uniform(-1, 1, size=n)
    ld = np.sort(rng.choice(n, size=25, replace=False)).astype(np.int32)
    xd = np.zeros(n)
    xd[ld] = rng.uniform(-1, 1, size=ld.size)
    A = fo.RefMatrix(rowptr, col, val)
    out.update(u_rowptr=rowptr, u_col=col, u_val=val, u_x=x, u_rhs=rhs, u_ld=ld, u_xd=xd)
    out["u_mult"] = A.mult(x)
    for name, (sol, info) in dict(
            bicg=A.bicg(np.zeros(n), rhs, tol=1e-10, maxiter=200),
            bicg_dir=A.bicg_dir(np.zeros(n), rhs, ld, tol=1e-10, maxiter=200),
            bicg_dir_xd=A.bicg_dir(np.zeros(n), rhs, ld, tol=1e-10, maxiter=200, xd=xd)).items():
        out["u_%s_x" % name] = sol
        out["u_%s_info" % name] = np.array([info["status"], info["nit"], info["res"], info["rhsn"]])
    # (2) SPD: 1-D Laplacian + mass (like ut_algebra.cpp test_cg_dir's problem family), n=501
    n = 501
    rows = [np.array([j for j in (i - 1, i, i + 1) if 0 <= j < n], dtype=np.int32) for i in range(n)]
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum([r.size for r in rows])
    col = np.concatenate(rows).astype(np.int32)
    val = np.where(col == np.repeat(np.arange(n), np.diff(rowptr)), 2.0, -1.0)
    rhs = np.zeros(n)
    ld = np.array([0, n - 1], dtype=np.int32)
    xd = np.zeros(n)
    xd[n - 1] = 1.0
    A = fo.RefMatrix(rowptr, col, val)
    out.update(s_rowptr=rowptr, s_col=col, s_val=val, s_rhs=rhs, s_ld=ld, s_xd=xd)
    sol, info = A.cg_dir(np.zeros(n), rhs, xd, ld, tol=1e-10, maxiter=2000)
    out["s_cg_dir_x"], out["s_cg_dir_info"] = sol, np.array([info["status"], info["nit"], info["res"], info["rhsn"]])
    rhs2 = rng.uniform(-1, 1, size=n)
    val2 = val + np.where(val == 2.0, 0.5, 0.0)
    A2 = fo.RefMatrix(rowptr, col, val2)
    sol, info = A2.cg(np.zeros(n), rhs2, tol=1e-10, maxiter=2000)
    out.update(s_val2=val2, s_rhs2=rhs2, s_cg_x=sol,
               s_cg_info=np.array([info["status"], info["nit"], info["res"], info["rhsn"]]))
    # (3) timing (src/time_integration.h compiled as is)
    tt = [fo.ref_timing(2e-11, 1e-16, 5e-13, dt) for dt in (1e-16, 7.07e-15, 5e-13, 1e-13)]
    out["timing"] = np.array([[t["dt0"], t["prefactor0"], t["prefactor"]] for t in tt])
    np.savez_compressed(os.path.join(HERE, "ref_algebra.npz"), **out)
    print("ref_algebra: bicg nit", out["u_bicg_info"][1], "cg_dir nit", out["s_cg_dir_info"][1])


def llg_system():
    case = cases.small_cuboid()
    oc = cases.oracle_ctx(case)
    oc.use_reference_algebra(True)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    oc.base_projection(case.angle)
    oc.prepare_elements(case.Hext, case.dt, case.prefactor)
    failed = oc.solve(case.dt)
    val, rhs, _ = oc.system()
    info = oc.iter_info()
    u1, v1, _, _ = oc.get_state(1)
    Kp0, Lp0 = oc.element(0)
    np.savez_compressed(os.path.join(HERE, "llg_system.npz"), val=val, rhs=rhs, u1=u1, v1=v1,
                        info=np.array([info["status"], info["nit"], info["res"], info["rhsn"]]),
                        failed=failed, v_max=oc.v_max(), Kp0=Kp0, Lp0=Lp0)
    print("llg_system: failed", failed, info)


def timestepper_script(seed=5489, n=400):
    """Seeded call sequence for TimeStepper: (op, value) with op 0 = operator()(stride),
    1 = set_soft_limit(value); mimics the accept / reject / dumax traffic of time_integration."""
    rng = np.random.default_rng(seed)
    ops = []
    t, target = 0.0, 1e-12
    for _ in range(n):
        r = rng.random()
        if r < 0.6:
            stride = target - t
            ops.append((0, stride))
        elif r < 0.8:
            ops.append((1, 10 ** rng.uniform(-15, -12)))
        else:
            target += 1e-12
            ops.append((0, target - t))
        t += 10 ** rng.uniform(-15, -13)
        if t >= target:
            target = t + 1e-12
    return np.array(ops)


def ref_timestepper():
    """Outputs of the REFERENCE's own TimeStepper and LogStats (oracle/_ref) on seeded scripts."""
    out = {}
    for k, (init, mn, mx) in enumerate([(7.07e-15, 1e-16, 5e-13), (2.2e-13, 5e-14, 1e-12), (1e-13, 1e-15, 1e-13)]):
        ops = timestepper_script(5489 + k)
        ts = fo.RefTimeStepper(init, mn, mx)
        res = []
        for op, val in ops:
            if op == 0:
                res.append(ts(val))
            else:
                ts.set_soft_limit(val)
                res.append(np.nan)
        out["ts%d_prm" % k] = np.array([init, mn, mx])
        out["ts%d_ops" % k] = ops
        out["ts%d_out" % k] = np.array(res)
    rng = np.random.default_rng(5489)
    xs = 10 ** rng.uniform(-16, -12, size=257)
    ls = fo.RefLogStats()
    snap = []
    for x in xs:
        ls.add(x)
        snap.append(ls.get())
    out["ls_x"] = xs
    out["ls_out"] = np.array(snap)
    np.savez_compressed(os.path.join(HERE, "ref_timestepper.npz"), **out)
    print("ref_timestepper: %d arrays" % len(out))


def bad_cuboids():
    """The reference's known-answer meshes for mesh::controlTriangles (unit-tests/meshes/
    bad_cuboid_{1..4}.msh, used by ut_readMesh.cpp:78-133), parsed as they are (no node sort, no
    dMs): 8 nodes and a dozen elements each."""
    out = {}
    for k in range(1, 5):
        m = meshgen.read_msh("/root/reference/unit-tests/meshes/bad_cuboid_%d.msh" % k,
                             ["whole_volume"], ["whole_surface"], scale=1.0)
        for name in ("node_p", "tet_ind", "tet_reg", "tri_ind", "tri_reg"):
            out["m%d_%s" % (k, name)] = getattr(m, name)
    np.savez_compressed(os.path.join(HERE, "bad_cuboids.npz"), **out)
    print("bad_cuboids: %d arrays" % len(out))


if __name__ == "__main__":
    which = sys.argv[1:] or ["ellipsoid", "ref_algebra", "llg_system", "ref_timestepper", "bad_cuboids"]
    for name in which:
        {"ellipsoid": ellipsoid, "ref_algebra": ref_algebra, "llg_system": llg_system,
         "ref_timestepper": ref_timestepper, "bad_cuboids": bad_cuboids}[name]()
