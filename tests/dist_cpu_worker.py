"""world_size-2 gloo worker for tests/test_dist_partition.py: executes, on the CPU, the exchange
pattern of the multi-GPU solve (halo push of the SpMV input, rank-ordered all-reduce of partial
dots) on the partition computed by feellgood_b200.dist and checks it against the global result.
The oracle supplies the global matrix (test infrastructure)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from feellgood_b200.dist import Partition  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = cases.small_cuboid(nx=10, ny=4, nz=3)
    oc = cases.oracle_ctx(case)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    oc.base_projection(case.angle)
    oc.prepare_elements(case.Hext, case.dt, case.prefactor)
    oc.assemble()
    val, rhs, x0 = oc.system()
    rp, col = oc.csr()
    n = oc.n
    K = np.zeros((n, n))
    for i in range(n):
        K[i, col[rp[i]:rp[i + 1]]] = val[rp[i]:rp[i + 1]]
    rng = np.random.default_rng(42)
    x = rng.standard_normal(n)
    y_ref = K @ x

    for method in ("slab", "rcb"):
        run_partition(case, method, rank, world, x, y_ref, rp, col, val)
    dist.barrier()
    if rank == 0:
        print("DIST_CPU_OK world=%d" % world)
    dist.destroy_process_group()
    oc.close()


def run_partition(case, method, rank, world, x, y_ref, rp, col, val):
    P = Partition(case.mesh, world, method=method)
    lp = P.local(rank)
    nl, no = lp.mesh.NOD, lp.n_owned
    # local vector: owned entries known, ghost tail filled by the neighbours' pushes
    xl = np.zeros((nl, 2))
    xl[:no] = x.reshape(-1, 2)[lp.l2g[:no]]
    reqs, bufs = [], {}
    for q in range(world):
        a, b = lp.send_ptr[q], lp.send_ptr[q + 1]
        if b > a:
            t = torch.from_numpy(np.ascontiguousarray(xl[lp.send_nodes[a:b]]))
            reqs.append(dist.isend(t, q))
        if lp.recv_from[q]:
            # my segment written by q = the ghosts owned by q (contiguous: sorted global ids)
            gh = lp.l2g[no:]
            sel = np.where(P.owner_of(gh) == q)[0]
            bufs[q] = (sel, torch.zeros((sel.size, 2), dtype=torch.float64))
            reqs.append(dist.irecv(bufs[q][1], q))
    for r in reqs:
        r.wait()
    for q, (sel, t) in bufs.items():
        assert np.array_equal(sel, np.arange(sel[0], sel[0] + sel.size))   # contiguous segment
        xl[no + sel] = t.numpy()
    assert np.array_equal(xl, x.reshape(-1, 2)[lp.l2g])                    # ghosts arrived in place
    # owned rows of K restricted to local columns reproduce the global product
    g2l = np.full(case.mesh.NOD, -1)
    g2l[lp.l2g] = np.arange(nl)
    yl = np.zeros((no, 2))
    for a in range(no):
        ga = lp.l2g[a]
        for d in range(2):
            i = 2 * ga + d
            cols = col[rp[i]:rp[i + 1]]
            lc = g2l[cols // 2]
            assert np.all(lc >= 0)                                         # every column is local
            yl[a, d] = np.dot(val[rp[i]:rp[i + 1]], xl[lc, cols % 2])
    assert np.allclose(yl, y_ref.reshape(-1, 2)[lp.l2g[:no]], rtol=1e-13, atol=1e-30)
    # rank-ordered all-reduce of the partial dots = what every rank of the CUDA path computes
    part = torch.tensor([float(np.sum(yl * yl))], dtype=torch.float64)
    parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, part)
    tot = 0.0
    for p in parts:
        tot += float(p)
    assert abs(tot - float(y_ref @ y_ref)) <= 1e-12 * float(y_ref @ y_ref)
    dist.barrier()


if __name__ == "__main__":
    main()
