"""Host-side logic of the multi-GPU path (no GPU): the slab partition, local meshes and halo plan
of feellgood_b200/dist.py, serially and as a world_size-2 gloo job that performs the same halo
exchange + scalar all-reduce pattern the CUDA path performs (with torch.distributed standing in for
the NVLink peer stores) and checks a partitioned SpMV / dot product against the global one."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from feellgood_b200.dist import Partition


def _plans(mesh, world, method="slab"):
    P = Partition(mesh, world, method=method)
    return P, [P.local(r) for r in range(world)]


@pytest.mark.parametrize("method", ["slab", "rcb"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_invariants(world, method):
    case = cases.small_cuboid(nx=12, ny=5, nz=3)
    mesh = case.mesh
    P, lps = _plans(mesh, world, method)
    if method == "slab":
        assert P.cuts[0] == 0 and P.cuts[-1] == mesh.NOD and np.all(np.diff(P.cuts) >= 0)
    owned = np.concatenate([lp.l2g[:lp.n_owned] for lp in lps])
    assert np.array_equal(np.sort(owned), np.arange(mesh.NOD))   # every node owned exactly once
    if method == "slab":
        assert np.array_equal(owned, np.arange(mesh.NOD))        # ... in contiguous blocks
    tets_seen = np.zeros(mesh.NT, dtype=int)
    for lp in lps:
        # local meshes are consistent renumberings of global tets, all touching an owned node
        gt = lp.l2g[lp.mesh.tet_ind]
        assert np.all((lp.mesh.tet_ind < lp.n_owned).any(axis=1))
        key = {tuple(t) for t in mesh.tet_ind.tolist()}
        assert all(tuple(t) in key for t in gt.tolist())
        assert np.array_equal(lp.mesh.node_p, mesh.node_p[lp.l2g])
        # ghosts are exactly the non-owned nodes of those tets, grouped by owner, ascending inside a group
        # (a slab partition: plain ascending global id)
        gh = lp.l2g[lp.n_owned:]
        key2 = P.owner_of(gh).astype(np.int64) * mesh.NOD + gh
        assert np.all(np.diff(key2) > 0)
        if method == "slab":
            assert np.all(np.diff(gh) > 0)
        assert set(gh.tolist()) == set(np.unique(gt).tolist()) - set(lp.l2g[:lp.n_owned].tolist())
        mask = (P.owner[mesh.tet_ind] == lp.rank).any(axis=1)
        tets_seen += mask
        assert lp.mesh.NT == mask.sum()
        # triangles: all nodes local, at least one owned
        if lp.mesh.NF:
            assert np.all((lp.mesh.tri_ind < lp.n_owned).any(axis=1))
    assert np.all(tets_seen >= 1)
    # halo plan: every ghost slot of every rank is written by exactly one sender, with the right node
    for q, lq in enumerate(lps):
        cover = np.zeros(lq.n_ghost, dtype=int)
        gq = lq.l2g[lq.n_owned:]
        for lp in lps:
            a, b = lp.send_ptr[q], lp.send_ptr[q + 1]
            if b > a:
                assert lp.rank != q and lq.recv_from[lp.rank] == 1
                assert np.array_equal(lp.l2g[lp.send_nodes[a:b]], gq[lp.send_dst[q]:lp.send_dst[q] + b - a])
                cover[lp.send_dst[q]:lp.send_dst[q] + b - a] += 1
            else:
                assert lq.recv_from[lp.rank] == 0
        assert np.all(cover == 1)


def test_balance_and_slab_neighbours():
    case = cases.film(64, 16, 2)
    P, lps = _plans(case.mesh, 4)
    sizes = np.diff(P.cuts)
    assert sizes.max() <= 1.25 * sizes.min()
    for lp in lps:                                   # slabs: at most two neighbours
        nb = {q for q in range(4) if lp.send_ptr[q + 1] > lp.send_ptr[q]}
        assert nb <= {lp.rank - 1, lp.rank + 1}
        assert np.all(P.cuts[1:-1] % 32 == 0)


def test_rcb_is_compact_and_balanced():
    """On a cube-shaped body the geometric k-way partition has a smaller halo than slabs, with balanced
    weights; on a film both are balanced."""
    from feellgood_b200 import meshgen
    m = meshgen.cuboid([0, 0, 0], [16.0, 16.0, 16.0], 16, 16, 16, scale=1e-9, with_surface=False)
    meshgen.sort_nodes(m)
    ghosts = {}
    for method in ("slab", "rcb"):
        P, lps = _plans(m, 8, method)
        w = 1.0 + np.bincount(m.tet_ind.ravel(), minlength=m.NOD)
        loads = np.array([w[lp.l2g[:lp.n_owned]].sum() for lp in lps])
        assert loads.max() <= 1.15 * loads.mean(), (method, loads)
        ghosts[method] = sum(lp.n_ghost for lp in lps)
    assert ghosts["rcb"] < 0.75 * ghosts["slab"], ghosts


def test_gloo_world2_halo_and_allreduce(tmp_path):
    """Two processes, gloo: partitioned y = K x and <y, y> with ghost exchange per the halo plan."""
    worker = os.path.join(cases.ROOT, "tests", "dist_cpu_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", worker]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=cases.ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_CPU_OK" in r.stdout
