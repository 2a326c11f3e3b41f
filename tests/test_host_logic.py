"""Host-side logic that needs no GPU: timing, the basis-angle random stream, mesh builders and
readers, workload builders and the algorithmic-byte model."""
import math
import os

import numpy as np

import cases
from feellgood_b200 import meshgen, workloads
from feellgood_b200.linear_algebra import M_2_PI, c_srand, mt19937_uniform01, timing, _c_rand


def test_timing_matches_reference_header():
    # src/time_integration.h:11-15,37-42 ; unit-tests/ut_time_int.cpp:19-43 (dt0 = sqrt(a b))
    t = timing(2e-11, 1e-16, 5e-13)
    assert t.get_dt() == math.sqrt(1e-16 * 5e-13) and t.TAUR == 100 * 5e-13
    tt = t.get_dt() / t.TAUR
    assert t.prefactor == 1.0 + tt * abs(math.log(tt))
    t.set_dt(1e-13)
    assert not t.is_dt_TooSmall()
    t.set_dt(1e-17)
    assert t.is_dt_TooSmall()
    t.inc_t()
    assert t.get_t() == 1e-17


def test_mt19937_first_uniform_draw():
    # libstdc++ generate_canonical<double,53> of mt19937(5489): first two outputs 3499211612, 581869302
    r = mt19937_uniform01(5489)
    assert r == (3499211612 + 581869302 * 4294967296.0) / 18446744073709551616.0
    assert 0.0 <= r < 1.0 and M_2_PI == 2.0 / math.pi
    c_srand(2)
    a = [_c_rand() for _ in range(3)]
    c_srand(2)
    assert a == [_c_rand() for _ in range(3)]            # --seed N reproducibility (src/main.cpp:211)


def test_cuboid_counts_and_orientation():
    m = meshgen.cuboid([0, 0, 0], [4, 3, 2], 4, 3, 2, scale=1e-9)
    assert m.NOD == 5 * 4 * 3 and m.NT == 4 * 3 * 2 * 6
    assert m.NF == 2 * 2 * (4 * 3 + 4 * 2 + 3 * 2)       # two triangles per boundary quad
    p = m.node_p[m.tet_ind]
    vol = np.einsum("ij,ij->i", p[:, 1] - p[:, 0], np.cross(p[:, 2] - p[:, 0], p[:, 3] - p[:, 0])) / 6
    assert abs(np.abs(vol).sum() - 24e-27) < 1e-36
    # boundary triangles point outwards: sum of area vectors vanishes, divergence theorem gives V
    q = m.node_p[m.tri_ind]
    nrm = 0.5 * np.cross(q[:, 1] - q[:, 0], q[:, 2] - q[:, 0])
    assert np.max(np.abs(nrm.sum(axis=0))) < 1e-30
    assert abs(np.einsum("ij,ij->", q.mean(axis=1), nrm) / 3 - 24e-27) < 1e-36


def test_sort_nodes_is_a_consistent_renumbering():
    m = meshgen.cuboid([0, 0, 0], [2, 7, 3], 2, 5, 3, scale=1e-9)
    p0, t0 = m.node_p.copy(), m.tet_ind.copy()
    meshgen.sort_nodes(m)
    assert np.all(np.diff(m.node_p[:, 1]) >= 0)          # longest axis = y
    assert np.array_equal(m.node_p[m.node_index[t0]], p0[t0])
    assert np.array_equal(m.node_p[m.tet_ind], p0[t0])


def test_dMs_of_closed_surface():
    m = meshgen.cuboid([0, 0, 0], [2, 2, 2], 2, 2, 2, scale=1e-9)
    d = meshgen.compute_dMs(m, [0.0, 8e5])
    assert np.all(d == 8e5)                              # outward surface of a single region
    m.tri_ind = m.tri_ind[:, [0, 2, 1]].copy()
    assert np.all(meshgen.compute_dMs(m, [0.0, 8e5]) == -8e5)


def test_read_msh_v2_and_v4(tmp_path):
    v2 = """$MeshFormat
2.2 0 8
$EndMeshFormat
$PhysicalNames
2
2 200 "surf"
3 300 "vol"
$EndPhysicalNames
$Nodes
5
1 0 0 0
2 1 0 0
3 0 1 0
4 0 0 1
5 1 1 1
$EndNodes
$Elements
4
1 2 2 200 1 1 3 2
2 4 2 300 1 1 2 3 4
3 4 2 300 1 2 3 4 5
4 15 2 1 1 1
$EndElements
"""
    f = tmp_path / "a.msh"
    f.write_text(v2)
    m = meshgen.read_msh(str(f), ["vol"], ["surf"], scale=1e-9)
    assert (m.NOD, m.NT, m.NF) == (5, 2, 1)
    assert np.array_equal(m.tet_ind, [[0, 1, 2, 3], [1, 2, 3, 4]]) and np.all(m.tet_reg == 1)
    assert np.array_equal(m.tri_ind, [[0, 2, 1]]) and m.node_p[4, 2] == 1e-9
    assert meshgen.read_msh(str(f), ["other"], [], scale=1.0).NT == 0     # unknown region dropped
    z = np.load(os.path.join(cases.GOLDEN, "ellipsoid_mesh.npz"))
    assert (z["node_p"].shape[0], z["tet_ind"].shape[0], z["tri_ind"].shape[0]) == (167, 499, 274)


def test_workload_builders_small():
    for name, scale, nt in (("film20m", 0.02, 26 * 26 * 2 * 6), ("sp4", 0.1, 25 * 6 * 2 * 6),
                            ("tube5m", 0.01, 8 * 152 * 7 * 6), ("ellipsoid", 1.0, 499)):
        w = workloads.build(name, scale=scale)
        assert w.mesh.NT == nt, name
        assert np.max(np.abs(np.linalg.norm(w.u, axis=1) - 1)) < 1e-14
        assert w.mesh.tet_ind.min() == 0 and w.mesh.tet_ind.max() == w.mesh.NOD - 1
        t = w.timing()
        assert t.get_dt() == w.dt and t.prefactor > 1.0
    d = workloads.build("disk1m", scale=0.1)
    r = np.hypot(d.mesh.node_p[:, 0], d.mesh.node_p[:, 1])
    assert r.max() < 250e-9 * 1.1 and d.mesh.NT % 6 == 0


def test_algorithmic_byte_model():
    n, nnz = 2000, 60000
    assert workloads.spmv_bytes_csr(n, nnz) == 12 * nnz + 20 * n
    assert workloads.spmv_bytes_blocks(n, nnz) == 9 * nnz + 4 * (n // 2) + 16 * n   # assembled 2x2 blocks
    # matrix-free operator: 12 B (10 B with 16-bit column offsets) per node pair, 113 B per node
    # (+16 B when x itself is read: setup stage, taps), one slice pointer per 32 nodes
    assert workloads.spmv_bytes(n, nnz) == 3 * nnz + 113 * (n // 2) + 4 * (n // 64)
    assert workloads.spmv_bytes(n, nnz, 2) == 10 * (nnz // 4) + 113 * (n // 2) + 4 * (n // 64)
    assert workloads.spmv_bytes(n, nnz, 4, True) == 3 * nnz + 129 * (n // 2) + 4 * (n // 64)
    assert workloads.spmv_bytes(n, nnz) < workloads.spmv_bytes_blocks(n, nnz) < workloads.spmv_bytes_csr(n, nnz)
    assert workloads.iter_bytes(n, nnz) == 2 * workloads.spmv_bytes(n, nnz) + 200 * n
    # SURVEY.md §8d B_step, term by term
    assert workloads.step_bytes_survey(1000, 6000, n, nnz, 2) == (
        72 * 1000 + (124 * 6000 + 112 * 1000 + 8 * nnz + 8 * n) + 88 * 1000 + (12 * nnz + 20 * n + 80 * n)
        + 2 * (2 * (12 * nnz + 20 * n) + 144 * n) + 136 * 1000)
    b0 = workloads.step_bytes(1000, 6000, n, nnz, 0)
    assert workloads.step_bytes(1000, 6000, n, nnz, 10) - b0 == 10 * workloads.iter_bytes(n, nnz)


def test_persistent_kernel_launch_shape():
    """Launch shape of the persistent solve kernel (fg_krylov.cu pk_plan / pk_warps through
    fg_solver_launch_shape, no device needed): the variant per mesh size, and the warp count of a
    1024-thread launch -- fewest rounds of the slice front first, then the fullest last round.  The
    interior and end ranks of the 4-GPU partition of the 20 M-tet mesh (39 035 / 39 091 slices) must get
    the same shape: the first heuristic gave them 24 warps / 11 rounds against 30 / 9."""
    from feellgood_b200 import capi
    S = 148
    shape = lambda n: capi.solver_launch_shape(n, S)
    # variants: one 256-thread CTA up to 8 slices; 512 threads + register-held heads up to 16 warps per SM;
    # 256 threads (4 CTAs per SM) below one slice per warp of a full 1024-thread grid; 1024 threads beyond
    assert shape(6) == dict(block=256, warps=8, grid=1, head=False)
    assert shape(8)["grid"] == 1 and not shape(8)["head"]
    assert shape(9) == dict(block=512, warps=16, grid=1, head=True)
    assert shape(1483) == dict(block=512, warps=16, grid=93, head=True)          # sp4
    assert shape(S * 16) == dict(block=512, warps=16, grid=S, head=True)
    assert shape(S * 16 + 1) == dict(block=256, warps=8, grid=297, head=False)
    assert shape(S * 32 - 1) == dict(block=256, warps=8, grid=4 * S, head=False)
    assert shape(S * 32) == dict(block=1024, warps=32, grid=S, head=False)
    assert shape(156252) == dict(block=1024, warps=32, grid=S, head=False)       # film20m
    assert shape(78126)["warps"] == 32                                           # its 2-GPU partition
    assert shape(39035) == shape(39091) == dict(block=1024, warps=30, grid=S, head=False)   # 4 GPUs
    assert shape(19517) == shape(19574) == dict(block=1024, warps=27, grid=S, head=False)   # 8 GPUs
    rounds = lambda n, nw: -(-n // (S * nw))
    for n in list(range(S * 32, S * 32 * 16, 997)) + [S * 32 * 16 - 1]:
        s = shape(n)
        assert s["block"] == 1024 and s["grid"] == S and 24 <= s["warps"] <= 32
        best = min(rounds(n, nw) for nw in range(24, 33))
        assert rounds(n, s["warps"]) == best == rounds(n, 32), n       # never more rounds than 32 warps
        assert s["warps"] == min(nw for nw in range(24, 33) if rounds(n, nw) == best), n
    # every slice has a warp: warps of the grid x rounds covers the mesh; small meshes get one slice per warp
    for n in (1, 5, 33, 500, 2368, 2369, 4000, 4735):
        s = shape(n)
        assert s["grid"] * s["warps"] >= n or s["block"] == 256 and s["grid"] == 4 * S
    # another SM count (the shape follows the device, not a constant)
    assert capi.solver_launch_shape(132 * 32, 132) == dict(block=1024, warps=32, grid=132, head=False)
    import pytest
    with pytest.raises(capi.FgError):
        capi.solver_launch_shape(0)
