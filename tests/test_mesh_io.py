"""Mesh-side inputs and on-disk formats: mesh::controlTriangles against the
reference's own known-answer meshes and messages (unit-tests/ut_readMesh.cpp:78-133), the counts of
ut_readMesh.cpp:19-76 for examples/ellipsoid.msh, and the .sol / .evol round trips."""
import os

import numpy as np

import cases
from feellgood_b200 import io as fio
from feellgood_b200 import meshgen

G = np.load(os.path.join(cases.GOLDEN, "bad_cuboids.npz"))
EXPECTED = [
    "Error: bad mesh. 2 instances of the same surface triangle have been found\n",
    "Error: bad mesh. A triangle which belongs to 1 surface region and no tetrahedron has been found\n",
    "Error: bad mesh. A triangular face shared by 3 tetrahedrons has been found\n",
    "Error: bad mesh. An internal triangle has been found in the surface region 1\n",
]


def test_control_triangles_known_answers():
    for k in range(1, 5):
        m = meshgen.Mesh(**{n: G["m%d_%s" % (k, n)] for n in ("node_p", "tet_ind", "tet_reg", "tri_ind", "tri_reg")},
                         tri_dMs=np.zeros(len(G["m%d_tri_reg" % k])))
        ok, msg, _ = meshgen.control_triangles(m, ["__default__", "whole_volume"],
                                               ["__default__", "whole_surface"], [0.0, 8e5])
        assert not ok and msg == EXPECTED[k - 1]


def test_ellipsoid_counts_and_dMs():
    """ut_readMesh.cpp:35,59,68: 167 nodes, 499 tetrahedra, 274 triangles; a complete closed surface
    needs no extra triangle and every dMs equals Ms (outward-oriented surface elements)."""
    case = cases.ellipsoid()
    m = case.mesh
    assert (m.NOD, m.NT, m.NF) == (167, 499, 274)
    m2 = meshgen.Mesh(m.node_p, m.tet_ind, m.tet_reg, m.tri_ind.copy(), m.tri_reg.copy(), np.zeros(m.NF))
    ok, msg, names = meshgen.control_triangles(m2, ["__default__", "ellipsoid_volume"],
                                               ["__default__", "ellipsoid_surface"], [0.0, 795774.7])
    assert ok and msg == "" and names == ["__default__", "ellipsoid_surface"] and m2.NF == 274
    assert np.array_equal(m2.tri_dMs, m.tri_dMs) and np.all(m2.tri_dMs == 795774.7)


def test_missing_surfaces_and_interfaces_are_created():
    """src/mesh.cpp:176-225: boundary faces not listed in the file get a 'surface(vol)' region,
    faces between two volume regions an 'interface(a, b)' region (first = larger region index)."""
    c = meshgen.cuboid([0, 0, 0], [4, 3, 2], 4, 3, 2, scale=1e-9, with_surface=False)
    c.tet_reg[:] = 1
    c.tet_reg[c.node_p[c.tet_ind].mean(axis=1)[:, 0] > 2e-9] = 2
    ok, msg, names = meshgen.control_triangles(c, ["__default__", "a", "b"], ["__default__"], [0.0, 8e5, 5e5])
    assert ok and names == ["__default__", "surface(a)", "interface(b, a)", "surface(b)"]
    nb = np.bincount(c.tri_reg, minlength=4)
    assert nb[1] == nb[3] == 2 * (3 * 2 + 2 * 2 * 2 + 2 * 2 * 3)   # each half: 1 x-face, 2 y, 2 z
    assert nb[2] == 2 * 3 * 2                                  # the x = 2 nm plane: 3 x 2 cells
    assert set(np.abs(c.tri_dMs[c.tri_reg == 1])) == {8e5}
    assert set(np.abs(c.tri_dMs[c.tri_reg == 3])) == {5e5}
    assert set(np.abs(c.tri_dMs[c.tri_reg == 2])) == {3e5}    # |Ms_a - Ms_b| across the interface
    # a second call finds nothing left to add
    nf = c.NF
    ok2, _, names2 = meshgen.control_triangles(c, ["__default__", "a", "b"], names, [0.0, 8e5, 5e5])
    assert ok2 and c.NF == nf and names2 == names


def test_sol_round_trip(tmp_path):
    """mesh::savesol -> mesh::readSol: rows are in the mesh file's node numbering (node_index)."""
    rng = np.random.default_rng(3)
    c = meshgen.cuboid([0, 0, 0], [5, 2, 1], 5, 2, 1, scale=1e-9)
    meshgen.sort_nodes(c)
    u = cases.unit_rows(rng.standard_normal((c.NOD, 3)))
    phi = rng.standard_normal(c.NOD) * 1e3
    path = str(tmp_path / "snap_iter7.sol")
    fio.savesol(path, u, phi, node_index=c.node_index, t=3.5e-11, precision=7)
    head = open(path).read().split("\n")[:3]
    assert head[0].startswith("## time: 3.5") and head[1] == "## columns: idx\tmx\tmy\tmz\tphi"
    assert head[2].split("\t")[0] == "0" and len(head[2].split("\t")) == 5
    t, u2, phi2 = fio.read_sol(path, c.NOD, node_index=c.node_index)
    assert t == 3.5e-11
    assert np.max(np.abs(u2 - u)) < 1e-7 and cases.rel_max(phi2, phi) < 1e-7   # 7 digits
    # a file without the time tag is rejected like readSol does
    bad = str(tmp_path / "bad.sol")
    open(bad, "w").write("0\t1\t0\t0\t0\n")
    try:
        fio.read_sol(bad, 1)
        assert False
    except ValueError as e:
        assert "no ## time: tag" in str(e)


def test_evol_round_trip(tmp_path):
    from feellgood_b200 import Settings
    from feellgood_b200.fem import Fem
    s = Settings([], evol_columns=["iter", "t", "<Mx>", "E_tot"])
    fem = Fem(s, None)
    fem.evol = [[0, 0.0, 0.25, -1.5e-19], [1, 1e-12, 0.125, -1.625e-19]]
    path = str(tmp_path / "run.evol")
    fem.write_evol(path)
    cols, rows = fio.read_evol(path)
    assert cols == ["iter", "t", "<Mx>", "E_tot"]
    assert np.array_equal(rows, np.array(fem.evol, dtype=float))     # %.16g round-trips doubles
