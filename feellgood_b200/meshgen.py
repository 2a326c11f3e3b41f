"""Host-side mesh builders for the LLG hot path (inputs only; not on the timed path).

The reference reads gmsh files through libgmsh (src/read.cpp:12-197) and post-processes them in
the Mesh::mesh constructor (src/mesh.h:38-135): node sort along the longest axis
(src/mesh.cpp:334-367), surface/interface detection with dMs (src/mesh.cpp:121-245).  libgmsh is
not available here, so this module provides

* ``read_msh``      — ASCII gmsh 2.2 / 4.1 reader for first-order tets + triangles,
* ``hex_grid``      — the gmsh-free ``Cuboid`` scheme of python-modules/meshMaker.py:440-519
                      (6 tets per hexahedron), vectorised, with an optional cell mask (disk) and a
                      periodic second axis + coordinate map (tube), used to synthesise the
                      BASELINE.json configs 2-5,
* ``sort_nodes``    — restatement of mesh::sortNodes,
* ``compute_dMs``   — restatement of the dMs loop of mesh::controlTriangles.

Everything returns a ``Mesh`` with zero-based int32 connectivity and float64 coordinates already
multiplied by the length unit (src/read.cpp:77-96).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    node_p: np.ndarray            # (NOD, 3) float64
    tet_ind: np.ndarray           # (NT, 4) int32
    tet_reg: np.ndarray           # (NT,) int32 index into the volume-region parameter list
    tri_ind: np.ndarray           # (NF, 3) int32
    tri_reg: np.ndarray           # (NF,) int32 index into the surface-region parameter list
    tri_dMs: np.ndarray           # (NF,) float64
    node_index: np.ndarray | None = None   # file index -> stored index (mesh::node_index)
    vol_names: list = field(default_factory=list)
    surf_names: list = field(default_factory=list)

    @property
    def NOD(self):
        return int(self.node_p.shape[0])

    @property
    def NT(self):
        return int(self.tet_ind.shape[0])

    @property
    def NF(self):
        return int(self.tri_ind.shape[0])


# --------------------------------------------------------------------------------------------
# Cuboid scheme (python-modules/meshMaker.py:440-519)
# --------------------------------------------------------------------------------------------
# corners: A=(i,j,k) B=(i+1,j,k) C=(i+1,j+1,k) D=(i,j+1,k) E=(i,j,k+1) F=(i+1,j,k+1)
#          G=(i+1,j+1,k+1) H=(i,j+1,k+1)
_CORNER = dict(A=(0, 0, 0), B=(1, 0, 0), C=(1, 1, 0), D=(0, 1, 0),
               E=(0, 0, 1), F=(1, 0, 1), G=(1, 1, 1), H=(0, 1, 1))
_TETS = ["ABDH", "ABEH", "EFHB", "BCDH", "BCGH", "FBGH"]


def hex_grid(nx, ny, nz, coords=None, cell_mask=None, periodic_y=False, scale=1.0,
             pt_min=(0.0, 0.0, 0.0), pt_max=(1.0, 1.0, 1.0), with_surface=True):
    """6-tets-per-hex mesh of an nx×ny×nz cell grid.

    coords: optional callable (X, Y, Z) -> (x, y, z) applied to the regular grid coordinates
            (for periodic_y the Y coordinate runs over [pt_min[1], pt_max[1]) without the seam).
    cell_mask: optional bool array (nx, ny, nz) selecting the cells to keep.
    periodic_y: identify node layer j=ny with j=0 (closed annulus).
    Surface triangles (the faces used exactly once, oriented outwards) go to surface region 1.
    """
    nyn = ny if periodic_y else ny + 1
    gx = np.linspace(pt_min[0], pt_max[0], nx + 1)
    gy = pt_min[1] + (pt_max[1] - pt_min[1]) * np.arange(nyn) / ny
    gz = np.linspace(pt_min[2], pt_max[2], nz + 1)
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    if coords is not None:
        X, Y, Z = coords(X, Y, Z)
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1) * scale

    def nid(i, j, k):
        return ((nz + 1) * nyn * i + (nz + 1) * (j % nyn if periodic_y else j) + k)

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    if cell_mask is not None:
        keep = cell_mask.astype(bool)
        I, J, K = I[keep], J[keep], K[keep]
    else:
        I, J, K = I.ravel(), J.ravel(), K.ravel()
    corner = {c: nid(I + o[0], J + o[1], K + o[2]).astype(np.int64) for c, o in _CORNER.items()}
    ncell = I.size
    tets = np.empty((ncell, 6, 4), dtype=np.int64)
    for t, names in enumerate(_TETS):
        for c, name in enumerate(names):
            tets[:, t, c] = corner[name]
    tets = tets.reshape(-1, 4)

    # compact node numbering when cells were masked out
    if cell_mask is not None:
        used = np.zeros(pts.shape[0], dtype=bool)
        used[tets.ravel()] = True
        remap = np.cumsum(used) - 1
        tets = remap[tets]
        pts = pts[used]

    mesh = Mesh(node_p=np.ascontiguousarray(pts, dtype=np.float64),
                tet_ind=np.ascontiguousarray(tets, dtype=np.int32),
                tet_reg=np.ones(tets.shape[0], dtype=np.int32),
                tri_ind=np.zeros((0, 3), dtype=np.int32), tri_reg=np.zeros(0, dtype=np.int32),
                tri_dMs=np.zeros(0, dtype=np.float64),
                vol_names=["__default__", "volume"], surf_names=["__default__", "surface"])
    if with_surface:
        tri = boundary_faces(mesh.node_p, mesh.tet_ind)
        mesh.tri_ind = tri
        mesh.tri_reg = np.ones(tri.shape[0], dtype=np.int32)
        mesh.tri_dMs = np.zeros(tri.shape[0], dtype=np.float64)
    return mesh


def cuboid(pt_min, pt_max, nx, ny, nz, scale=1e-9, with_surface=True):
    """``Cuboid(pt_min, pt_max, nbX, nbY, nbZ)`` of python-modules/meshMaker.py:440-519."""
    return hex_grid(nx, ny, nz, pt_min=pt_min, pt_max=pt_max, scale=scale,
                    with_surface=with_surface)


def disk(radius, thickness, ncell_diam, nz, scale=1e-9, with_surface=True):
    """Disk of given radius/thickness: Cartesian hex grid clipped to the circle (staircase rim)."""
    n = int(ncell_diam)
    h = 2.0 * radius / n
    cx = -radius + (np.arange(n) + 0.5) * h
    CX, CY = np.meshgrid(cx, cx, indexing="ij")
    inside = (CX * CX + CY * CY) <= radius * radius
    mask = np.repeat(inside[:, :, None], nz, axis=2)
    return hex_grid(n, n, nz, cell_mask=mask, pt_min=(-radius, -radius, 0.0),
                    pt_max=(radius, radius, thickness), scale=scale, with_surface=with_surface)


def tube(r1, r2, length, nr, ntheta, nz, scale=1e-9, with_surface=True):
    """Closed annular tube r1<r<r2 (examples/tube.py geometry), structured (r, theta, z) grid with
    the theta axis periodic; the tube axis is z like in the reference example."""
    def to_xyz(R, T, Z):
        return R * np.cos(T), R * np.sin(T), Z
    return hex_grid(nr, ntheta, nz, coords=to_xyz, periodic_y=True, pt_min=(r1, 0.0, -0.5 * length),
                    pt_max=(r2, 2.0 * np.pi, 0.5 * length), scale=scale, with_surface=with_surface)


# --------------------------------------------------------------------------------------------
# topology helpers
# --------------------------------------------------------------------------------------------
def _oriented(node_p, tet_ind):
    """Tet::orientate (src/tetra.cpp:410-424): swap ind[2], ind[3] when the mixed product < 0."""
    p = node_p[tet_ind]
    a, b, c = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]
    mixed = np.einsum("ij,ij->i", a, np.cross(b, c))
    t = tet_ind.copy()
    neg = mixed < 0
    t[neg, 2], t[neg, 3] = tet_ind[neg, 3], tet_ind[neg, 2]
    return t


def _outward_faces(tet_oriented):
    """The four faces of each tet, oriented outwards (src/mesh.cpp:137-140)."""
    t = tet_oriented
    return np.concatenate([t[:, [0, 2, 1]], t[:, [1, 2, 3]], t[:, [2, 0, 3]], t[:, [3, 0, 1]]])


def _perm_parity(tri):
    """0 when (a,b,c) is an even permutation of its sorted order, 1 when odd."""
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    inv = (a > b).astype(np.int8) + (a > c).astype(np.int8) + (b > c).astype(np.int8)
    return inv & 1


def _face_keys(tri, nod):
    s = np.sort(tri.astype(np.int64), axis=1)
    return (s[:, 0] * nod + s[:, 1]) * nod + s[:, 2]


def boundary_faces(node_p, tet_ind):
    """Faces belonging to exactly one tetrahedron, oriented outwards."""
    t = _oriented(node_p, tet_ind)
    faces = _outward_faces(t)
    keys = _face_keys(faces, node_p.shape[0])
    _, first, counts = np.unique(keys, return_index=True, return_counts=True)
    sel = np.sort(first[counts == 1])
    return np.ascontiguousarray(faces[sel], dtype=np.int32)


def compute_dMs(mesh, Ms_per_vol_region):
    """dMs loop of mesh::controlTriangles (src/mesh.cpp:227-242): for every tet face that matches
    a surface element, dMs += Ms of the tet's region, negated when the orientations differ."""
    if mesh.NF == 0:
        return np.zeros(0)
    nod = mesh.NOD
    t = _oriented(mesh.node_p, mesh.tet_ind)
    faces = _outward_faces(t)
    fkeys = _face_keys(faces, nod)
    fpar = _perm_parity(faces)
    fMs = np.tile(np.asarray(Ms_per_vol_region, dtype=np.float64)[mesh.tet_reg], 4)
    order = np.argsort(fkeys, kind="stable")
    fkeys, fpar, fMs = fkeys[order], fpar[order], fMs[order]
    tkeys = _face_keys(mesh.tri_ind, nod)
    tpar = _perm_parity(mesh.tri_ind)
    dMs = np.zeros(mesh.NF)
    lo = np.searchsorted(fkeys, tkeys, side="left")
    hi = np.searchsorted(fkeys, tkeys, side="right")
    for shift in range(2):                      # a face is shared by at most two tets
        idx = lo + shift
        ok = idx < hi
        ii = np.where(ok, idx, 0)
        flipped = (tpar ^ fpar[ii]).astype(bool)
        dMs += np.where(ok, np.where(flipped, -fMs[ii], fMs[ii]), 0.0)
    return dMs


def control_triangles(mesh, vol_names, surf_names, Ms_per_vol_region):
    """mesh::controlTriangles (src/mesh.cpp:121-245) with diffTriHandler / assertNoErrInTriangle
    (src/mesh.cpp:247-332): validates the surface elements against the faces of the tetrahedra,
    creates the surface / interface triangles the mesh file does not list (new regions named
    ``surface(vol)`` / ``interface(vol1, vol2)`` with the default surface parameters) and sets dMs.

    vol_names / surf_names: region names, index 0 = "__default__" (Settings::paramTetra[0] /
    paramTriangle[0]).  Must run BEFORE sort_nodes, like in the reference's mesh constructor.
    Returns (ok, message, surf_names): on failure `message` is the reference's stderr text and the
    mesh is untouched; on success surf_names includes the created regions, mesh.tri_ind / tri_reg are
    extended and mesh.tri_dMs is set."""
    if mesh.NT == 0:
        return False, "Error: not a single tetrahedron is present in the mesh\n", list(surf_names)
    nod = mesh.NOD
    faces = _outward_faces(_oriented(mesh.node_p, mesh.tet_ind)).astype(np.int64)
    f_reg = np.tile(mesh.tet_reg.astype(np.int64), 4)
    tri = mesh.tri_ind.astype(np.int64).reshape(-1, 3)
    allv = np.concatenate([faces, tri])
    is_surf = np.concatenate([np.zeros(len(faces), bool), np.ones(len(tri), bool)])
    reg = np.concatenate([f_reg, mesh.tri_reg.astype(np.int64)])
    srt = np.sort(allv, axis=1)
    flipped = _perm_parity(allv).astype(bool)
    key = (srt[:, 0] * nod + srt[:, 1]) * nod + srt[:, 2]
    order = np.argsort(key, kind="stable")
    key, srt, is_surf, reg, flipped = key[order], srt[order], is_surf[order], reg[order], flipped[order]
    start = np.flatnonzero(np.concatenate([[True], key[1:] != key[:-1]]))
    gid = np.cumsum(np.concatenate([[True], key[1:] != key[:-1]])) - 1
    ng = start.size
    nb_surf = np.bincount(gid, weights=is_surf, minlength=ng).astype(np.int64)
    nb_face = np.bincount(gid, weights=~is_surf, minlength=ng).astype(np.int64)
    # sorted pair of the volume regions around the face: first >= second, -1 = none
    big = np.full(ng, -1, dtype=np.int64)
    np.maximum.at(big, gid[~is_surf], reg[~is_surf])
    small_src = np.where(~is_surf, reg, np.iinfo(np.int64).max)
    small = np.full(ng, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(small, gid, small_src)
    second = np.where(nb_face >= 2, small, -1)
    surf_reg = np.full(ng, -1, dtype=np.int64)
    surf_reg[gid[is_surf]] = reg[is_surf]                 # the last one of the group, like the walk
    # first group (in the sorted walk) that violates a rule decides the message
    e1 = nb_surf > 1
    e2 = ~e1 & (nb_face == 0)
    e3 = ~e1 & ~e2 & (nb_face > 2)
    e4 = ~e1 & ~e2 & ~e3 & (nb_face == 2) & (nb_surf == 1) & (big == second)
    bad = np.flatnonzero(e1 | e2 | e3 | e4)
    if bad.size:
        g = bad[0]
        if e1[g]:
            msg = "Error: bad mesh. %d instances of the same surface triangle have been found\n" % nb_surf[g]
        elif e2[g]:
            msg = ("Error: bad mesh. A triangle which belongs to 1 surface region and no tetrahedron "
                   "has been found\n")
        elif e3[g]:
            msg = "Error: bad mesh. A triangular face shared by %d tetrahedrons has been found" % nb_face[g]
            msg += (" in the surface region %d\n" % surf_reg[g]) if nb_surf[g] == 1 else "\n"
        else:
            msg = "Error: bad mesh. An internal triangle has been found in the surface region %d\n" % surf_reg[g]
        return False, msg, list(surf_names)
    surf_names = list(surf_names)
    if not surf_names or surf_names[0] != "__default__":
        return False, "Internal Error: could not find default surface region.\n", surf_names
    # missing surface / interface triangles, in walk order; node order = sorted indices
    add = np.flatnonzero((nb_surf == 0) & ((nb_face == 1) | ((nb_face == 2) & (big != second))))
    new_tri = srt[start[add]]
    pairs = list(zip(big[add].tolist(), second[add].tolist()))
    pair_to_region = {}
    new_reg = np.zeros(len(pairs), dtype=np.int32)
    for k, pr in enumerate(pairs):
        if pr not in pair_to_region:
            base = ("surface(%s)" % vol_names[pr[0]]) if pr[1] == -1 else \
                   ("interface(%s, %s)" % (vol_names[pr[0]], vol_names[pr[1]]))
            name, i = base, 0
            while name in surf_names:
                i += 1
                name = "%s.%d" % (base, i)
            surf_names.append(name)
            pair_to_region[pr] = len(surf_names) - 1
        new_reg[k] = pair_to_region[pr]
    if len(add):
        mesh.tri_ind = np.ascontiguousarray(np.concatenate([tri, new_tri]).astype(np.int32))
        mesh.tri_reg = np.ascontiguousarray(np.concatenate([mesh.tri_reg.astype(np.int32), new_reg]))
    mesh.tri_dMs = compute_dMs(mesh, Ms_per_vol_region)
    return True, "", surf_names


def sort_nodes(mesh):
    """mesh::sortNodes (src/mesh.cpp:334-367) along the longest axis chosen as in
    src/mesh.h:61-79.  A stable sort is used (std::sort leaves ties unspecified)."""
    p = mesh.node_p
    l = p.max(axis=0) - p.min(axis=0)
    if l[0] > l[1]:
        axis = 0 if l[0] > l[2] else 2
    else:
        axis = 1 if l[1] > l[2] else 2
    perm = np.argsort(p[:, axis], kind="stable")
    node_index = np.empty_like(perm)
    node_index[perm] = np.arange(perm.size)
    mesh.node_p = np.ascontiguousarray(p[perm])
    mesh.tet_ind = np.ascontiguousarray(node_index[mesh.tet_ind].astype(np.int32))
    if mesh.NF:
        mesh.tri_ind = np.ascontiguousarray(node_index[mesh.tri_ind].astype(np.int32))
    mesh.node_index = node_index.astype(np.int64)
    return mesh


# --------------------------------------------------------------------------------------------
# gmsh ASCII reader (2.2 and 4.1), first-order tets (type 4) and triangles (type 2) only
# --------------------------------------------------------------------------------------------
def read_msh(path, vol_regions, surf_regions, scale=1e-9):
    """Read the named volume / surface regions of an ASCII .msh file.

    vol_regions / surf_regions are the region-name lists of the settings; index 0 of the returned
    region ids is reserved for ``__default__`` like Settings::paramTetra[0] / paramTriangle[0]
    (src/mesh.cpp:172-176).  Elements outside the named regions are dropped like
    mesh::readTetraedrons / readTriangles do.
    """
    with open(path, "r") as f:
        tok = f.read().split("\n")
    sections = {}
    i = 0
    while i < len(tok):
        line = tok[i].strip()
        if line.startswith("$") and not line.startswith("$End"):
            name = line[1:]
            j = i + 1
            while tok[j].strip() != "$End" + name:
                j += 1
            sections[name] = tok[i + 1:j]
            i = j
        i += 1
    version = float(sections["MeshFormat"][0].split()[0])
    phys = {}  # (dim, tag) -> name
    for ln in sections.get("PhysicalNames", [])[1:]:
        parts = ln.split(None, 2)
        phys[(int(parts[0]), int(parts[1]))] = parts[2].strip().strip('"')
    vol_id = {n: k + 1 for k, n in enumerate(vol_regions)}
    surf_id = {n: k + 1 for k, n in enumerate(surf_regions)}

    tets, treg, tris, freg = [], [], [], []
    if version >= 4.0:
        ent = sections["Entities"]
        npnt, ncur, nsur, nvol = (int(x) for x in ent[0].split())
        ent_phys = {}
        row = 1 + npnt + ncur
        for dim, cnt in ((2, nsur), (3, nvol)):
            for _ in range(cnt):
                parts = ent[row].split()
                row += 1
                tag = int(parts[0])
                nphys = int(parts[7])
                ent_phys[(dim, tag)] = [int(x) for x in parts[8:8 + nphys]]
        nd = sections["Nodes"]
        nblocks, nnodes = (int(x) for x in nd[0].split()[:2])
        coords = np.zeros((nnodes, 3))
        row = 1
        for _ in range(nblocks):
            _, _, _, nb = (int(x) for x in nd[row].split())
            tags = [int(nd[row + 1 + k]) for k in range(nb)]
            for k in range(nb):
                coords[tags[k] - 1] = [float(x) for x in nd[row + 1 + nb + k].split()[:3]]
            row += 1 + 2 * nb
        el = sections["Elements"]
        nblocks = int(el[0].split()[0])
        row = 1
        for _ in range(nblocks):
            edim, etag, etype, nb = (int(x) for x in el[row].split())
            names = [phys.get((edim, pt)) for pt in ent_phys.get((edim, etag), [])]
            for k in range(nb):
                parts = [int(x) for x in el[row + 1 + k].split()]
                for name in names:
                    if etype == 4 and name in vol_id:
                        tets.append(parts[1:5])
                        treg.append(vol_id[name])
                    elif etype == 2 and name in surf_id:
                        tris.append(parts[1:4])
                        freg.append(surf_id[name])
            row += 1 + nb
    else:
        nd = sections["Nodes"]
        nnodes = int(nd[0])
        coords = np.zeros((nnodes, 3))
        for ln in nd[1:1 + nnodes]:
            parts = ln.split()
            coords[int(parts[0]) - 1] = [float(x) for x in parts[1:4]]
        el = sections["Elements"]
        for ln in el[1:1 + int(el[0])]:
            parts = ln.split()
            etype, ntags = int(parts[1]), int(parts[2])
            name = phys.get((3 if etype == 4 else 2, int(parts[3])), None)
            if name is None:  # meshMaker.py writes sub-surface names verbatim in the tag column
                name = parts[3]
            nodes = [int(x) for x in parts[3 + ntags:]]
            if etype == 4 and name in vol_id:
                tets.append(nodes[:4])
                treg.append(vol_id[name])
            elif etype == 2 and name in surf_id:
                tris.append(nodes[:3])
                freg.append(surf_id[name])
    mesh = Mesh(node_p=np.ascontiguousarray(coords * scale),
                tet_ind=np.asarray(tets, dtype=np.int32).reshape(-1, 4) - 1,
                tet_reg=np.asarray(treg, dtype=np.int32),
                tri_ind=np.asarray(tris, dtype=np.int32).reshape(-1, 3) - 1,
                tri_reg=np.asarray(freg, dtype=np.int32),
                tri_dMs=np.zeros(len(tris)),
                vol_names=["__default__"] + list(vol_regions),
                surf_names=["__default__"] + list(surf_regions))
    mesh.tet_ind = np.ascontiguousarray(mesh.tet_ind)
    mesh.tri_ind = np.ascontiguousarray(mesh.tri_ind)
    return mesh
