"""On-disk formats of the reference around the hot path.

    savesol / read_sol    mesh::savesol (src/save.cpp:163-186), mesh::readSol (src/read.cpp:198-246):
                          the .sol magnetisation snapshot, "## time:" / "## columns:" metadata lines
                          (src/tags.h:11-18), one row `idx mx my mz phi` per node in the ORIGINAL
                          node numbering of the mesh file (node_index of mesh::sortNodes).
    read_evol             the .evol table written by Fem::saver (feellgood_b200.fem.Fem.write_evol).

The gmsh .msh reader and mesh::controlTriangles live in meshgen.py (read_msh, control_triangles).
"""
from __future__ import annotations

import numpy as np

TAG_TIME = "## time:"            # tags::sol::time
TAG_COLUMNS = "## columns:"      # tags::sol::columns / tags::evol::columns
SOL_COLUMNS = "idx\tmx\tmy\tmz\tphi"   # tags::sol::defaultColumnsTitle


def sol_metadata(t, extra=()):
    """Settings::solMetadata (src/settings.cpp:303-312) without the host-specific common lines."""
    lines = list(extra) + ["%s %e" % (TAG_TIME, t), "%s %s" % (TAG_COLUMNS, SOL_COLUMNS)]
    return "\n".join(lines) + "\n"


def savesol(path, u, phi, node_index=None, t=0.0, precision=7, metadata=None):
    """mesh::savesol: row i holds the node whose index in the mesh file was i, i.e. sorted node
    node_index[i]; std::scientific with `precision` digits (Settings::precision = 7)."""
    u, phi = np.asarray(u, dtype=np.float64), np.asarray(phi, dtype=np.float64)
    n = u.shape[0]
    j = np.arange(n) if node_index is None else np.asarray(node_index)
    fmt = "%%.%de" % precision
    with open(path, "w") as f:
        f.write(sol_metadata(t) if metadata is None else metadata)
        for i in range(n):
            k = j[i]
            f.write("%d\t%s\t%s\t%s\t%s\n" % (i, fmt % u[k, 0], fmt % u[k, 1], fmt % u[k, 2], fmt % phi[k]))


def read_sol(path, n_nodes, node_index=None):
    """mesh::readSol: returns (t, u, phi) in the SORTED node numbering; raises like the reference
    exits (missing "## time:" tag, node index mismatch)."""
    t, have_t = 0.0, False
    rows = []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("#") or line.strip() == "":
                k = line.find(TAG_TIME)
                if k >= 0:
                    t = float(line[k + len(TAG_TIME):])
                    have_t = True
                continue
            rows.append(line.split())
    if not have_t:
        raise ValueError("error: no ## time: tag in input .sol file %s" % path)
    if len(rows) < n_nodes:
        raise ValueError("error: .sol file %s has %d rows for %d nodes" % (path, len(rows), n_nodes))
    j = np.arange(n_nodes) if node_index is None else np.asarray(node_index)
    u, phi = np.zeros((n_nodes, 3)), np.zeros(n_nodes)
    for i in range(n_nodes):
        r = rows[i]
        if int(r[0]) != i:
            raise ValueError("error: mesh node index mismatch between mesh and input .sol file")
        u[j[i]] = [float(r[1]), float(r[2]), float(r[3])]
        phi[j[i]] = float(r[4])
    return t, u, phi


def read_evol(path):
    """(columns, rows) of an .evol file: metadata lines start with '#', the last '## columns:' line
    names the tab-separated columns."""
    cols, rows = None, []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("#"):
                if line.startswith(TAG_COLUMNS):
                    cols = line[len(TAG_COLUMNS):].strip().split("\t")
                continue
            if line.strip():
                rows.append([float(x) for x in line.split()])
    return cols, np.array(rows)
