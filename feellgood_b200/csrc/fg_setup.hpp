// fg_setup.hpp — once-per-mesh host preprocessing for the B200 LLG path (pure C++17, OpenMP).
//
// Builds, from what Mesh::mesh hands to LinAlgebra, everything the per-step kernels read:
// oriented connectivity and geometry tables (Tet ctor), magnetic masks (mesh ctor), the node-level
// sparsity of K (solver<2>::build_shape with the LinAlgebra edge filter), the node->element
// incidence lists that make the per-step assembly a race-free gather, and the two per-mesh
// constants of the factorised stiffness: S_ab = sum_T Abis_T vol_T (grad a_a . grad a_b) and the
// lumped mass Aw_a.  References are cited at each step in fg_setup.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/feellgood_b200.h"

namespace fg
{
constexpr double FG_MU0_HOST = 1.25663706127e-6;  // src/config.h.in:31

struct HostSetup
    {
    int NOD = 0, NT = 0, NF = 0;
    int npi_tet = 5, npi_tri = 4;
    std::vector<double> node_p;    // NOD x 3 (Gauss points of the charges, Tri::potential)
    // all tets, in caller order (taps + energies)
    std::vector<int> tet_ind;      // NT x 4, oriented
    std::vector<double> tet_da;    // NT x 12
    std::vector<double> tet_detJ;  // NT
    std::vector<int> tet_reg;      // NT
    std::vector<int> magTet;       // indices of magnetic tets (Ms > 0)
    std::vector<int> tet_to_mag;   // NT: compact index or -1
    std::vector<unsigned char> magNode;  // NOD
    // triangles
    std::vector<int> tri_ind;      // NF x 3
    std::vector<int> tri_reg;      // NF
    std::vector<double> tri_dMs, tri_surf;  // NF
    std::vector<double> tri_nrm;   // NF x 3 unit normals (demag energy of the surface, src/triangle.cpp:80-85)
    std::vector<int> magTri;       // magnetic && !suppress_charges (rhs contributors)
    std::vector<int> actTri;       // magTri with Ks != 0 : the only ones whose Lp reaches the rhs
    // node-level pattern of K
    std::vector<int> nptr, ncol;   // NOD+1, nnzb
    std::vector<double> S;         // nnzb
    std::vector<double> Aw;        // NOD
    long long n_edges = 0, n_edges_mag = 0;
    std::vector<int> extra_edges;  // 2 x (n_edges - n_edges_mag): mesh edges with a non-magnetic end (max_angle)
    // incidences: entry = 4*tm + i (tm = compact magnetic tet index) / 3*fa + i
    std::vector<int> inc_ptr, inc;          // NOD+1, 4*n_magTet
    std::vector<int> inc_tri_ptr, inc_tri;  // NOD+1, 3*n_actTri
    std::vector<int> lvd;          // masked dofs (src/linear_algebra.h:55-63)

    // ---- device ordering ------------------------------------------------------
    // Rows are permuted inside windows of SELL_WINDOW nodes by descending block count (stable, so
    // locality survives) and cut into slices of 32 rows: SELL-32-sigma with 2x2 blocks.  All
    // device arrays indexed by node use the device row number.
    int n_owned = 0;               // nodes [0, n_owned) have rows here; the rest are ghosts (multi-GPU)
    int NODp = 0;                  // n_owned rounded up to a multiple of 32 (pad rows: perm = -1)
    int NODt = 0;                  // NODp + ghosts: length of every device node array
    std::vector<int> perm, iperm;  // device row -> node | node -> device row
    int nslice = 0;
    std::vector<int> sptr;         // nslice+1 : first block-column of each slice
    std::vector<int> sdeg;         // NODp : blocks in the row (0 for pad rows)
    std::vector<int> scol;         // sptr[nslice]*32 : device row of the column node (pad: own row)
    std::vector<double> sS;        // sptr[nslice]*32 : S in SELL order (pad: 0)
    std::vector<int> iptr, sinc;   // SELL incidence lists: nslice+1 | iptr[nslice]*32 (record index, -1 pad)
    std::vector<int> itptr, sinct; // same for the active triangles
    std::vector<int> tet_dev_ind;  // 4*n_magTet : device rows of the magnetic tets, device tet order
    std::vector<int> tet_slot;     // 4*n_magTet : where (tet, local node) writes its record = its
                                   // position in the node's SELL incidence list (-1: no row)
    // ---- gather blocks of the matrix-free SpMV ----------------------------------
    // With the compact ordering (order_kind == 1) the rows are cut into blocks of GATHER_BLOCK = 256
    // consecutive device rows (8 slices) that are spatially compact (recursive coordinate bisection), so that
    // the images a block gathers are its own 256 rows plus a small halo.  A thread group stages them once
    // in shared memory and the stored pairs address them with 16-bit LOCAL indices, whatever the distance
    // of the column in the global numbering (ghost columns of a partition included):
    //   local index < 256            : row 256 b + index of the block itself
    //   local index = 256 + k        : row bhalo[bptr[b] + k]
    int order_kind = 0;            // 0 window-local sort of the caller's order | 1 compact blocks (RCB)
    int nblock = 0;                // ceil(nslice / 8)
    int stage_cap = 0;             // 256 + largest halo of a block (0: no block tables)
    std::vector<int> bptr;         // nblock+1
    std::vector<int> bhalo;        // bptr[nblock] device rows, ascending inside a block
    std::vector<unsigned short> lcol;   // sptr[nslice]*32 : local index of the column node (pad: own row)
    std::vector<unsigned char> bghost;  // nblock : 1 = the block's halo has a ghost row (>= NODp)
    };
constexpr int GATHER_BLOCK = 256;
constexpr int SELL_C = 32;
constexpr int SELL_WINDOW_DEFAULT = 1024;
// window of the row permutation (multiple of 32); FG_SELL_WINDOW overrides the default (experiments)
int sell_window();
#define SELL_WINDOW (fg::sell_window())

// returns FG_OK or FG_ERR_*; message in err
// n_owned < 0: all nodes are owned (single GPU)
int host_setup(const fg_mesh &mesh, const fg_params &prm, int n_owned, HostSetup &out, std::string &err);

// Gauss tables (src/tetra.h:29-81, src/triangle.h:21-65): a[i*npi+g], pds[g]
void tet_tables(int npi, double a[20], double pds[5]);
void tri_tables(int npi, double a[12], double pds[4]);

}  // namespace fg
