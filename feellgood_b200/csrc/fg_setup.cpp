#include <cstdlib>
// fg_setup.cpp — once-per-mesh host preprocessing (see fg_setup.hpp).
#include "fg_setup.hpp"

#include <algorithm>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <numeric>

namespace fg
{
int sell_window()
    {
    static const int w = []
        {
        const char *e = getenv("FG_SELL_WINDOW");
        int v = e ? atoi(e) : SELL_WINDOW_DEFAULT;
        if (v < SELL_C) v = SELL_C;
        return (v / SELL_C) * SELL_C;
        }();
    return w;
    }

// Gauss tables, reference src/tetra.h:29-81.  The barycentric weights are evaluated with the same
// expression as the reference (1 - u - v - w) so that they round identically.
void tet_tables(int npi, double a[20], double pds[5])
    {
    if (npi == 1)
        {
        const double q = 1. / 4.;
        a[0] = 1. - q - q - q;
        a[1] = a[2] = a[3] = q;
        pds[0] = 1. / 6.;
        return;
        }
    const double A = 1. / 4., B = 1. / 6., C = 1. / 2., D = -2. / 15., E = 3. / 40.;
    const double u[5] = {A, B, B, B, C}, v[5] = {A, B, B, C, B}, w[5] = {A, B, C, B, B};
    const double p[5] = {D, E, E, E, E};
    for (int g = 0; g < 5; g++)
        {
        a[0 * 5 + g] = 1. - u[g] - v[g] - w[g];
        a[1 * 5 + g] = u[g];
        a[2 * 5 + g] = v[g];
        a[3 * 5 + g] = w[g];
        pds[g] = p[g];
        }
    }

// reference src/triangle.h:21-65
void tri_tables(int npi, double a[12], double pds[4])
    {
    if (npi == 1)
        {
        a[0] = 1. - 1. / 3. - 1. / 3.;
        a[1] = a[2] = 1. / 3.;
        pds[0] = 1. / 2.;
        return;
        }
    const double u[4] = {1 / 3., 1 / 5., 3 / 5., 1 / 5.}, v[4] = {1 / 3., 1 / 5., 1 / 5., 3 / 5.};
    const double p[4] = {-27 / 96., 25 / 96., 25 / 96., 25 / 96.};
    for (int g = 0; g < 4; g++)
        {
        a[0 * 4 + g] = 1. - u[g] - v[g];
        a[1 * 4 + g] = u[g];
        a[2 * 4 + g] = v[g];
        pds[g] = p[g];
        }
    }

namespace
{
struct V3
    {
    double x, y, z;
    };
inline V3 sub(const double *a, const double *b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline V3 cross(const V3 &a, const V3 &b)
    { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// Tet ctor (src/tetra.h:140-163): orientate (src/tetra.cpp:410-424), Jacobian (:393-408),
// da = dadu * J^-1 with Eigen's cofactor inverse.  Returns false on a singular element.
bool tet_geometry(const double *P, int ind[4], double da[12], double &detJ)
    {
    const double *p0 = P + 3 * (size_t)ind[0];
    V3 e1 = sub(P + 3 * (size_t)ind[1], p0), e2 = sub(P + 3 * (size_t)ind[2], p0),
       e3 = sub(P + 3 * (size_t)ind[3], p0);
    const double mixed = dot(e1, cross(e2, e3));
    if (std::fabs(mixed) < 1e-40) return false;
    if (mixed < 0.0)
        {
        std::swap(ind[2], ind[3]);
        std::swap(e2, e3);
        }
    // J = [e1 e2 e3] (columns)
    const double J[3][3] = {{e1.x, e2.x, e3.x}, {e1.y, e2.y, e3.y}, {e1.z, e2.z, e3.z}};
    detJ = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
           - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
           + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    double cof[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            cof[i][j] = J[i1][j1] * J[i2][j2] - J[i1][j2] * J[i2][j1];
            }
    const double det = cof[0][0] * J[0][0] + cof[0][1] * J[0][1] + cof[0][2] * J[0][2];
    const double inv = 1.0 / det;
    double Ji[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Ji[r][c] = cof[c][r] * inv;
    for (int d = 0; d < 3; d++)
        {
        da[3 * 1 + d] = Ji[0][d];
        da[3 * 2 + d] = Ji[1][d];
        da[3 * 3 + d] = Ji[2][d];
        da[3 * 0 + d] = (-1. * Ji[0][d] + -1. * Ji[1][d]) + -1. * Ji[2][d];
        }
    return true;
    }
}  // namespace

int host_setup(const fg_mesh &m, const fg_params &prm, int n_owned, HostSetup &h, std::string &err)
    {
    char buf[256];
    if (m.NOD <= 0 || m.NT < 0 || m.NF < 0 || !m.node_p || (m.NT > 0 && (!m.tet_ind || !m.tet_reg))
        || (m.NF > 0 && (!m.tri_ind || !m.tri_reg || !m.tri_dMs)))
        {
        err = "fg_create: null or empty mesh arrays";
        return FG_ERR_INVALID;
        }
    if ((prm.npi_tet != 5 && prm.npi_tet != 1) || (prm.npi_tri != 4 && prm.npi_tri != 1)
        || prm.nreg_tet <= 0 || !prm.prm_tet || (prm.nreg_tri > 0 && !prm.prm_tri))
        {
        err = "fg_create: npi_tet must be 5|1, npi_tri 4|1, and region tables non-empty";
        return FG_ERR_INVALID;
        }
    h.NOD = m.NOD;
    h.node_p.assign(m.node_p, m.node_p + 3 * (size_t)m.NOD);
    if (n_owned < 0 || n_owned > m.NOD) n_owned = m.NOD;
    h.n_owned = n_owned;
    h.NT = m.NT;
    h.NF = m.NF;
    h.npi_tet = prm.npi_tet;
    h.npi_tri = prm.npi_tri;
    const int NOD = m.NOD, NT = m.NT, NF = m.NF;

    // ---- tets: connectivity checks, orientation, geometry -----------------------------------
    h.tet_ind.assign(m.tet_ind, m.tet_ind + 4 * (size_t)NT);
    h.tet_reg.assign(m.tet_reg, m.tet_reg + (size_t)NT);
    h.tet_da.resize(12 * (size_t)NT);
    h.tet_detJ.resize((size_t)NT);
    int bad = -1, bad_kind = 0;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < NT; t++)
        {
        int *ind = &h.tet_ind[4 * (size_t)t];
        bool ok = h.tet_reg[t] >= 0 && h.tet_reg[t] < prm.nreg_tet;
        for (int i = 0; i < 4; i++) ok = ok && ind[i] >= 0 && ind[i] < NOD;
        if (!ok)
            {
#pragma omp critical
                { bad = t; bad_kind = 1; }
            continue;
            }
        if (!tet_geometry(m.node_p, ind, &h.tet_da[12 * (size_t)t], h.tet_detJ[t]))
            {
#pragma omp critical
                { if (bad_kind == 0) { bad = t; bad_kind = 2; } }
            }
        }
    if (bad >= 0)
        {
        snprintf(buf, sizeof buf, "fg_create: tetrahedron %d %s", bad,
                 bad_kind == 1 ? "has an index out of range" : "is singular (Tet::orientate)");
        err = buf;
        return FG_ERR_MESH;
        }

    // ---- magnetic masks (src/mesh.h:115-131, :331-336) ---------------------------------------
    h.magNode.assign((size_t)NOD, 0);
    h.tet_to_mag.assign((size_t)NT, -1);
    h.magTet.clear();
    for (int t = 0; t < NT; t++)
        if (prm.prm_tet[h.tet_reg[t]].Ms > 0)
            {
            h.tet_to_mag[t] = (int)h.magTet.size();
            h.magTet.push_back(t);
            for (int i = 0; i < 4; i++) h.magNode[h.tet_ind[4 * (size_t)t + i]] = 1;
            }
    const int NTm = (int)h.magTet.size();

    // ---- triangles (src/triangle.h:113-127,227) ------------------------------------------------
    h.tri_ind.assign(m.tri_ind, m.tri_ind + 3 * (size_t)NF);
    h.tri_reg.assign(m.tri_reg, m.tri_reg + (size_t)NF);
    h.tri_dMs.assign(m.tri_dMs, m.tri_dMs + (size_t)NF);
    h.tri_surf.resize((size_t)NF);
    h.tri_nrm.resize(3 * (size_t)NF);
    h.magTri.clear();
    h.actTri.clear();
    for (int f = 0; f < NF; f++)
        {
        const int *ind = &h.tri_ind[3 * (size_t)f];
        bool ok = h.tri_reg[f] >= 0 && h.tri_reg[f] < prm.nreg_tri;
        for (int i = 0; i < 3; i++) ok = ok && ind[i] >= 0 && ind[i] < NOD;
        if (!ok)
            {
            snprintf(buf, sizeof buf, "fg_create: triangle %d has an index out of range", f);
            err = buf;
            return FG_ERR_MESH;
            }
        const double *p0 = m.node_p + 3 * (size_t)ind[0];
        const V3 nrm = cross(sub(m.node_p + 3 * (size_t)ind[1], p0),
                             sub(m.node_p + 3 * (size_t)ind[2], p0));
        h.tri_surf[f] = 0.5 * std::sqrt(dot(nrm, nrm));
            {  // Tri::calc_norm, src/triangle.h:200-206 (Eigen normalize: untouched when the norm is 0)
            const double z = dot(nrm, nrm), sq = z > 0 ? std::sqrt(z) : 1.0;
            h.tri_nrm[3 * (size_t)f] = nrm.x / sq;
            h.tri_nrm[3 * (size_t)f + 1] = nrm.y / sq;
            h.tri_nrm[3 * (size_t)f + 2] = nrm.z / sq;
            }
        const bool mag = h.magNode[ind[0]] && h.magNode[ind[1]] && h.magNode[ind[2]];
        const fg_tri_prm &tp = prm.prm_tri[h.tri_reg[f]];
        if (mag && !tp.suppress_charges)
            {
            h.magTri.push_back(f);
            if (tp.Ks != 0) h.actTri.push_back(f);
            }
        }

    // ---- node adjacency over ALL tets (src/mesh.h:100-114: sorted unique edge list) ----------
    std::vector<int64_t> aptr((size_t)NOD + 1, 0);
    for (int t = 0; t < NT; t++)
        for (int i = 0; i < 4; i++) aptr[(size_t)h.tet_ind[4 * (size_t)t + i] + 1] += 3;
    for (int a = 0; a < NOD; a++) aptr[a + 1] += aptr[a];
    std::vector<int> adj((size_t)aptr[NOD]);
        {
        std::vector<int64_t> fill(aptr.begin(), aptr.end() - 1);
        for (int t = 0; t < NT; t++)
            {
            const int *ind = &h.tet_ind[4 * (size_t)t];
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++)
                    if (i != j) adj[(size_t)fill[ind[i]]++] = ind[j];
            }
        }
    // sort + unique each row; count edges; keep magnetic neighbours (+ self) for the pattern
    // (src/solver.h:75-104 with the filter of src/linear_algebra.h:43)
    std::vector<int> deg_all((size_t)NOD), deg_mag((size_t)NOD);
    h.extra_edges.clear();
#pragma omp parallel
        {
        std::vector<int> extra;  // edges with a non-magnetic end: in msh.edges, never in the pattern
#pragma omp for schedule(dynamic, 1024) nowait
        for (int a = 0; a < NOD; a++)
            {
            int *b = adj.data() + aptr[a], *e = adj.data() + aptr[a + 1];
            std::sort(b, e);
            e = std::unique(b, e);
            deg_all[a] = (int)(e - b);
            for (int *q = b; q < e; ++q)
                if (*q > a && !(h.magNode[a] && h.magNode[*q]))
                    {
                    extra.push_back(a);
                    extra.push_back(*q);
                    }
            int k = 0;
            if (h.magNode[a])
                for (int *q = b; q < e; ++q)
                    if (h.magNode[*q]) b[k++] = *q;
            deg_mag[a] = k;
            }
#pragma omp critical
        h.extra_edges.insert(h.extra_edges.end(), extra.begin(), extra.end());
        }
    h.n_edges = 0;
    h.n_edges_mag = 0;
    for (int a = 0; a < NOD; a++)
        {
        h.n_edges += deg_all[a];
        h.n_edges_mag += deg_mag[a];
        }
    h.n_edges /= 2;
    h.n_edges_mag /= 2;
    h.nptr.assign((size_t)NOD + 1, 0);
    for (int a = 0; a < NOD; a++)
        {
        const int64_t nx = (int64_t)h.nptr[a] + deg_mag[a] + 1;
        if (4 * nx > INT32_MAX)
            {
            err = "fg_create: more than 2^31 matrix entries on one device; partition the mesh";
            return FG_ERR_INVALID;
            }
        h.nptr[a + 1] = (int)nx;
        }
    h.ncol.resize((size_t)h.nptr[NOD]);
#pragma omp parallel for schedule(static)
    for (int a = 0; a < NOD; a++)
        {
        const int *b = adj.data() + aptr[a];
        int *o = &h.ncol[(size_t)h.nptr[a]];
        int k = 0, j = 0;
        while (j < deg_mag[a] && b[j] < a) o[k++] = b[j++];
        o[k++] = a;
        while (j < deg_mag[a]) o[k++] = b[j++];
        }
    std::vector<int>().swap(adj);

    // ---- device ordering: window-local stable sort by descending block count, slices of 32 -----
    const int NOW = h.n_owned;   // rows exist for the owned nodes only; ghosts follow the padded rows
    const int NODp = ((NOW + SELL_C - 1) / SELL_C) * SELL_C;
    h.NODp = NODp;
    h.NODt = NODp + (NOD - NOW);
    h.nslice = NODp / SELL_C;
    h.perm.assign((size_t)h.NODt, -1);
    h.iperm.assign((size_t)NOD, -1);
    for (int a = NOW; a < NOD; a++)
        {
        h.perm[(size_t)NODp + (a - NOW)] = a;
        h.iperm[a] = NODp + (a - NOW);
        }
    // Default: the caller's (reference's) node order, sorted inside windows.  FG_ORDER=rcb builds spatially
    // compact blocks of GATHER_BLOCK rows first (recursive coordinate bisection of the owned nodes: the range
    // is split at a multiple of the block size along the longest axis of its bounding box), then sorts inside
    // each block by descending pair count; the persistent kernel then stages the gathered images of a block
    // in shared memory (fg_solve_pk.cuh).  Measured on the 20 M-tet film (profiles/r02e_*): window order +
    // 16-bit global offsets 223 us per product, compact blocks through L1 with 32-bit columns 227 us, compact
    // blocks staged in shared memory (single-buffered) 278 us -- the staged variant is latency-bound until
    // its staging is pipelined across blocks, so it is not the default.
        {
        const char *eo = getenv("FG_ORDER");
        h.order_kind = (eo && !strcmp(eo, "rcb")) ? 1 : 0;
        }
    std::vector<int> base((size_t)NOW);  // position p of the pre-order holds node base[p]
    std::iota(base.begin(), base.end(), 0);
    int win = SELL_WINDOW;
    if (h.order_kind == 1)
        {
        win = GATHER_BLOCK;
        struct Range { int lo, hi; };
        std::vector<Range> todo;
        todo.push_back({0, NOW});
        const double *P = h.node_p.data();
        // breadth-first: the ranges of one level are independent
        while (!todo.empty())
            {
            std::vector<Range> next((size_t)2 * todo.size());
            std::vector<char> used((size_t)2 * todo.size(), 0);
#pragma omp parallel for schedule(dynamic, 1)
            for (long long q = 0; q < (long long)todo.size(); q++)
                {
                const int lo = todo[(size_t)q].lo, hi = todo[(size_t)q].hi;
                if (hi - lo <= GATHER_BLOCK) continue;
                double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
                for (int p = lo; p < hi; p++)
                    for (int c = 0; c < 3; c++)
                        {
                        const double v = P[3 * (size_t)base[(size_t)p] + c];
                        mn[c] = std::min(mn[c], v);
                        mx[c] = std::max(mx[c], v);
                        }
                int ax = 0;
                if (mx[1] - mn[1] > mx[ax] - mn[ax]) ax = 1;
                if (mx[2] - mn[2] > mx[ax] - mn[ax]) ax = 2;
                const int nb = (hi - lo + GATHER_BLOCK - 1) / GATHER_BLOCK;
                const int mid = lo + (nb / 2) * GATHER_BLOCK;  // left part: whole blocks
                std::nth_element(base.begin() + lo, base.begin() + mid, base.begin() + hi, [&](int x, int y)
                    {
                    const double vx = P[3 * (size_t)x + ax], vy = P[3 * (size_t)y + ax];
                    return vx < vy || (vx == vy && x < y);
                    });
                next[(size_t)2 * q] = {lo, mid};
                next[(size_t)2 * q + 1] = {mid, hi};
                used[(size_t)2 * q] = used[(size_t)2 * q + 1] = 1;
                }
            todo.clear();
            for (size_t q = 0; q < next.size(); q++)
                if (used[q]) todo.push_back(next[q]);
            }
        }
        {
        const int nwin = (NOW + win - 1) / win;
#pragma omp parallel for schedule(static)
        for (int wdx = 0; wdx < nwin; wdx++)
            {
            const int b = wdx * win, e = std::min(NOW, b + win);
            int *q = &h.perm[(size_t)b];
            std::copy(base.begin() + b, base.begin() + e, q);
            if (h.order_kind == 1) std::sort(q, q + (e - b));  // deterministic start: ascending node number
            std::stable_sort(q, q + (e - b), [&](int x, int y)
                { return h.nptr[x + 1] - h.nptr[x] > h.nptr[y + 1] - h.nptr[y]; });
            for (int r = b; r < e; r++) h.iperm[(size_t)h.perm[(size_t)r]] = r;
            }
        }
    std::vector<int>().swap(base);
    // magnetic tets follow the node order (sorted by their smallest device row) so that
    // neighbouring threads of the element kernel gather neighbouring node records
        {
        std::vector<int> key((size_t)NTm);
        for (int tm = 0; tm < NTm; tm++)
            {
            const int *ind = &h.tet_ind[4 * (size_t)h.magTet[tm]];
            key[tm] = std::min(std::min(h.iperm[ind[0]], h.iperm[ind[1]]),
                               std::min(h.iperm[ind[2]], h.iperm[ind[3]]));
            }
        std::vector<int> order((size_t)NTm);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
        std::vector<int> mt((size_t)NTm);
        for (int k = 0; k < NTm; k++) mt[k] = h.magTet[order[k]];
        h.magTet.swap(mt);
        for (int tm = 0; tm < NTm; tm++) h.tet_to_mag[h.magTet[tm]] = tm;
        h.tet_dev_ind.resize(4 * (size_t)NTm);
        for (int tm = 0; tm < NTm; tm++)
            for (int i = 0; i < 4; i++)
                h.tet_dev_ind[4 * (size_t)tm + i] = h.iperm[h.tet_ind[4 * (size_t)h.magTet[tm] + i]];
        }

    // ---- incidence lists node -> (magnetic tet, local node) ------------------------------------
    h.inc_ptr.assign((size_t)NOD + 1, 0);
    for (int tm = 0; tm < NTm; tm++)
        for (int i = 0; i < 4; i++) h.inc_ptr[(size_t)h.tet_ind[4 * (size_t)h.magTet[tm] + i] + 1]++;
    for (int a = 0; a < NOD; a++) h.inc_ptr[a + 1] += h.inc_ptr[a];
    h.inc.resize(4 * (size_t)NTm);
        {
        std::vector<int> fill(h.inc_ptr.begin(), h.inc_ptr.end() - 1);
        for (int tm = 0; tm < NTm; tm++)  // ascending tm inside each node's list
            for (int i = 0; i < 4; i++)
                h.inc[(size_t)fill[h.tet_ind[4 * (size_t)h.magTet[tm] + i]]++] = 4 * tm + i;
        }
    const int NFa = (int)h.actTri.size();
    h.inc_tri_ptr.assign((size_t)NOD + 1, 0);
    for (int fa = 0; fa < NFa; fa++)
        for (int i = 0; i < 3; i++) h.inc_tri_ptr[(size_t)h.tri_ind[3 * (size_t)h.actTri[fa] + i] + 1]++;
    for (int a = 0; a < NOD; a++) h.inc_tri_ptr[a + 1] += h.inc_tri_ptr[a];
    h.inc_tri.resize(3 * (size_t)NFa);
        {
        std::vector<int> fill(h.inc_tri_ptr.begin(), h.inc_tri_ptr.end() - 1);
        for (int fa = 0; fa < NFa; fa++)
            for (int i = 0; i < 3; i++)
                h.inc_tri[(size_t)fill[h.tri_ind[3 * (size_t)h.actTri[fa] + i]]++] = 3 * fa + i;
        }

    // ---- per-mesh constants of the stiffness (src/tetra.cpp:108-140) -------------
    // E_ab(step) = prefactor*s_dt * S_ab + delta_ab * Malpha_a(step);   a_w lumped into Aw_a.
    double ta[20], tp[5];
    tet_tables(prm.npi_tet, ta, tp);
    const int npi = prm.npi_tet;
    h.S.assign(h.ncol.size(), 0.0);
    h.Aw.assign((size_t)NOD, 0.0);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int a = 0; a < NOD; a++)
        {
        const int rb = h.nptr[a], re = h.nptr[a + 1];
        double aw = 0.0;
        for (int q = h.inc_ptr[a]; q < h.inc_ptr[a + 1]; q++)
            {
            const int tm = h.inc[q] >> 2, i = h.inc[q] & 3, t = h.magTet[tm];
            const fg_tet_prm &rp = prm.prm_tet[h.tet_reg[t]];
            const double Abis = 2.0 * rp.A / (FG_MU0_HOST * rp.Ms);
            const double detJ = h.tet_detJ[t];
            double wsum = 0.0, a_w = 0.0;
            for (int g = 0; g < npi; g++)
                {
                const double w = detJ * tp[g];
                wsum += w;
                a_w += ta[i * npi + g] * w;
                }
            aw += a_w;
            const double *da = &h.tet_da[12 * (size_t)t];
            const int *ind = &h.tet_ind[4 * (size_t)t];
            for (int j = 0; j < 4; j++)
                {
                const double dd = da[3 * i] * da[3 * j] + da[3 * i + 1] * da[3 * j + 1]
                                  + da[3 * i + 2] * da[3 * j + 2];
                const int pos = (int)(std::lower_bound(&h.ncol[rb], &h.ncol[rb] + (re - rb), ind[j])
                                      - &h.ncol[0]);
                h.S[pos] += dd * (Abis * wsum);
                }
            }
        h.Aw[a] = aw;
        }

    // ---- SELL-32 images of the pattern, of S and of the incidence lists ------------------------
        {
        const int ns = h.nslice;
        h.sdeg.assign((size_t)NODp, 0);
        h.sptr.assign((size_t)ns + 1, 0);
        h.iptr.assign((size_t)ns + 1, 0);
        h.itptr.assign((size_t)ns + 1, 0);
        for (int s = 0; s < ns; s++)
            {
            int wk = 0, wi = 0, wt = 0;
            for (int l = 0; l < SELL_C; l++)
                {
                const int a = h.perm[(size_t)s * SELL_C + l];
                if (a < 0) continue;
                h.sdeg[(size_t)s * SELL_C + l] = h.nptr[a + 1] - h.nptr[a];
                wk = std::max(wk, h.nptr[a + 1] - h.nptr[a]);
                wi = std::max(wi, h.inc_ptr[a + 1] - h.inc_ptr[a]);
                wt = std::max(wt, h.inc_tri_ptr[a + 1] - h.inc_tri_ptr[a]);
                }
            const int64_t nk = (int64_t)h.sptr[s] + wk;
            if (nk * SELL_C * 4 > INT32_MAX)
                {
                err = "fg_create: more than 2^31 matrix entries on one device; partition the mesh";
                return FG_ERR_INVALID;
                }
            h.sptr[s + 1] = (int)nk;
            h.iptr[s + 1] = h.iptr[s] + wi;
            h.itptr[s + 1] = h.itptr[s] + wt;
            }
        h.scol.assign((size_t)h.sptr[ns] * SELL_C, 0);
        h.sS.assign((size_t)h.sptr[ns] * SELL_C, 0.0);
        h.sinc.assign((size_t)h.iptr[ns] * SELL_C, -1);
        h.sinct.assign((size_t)h.itptr[ns] * SELL_C, -1);
        h.tet_slot.assign(4 * (size_t)NTm, -1);
#pragma omp parallel for schedule(static)
        for (int s = 0; s < ns; s++)
            for (int l = 0; l < SELL_C; l++)
                {
                const int r = s * SELL_C + l;
                const int a = h.perm[(size_t)r];
                const int wk = h.sptr[s + 1] - h.sptr[s];
                const int deg = a < 0 ? 0 : h.nptr[a + 1] - h.nptr[a];
                for (int j = 0; j < wk; j++)
                    {
                    const size_t pos = ((size_t)h.sptr[s] + j) * SELL_C + l;
                    if (j < deg)
                        {
                        h.scol[pos] = h.iperm[(size_t)h.ncol[(size_t)h.nptr[a] + j]];
                        h.sS[pos] = h.S[(size_t)h.nptr[a] + j];
                        }
                    else
                        h.scol[pos] = r;
                    }
                if (a < 0) continue;
                for (int q = h.inc_ptr[a]; q < h.inc_ptr[a + 1]; q++)
                    {
                    const size_t slot = ((size_t)h.iptr[s] + (q - h.inc_ptr[a])) * SELL_C + l;
                    h.sinc[slot] = h.inc[q];
                    h.tet_slot[(size_t)h.inc[q]] = (int)slot;  // each (tet, local node) occurs once
                    }
                for (int q = h.inc_tri_ptr[a]; q < h.inc_tri_ptr[a + 1]; q++)
                    h.sinct[((size_t)h.itptr[s] + (q - h.inc_tri_ptr[a])) * SELL_C + l] = h.inc_tri[q];
                }
        }

    // ---- gather blocks: halo lists and 16-bit local column indices (fg_setup.hpp) ---------------
    h.nblock = 0;
    h.stage_cap = 0;
    if (h.order_kind == 1)
        {
        const int SPB = GATHER_BLOCK / SELL_C;  // slices per block
        const int nb = (h.nslice + SPB - 1) / SPB;
        h.nblock = nb;
        h.bptr.assign((size_t)nb + 1, 0);
        h.bghost.assign((size_t)nb, 0);
        std::vector<std::vector<int>> halo((size_t)nb);
#pragma omp parallel for schedule(dynamic, 64)
        for (int b = 0; b < nb; b++)
            {
            const int r0 = b * GATHER_BLOCK, r1 = r0 + GATHER_BLOCK;
            const int s0 = b * SPB, s1 = std::min(h.nslice, s0 + SPB);
            std::vector<int> &hl = halo[(size_t)b];
            for (size_t pos = (size_t)h.sptr[s0] * SELL_C; pos < (size_t)h.sptr[s1] * SELL_C; pos++)
                {
                const int c = h.scol[pos];
                if (c < r0 || c >= r1) hl.push_back(c);
                }
            std::sort(hl.begin(), hl.end());
            hl.erase(std::unique(hl.begin(), hl.end()), hl.end());
            if (!hl.empty() && hl.back() >= NODp) h.bghost[(size_t)b] = 1;
            }
        int worst = 0;
        for (int b = 0; b < nb; b++)
            {
            h.bptr[(size_t)b + 1] = h.bptr[(size_t)b] + (int)halo[(size_t)b].size();
            worst = std::max(worst, (int)halo[(size_t)b].size());
            }
        if (GATHER_BLOCK + worst <= 65535)
            {
            h.stage_cap = GATHER_BLOCK + worst;
            h.bhalo.resize((size_t)h.bptr[(size_t)nb]);
            h.lcol.assign(h.scol.size(), 0);
#pragma omp parallel for schedule(dynamic, 64)
            for (int b = 0; b < nb; b++)
                {
                const std::vector<int> &hl = halo[(size_t)b];
                std::copy(hl.begin(), hl.end(), h.bhalo.begin() + h.bptr[(size_t)b]);
                const int r0 = b * GATHER_BLOCK, r1 = r0 + GATHER_BLOCK;
                const int s0 = b * SPB, s1 = std::min(h.nslice, s0 + SPB);
                for (size_t pos = (size_t)h.sptr[s0] * SELL_C; pos < (size_t)h.sptr[s1] * SELL_C; pos++)
                    {
                    const int c = h.scol[pos];
                    if (c >= r0 && c < r1)
                        h.lcol[pos] = (unsigned short)(c - r0);
                    else
                        h.lcol[pos] = (unsigned short)(GATHER_BLOCK
                                                       + (int)(std::lower_bound(hl.begin(), hl.end(), c) - hl.begin()));
                    }
                }
            }
        }

    // ---- masked dofs (src/linear_algebra.h:55-63) ---------------------------------------------
    h.lvd.clear();
    for (int a = 0; a < NOD; a++)
        if (!h.magNode[a])
            {
            h.lvd.push_back(2 * a);
            h.lvd.push_back(2 * a + 1);
            }
    return FG_OK;
    }

}  // namespace fg
