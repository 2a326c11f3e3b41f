// ut_oracle.cpp — the reference's unit-test strategy applied to the CPU oracle.
//
// Each case re-derives a formula in plain scalar C++ from its textbook definition (the same
// "ref code vs code to test" pattern as /root/reference/unit-tests, which is cited per case) on
// the reference's own fixtures: the unit tetrahedron (0,0,0),(1,0,0),(0,1,0),(0,0,1), the
// triangle (1,0,0),(0,1,0),(1,1,0), random unit vectors from std::mt19937(5489) through
// uniform_real_distribution (unit-tests/ut_tools.h:5-44, ut_config.h.in:8-24), and the
// reference's tolerance UT_TOL = 5e-16 (relaxed ×10 / ×1000 exactly where the reference does).
// Exit status 0 = all cases pass.  Built and run by tests/test_oracle_ut.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../../oracle/fg_oracle.h"

static const double UT_TOL = 5e-16;
static int n_fail = 0, n_check = 0;

// Boost.Test's tolerance is relative (strong check); small values are compared absolutely.
static void check_close(double a, double b, double tol, const char *what, int line)
    {
    n_check++;
    double d = std::fabs(a - b), m = std::max(std::fabs(a), std::fabs(b));
    bool ok = (d <= tol * m) || (d <= tol);
    if (!ok)
        {
        n_fail++;
        std::printf("FAIL line %d: %s: %.17g vs %.17g (diff %.3g)\n", line, what, a, b, d);
        }
    }
static void check_true(bool ok, const char *what, int line)
    {
    n_check++;
    if (!ok)
        {
        n_fail++;
        std::printf("FAIL line %d: %s\n", line, what);
        }
    }
#define CLOSE(a, b, tol) check_close((a), (b), (tol), #a " == " #b, __LINE__)
#define TRUE_(c) check_true((c), #c, __LINE__)

struct V3
    {
    double x[3];
    };
static V3 sph(double theta, double phi)  // ut_tools.h:37-41 rand_vec3d
    {
    const double si_t = std::sin(theta);
    return V3{{si_t * std::cos(phi), si_t * std::sin(phi), std::cos(theta)}};
    }
static V3 cyl(double theta, double z)  // ut_node.cpp:51-55 unit_vector
    {
    double r = std::sqrt(1 - z * z);
    return V3{{r * std::cos(theta), r * std::sin(theta), z}};
    }
static double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static const double unit_tet[12] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
static const double unit_tri[9] = {1, 0, 0, 0, 1, 0, 1, 1, 0};

struct TetFix
    {
    int ind[4] = {0, 1, 2, 3};
    double da[12], w[5], detJ;
    explicit TetFix(int npi = 5)
        {
        fgo_tet_orientate(unit_tet, ind);
        detJ = fgo_tet_setup(unit_tet, ind, npi, da, w);
        }
    };

// --- ut_time_int.cpp:19-37 timing_constructor -------------------------------------------------
static void t_timing_constructor()
    {
    std::mt19937 gen(5489u);
    std::uniform_real_distribution<> distrib(0.0, 1.0);
    double a = distrib(gen), b = distrib(gen);
    CLOSE(fgo_timing_dt0(std::min(a, b), std::max(a, b)), std::sqrt(a * b), UT_TOL);
    }

// --- ut_time_int.cpp:45-96 calc_alpha_eff -------------------------------------------------------
static void t_calc_alpha_eff()
    {
    std::mt19937 gen(5489u);
    std::uniform_real_distribution<> distrib(0.0, 1.0);
    for (int rep = 0; rep < 8; rep++)
        {
        double alpha = distrib(gen), X = distrib(gen);
        if (rep & 1) X = -X;           // the reference only draws X>0; also cover the h<=0 branch
        if (rep >= 4) X *= 1e12;       // ... and the saturated |h|>M branches
        double dt = fgo_timing_dt0(1e-14, 1e-9);
        double reduced_dt = FGO_GAMMA0 * dt, r = 0.1, M = 2. * alpha * r / reduced_dt, alfa;
        if (X > 0.)
            alfa = (X > M) ? alpha + reduced_dt / 2. * M : alpha + reduced_dt / 2. * X;
        else
            alfa = (X < -M) ? alpha / (1. + reduced_dt / (2. * alpha) * M)
                            : alpha / (1. - reduced_dt / (2. * alpha) * X);
        double uH[5] = {X, X, X, X, X}, out[5];
        fgo_calc_alpha_eff(5, dt, alpha, uH, out);
        for (int g = 0; g < 5; g++) CLOSE(out[g], alfa, UT_TOL);
        }
    }

// --- ut_tetra.cpp:21-101 Tet_inner_tables, :103-120 Tet_calc_vol --------------------------------
static void t_tet_inner_tables()
    {
    // a general (non unit) tetrahedron as well as the unit one
    std::mt19937 gen(5489u);
    std::uniform_real_distribution<> distrib(0.0, 1.0);
    for (int rep = 0; rep < 4; rep++)
        {
        double p[12];
        for (int k = 0; k < 12; k++) p[k] = (rep == 0) ? unit_tet[k] : unit_tet[k] + 0.3 * distrib(gen);
        int ind[4] = {0, 1, 2, 3};
        fgo_tet_orientate(p, ind);
        double da[12], w[5];
        double detJ_o = fgo_tet_setup(p, ind, 5, da, w);
        // hand formula: J = nod * dadu, cofactor inverse divided by det (ut_tools.h:73-104)
        double J[3][3];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) J[r][c] = p[3 * ind[c + 1] + r] - p[3 * ind[0] + r];
        double detJ = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
                      - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                      + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        double I[3][3] = {
            {(J[1][1] * J[2][2] - J[1][2] * J[2][1]) / detJ, (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / detJ,
             (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / detJ},
            {(J[1][2] * J[2][0] - J[1][0] * J[2][2]) / detJ, (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / detJ,
             (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / detJ},
            {(J[1][0] * J[2][1] - J[1][1] * J[2][0]) / detJ, (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / detJ,
             (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / detJ}};
        const double dadu[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        double err2 = 0;
        for (int i = 0; i < 4; i++)
            for (int d = 0; d < 3; d++)
                {
                double s = 0;
                for (int k = 0; k < 3; k++) s += dadu[i][k] * I[k][d];
                err2 += (s - da[3 * i + d]) * (s - da[3 * i + d]);
                }
        // the reference tests sqrt(sum sq diff) == 0 at UT_TOL on the unit tet; da entries are O(1)
        TRUE_(std::sqrt(err2) <= (rep == 0 ? UT_TOL : 1e-13));
        CLOSE(detJ_o, detJ, 10 * UT_TOL);
        const double *pds = fgo_tet_pds(5);
        for (int g = 0; g < 5; g++) CLOSE(w[g], detJ * pds[g], 10 * UT_TOL);
        TRUE_(detJ_o > 0);  // orientate() guarantees a positive Jacobian
        }
    TetFix t;
    double vol = 0;
    for (int g = 0; g < 5; g++) vol += t.w[g];
    CLOSE(vol, 1 / 6.0, UT_TOL);
    TetFix t1(1);
    CLOSE(t1.w[0], 1 / 6.0, UT_TOL);
    }

// --- ut_tet_lumping.cpp:94-197 tet_lumping (includes :17-92 tet_exchange_lumping) ----------------
static void t_tet_lumping()
    {
    std::mt19937 gen(5489u);
    std::uniform_real_distribution<> distrib(0.0, 1.0);
    TetFix t;
    double u[12];
    for (int i = 0; i < 4; i++)
        {
        V3 v = sph(M_PI * distrib(gen), 2 * M_PI * distrib(gen));
        for (int d = 0; d < 3; d++) u[3 * i + d] = v.x[d];
        }
    double a = distrib(gen), b = distrib(gen);
    double dtmin = std::min(a, b), dtmax = std::max(a, b);
    double dt = fgo_timing_dt0(dtmin, dtmax), s_dt = FGO_THETA * dt, TAUR = 100. * dtmax;
    double uH[5], alfa[5];
    double X = distrib(gen);
    for (int g = 0; g < 5; g++) uH[g] = X;
    fgo_calc_alpha_eff(5, dt, 0.5, uH, alfa);
    double A = distrib(gen), Js = 0.5 + distrib(gen), Abis = 2.0 * A / Js;
    // explicit per-Gauss-point accumulation of the 12x12 operator
    const double *sh = fgo_tet_a(5);
    double AE[12][12] = {{0}};
    double R = dt / TAUR * std::fabs(std::log(dt / TAUR));
    for (int g = 0; g < 5; g++)
        {
        double w = t.w[g];
        for (int ie = 0; ie < 4; ie++)
            {
            double ai = sh[ie * 5 + g];
            for (int blk = 0; blk < 3; blk++) AE[blk * 4 + ie][blk * 4 + ie] += alfa[g] * ai * w;
            AE[ie][8 + ie] += +u[3 * ie + 1] * ai * w;
            AE[ie][4 + ie] += -u[3 * ie + 2] * ai * w;
            AE[4 + ie][ie] += +u[3 * ie + 2] * ai * w;
            AE[4 + ie][8 + ie] += -u[3 * ie + 0] * ai * w;
            AE[8 + ie][4 + ie] += +u[3 * ie + 0] * ai * w;
            AE[8 + ie][ie] += -u[3 * ie + 1] * ai * w;
            for (int je = 0; je < 4; je++)
                {
                double DD = dot(t.da + 3 * ie, t.da + 3 * je);
                for (int blk = 0; blk < 3; blk++)
                    AE[blk * 4 + ie][blk * 4 + je] += s_dt * (1. + R) * 2 * A / Js * DD * w;
                }
            }
        }
    double AEo[144] = {0};
    fgo_tet_lumping(5, t.da, t.w, u, alfa, fgo_timing_prefactor(dt, dtmax) * s_dt * Abis, AEo);
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 12; j++) CLOSE(AEo[i * 12 + j], AE[i][j], 10 * UT_TOL);
    }

// --- ut_anisotropy.cpp:22-119 anisotropy_uniax, :121-251 anisotropy_cubic ------------------------
static void t_anisotropy()
    {
    std::mt19937 gen(5489u);
    std::uniform_real_distribution<> distrib(0.0, 1.0);
    const double *sh = fgo_tet_a(5);
    double u[12], v[12], U[15], V[15];
    for (int i = 0; i < 4; i++)
        {
        V3 a = sph(M_PI * distrib(gen), 2 * M_PI * distrib(gen));
        V3 b = sph(M_PI * distrib(gen), 2 * M_PI * distrib(gen));
        for (int d = 0; d < 3; d++) { u[3 * i + d] = a.x[d]; v[3 * i + d] = b.x[d]; }
        }
    for (int d = 0; d < 3; d++)
        for (int g = 0; g < 5; g++)
            {
            double su = 0, sv = 0;
            for (int i = 0; i < 4; i++) { su += u[3 * i + d] * sh[i * 5 + g]; sv += v[3 * i + d] * sh[i * 5 + g]; }
            U[d * 5 + g] = su;
            V[d * 5 + g] = sv;
            }
    double dt = distrib(gen);
    // uniaxial
    {
    V3 uk = sph(M_PI * distrib(gen), 2 * M_PI * distrib(gen));
    double K = distrib(gen), Js = 0.5 + distrib(gen), Kbis = 2.0 * K / Js;
    double Ha[15] = {0}, contrib[5];
    fgo_calc_aniso_uniax(5, uk.x, Kbis, FGO_THETA * dt, U, V, Ha, contrib);
    for (int g = 0; g < 5; g++)
        {
        double uk_u = uk.x[0] * U[g] + uk.x[1] * U[5 + g] + uk.x[2] * U[10 + g];
        double uk_v = uk.x[0] * V[g] + uk.x[1] * V[5 + g] + uk.x[2] * V[10 + g];
        double err2 = 0;
        for (int d = 0; d < 3; d++)
            {
            double ref = 2 * K / Js * uk_u * uk.x[d] + FGO_THETA * dt * (2 * K / Js * uk_v * uk.x[d]);
            err2 += (ref - Ha[d * 5 + g]) * (ref - Ha[d * 5 + g]);
            }
        TRUE_(std::sqrt(err2) <= 10 * UT_TOL);
        CLOSE(contrib[g], 2 * K / Js * uk_u * uk_u, 10 * UT_TOL);
        }
    }
    // cubic
    {
    V3 rv = sph(M_PI * distrib(gen), 2 * M_PI * distrib(gen));
    V3 ex = sph(M_PI * distrib(gen), 2 * M_PI * distrib(gen));
    double ey[3] = {ex.x[1] * rv.x[2] - ex.x[2] * rv.x[1], ex.x[2] * rv.x[0] - ex.x[0] * rv.x[2],
                    ex.x[0] * rv.x[1] - ex.x[1] * rv.x[0]};
    double ez[3] = {ex.x[1] * ey[2] - ex.x[2] * ey[1], ex.x[2] * ey[0] - ex.x[0] * ey[2],
                    ex.x[0] * ey[1] - ex.x[1] * ey[0]};
    double ny = std::sqrt(dot(ey, ey)), nz = std::sqrt(dot(ez, ez));
    for (int d = 0; d < 3; d++) { ey[d] /= ny; ez[d] /= nz; }
    double K3 = distrib(gen), Js = 0.5 + distrib(gen), K3bis = 2.0 * K3 / Js;
    double Ha[15] = {0}, contrib[5];
    fgo_calc_aniso_cub(5, ex.x, ey, ez, K3bis, FGO_THETA * dt, U, V, Ha, contrib);
    const double *ax[3] = {ex.x, ey, ez};
    for (int g = 0; g < 5; g++)
        {
        double cu[3], cv[3];
        for (int k = 0; k < 3; k++)
            {
            cu[k] = ax[k][0] * U[g] + ax[k][1] * U[5 + g] + ax[k][2] * U[10 + g];
            cv[k] = ax[k][0] * V[g] + ax[k][1] * V[5 + g] + ax[k][2] * V[10 + g];
            }
        double uHa3u = -2 * K3 / Js
                       * (cu[0] * (1 - cu[0] * cu[0]) * cu[0] + cu[1] * (1 - cu[1] * cu[1]) * cu[1]
                          + cu[2] * (1 - cu[2] * cu[2]) * cu[2]);
        double err2 = 0;
        for (int d = 0; d < 3; d++)
            {
            // time-derivative part: component d uses direction d's cosine and ex[d] (literal quirk,
            // ut_anisotropy.cpp:216-219 / src/tetra.cpp:199-203)
            double Ht = -2 * K3 / Js * cv[d] * (1 - 3 * cu[d] * cu[d]) * ex.x[d];
            double H = 0;
            for (int k = 0; k < 3; k++) H += -2 * K3 / Js * cu[k] * (1 - cu[k] * cu[k]) * ax[k][d];
            double ref = H + FGO_THETA * dt * Ht;
            err2 += (ref - Ha[d * 5 + g]) * (ref - Ha[d * 5 + g]);
            }
        TRUE_(std::sqrt(err2) <= 10 * UT_TOL);
        CLOSE(contrib[g], uHa3u, 10 * UT_TOL);
        }
    }
    }

// --- ut_node.cpp:58-100 setBasis axis, :102-121 orthonormality, :123-149 node_evol --------------
static void t_node()
    {
    std::mt19937 gen(5489u);
    std::uniform_real_distribution<> distrib(-1.0, 1.0);
    for (int rep = 0; rep < 64; rep++)
        {
        V3 u = cyl(M_PI * distrib(gen), distrib(gen));
        double ep[3], eq[3], r = M_PI * distrib(gen);
        // axis choice: with r = 0 the un-rotated ep is the Gram-Schmidt image of the chosen axis
        double ep0[3], eq0[3];
        fgo_node_set_basis(u.x, 0.0, ep0, eq0);
        double ax = std::fabs(u.x[0]), ay = std::fabs(u.x[1]), az = std::fabs(u.x[2]);
        int ref = (ax < ay) ? ((ax < az) ? 0 : 2) : ((ay < az) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[ref] = 1.0;
        double proj = dot(e, u.x), g[3], ng;
        for (int d = 0; d < 3; d++) g[d] = e[d] - proj * u.x[d];
        ng = std::sqrt(dot(g, g));
        for (int d = 0; d < 3; d++) CLOSE(ep0[d], g[d] / ng, 10 * UT_TOL);
        fgo_node_set_basis(u.x, r, ep, eq);
        CLOSE(std::sqrt(dot(u.x, u.x)), 1.0, 10 * UT_TOL);
        CLOSE(std::sqrt(dot(ep, ep)), 1.0, 10 * UT_TOL);
        CLOSE(std::sqrt(dot(eq, eq)), 1.0, 10 * UT_TOL);
        CLOSE(dot(u.x, ep), 0.0, 10 * UT_TOL);
        CLOSE(dot(ep, eq), 0.0, 10 * UT_TOL);
        CLOSE(dot(eq, u.x), 0.0, 10 * UT_TOL);
        // rotation: (ep, eq) = R(r) (ep0, eq0)
        for (int d = 0; d < 3; d++)
            {
            CLOSE(ep[d], std::cos(r) * ep0[d] - std::sin(r) * eq0[d], 10 * UT_TOL);
            CLOSE(eq[d], std::sin(r) * ep0[d] + std::cos(r) * eq0[d], 10 * UT_TOL);
            }
        double vp = distrib(gen), vq = distrib(gen), dt = distrib(gen) + 1.0, u1[3], v1[3];
        fgo_node_make_evol(u.x, ep, eq, vp, vq, dt, u1, v1);
        double err2 = 0, un[3], nn;
        for (int d = 0; d < 3; d++)
            {
            double vv = vp * ep[d] + vq * eq[d];
            err2 += (v1[d] - vv) * (v1[d] - vv);
            un[d] = u.x[d] + dt * vv;
            }
        TRUE_(std::sqrt(err2) <= 1e3 * UT_TOL);
        nn = std::sqrt(dot(un, un));
        for (int d = 0; d < 3; d++) CLOSE(u1[d], un[d] / nn, 1e3 * UT_TOL);
        }
    }

// --- ut_tetra.cpp:365-409 Tet_Pcoeff, ut_triangle.cpp:360-404 Tri_Pcoeff ------------------------
static void t_pcoeff()
    {
    std::mt19937 gen(5489u);
    std::uniform_real_distribution<> distrib(0.0, 1.0);
    for (int N = 3; N <= 4; N++)
        {
        double ep[12], eq[12], P[8 * 12];
        for (int i = 0; i < N; i++)
            {
            V3 u = sph(M_PI * distrib(gen), 2 * M_PI * distrib(gen));
            fgo_node_set_basis(u.x, 2 * M_PI * distrib(gen), ep + 3 * i, eq + 3 * i);
            }
        fgo_build_matP(N, ep, eq, P);
        std::vector<double> Pref(2 * N * 3 * N, 0.0);
        for (int i = 0; i < N; i++)
            for (int d = 0; d < 3; d++)
                {
                Pref[i * 3 * N + d * N + i] = ep[3 * i + d];
                Pref[(N + i) * 3 * N + d * N + i] = eq[3 * i + d];
                }
        for (int k = 0; k < 2 * N * 3 * N; k++) TRUE_(P[k] == Pref[k]);
        }
    }

// --- ut_triangle.cpp: surface / weights of the fixture triangle ----------------------------------
static void t_tri_tables()
    {
    int ind[3] = {0, 1, 2};
    double surf, n[3], w[4];
    fgo_tri_setup(unit_tri, ind, 4, &surf, n, w);
    CLOSE(surf, 0.5, UT_TOL);
    CLOSE(n[2], -1.0, UT_TOL);  // (p1-p0)x(p2-p0) of the fixture points to -z
    double s = 0;
    for (int g = 0; g < 4; g++) s += w[g];
    CLOSE(s, 0.5, 10 * UT_TOL);
    }

// --- ut_algebra.cpp:120-158 test_matrix_shape ----------------------------------------------------
struct Coef
    {
    int r, c;
    double v;
    };
struct Csr
    {
    int n;
    std::vector<int> rowptr, col;
    std::vector<double> val;
    };
static Csr build_csr(int n, const std::vector<Coef> &cs)  // unit-tests/sparse_matrix.h:25-39
    {
    std::vector<std::vector<int>> shape(n);
    for (auto &c : cs) shape[c.r].push_back(c.c);
    Csr A;
    A.n = n;
    A.rowptr.assign(n + 1, 0);
    for (int i = 0; i < n; i++)
        {
        std::sort(shape[i].begin(), shape[i].end());
        shape[i].erase(std::unique(shape[i].begin(), shape[i].end()), shape[i].end());
        A.rowptr[i + 1] = A.rowptr[i] + (int)shape[i].size();
        for (int c : shape[i]) A.col.push_back(c);
        }
    A.val.assign(A.col.size(), 0.0);
    for (auto &c : cs)
        {
        auto b = A.col.begin() + A.rowptr[c.r], e = A.col.begin() + A.rowptr[c.r + 1];
        A.val[std::lower_bound(b, e, c.c) - A.col.begin()] += c.v;
        }
    return A;
    }
static double resid(const Csr &A, const std::vector<double> &x, const std::vector<double> &b)
    {
    std::vector<double> y(A.n);
    fgo_spmv(A.n, A.rowptr.data(), A.col.data(), A.val.data(), x.data(), y.data());
    double s = 0;
    for (int i = 0; i < A.n; i++) s += (y[i] - b[i]) * (y[i] - b[i]);
    return std::sqrt(s);
    }

static void t_algebra_small()
    {
    Csr m = build_csr(4, {{1, 1, 3.14}, {0, 0, 1}, {2, 2, 5}, {3, 3, 42}, {1, 3, -10}, {1, 3, 10}, {0, 3, 0.5}});
    std::vector<double> x{1, 1, 1, 1}, y(4);
    fgo_spmv(4, m.rowptr.data(), m.col.data(), m.val.data(), x.data(), y.data());
    TRUE_(y[0] == 1.5);
    TRUE_(y[1] == 3.14);
    TRUE_(y[2] == 5.0);
    TRUE_(y[3] == 42.0);
    // ut_algebra.cpp:160-190 test_cg, :245-275 test_bicg
    std::vector<double> b{1, 1, 1, 1};
    for (int which = 0; which < 2; which++)
        {
        std::vector<double> xs(4, 0.0);
        fgo_iter it{1e-6, 700, 0, 0, 0, 0};
        if (which == 0) fgo_cg(&it, 4, m.rowptr.data(), m.col.data(), m.val.data(), xs.data(), b.data());
        else fgo_bicg(&it, 4, m.rowptr.data(), m.col.data(), m.val.data(), xs.data(), b.data());
        TRUE_(it.res < 1e-6);
        fgo_spmv(4, m.rowptr.data(), m.col.data(), m.val.data(), xs.data(), y.data());
        for (int i = 0; i < 4; i++) TRUE_((y[i] - b[i]) * (y[i] - b[i]) <= 10 * UT_TOL);
        }
    }

// --- ut_algebra.cpp:192-243 test_cg_dir, :277-330 test_bicg_dir: 1-D Laplacian, Dirichlet ramp ----
static void t_laplacian_dir()
    {
    const int NOD = 1001, MAXITER = 5000;
    const double tol = 1e-6;
    std::vector<Coef> cs = {{0, 0, 1.0}, {0, 1, -1.0}, {NOD - 1, NOD - 2, -1.0}, {NOD - 1, NOD - 1, 1.0}};
    for (int n = 1; n < NOD - 1; ++n)
        {
        cs.push_back({n, n - 1, -1.0});
        cs.push_back({n, n, 2.0});
        cs.push_back({n, n + 1, -1.0});
        }
    Csr K = build_csr(NOD, cs);
    std::vector<int> ld{0, NOD - 1};
    std::vector<double> Vd(NOD, 0.0), L(NOD, 0.0);
    Vd[NOD - 1] = 1.0;
    {
    std::vector<double> X(NOD, 0.0);
    fgo_iter it{tol, MAXITER, 0, 0, 0, 0};
    fgo_cg_dir(&it, NOD, K.rowptr.data(), K.col.data(), K.val.data(), X.data(), L.data(), Vd.data(), ld.data(), 2);
    for (int i = 0; i < NOD; i += 50) TRUE_(std::fabs(X[i] - i / (double)(NOD - 1)) < tol);
    }
    {
    std::vector<double> X(NOD, 0.0);
    fgo_iter it{tol, MAXITER, 0, 0, 0, 0};
    fgo_bicg_dir_xd(&it, NOD, K.rowptr.data(), K.col.data(), K.val.data(), X.data(), L.data(), Vd.data(), ld.data(), 2);
    for (int i = 0; i < NOD; i += 50)
        {
        double d = X[i] - i / (double)(NOD - 1);
        TRUE_(d * d < 10.0 * tol);
        }
    }
    }

// --- ut_algebra_bicg.cpp:19-117 ------------------------------------------------------------------
static void t_bicg_problems()
    {
    const int N = 10000;
    const double tol = 1e-8;
    {  // silly_problem_solver: identity, 0 iterations, exact
    std::vector<Coef> cs;
    for (int i = 0; i < N; i++) cs.push_back({i, i, 1.0});
    Csr A = build_csr(N, cs);
    std::vector<double> x(N, 0.0), b(N, 1.0);
    fgo_iter it{tol, 100, 0, 0, 0, 0};
    fgo_bicg(&it, N, A.rowptr.data(), A.col.data(), A.val.data(), x.data(), b.data());
    TRUE_(it.nit == 0);
    TRUE_(resid(A, x, b) == 0.0);
    }
    for (int asym = 0; asym < 2; asym++)
        {  // rand_sp_mat_problem_solver / rand_asym_sp_mat_problem_solver
        std::vector<Coef> cs;
        for (int i = 0; i < N; i++) cs.push_back({i, i, 1.0});
        std::mt19937 gen(5489u);
        std::uniform_int_distribution<> distrib(0, N - 1);
        for (int nb = 0; nb < (asym ? 800 : 400); nb++)
            {
            int i = distrib(gen), j = distrib(gen);
            cs.push_back({i, j, 1.0});
            if (!asym) cs.push_back({j, i, 1.0});
            }
        Csr A = build_csr(N, cs);
        std::vector<double> x(N, 2.0), b(N, 1.0);
        fgo_iter it{tol, 2000, 0, 0, 0, 0};
        fgo_bicg(&it, N, A.rowptr.data(), A.col.data(), A.val.data(), x.data(), b.data());
        TRUE_(it.nit > 1);
        TRUE_(resid(A, x, b) < tol);
        }
    }

int main()
    {
    t_timing_constructor();
    t_calc_alpha_eff();
    t_tet_inner_tables();
    t_tet_lumping();
    t_anisotropy();
    t_node();
    t_pcoeff();
    t_tri_tables();
    t_algebra_small();
    t_laplacian_dir();
    t_bicg_problems();
    std::printf("ut_oracle: %d checks, %d failures\n", n_check, n_fail);
    return n_fail ? 1 : 0;
    }
