/* fg_oracle.c — CPU ORACLE (test infrastructure, not product code): plain-C restatement of the
 * FeeLLGood per-time-step LLG hot path.  See fg_oracle.h for scope and pinning; every function
 * cites the reference file:line (relative to /root/reference) it follows.
 *
 * Threading: loops the reference runs under EXEC_POL (std::execution::par → TBB) carry an OpenMP
 * `parallel for` as the threaded stand-in; everything the reference runs serially stays serial
 * (BLAS-1, rhs scatter, node update), see SURVEY.md §8d.
 */
#include "fg_oracle.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* small 3-vector helpers (Eigen fixed-size semantics)                                          */
/* ------------------------------------------------------------------------------------------ */
static inline double dot3(const double *a, const double *b)
    { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static inline void cross3(const double *a, const double *b, double *r)
    {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
    }

/* Eigen MatrixBase::normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z) */
static inline void normalize3(double *a)
    {
    double z = dot3(a, a);
    if (z > 0.0)
        {
        double s = sqrt(z);
        a[0] /= s;
        a[1] /= s;
        a[2] /= s;
        }
    }

/* ------------------------------------------------------------------------------------------ */
/* timing, src/time_integration.h:11-15,37-42                                                  */
/* ------------------------------------------------------------------------------------------ */
double fgo_timing_dt0(double dtmin, double dtmax) { return sqrt(dtmin * dtmax); }

double fgo_timing_prefactor(double dt, double dtmax)
    {
    const double TAUR = 100. * dtmax;
    double t_tilde = dt / TAUR;
    return (1. + t_tilde * fabs(log(t_tilde)));
    }

/* ------------------------------------------------------------------------------------------ */
/* Gauss tables, src/tetra.h:29-81, src/triangle.h:21-65                                        */
/* ------------------------------------------------------------------------------------------ */
static double TET_A5[4 * 5], TET_A1[4 * 1], TET_PDS5[5], TET_PDS1[1];
static double TRI_A4[3 * 4], TRI_A1[3 * 1], TRI_PDS4[4], TRI_PDS1[1];
static int tables_ready = 0;

static void init_tables(void)
    {
    if (tables_ready) return;
    /* tetra.h:47-69 */
    const double A = 1. / 4., B = 1. / 6., C = 1. / 2., D = -2. / 15., E = 3. / 40.;
    const double u[5] = {A, B, B, B, C}, v[5] = {A, B, B, C, B}, w[5] = {A, B, C, B, B};
    const double pds[5] = {D, E, E, E, E};
    for (int j = 0; j < 5; j++)
        {
        TET_A5[0 * 5 + j] = 1. - u[j] - v[j] - w[j];
        TET_A5[1 * 5 + j] = u[j];
        TET_A5[2 * 5 + j] = v[j];
        TET_A5[3 * 5 + j] = w[j];
        TET_PDS5[j] = pds[j];
        }
    /* tetra.h:29-45 */
    TET_A1[0] = 1. - A - A - A;
    TET_A1[1] = A;
    TET_A1[2] = A;
    TET_A1[3] = A;
    TET_PDS1[0] = 1. / 6.;
    /* triangle.h:41-65 */
    const double tu[4] = {1 / 3., 1 / 5., 3 / 5., 1 / 5.}, tv[4] = {1 / 3., 1 / 5., 1 / 5., 3 / 5.};
    const double tp[4] = {-27 / 96., 25 / 96., 25 / 96., 25 / 96.};
    for (int j = 0; j < 4; j++)
        {
        TRI_A4[0 * 4 + j] = 1. - tu[j] - tv[j];
        TRI_A4[1 * 4 + j] = tu[j];
        TRI_A4[2 * 4 + j] = tv[j];
        TRI_PDS4[j] = tp[j];
        }
    /* triangle.h:21-39 */
    TRI_A1[0] = 1. - 1. / 3. - 1. / 3.;
    TRI_A1[1] = 1. / 3.;
    TRI_A1[2] = 1. / 3.;
    TRI_PDS1[0] = 1. / 2.;
    tables_ready = 1;
    }

const double *fgo_tet_a(int npi) { init_tables(); return npi == 1 ? TET_A1 : TET_A5; }
const double *fgo_tet_pds(int npi) { init_tables(); return npi == 1 ? TET_PDS1 : TET_PDS5; }
const double *fgo_tri_a(int npi) { init_tables(); return npi == 1 ? TRI_A1 : TRI_A4; }
const double *fgo_tri_pds(int npi) { init_tables(); return npi == 1 ? TRI_PDS1 : TRI_PDS4; }

/* ------------------------------------------------------------------------------------------ */
/* Node, src/node.h:73-122                                                                     */
/* ------------------------------------------------------------------------------------------ */
void fgo_node_set_basis(const double u[3], double r, double ep[3], double eq[3])
    {
    /* node.h:76-78: axis with the smallest |u_k| (Eigen minCoeff keeps the first minimum) */
    int minIdx = 0;
    double m = fabs(u[0]);
    if (fabs(u[1]) < m) { m = fabs(u[1]); minIdx = 1; }
    if (fabs(u[2]) < m) { m = fabs(u[2]); minIdx = 2; }
    ep[0] = ep[1] = ep[2] = 0.0;
    ep[minIdx] = 1.0;
    /* node.h:80-81: Gram-Schmidt against u */
    double d = dot3(ep, u);
    ep[0] -= d * u[0];
    ep[1] -= d * u[1];
    ep[2] -= d * u[2];
    normalize3(ep);
    /* node.h:84 */
    cross3(u, ep, eq);
    /* node.h:87-89: rotation by r */
    const double c = cos(r), s = sin(r);
    double new_ep[3];
    for (int k = 0; k < 3; k++) new_ep[k] = c * ep[k] - s * eq[k];
    for (int k = 0; k < 3; k++) eq[k] = s * ep[k] + c * eq[k];
    for (int k = 0; k < 3; k++) ep[k] = new_ep[k];
    }

void fgo_node_make_evol(const double u0[3], const double ep[3], const double eq[3], double vp,
                        double vq, double dt, double u1[3], double v1[3])
    {
    /* node.h:119-121 */
    for (int k = 0; k < 3; k++) v1[k] = vp * ep[k] + vq * eq[k];
    for (int k = 0; k < 3; k++) u1[k] = u0[k] + dt * v1[k];
    normalize3(u1);
    }

/* ------------------------------------------------------------------------------------------ */
/* Tet geometry, src/tetra.h:140-163, src/tetra.cpp:393-424                                    */
/* ------------------------------------------------------------------------------------------ */
int fgo_tet_orientate(const double *node_p, int ind[4])
    {
    const double *p0 = node_p + 3 * ind[0], *p1 = node_p + 3 * ind[1];
    const double *p2 = node_p + 3 * ind[2], *p3 = node_p + 3 * ind[3];
    double a[3], b[3], c[3], bc[3];
    for (int k = 0; k < 3; k++) { a[k] = p1[k] - p0[k]; b[k] = p2[k] - p0[k]; c[k] = p3[k] - p0[k]; }
    cross3(b, c, bc);
    const double mixed_prod = dot3(a, bc);
    if (fabs(mixed_prod) < FGO_EPSILON) return -1;
    if (mixed_prod < 0.0)
        {
        int t = ind[2];
        ind[2] = ind[3];
        ind[3] = t;
        return 1;
        }
    return 0;
    }

double fgo_tet_setup(const double *node_p, const int ind[4], int npi, double da[12],
                     double *weight)
    {
    init_tables();
    const double *p0 = node_p + 3 * ind[0], *p1 = node_p + 3 * ind[1];
    const double *p2 = node_p + 3 * ind[2], *p3 = node_p + 3 * ind[3];
    /* tetra.cpp:393-408: J(r,c) = (p_{c+1} - p_0)[r] */
    double J[3][3];
    for (int r = 0; r < 3; r++)
        {
        J[r][0] = p1[r] - p0[r];
        J[r][1] = p2[r] - p0[r];
        J[r][2] = p3[r] - p0[r];
        }
    /* Eigen 3x3 determinant (bruteforce_det3_helper) */
    const double detJ = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
                        - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                        + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    /* Eigen 3x3 inverse: cofactors, det from first column, multiply by 1/det */
    double cof[3][3]; /* cof[i][j] = cofactor(i,j) of J */
    cof[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    cof[1][0] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    cof[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    cof[0][1] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    cof[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    cof[2][1] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    cof[0][2] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    cof[1][2] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    cof[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det2 = cof[0][0] * J[0][0] + cof[0][1] * J[0][1] + cof[0][2] * J[0][2];
    const double invdet = 1.0 / det2;
    double Ji[3][3]; /* inverse(r,c) = cofactor(c,r) / det */
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Ji[r][c] = cof[c][r] * invdet;
    /* tetra.h:156-157: da = dadu * J^-1 */
    static const double dadu[4][3] = {{-1., -1., -1.}, {1., 0., 0.}, {0., 1., 0.}, {0., 0., 1.}};
    for (int i = 0; i < 4; i++)
        for (int d = 0; d < 3; d++)
            {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += dadu[i][k] * Ji[k][d];
            da[3 * i + d] = s;
            }
    const double *pds = fgo_tet_pds(npi);
    for (int j = 0; j < npi; j++) weight[j] = detJ * pds[j]; /* tetra.h:159-160 */
    return detJ;
    }

/* src/triangle.h:113-127 (ctor), :206-213 (calc_norm), :227 (calc_surf), :236-242 */
void fgo_tri_setup(const double *node_p, const int ind[3], int npi, double *surf, double n[3],
                   double *weight)
    {
    init_tables();
    const double *p0 = node_p + 3 * ind[0], *p1 = node_p + 3 * ind[1], *p2 = node_p + 3 * ind[2];
    double a[3], b[3];
    for (int k = 0; k < 3; k++) { a[k] = p1[k] - p0[k]; b[k] = p2[k] - p0[k]; }
    cross3(a, b, n);
    *surf = 0.5 * sqrt(dot3(n, n));
    normalize3(n);
    const double *pds = fgo_tri_pds(npi);
    for (int i = 0; i < npi; i++) weight[i] = 2.0 * (*surf) * pds[i];
    }

/* ------------------------------------------------------------------------------------------ */
/* src/tetra.cpp:47-75                                                                          */
/* ------------------------------------------------------------------------------------------ */
void fgo_calc_alpha_eff(int npi, double dt, double alpha, const double *uHeff, double *a_eff)
    {
    double reduced_dt = FGO_GAMMA0 * dt;
    const double r = 0.1;
    const double M = 2. * alpha * r / reduced_dt;
    for (int g = 0; g < npi; g++)
        {
        double h = uHeff[g];
        a_eff[g] = alpha;
        if (h > 0.)
            {
            if (h > M)
                a_eff[g] = alpha + reduced_dt / 2. * M;
            else
                a_eff[g] = alpha + reduced_dt / 2. * h;
            }
        else
            {
            if (h < -M)
                a_eff[g] = alpha / (1. + reduced_dt / (2. * alpha) * M);
            else
                a_eff[g] = alpha / (1. - reduced_dt / (2. * alpha) * h);
            }
        }
    }

/* ------------------------------------------------------------------------------------------ */
/* Tet::lumping + calcDiagBlock + calcOffDiagBlock, src/tetra.cpp:108-148                       */
/* ------------------------------------------------------------------------------------------ */
#define AE_(r, c) AE[(r) * 12 + (c)]
void fgo_tet_lumping(int npi, const double da[12], const double *weight, const double *u_nod,
                     const double *alpha_eff, double prefactor, double AE[144])
    {
    const int N = 4;
    const double *a = fgo_tet_a(npi);
    /* tetra.cpp:114: contrib = eigen_a * weight.cwiseProduct(alpha_eff) */
    double contrib[4], a_w[4], wsum = 0.0;
    for (int i = 0; i < N; i++)
        {
        double s = 0.0, t = 0.0;
        for (int g = 0; g < npi; g++)
            {
            s += a[i * npi + g] * (weight[g] * alpha_eff[g]);
            t += a[i * npi + g] * weight[g]; /* tetra.cpp:121: a_w = eigen_a * weight */
            }
        contrib[i] = s;
        a_w[i] = t;
        }
    for (int g = 0; g < npi; g++) wsum += weight[g];
    /* tetra.cpp:133-140: result = da*da^T; result *= c*weight.sum(); diagonal += x */
    double blk[4][4];
    const double cw = prefactor * wsum;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            {
            double s = 0.0;
            for (int d = 0; d < 3; d++) s += da[3 * i + d] * da[3 * j + d];
            blk[i][j] = s * cw;
            }
    for (int i = 0; i < N; i++) blk[i][i] += contrib[i];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            {
            AE_(i, j) += blk[i][j];
            AE_(N + i, N + j) += blk[i][j];
            AE_(2 * N + i, 2 * N + j) += blk[i][j];
            }
    /* tetra.cpp:122-130: off-diagonal blocks, diag = a_w ∘ m_{x,y,z} at the nodes */
    for (int i = 0; i < N; i++)
        {
        double dx = a_w[i] * u_nod[3 * i + 0];
        double dy = a_w[i] * u_nod[3 * i + 1];
        double dz = a_w[i] * u_nod[3 * i + 2];
        AE_(N + i, 2 * N + i) -= dx;
        AE_(2 * N + i, N + i) += dx;
        AE_(i, 2 * N + i) += dy;
        AE_(2 * N + i, i) -= dy;
        AE_(i, N + i) -= dz;
        AE_(N + i, i) += dz;
        }
    }
#undef AE_

/* src/tetra.cpp:171-181 */
void fgo_calc_aniso_uniax(int npi, const double uk[3], double Kbis, double s_dt, const double *U,
                          const double *V, double *H_aniso, double *out)
    {
    for (int g = 0; g < npi; g++)
        {
        double t[3];
        for (int d = 0; d < 3; d++) t[d] = U[d * npi + g] + s_dt * V[d * npi + g];
        double f = Kbis * dot3(uk, t);
        for (int d = 0; d < 3; d++) H_aniso[d * npi + g] += f * uk[d];
        }
    for (int g = 0; g < npi; g++)
        {
        double s = 0.0; /* (U^T * uk)(g) */
        for (int d = 0; d < 3; d++) s += U[d * npi + g] * uk[d];
        out[g] = Kbis * (s * s);
        }
    }

/* src/tetra.cpp:183-208 */
void fgo_calc_aniso_cub(int npi, const double ex[3], const double ey[3], const double ez[3],
                        double K3bis, double s_dt, const double *U, const double *V,
                        double *H_aniso, double *out)
    {
    for (int g = 0; g < npi; g++)
        {
        double Ug[3], Vg[3];
        for (int d = 0; d < 3; d++) { Ug[d] = U[d * npi + g]; Vg[d] = V[d * npi + g]; }
        double uk_u[3] = {dot3(ex, Ug), dot3(ey, Ug), dot3(ez, Ug)};
        double uk_v[3] = {dot3(ex, Vg), dot3(ey, Vg), dot3(ez, Vg)};
        double uk_uuu[3];
        for (int k = 0; k < 3; k++) uk_uuu[k] = uk_u[k] * (1.0 - uk_u[k] * uk_u[k]);
        double tmp[3]; /* tetra.cpp:199: uk_v.cwiseProduct(ex) — literal */
        for (int k = 0; k < 3; k++) tmp[k] = uk_v[k] * ex[k];
        for (int d = 0; d < 3; d++)
            {
            double inner = uk_uuu[0] * ex[d] + uk_uuu[1] * ey[d] + uk_uuu[2] * ez[d]
                           + s_dt * (tmp[d] * (1.0 - 3 * (uk_u[d] * uk_u[d])));
            H_aniso[d * npi + g] += -K3bis * inner;
            }
        out[g] = -K3bis * dot3(uk_u, uk_uuu);
        }
    }

/* element::buildMatP src/element.h:81-96 */
void fgo_build_matP(int N, const double *ep, const double *eq, double *P)
    {
    const int C = 3 * N;
    memset(P, 0, sizeof(double) * 2 * N * C);
    for (int i = 0; i < N; i++)
        for (int d = 0; d < 3; d++)
            {
            P[i * C + d * N + i] = ep[3 * i + d];
            P[(N + i) * C + d * N + i] = eq[3 * i + d];
            }
    }

/* add_drift_BE src/tetra.cpp:150-169 ; BE is [d*4+i] */
static void add_drift_BE(int npi, const double *weight, double alpha, double s_dt, double Vdrift,
                         const double *U, const double *V, const double *dUd_, const double *dVd_,
                         double *BE)
    {
    const double *a = fgo_tet_a(npi);
    for (int g = 0; g < npi; g++)
        {
        double Ug[3], Vg[3], dU[3], dV[3], c1[3], c2[3], c3[3];
        for (int d = 0; d < 3; d++)
            {
            Ug[d] = U[d * npi + g];
            Vg[d] = V[d * npi + g];
            dU[d] = dUd_[d * npi + g];
            dV[d] = dVd_[d * npi + g];
            }
        cross3(Ug, dU, c1);
        cross3(Ug, dV, c2);
        cross3(Vg, dU, c3);
        for (int i = 0; i < 4; i++)
            for (int d = 0; d < 3; d++)
                {
                double interim = a[i * npi + g]
                                 * (alpha * dU[d] + c1[d] + s_dt * (alpha * dV[d] + c2[d] + c3[d]));
                BE[d * 4 + i] += Vdrift * weight[g] * interim;
                }
        }
    }

/* ------------------------------------------------------------------------------------------ */
/* Tet::integrales src/tetra.cpp:210-307                                                        */
/* ------------------------------------------------------------------------------------------ */
void fgo_tet_integrales(int npi, const fgo_tet_prm *prm, double dt, double prefactor,
                        const double da[12], const double *weight, const double *u_nod,
                        const double *v_nod, const double *phi_nod, const double *phiv_nod,
                        const double *ep_nod, const double *eq_nod, const double *Hext, int idx_dir,
                        double Vdrift, double Kp[64], double Lp[8])
    {
    const int N = 4;
    const double *a = fgo_tet_a(npi);
    const double alpha = prm->alpha_LLG;
    const double Ms = prm->Ms;
    const double Abis = 2.0 * prm->A / (FGO_MU0 * Ms);
    const double s_dt = FGO_THETA * dt * FGO_GAMMA0;

    /* interpolation, tetra.h:183-218 ; all Gauss-point matrices are [d*npi+g] */
    double U[15], V[15], dUdx[15], dUdy[15], dUdz[15], Hd[15], Hv[15];
    for (int d = 0; d < 3; d++)
        for (int g = 0; g < npi; g++)
            {
            double su = 0.0, sv = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
            for (int i = 0; i < N; i++)
                {
                su += u_nod[3 * i + d] * a[i * npi + g];
                sv += v_nod[3 * i + d] * a[i * npi + g];
                sx += u_nod[3 * i + d] * da[3 * i + 0];
                sy += u_nod[3 * i + d] * da[3 * i + 1];
                sz += u_nod[3 * i + d] * da[3 * i + 2];
                }
            U[d * npi + g] = su;
            V[d * npi + g] = sv;
            dUdx[d * npi + g] = sx;
            dUdy[d * npi + g] = sy;
            dUdz[d * npi + g] = sz;
            }
    for (int g = 0; g < npi; g++)
        for (int d = 0; d < 3; d++)
            {
            double hd = 0.0, hv = 0.0; /* tetra.h:208-217: X.col(j) -= scalar_nod[i]*da.row(i) */
            for (int i = 0; i < N; i++)
                {
                hd -= phi_nod[i] * da[3 * i + d];
                hv -= phiv_nod[i] * da[3 * i + d];
                }
            Hd[d * npi + g] = hd;
            Hv[d * npi + g] = hv;
            }

    /* tetra.cpp:232-233 */
    double uHeff[5];
    for (int g = 0; g < npi; g++)
        {
        double nx = 0.0, ny = 0.0, nz = 0.0;
        for (int d = 0; d < 3; d++)
            {
            nx += dUdx[d * npi + g] * dUdx[d * npi + g];
            ny += dUdy[d * npi + g] * dUdy[d * npi + g];
            nz += dUdz[d * npi + g] * dUdz[d * npi + g];
            }
        uHeff[g] = -Abis * (nx + ny + nz);
        }
    double H_aniso[15];
    for (int k = 0; k < 3 * npi; k++) H_aniso[k] = 0.0;

    /* tetra.cpp:237-246 */
    if (prm->K != 0)
        {
        double Kbis = 2.0 * prm->K / (FGO_MU0 * Ms), add[5];
        fgo_calc_aniso_uniax(npi, prm->uk, Kbis, s_dt / FGO_GAMMA0, U, V, H_aniso, add);
        for (int g = 0; g < npi; g++) uHeff[g] += add[g];
        }
    if (prm->K3 != 0)
        {
        double K3bis = 2.0 * prm->K3 / (FGO_MU0 * Ms), add[5];
        fgo_calc_aniso_cub(npi, prm->ex, prm->ey, prm->ez, K3bis, s_dt / FGO_GAMMA0, U, V, H_aniso,
                           add);
        for (int g = 0; g < npi; g++) uHeff[g] += add[g];
        }

    /* tetra.cpp:248-255 ; Hst = 0 (extraField is a no-op unless spin accumulation is on) */
    double Heff[15], H[15], Hst[15];
    for (int k = 0; k < 3 * npi; k++)
        {
        Heff[k] = Hd[k] + Hext[k];
        H[k] = Heff[k];
        Hst[k] = 0.0;
        }
    for (int k = 0; k < 3 * npi; k++) Heff[k] += Hst[k];
    for (int g = 0; g < npi; g++)
        {
        double s = 0.0;
        for (int d = 0; d < 3; d++) s += U[d * npi + g] * Heff[d * npi + g];
        uHeff[g] += s;
        }

    /* tetra.cpp:257-261 */
    double a_eff[5];
    fgo_calc_alpha_eff(npi, dt, alpha, uHeff, a_eff);
    double AE[144];
    memset(AE, 0, sizeof(AE));
    fgo_tet_lumping(npi, da, weight, u_nod, a_eff, prefactor * s_dt * Abis, AE);

    /* tetra.cpp:263-274: Kp = Perm * P * AE * P^T, Perm.indices = {4,5,6,7,0,1,2,3}:
     * row i of (P*AE*P^T) lands on row Perm[i]. */
    double P[8 * 12], PA[8 * 12], PAPt[64];
    fgo_build_matP(N, ep_nod, eq_nod, P);
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 12; c++)
            {
            double s = 0.0;
            for (int k = 0; k < 12; k++) s += P[r * 12 + k] * AE[k * 12 + c];
            PA[r * 12 + c] = s;
            }
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++)
            {
            double s = 0.0;
            for (int k = 0; k < 12; k++) s += PA[r * 12 + k] * P[c * 12 + k];
            PAPt[r * 8 + c] = s;
            }
    static const int Perm[8] = {4, 5, 6, 7, 0, 1, 2, 3};
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) Kp[Perm[r] * 8 + c] = PAPt[r * 8 + c];

    /* tetra.cpp:277-303: BE, stored [d*4+i] (BE(d,i)) */
    double BE[12];
    for (int k = 0; k < 12; k++) BE[k] = 0.0;
    if (idx_dir != FGO_IDX_UNDEF)
        {
        double dVd_dir[15];
        for (int d = 0; d < 3; d++)
            for (int g = 0; g < npi; g++)
                {
                double s = 0.0;
                for (int i = 0; i < N; i++) s += v_nod[3 * i + d] * da[3 * i + idx_dir];
                dVd_dir[d * npi + g] = s;
                }
        const double *dUd = (idx_dir == FGO_IDX_Z) ? dUdz : (idx_dir == FGO_IDX_Y ? dUdy : dUdx);
        add_drift_BE(npi, weight, alpha, s_dt, Vdrift, U, V, dUd, dVd_dir, BE);
        }
    for (int k = 0; k < 3 * npi; k++) H[k] += H_aniso[k] + (s_dt / FGO_GAMMA0) * Hv[k];

    for (int g = 0; g < npi; g++)
        {
        const double w = weight[g];
        double scal_Hst_u = 0.0;
        for (int d = 0; d < 3; d++) scal_Hst_u += Hst[d * npi + g] * U[d * npi + g];
        for (int i = 0; i < N; i++)
            {
            const double ai_w = w * a[i * npi + g];
            for (int d = 0; d < 3; d++)
                {
                BE[d * 4 + i] -= w * Abis
                                 * (da[3 * i + 0] * dUdx[d * npi + g] + da[3 * i + 1] * dUdy[d * npi + g]
                                    + da[3 * i + 2] * dUdz[d * npi + g]);
                BE[d * 4 + i] += ai_w * (H[d * npi + g] + Hst[d * npi + g]);
                BE[d * 4 + i] -= ai_w * scal_Hst_u * s_dt * V[d * npi + g];
                }
            }
        }
    /* tetra.cpp:306: Lp = Perm * P * BE.reshaped<RowMajor>() ; row-major reshape of BE(d,i)
     * is exactly the [d*4+i] storage. */
    double PL[8];
    for (int r = 0; r < 8; r++)
        {
        double s = 0.0;
        for (int k = 0; k < 12; k++) s += P[r * 12 + k] * BE[k];
        PL[r] = s;
        }
    for (int r = 0; r < 8; r++) Lp[Perm[r]] = PL[r];
    }

/* ------------------------------------------------------------------------------------------ */
/* Tri::integrales src/triangle.cpp:6-36                                                        */
/* ------------------------------------------------------------------------------------------ */
void fgo_tri_integrales(int npi, const fgo_tri_prm *prm, double dMs, const double *weight,
                        const double *u_nod, const double *ep_nod, const double *eq_nod,
                        double Lp[6])
    {
    const int N = 3;
    const double *a = fgo_tri_a(npi);
    double Kbis = 2.0 * prm->Ks / dMs;
    double u[12]; /* [d*npi+g] */
    for (int d = 0; d < 3; d++)
        for (int g = 0; g < npi; g++)
            {
            double s = 0.0;
            for (int i = 0; i < N; i++) s += u_nod[3 * i + d] * a[i * npi + g];
            u[d * npi + g] = s;
            }
    double BE[9]; /* BE(k,i) at [k*3+i] */
    for (int k = 0; k < 9; k++) BE[k] = 0.0;
    for (int g = 0; g < npi; g++)
        {
        double ug[3] = {u[0 * npi + g], u[1 * npi + g], u[2 * npi + g]};
        double _prefactor = weight[g] * Kbis * dot3(prm->uk, ug);
        for (int i = 0; i < N; i++)
            for (int k = 0; k < 3; k++) BE[k * 3 + i] += _prefactor * a[i * npi + g] * prm->uk[k];
        }
    double P[6 * 9], PL[6];
    fgo_build_matP(N, ep_nod, eq_nod, P);
    for (int r = 0; r < 6; r++)
        {
        double s = 0.0;
        for (int k = 0; k < 9; k++) s += P[r * 9 + k] * BE[k];
        PL[r] = s;
        }
    static const int Perm[6] = {3, 4, 5, 0, 1, 2};
    for (int r = 0; r < 6; r++) Lp[Perm[r]] = PL[r];
    }

/* ------------------------------------------------------------------------------------------ */
/* src/algebra: BLAS-1 (algebra.h:35-77, algebraCore.h:10-17), all serial like the reference    */
/* ------------------------------------------------------------------------------------------ */
static double v_dot(int n, const double *X, const double *Y)
    {
    double s = 0.0; /* std::inner_product left fold */
    for (int i = 0; i < n; i++) s = s + X[i] * Y[i];
    return s;
    }
static double v_norm(int n, const double *X) { return sqrt(fabs(v_dot(n, X, X))); }
static void v_scaled(int n, double alpha, double *Y) { for (int i = 0; i < n; i++) Y[i] *= alpha; }
static void v_pdirect(int n, const double *X, const double *Y, double *Z)
    { for (int i = 0; i < n; i++) Z[i] = X[i] * Y[i]; }
static void v_add(int n, const double *X, double *Y) { for (int i = 0; i < n; i++) Y[i] += X[i]; }
static void v_sub(int n, const double *X, double *Y) { for (int i = 0; i < n; i++) Y[i] -= X[i]; }
static void v_scaled_add(int n, const double *X, double alpha, double *Y)
    { for (int i = 0; i < n; i++) Y[i] += alpha * X[i]; }
static void v_mask(const int *ld, int nld, double *X) { for (int k = 0; k < nld; k++) X[ld[k]] = 0.0; }

/* SparseMatrix::mult sparseMat.h:158-170 (parallel over rows under EXEC_POL) */
void fgo_spmv(int n, const int *rowptr, const int *col, const double *val, const double *x,
              double *y)
    {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++)
        {
        double v = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; k++) v += val[k] * x[col[k]];
        y[i] = v;
        }
    }

/* SparseMatrix::operator()(i,i) + build_diag_precond sparseMat.h:143-183 */
static void build_diag_precond(int n, const int *rowptr, const int *col, const double *val,
                               double *D)
    {
    for (int i = 0; i < n; i++)
        {
        double c = 0.0;
        int lo = rowptr[i], hi = rowptr[i + 1];
        while (lo < hi) /* lower_bound */
            {
            int mid = lo + (hi - lo) / 2;
            if (col[mid] < i) lo = mid + 1; else hi = mid;
            }
        if (lo < rowptr[i + 1] && col[lo] == i) c = val[lo];
        D[i] = 1.0 / c;
        }
    }

/* iteration<T>, src/algebra/iter.h:37-179 */
static void it_reset(fgo_iter *it)
    {
    it->rhsn = 1.0;
    it->nit = 0;
    it->res = 1.7976931348623157e308; /* numeric_limits<double>::max() */
    it->status = FGO_UNDEFINED;
    }
static int it_finished(fgo_iter *it, double nr)
    {
    it->res = fabs(nr);
    if (isnan(it->res)) { it->status = FGO_CANNOT_CONVERGE; return 0; }
    if (it->res <= it->rhsn * it->resmax) { it->status = FGO_CONVERGED; return 1; }
    return 0;
    }
static void it_inc(fgo_iter *it)
    {
    it->nit++;
    if (it->nit >= it->maxiter) it->status = FGO_ITER_OVERFLOW;
    }

/* common body of bicg / bicg_dir / bicg_dir(xd): src/algebra/bicg.h:14-72,83-154,163-234.
 * ld == NULL → plain bicg (no masking); xd != NULL → Dirichlet values variant. */
static void bicg_core(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                      double *x, const double *rhs, const double *xd, const int *ld, int nld)
    {
    double rho_1 = 0.0, rho_2 = 0.0, alpha = 0.0, beta = 0.0, omega = 0.0;
    double *buf = (double *)malloc(sizeof(double) * 10 * (size_t)n);
    double *p = buf, *phat = buf + (size_t)n, *shat = buf + 2 * (size_t)n, *r = buf + 3 * (size_t)n;
    double *rt = buf + 4 * (size_t)n, *s = buf + 5 * (size_t)n, *t = buf + 6 * (size_t)n;
    double *v = buf + 7 * (size_t)n, *D = buf + 8 * (size_t)n, *b = buf + 9 * (size_t)n;
    memcpy(b, rhs, sizeof(double) * n);
    for (int i = 0; i < n; i++) phat[i] = 0.0;

    build_diag_precond(n, rowptr, col, val, D);
    if (xd)
        {
        fgo_spmv(n, rowptr, col, val, xd, v);
        v_sub(n, v, b); /* b -= A xd */
        }
    if (ld) { v_mask(ld, nld, b); v_mask(ld, nld, D); }
    it->rhsn = v_norm(n, b);
    memcpy(r, b, sizeof(double) * n);
    fgo_spmv(n, rowptr, col, val, x, v);
    v_sub(n, v, r);
    if (ld) v_mask(ld, nld, r);
    memcpy(rt, r, sizeof(double) * n);
    memcpy(p, r, sizeof(double) * n);
    while (!it_finished(it, v_norm(n, r)))
        {
        rho_1 = v_dot(n, rt, r);
        if (it->nit > 0)
            {
            if ((rho_2 == 0) || (omega == 0))
                {
                it->status = FGO_CANNOT_CONVERGE;
                break;
                }
            beta = (rho_1 / rho_2) * (alpha / omega);
            v_scaled(n, omega, v);
            v_sub(n, v, p);
            v_scaled(n, beta, p);
            v_add(n, r, p);
            }
        v_pdirect(n, D, p, phat);
        fgo_spmv(n, rowptr, col, val, phat, v);
        if (ld) v_mask(ld, nld, v);
        alpha = rho_1 / v_dot(n, v, rt);
        memcpy(s, r, sizeof(double) * n);
        v_scaled_add(n, v, -alpha, s);
        if (it_finished(it, v_norm(n, s)))
            {
            v_scaled_add(n, phat, alpha, x);
            break;
            }
        else if ((it->status == FGO_ITER_OVERFLOW) || (it->status == FGO_CANNOT_CONVERGE))
            { break; }
        v_pdirect(n, D, s, shat);
        fgo_spmv(n, rowptr, col, val, shat, t);
        if (ld) v_mask(ld, nld, t);
        omega = v_dot(n, t, s) / v_dot(n, t, t);
        v_scaled_add(n, phat, alpha, x);
        v_scaled_add(n, shat, omega, x);
        v_scaled(n, omega, t);
        memcpy(r, s, sizeof(double) * n);
        v_sub(n, t, r);
        rho_2 = rho_1;
        it_inc(it);
        }
    if (xd) v_add(n, xd, x);
    free(buf);
    }

void fgo_bicg(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val, double *x,
              const double *rhs)
    {
    it_reset(it);
    bicg_core(it, n, rowptr, col, val, x, rhs, NULL, NULL, 0);
    }

double fgo_bicg_dir(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                    double *x, const double *rhs, const int *ld, int nld)
    {
    static const int none = 0;
    it_reset(it);
    bicg_core(it, n, rowptr, col, val, x, rhs, NULL, ld ? ld : &none, ld ? nld : 0);
    return it->res / it->rhsn;
    }

void fgo_bicg_dir_xd(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                     double *x, const double *rhs, const double *xd, const int *ld, int nld)
    {
    static const int none = 0;
    it_reset(it);
    bicg_core(it, n, rowptr, col, val, x, rhs, xd, ld ? ld : &none, ld ? nld : 0);
    }

/* cg / cg_dir: src/algebra/cg.h:15-58,68-121 */
static void cg_core(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                    double *x, const double *rhs, const double *xd, const int *ld, int nld)
    {
    double rho, rho_1 = 0.0;
    double *buf = (double *)malloc(sizeof(double) * 7 * (size_t)n);
    double *p = buf, *q = buf + (size_t)n, *r = buf + 2 * (size_t)n, *z = buf + 3 * (size_t)n;
    double *D = buf + 4 * (size_t)n, *b = buf + 5 * (size_t)n, *v_temp = buf + 6 * (size_t)n;
    memcpy(b, rhs, sizeof(double) * n);
    build_diag_precond(n, rowptr, col, val, D);
    if (xd)
        {
        fgo_spmv(n, rowptr, col, val, xd, z);
        v_sub(n, z, b);
        }
    if (ld) { v_mask(ld, nld, b); v_mask(ld, nld, D); }
    it->rhsn = v_norm(n, b);
    memcpy(r, b, sizeof(double) * n);
    fgo_spmv(n, rowptr, col, val, x, v_temp);
    v_sub(n, v_temp, r);
    if (ld) v_mask(ld, nld, r);
    v_pdirect(n, D, r, z);
    rho = v_dot(n, z, r);
    memcpy(p, z, sizeof(double) * n);
    while (!it_finished(it, v_norm(n, r)) && (it->status != FGO_ITER_OVERFLOW)
           && (it->status != FGO_CANNOT_CONVERGE))
        {
        if (it->nit > 0)
            {
            v_pdirect(n, D, r, z);
            rho = v_dot(n, z, r);
            v_scaled(n, rho / rho_1, p);
            v_add(n, z, p);
            }
        fgo_spmv(n, rowptr, col, val, p, q);
        if (ld) v_mask(ld, nld, q);
        double q_dot_p = v_dot(n, q, p);
        if (q_dot_p == 0.0)
            {
            it->status = FGO_CANNOT_CONVERGE;
            break;
            }
        double a = rho / q_dot_p;
        v_scaled_add(n, p, +a, x);
        v_scaled_add(n, q, -a, r);
        rho_1 = rho;
        it_inc(it);
        }
    if (xd) v_add(n, xd, x);
    free(buf);
    }

void fgo_cg(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val, double *x,
            const double *rhs)
    {
    it_reset(it);
    cg_core(it, n, rowptr, col, val, x, rhs, NULL, NULL, 0);
    }

void fgo_cg_dir(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                double *x, const double *rhs, const double *xd, const int *ld, int nld)
    {
    static const int none = 0;
    it_reset(it);
    cg_core(it, n, rowptr, col, val, x, rhs, xd, ld ? ld : &none, ld ? nld : 0);
    }

/* ------------------------------------------------------------------------------------------ */
/* mesh-level context                                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct ref_algebra_vt
    {
    void *handle;
    void *(*create)(int n, const int *rowptr, const int *col);
    void (*destroy)(void *m);
    void (*clear)(void *m);
    void (*add)(void *m, int i, int j, double v);
    void (*set)(void *m, int i, int j, double v);
    void (*get_values)(void *m, double *val);
    double (*bicg_dir)(void *m, double *x, const double *rhs, int n, const int *ld, int nld,
                       double tol, int maxiter, int *status, int *nit, double *res, double *rhsn);
    } ref_algebra_vt;

struct fgo_ctx
    {
    int NOD, NT, NF, npi, npi_tri, nreg_tet, nreg_tri, nthreads;
    double *p;                     /* NOD x 3 */
    double *ep, *eq;               /* NOD x 3 */
    double *u[2], *v[2], *phi[2], *phiv[2]; /* dataNode d[CURRENT|NEXT], node.h:47-70 */
    int *tet_ind, *tet_reg;        /* NT x 4 (oriented), NT */
    double *tet_da, *tet_w;        /* NT x 12, NT x npi */
    double *tet_Kp, *tet_Lp;       /* NT x 64, NT x 8 : stored in the element, element.h:62,65 */
    int *tri_ind, *tri_reg;
    double *tri_dMs, *tri_w, *tri_Lp; /* NF, NF x npi_tri, NF x 6 */
    fgo_tet_prm *prm_tet;
    fgo_tri_prm *prm_tri;
    unsigned char *magNode;        /* mesh.h:115-127 */
    int *magTet, n_magTet;
    int *magTri, n_magTri;
    int *edges, n_edges, n_edges_mag; /* sorted unique (first<second), mesh.h:100-114 */
    int n, nnz, *rowptr, *col;     /* solver.h:75-104 shape of K */
    double *K, *L_rhs, *Xw;
    int *lvd, nlvd;                /* linear_algebra.h:55-63 */
    double *extSpaceField;         /* NT x 3 x npi or NULL */
    fgo_iter iter;
    double v_max;
    ref_algebra_vt ref;            /* optional: the reference's own SparseMatrix + bicg_dir */
    void *refK;
    };

static int cmp_edge(const void *a, const void *b)
    {
    const int *x = (const int *)a, *y = (const int *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
    }
static int cmp_int(const void *a, const void *b)
    {
    int x = *(const int *)a, y = *(const int *)b;
    return x < y ? -1 : (x > y);
    }

static int is_magnetic_tet(const fgo_ctx *c, int t) { return c->prm_tet[c->tet_reg[t]].Ms > 0; }

fgo_ctx *fgo_create(int NOD, const double *node_p, int NT, const int *tet_ind, const int *tet_reg,
                    int NF, const int *tri_ind, const int *tri_reg, const double *tri_dMs,
                    int nreg_tet, const fgo_tet_prm *prm_tet, int nreg_tri,
                    const fgo_tri_prm *prm_tri, int npi_tet, int npi_tri, double tol, int maxiter)
    {
    init_tables();
    fgo_ctx *c = (fgo_ctx *)calloc(1, sizeof(fgo_ctx));
    c->NOD = NOD; c->NT = NT; c->NF = NF; c->npi = npi_tet; c->npi_tri = npi_tri;
    c->nreg_tet = nreg_tet; c->nreg_tri = nreg_tri; c->nthreads = 1;
    c->iter.resmax = tol; c->iter.maxiter = maxiter;
    it_reset(&c->iter);
    size_t N3 = sizeof(double) * 3 * (size_t)NOD, N1 = sizeof(double) * (size_t)NOD;
    c->p = (double *)malloc(N3); memcpy(c->p, node_p, N3);
    c->ep = (double *)calloc(1, N3); c->eq = (double *)calloc(1, N3);
    for (int k = 0; k < 2; k++)
        {
        c->u[k] = (double *)calloc(1, N3); c->v[k] = (double *)calloc(1, N3);
        c->phi[k] = (double *)calloc(1, N1); c->phiv[k] = (double *)calloc(1, N1);
        }
    c->prm_tet = (fgo_tet_prm *)malloc(sizeof(fgo_tet_prm) * nreg_tet);
    memcpy(c->prm_tet, prm_tet, sizeof(fgo_tet_prm) * nreg_tet);
    c->prm_tri = (fgo_tri_prm *)malloc(sizeof(fgo_tri_prm) * (nreg_tri > 0 ? nreg_tri : 1));
    if (nreg_tri > 0) memcpy(c->prm_tri, prm_tri, sizeof(fgo_tri_prm) * nreg_tri);

    /* tets: Tet ctor (orientate, da, weight) */
    c->tet_ind = (int *)malloc(sizeof(int) * 4 * (size_t)NT);
    c->tet_reg = (int *)malloc(sizeof(int) * (size_t)NT);
    memcpy(c->tet_ind, tet_ind, sizeof(int) * 4 * (size_t)NT);
    memcpy(c->tet_reg, tet_reg, sizeof(int) * (size_t)NT);
    c->tet_da = (double *)malloc(sizeof(double) * 12 * (size_t)NT);
    c->tet_w = (double *)malloc(sizeof(double) * npi_tet * (size_t)NT);
    c->tet_Kp = (double *)calloc((size_t)NT * 64, sizeof(double));
    c->tet_Lp = (double *)calloc((size_t)NT * 8, sizeof(double));
    for (int t = 0; t < NT; t++)
        {
        if (fgo_tet_orientate(c->p, c->tet_ind + 4 * t) < 0)
            {
            fprintf(stderr, "fg_oracle: singular tetrahedron %d\n", t);
            exit(1);
            }
        fgo_tet_setup(c->p, c->tet_ind + 4 * t, npi_tet, c->tet_da + 12 * (size_t)t,
                      c->tet_w + (size_t)npi_tet * t);
        }
    /* tris */
    c->tri_ind = (int *)malloc(sizeof(int) * 3 * (size_t)(NF > 0 ? NF : 1));
    c->tri_reg = (int *)malloc(sizeof(int) * (size_t)(NF > 0 ? NF : 1));
    c->tri_dMs = (double *)malloc(sizeof(double) * (size_t)(NF > 0 ? NF : 1));
    c->tri_w = (double *)malloc(sizeof(double) * npi_tri * (size_t)(NF > 0 ? NF : 1));
    c->tri_Lp = (double *)calloc((size_t)(NF > 0 ? NF : 1) * 6, sizeof(double));
    if (NF > 0)
        {
        memcpy(c->tri_ind, tri_ind, sizeof(int) * 3 * (size_t)NF);
        memcpy(c->tri_reg, tri_reg, sizeof(int) * (size_t)NF);
        memcpy(c->tri_dMs, tri_dMs, sizeof(double) * (size_t)NF);
        for (int f = 0; f < NF; f++)
            {
            double surf, nrm[3];
            fgo_tri_setup(c->p, c->tri_ind + 3 * f, npi_tri, &surf, nrm, c->tri_w + (size_t)npi_tri * f);
            }
        }

    /* mesh.h:100-114: all tet edges, sorted unique */
    int *ed = (int *)malloc(sizeof(int) * 2 * 6 * (size_t)NT);
    size_t ne = 0;
    for (int t = 0; t < NT; t++)
        for (int i = 0; i < 3; ++i)
            for (int j = i + 1; j < 4; ++j)
                {
                int a = c->tet_ind[4 * t + i], b = c->tet_ind[4 * t + j];
                ed[2 * ne] = a < b ? a : b;
                ed[2 * ne + 1] = a < b ? b : a;
                ne++;
                }
    qsort(ed, ne, 2 * sizeof(int), cmp_edge);
    size_t nu = 0;
    for (size_t k = 0; k < ne; k++)
        if (k == 0 || ed[2 * k] != ed[2 * nu - 2] || ed[2 * k + 1] != ed[2 * nu - 1])
            {
            ed[2 * nu] = ed[2 * k];
            ed[2 * nu + 1] = ed[2 * k + 1];
            nu++;
            }
    c->edges = (int *)realloc(ed, sizeof(int) * 2 * (nu > 0 ? nu : 1));
    c->n_edges = (int)nu;

    /* mesh.h:115-131: magNode, magTet, magTri */
    c->magNode = (unsigned char *)calloc((size_t)NOD, 1);
    c->magTet = (int *)malloc(sizeof(int) * (size_t)(NT > 0 ? NT : 1));
    for (int t = 0; t < NT; t++)
        if (is_magnetic_tet(c, t))
            {
            c->magTet[c->n_magTet++] = t;
            for (int i = 0; i < 4; i++) c->magNode[c->tet_ind[4 * t + i]] = 1;
            }
    c->magTri = (int *)malloc(sizeof(int) * (size_t)(NF > 0 ? NF : 1));
    for (int f = 0; f < NF; f++)
        {
        const int *ind = c->tri_ind + 3 * f;
        int mag = c->magNode[ind[0]] && c->magNode[ind[1]] && c->magNode[ind[2]];
        if (mag && !c->prm_tri[c->tri_reg[f]].suppress_charges) c->magTri[c->n_magTri++] = f;
        }

    /* solver.h:75-104 with the LinAlgebra edge filter (linear_algebra.h:43): 2x2 block per node
     * and per directed magnetic edge; rows sorted (std::set). */
    const int n = 2 * NOD;
    c->n = n;
    int *deg = (int *)calloc((size_t)NOD + 1, sizeof(int));
    for (int i = 0; i < NOD; i++) deg[i] = 1;
    for (int e = 0; e < c->n_edges; e++)
        {
        int a = c->edges[2 * e], b = c->edges[2 * e + 1];
        if (c->magNode[a] && c->magNode[b]) { deg[a]++; deg[b]++; c->n_edges_mag++; }
        }
    int *nptr = (int *)malloc(sizeof(int) * ((size_t)NOD + 1));
    nptr[0] = 0;
    for (int i = 0; i < NOD; i++) nptr[i + 1] = nptr[i] + deg[i];
    int *ncol = (int *)malloc(sizeof(int) * (size_t)nptr[NOD]);
    int *fill = (int *)calloc((size_t)NOD, sizeof(int));
    for (int i = 0; i < NOD; i++) ncol[nptr[i] + fill[i]++] = i;
    for (int e = 0; e < c->n_edges; e++)
        {
        int a = c->edges[2 * e], b = c->edges[2 * e + 1];
        if (c->magNode[a] && c->magNode[b])
            {
            ncol[nptr[a] + fill[a]++] = b;
            ncol[nptr[b] + fill[b]++] = a;
            }
        }
    for (int i = 0; i < NOD; i++) qsort(ncol + nptr[i], deg[i], sizeof(int), cmp_int);
    c->nnz = 4 * nptr[NOD];
    c->rowptr = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    c->col = (int *)malloc(sizeof(int) * (size_t)c->nnz);
    c->rowptr[0] = 0;
    for (int i = 0; i < NOD; i++)
        for (int k = 0; k < 2; k++)
            {
            int row = 2 * i + k, base = c->rowptr[row];
            for (int j = 0; j < deg[i]; j++)
                {
                c->col[base + 2 * j] = 2 * ncol[nptr[i] + j];
                c->col[base + 2 * j + 1] = 2 * ncol[nptr[i] + j] + 1;
                }
            c->rowptr[row + 1] = base + 2 * deg[i];
            }
    free(deg); free(nptr); free(ncol); free(fill);
    c->K = (double *)calloc((size_t)c->nnz, sizeof(double));
    c->L_rhs = (double *)calloc((size_t)n, sizeof(double));
    c->Xw = (double *)calloc((size_t)n, sizeof(double));

    /* linear_algebra.h:55-63 */
    c->lvd = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < NOD; i++)
        if (!c->magNode[i])
            {
            c->lvd[c->nlvd++] = 2 * i;
            c->lvd[c->nlvd++] = 2 * i + 1;
            }
    return c;
    }

void fgo_destroy(fgo_ctx *c)
    {
    if (!c) return;
    if (c->refK && c->ref.destroy) c->ref.destroy(c->refK);
    if (c->ref.handle) dlclose(c->ref.handle);
    free(c->p); free(c->ep); free(c->eq);
    for (int k = 0; k < 2; k++) { free(c->u[k]); free(c->v[k]); free(c->phi[k]); free(c->phiv[k]); }
    free(c->tet_ind); free(c->tet_reg); free(c->tet_da); free(c->tet_w); free(c->tet_Kp);
    free(c->tet_Lp); free(c->tri_ind); free(c->tri_reg); free(c->tri_dMs); free(c->tri_w);
    free(c->tri_Lp); free(c->prm_tet); free(c->prm_tri); free(c->magNode); free(c->magTet);
    free(c->magTri); free(c->edges); free(c->rowptr); free(c->col); free(c->K); free(c->L_rhs);
    free(c->Xw); free(c->lvd); free(c->extSpaceField);
    free(c);
    }

void fgo_set_num_threads(fgo_ctx *c, int nthreads)
    {
    c->nthreads = nthreads > 0 ? nthreads : 1;
#ifdef _OPENMP
    omp_set_num_threads(c->nthreads);
#endif
    }

int fgo_use_reference_algebra(fgo_ctx *c, const char *so_path)
    {
    if (c->refK && c->ref.destroy) { c->ref.destroy(c->refK); c->refK = NULL; }
    if (c->ref.handle) { dlclose(c->ref.handle); memset(&c->ref, 0, sizeof(c->ref)); }
    if (!so_path) return 0;
    void *h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "fg_oracle: dlopen(%s): %s\n", so_path, dlerror()); return -1; }
    c->ref.handle = h;
    *(void **)&c->ref.create = dlsym(h, "fgref_matrix_create");
    *(void **)&c->ref.destroy = dlsym(h, "fgref_matrix_destroy");
    *(void **)&c->ref.clear = dlsym(h, "fgref_matrix_clear");
    *(void **)&c->ref.add = dlsym(h, "fgref_matrix_add");
    *(void **)&c->ref.set = dlsym(h, "fgref_matrix_set");
    *(void **)&c->ref.get_values = dlsym(h, "fgref_matrix_get_values");
    *(void **)&c->ref.bicg_dir = dlsym(h, "fgref_bicg_dir");
    if (!c->ref.create || !c->ref.destroy || !c->ref.clear || !c->ref.add || !c->ref.set
        || !c->ref.get_values || !c->ref.bicg_dir)
        {
        fprintf(stderr, "fg_oracle: %s lacks fgref_* symbols\n", so_path);
        dlclose(h);
        memset(&c->ref, 0, sizeof(c->ref));
        return -2;
        }
    c->refK = c->ref.create(c->n, c->rowptr, c->col);
    return 0;
    }

void fgo_sizes(const fgo_ctx *c, long long out[10])
    {
    out[0] = c->NOD; out[1] = c->NT; out[2] = c->NF; out[3] = c->n_magTet; out[4] = c->n_magTri;
    out[5] = c->n_edges; out[6] = c->n_edges_mag; out[7] = c->n; out[8] = c->nnz; out[9] = c->nlvd;
    }

void fgo_set_state(fgo_ctx *c, const double *u, const double *v, const double *phi,
                   const double *phiv)
    {
    size_t N3 = sizeof(double) * 3 * (size_t)c->NOD, N1 = sizeof(double) * (size_t)c->NOD;
    for (int k = 0; k < 2; k++)
        {
        memcpy(c->u[k], u, N3);
        if (v) memcpy(c->v[k], v, N3); else memset(c->v[k], 0, N3);
        if (phi) memcpy(c->phi[k], phi, N1); else memset(c->phi[k], 0, N1);
        if (phiv) memcpy(c->phiv[k], phiv, N1); else memset(c->phiv[k], 0, N1);
        }
    }

void fgo_set_next_v(fgo_ctx *c, const double *v)
    { memcpy(c->v[1], v, sizeof(double) * 3 * (size_t)c->NOD); }

void fgo_set_potentials_next(fgo_ctx *c, const double *phi, const double *phiv)
    {
    size_t N1 = sizeof(double) * (size_t)c->NOD;
    memcpy(c->phi[1], phi, N1);
    memcpy(c->phiv[1], phiv, N1);
    }

void fgo_get_state(const fgo_ctx *c, int step, double *u, double *v, double *phi, double *phiv)
    {
    size_t N3 = sizeof(double) * 3 * (size_t)c->NOD, N1 = sizeof(double) * (size_t)c->NOD;
    if (u) memcpy(u, c->u[step], N3);
    if (v) memcpy(v, c->v[step], N3);
    if (phi) memcpy(phi, c->phi[step], N1);
    if (phiv) memcpy(phiv, c->phiv[step], N1);
    }

void fgo_get_basis(const fgo_ctx *c, double *ep, double *eq)
    {
    size_t N3 = sizeof(double) * 3 * (size_t)c->NOD;
    memcpy(ep, c->ep, N3);
    memcpy(eq, c->eq, N3);
    }

/* Node::evolution node.h:107 via mesh.h:189-193 */
void fgo_evolution(fgo_ctx *c)
    {
    size_t N3 = sizeof(double) * 3 * (size_t)c->NOD, N1 = sizeof(double) * (size_t)c->NOD;
    memcpy(c->u[0], c->u[1], N3);
    memcpy(c->v[0], c->v[1], N3);
    memcpy(c->phi[0], c->phi[1], N1);
    memcpy(c->phiv[0], c->phiv[1], N1);
    }

void fgo_set_ext_space_field(fgo_ctx *c, const double *field)
    {
    size_t sz = sizeof(double) * 3 * (size_t)c->npi * (size_t)c->NT;
    if (!c->extSpaceField) c->extSpaceField = (double *)malloc(sz);
    memcpy(c->extSpaceField, field, sz);
    }

/* linear_algebra.cpp:3-11 → mesh.h:178-182 (parallel for_each over nodes) */
void fgo_base_projection(fgo_ctx *c, double r)
    {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->NOD; i++)
        fgo_node_set_basis(c->u[0] + 3 * (size_t)i, r, c->ep + 3 * (size_t)i, c->eq + 3 * (size_t)i);
    }

static void integrate_tet(fgo_ctx *c, int t, const double *Hext, double dt, double prefactor,
                          int idx_dir, double Vdrift)
    {
    const int *ind = c->tet_ind + 4 * (size_t)t;
    double u[12], v[12], phi[4], phiv[4], ep[12], eq[12];
    for (int i = 0; i < 4; i++)
        {
        size_t k = (size_t)ind[i];
        for (int d = 0; d < 3; d++)
            {
            u[3 * i + d] = c->u[0][3 * k + d];
            v[3 * i + d] = c->v[0][3 * k + d];
            ep[3 * i + d] = c->ep[3 * k + d];
            eq[3 * i + d] = c->eq[3 * k + d];
            }
        phi[i] = c->phi[0][k];
        phiv[i] = c->phiv[0][k];
        }
    fgo_tet_integrales(c->npi, &c->prm_tet[c->tet_reg[t]], dt, prefactor, c->tet_da + 12 * (size_t)t,
                       c->tet_w + (size_t)c->npi * t, u, v, phi, phiv, ep, eq, Hext, idx_dir, Vdrift,
                       c->tet_Kp + 64 * (size_t)t, c->tet_Lp + 8 * (size_t)t);
    }

static void integrate_tris(fgo_ctx *c)
    {
    /* linear_algebra.cpp:45-51 : all tris that are magnetic and have Ks != 0 */
#pragma omp parallel for schedule(static)
    for (int f = 0; f < c->NF; f++)
        {
        const int *ind = c->tri_ind + 3 * (size_t)f;
        int mag = c->magNode[ind[0]] && c->magNode[ind[1]] && c->magNode[ind[2]];
        if (!(mag && c->prm_tri[c->tri_reg[f]].Ks != 0)) continue;
        double u[9], ep[9], eq[9];
        for (int i = 0; i < 3; i++)
            for (int d = 0; d < 3; d++)
                {
                u[3 * i + d] = c->u[0][3 * (size_t)ind[i] + d];
                ep[3 * i + d] = c->ep[3 * (size_t)ind[i] + d];
                eq[3 * i + d] = c->eq[3 * (size_t)ind[i] + d];
                }
        fgo_tri_integrales(c->npi_tri, &c->prm_tri[c->tri_reg[f]], c->tri_dMs[f],
                           c->tri_w + (size_t)c->npi_tri * f, u, ep, eq, c->tri_Lp + 6 * (size_t)f);
        }
    }

/* linear_algebra.cpp:26-52 */
void fgo_prepare_elements(fgo_ctx *c, const double Hext[3], double dt, double prefactor,
                          int idx_dir, double Vdrift)
    {
    double H[15];
    for (int d = 0; d < 3; d++)
        for (int g = 0; g < c->npi; g++) H[d * c->npi + g] = Hext[d]; /* H.colwise() = Hext */
#pragma omp parallel for schedule(static)
    for (int t = 0; t < c->NT; t++)
        if (is_magnetic_tet(c, t)) integrate_tet(c, t, H, dt, prefactor, idx_dir, Vdrift);
    integrate_tris(c);
    }

/* linear_algebra.cpp:54-81 */
void fgo_prepare_elements_space(fgo_ctx *c, double A_Hext, double dt, double prefactor,
                                int idx_dir, double Vdrift)
    {
    if (!c->extSpaceField) { fprintf(stderr, "fg_oracle: no extSpaceField set\n"); exit(1); }
    const int sz = 3 * c->npi;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < c->NT; t++)
        if (is_magnetic_tet(c, t))
            {
            double H[15];
            for (int k = 0; k < sz; k++) H[k] = A_Hext * c->extSpaceField[(size_t)sz * t + k];
            integrate_tet(c, t, H, dt, prefactor, idx_dir, Vdrift);
            }
    integrate_tris(c);
    }

/* SparseVector::operator[] sparseMat.h:68-75 */
static inline int csr_find(const fgo_ctx *c, int i, int j)
    {
    int lo = c->rowptr[i], hi = c->rowptr[i + 1];
    while (lo < hi)
        {
        int mid = lo + (hi - lo) / 2;
        if (c->col[mid] < j) lo = mid + 1; else hi = mid;
        }
    return lo;
    }

/* solver.cpp:9-48 + :59 */
void fgo_assemble(fgo_ctx *c)
    {
    const int use_ref = (c->refK != NULL);
    /* K.clear(), sparseMat.h:93-98 */
    if (use_ref) c->ref.clear(c->refK);
    else memset(c->K, 0, sizeof(double) * (size_t)c->nnz);
    /* solver.cpp:12-17 + solver.h:110-129: parallel over magTet, K.add = lower_bound + row mutex.
     * The OpenMP stand-in uses an atomic update in place of the mutex. */
#pragma omp parallel for schedule(static)
    for (int m = 0; m < c->n_magTet; m++)
        {
        const int t = c->magTet[m];
        const int *ind = c->tet_ind + 4 * (size_t)t;
        const double *Ke = c->tet_Kp + 64 * (size_t)t;
        for (int ie = 0; ie < 4; ie++)
            {
            int i_ = ind[ie];
            for (int je = 0; je < 4; je++)
                {
                int j_ = ind[je];
                for (int di = 0; di < 2; di++)
                    for (int dj = 0; dj < 2; dj++)
                        {
                        double val = Ke[(di * 4 + ie) * 8 + (dj * 4 + je)];
                        if (use_ref)
                            c->ref.add(c->refK, 2 * i_ + di, 2 * j_ + dj, val);
                        else
                            {
                            int k = csr_find(c, 2 * i_ + di, 2 * j_ + dj);
#pragma omp atomic
                            c->K[k] += val;
                            }
                        }
                }
            }
        }
    /* solver.cpp:25-42 + solver.h:135-143: serial rhs */
    memset(c->L_rhs, 0, sizeof(double) * (size_t)c->n);
    for (int m = 0; m < c->n_magTet; m++)
        {
        const int t = c->magTet[m];
        const int *ind = c->tet_ind + 4 * (size_t)t;
        const double *Le = c->tet_Lp + 8 * (size_t)t;
        for (int ie = 0; ie < 4; ie++)
            for (int di = 0; di < 2; di++) c->L_rhs[2 * ind[ie] + di] += Le[di * 4 + ie];
        }
    for (int m = 0; m < c->n_magTri; m++)
        {
        const int f = c->magTri[m];
        const int *ind = c->tri_ind + 3 * (size_t)f;
        const double *Le = c->tri_Lp + 6 * (size_t)f;
        for (int ie = 0; ie < 3; ie++)
            for (int di = 0; di < 2; di++) c->L_rhs[2 * ind[ie] + di] += Le[di * 3 + ie];
        }
    /* solver.cpp:46-48 */
    v_mask(c->lvd, c->nlvd, c->L_rhs);
    for (int k = 0; k < c->nlvd; k++)
        {
        int i = c->lvd[k];
        if (use_ref) c->ref.set(c->refK, i, i, 1.0);
        else c->K[csr_find(c, i, i)] = 1.0;
        }
    if (use_ref) c->ref.get_values(c->refK, c->K); /* keep the tap coherent */
    /* buildInitGuess linear_algebra.cpp:13-24 : proj of d[NEXT].v on (ep, eq), / gamma0 */
    for (int i = 0; i < c->n; i++) c->Xw[i] = 0;
    for (int i = 0; i < c->NOD; i++)
        if (c->magNode[i])
            {
            c->Xw[2 * i] = dot3(c->v[1] + 3 * (size_t)i, c->ep + 3 * (size_t)i) / FGO_GAMMA0;
            c->Xw[2 * i + 1] = dot3(c->v[1] + 3 * (size_t)i, c->eq + 3 * (size_t)i) / FGO_GAMMA0;
            }
    }

/* solver.cpp:6-90 */
int fgo_solve(fgo_ctx *c, double dt)
    {
    it_reset(&c->iter);
    fgo_assemble(c);
    if (c->refK)
        c->ref.bicg_dir(c->refK, c->Xw, c->L_rhs, c->n, c->lvd, c->nlvd, c->iter.resmax,
                        c->iter.maxiter, &c->iter.status, &c->iter.nit, &c->iter.res,
                        &c->iter.rhsn);
    else
        fgo_bicg_dir(&c->iter, c->n, c->rowptr, c->col, c->K, c->Xw, c->L_rhs, c->lvd, c->nlvd);

    if ((c->iter.status == FGO_ITER_OVERFLOW) || (c->iter.status == FGO_CANNOT_CONVERGE)
        || (c->iter.res > c->iter.resmax))
        return 1;

    double v2max = 0.0;
    for (int i = 0; i < c->NOD; i++)
        if (c->magNode[i])
            {
            double vp = c->Xw[2 * i], vq = c->Xw[2 * i + 1];
            double v2 = vp * vp + vq * vq;
            if (v2 > v2max) v2max = v2;
            /* mesh.h:185-186: make_evol(vp*gamma0, vq*gamma0, dt) */
            fgo_node_make_evol(c->u[0] + 3 * (size_t)i, c->ep + 3 * (size_t)i, c->eq + 3 * (size_t)i,
                               vp * FGO_GAMMA0, vq * FGO_GAMMA0, dt, c->u[1] + 3 * (size_t)i,
                               c->v[1] + 3 * (size_t)i);
            }
    c->v_max = FGO_GAMMA0 * sqrt(v2max);
    return 0;
    }

double fgo_get_v_max(const fgo_ctx *c) { return c->v_max; }
void fgo_get_iter(const fgo_ctx *c, fgo_iter *out) { *out = c->iter; }

void fgo_get_tet_ind(const fgo_ctx *c, int *ind)
    { memcpy(ind, c->tet_ind, sizeof(int) * 4 * (size_t)c->NT); }
void fgo_get_tet_geom(const fgo_ctx *c, double *da, double *weight)
    {
    if (da) memcpy(da, c->tet_da, sizeof(double) * 12 * (size_t)c->NT);
    if (weight) memcpy(weight, c->tet_w, sizeof(double) * (size_t)c->npi * (size_t)c->NT);
    }
void fgo_get_element(const fgo_ctx *c, int tet, double Kp[64], double Lp[8])
    {
    memcpy(Kp, c->tet_Kp + 64 * (size_t)tet, sizeof(double) * 64);
    memcpy(Lp, c->tet_Lp + 8 * (size_t)tet, sizeof(double) * 8);
    }
void fgo_get_tri_element(const fgo_ctx *c, int tri, double Lp[6])
    { memcpy(Lp, c->tri_Lp + 6 * (size_t)tri, sizeof(double) * 6); }
void fgo_get_csr(const fgo_ctx *c, int *rowptr, int *col)
    {
    memcpy(rowptr, c->rowptr, sizeof(int) * ((size_t)c->n + 1));
    memcpy(col, c->col, sizeof(int) * (size_t)c->nnz);
    }
void fgo_get_system(const fgo_ctx *c, double *val, double *rhs, double *x)
    {
    if (val) memcpy(val, c->K, sizeof(double) * (size_t)c->nnz);
    if (rhs) memcpy(rhs, c->L_rhs, sizeof(double) * (size_t)c->n);
    if (x) memcpy(x, c->Xw, sizeof(double) * (size_t)c->n);
    }
void fgo_get_masks(const fgo_ctx *c, unsigned char *magNode, int *lvd)
    {
    if (magNode) memcpy(magNode, c->magNode, (size_t)c->NOD);
    if (lvd) memcpy(lvd, c->lvd, sizeof(int) * (size_t)c->nlvd);
    }
void fgo_get_edges(const fgo_ctx *c, int *edges)
    { memcpy(edges, c->edges, sizeof(int) * 2 * (size_t)c->n_edges); }

/* ------------------------------------------------------------------------------------------ */
/* "next" rows: energies (src/energy.cpp:5-68, src/tetra.cpp:309-391, src/triangle.cpp:38-43,   */
/* 80-85), averages (src/mesh.cpp:89-106), max_angle (src/mesh.h:295-306)                       */
/* ------------------------------------------------------------------------------------------ */
static void energy_core(const fgo_ctx *c, const double *Hext, int space, double fieldAmp, double E[4])
    {
    const int npi = c->npi;
    const double *a = fgo_tet_a(npi);
    E[0] = E[1] = E[2] = E[3] = 0.0;
    for (int m = 0; m < c->n_magTet; m++)
        {
        const int t = c->magTet[m];
        const int *ind = c->tet_ind + 4 * (size_t)t;
        const double *da = c->tet_da + 12 * (size_t)t, *w = c->tet_w + (size_t)npi * t;
        const fgo_tet_prm *prm = &c->prm_tet[c->tet_reg[t]];
        double u[15], dudx[3], dudy[3], dudz[3], phi[5];
        for (int d = 0; d < 3; d++)
            {
            double sx = 0, sy = 0, sz = 0;
            for (int i = 0; i < 4; i++)
                {
                double un = c->u[1][3 * (size_t)ind[i] + d];
                sx += un * da[3 * i + 0];
                sy += un * da[3 * i + 1];
                sz += un * da[3 * i + 2];
                }
            dudx[d] = sx; dudy[d] = sy; dudz[d] = sz;
            for (int g = 0; g < npi; g++)
                {
                double s = 0;
                for (int i = 0; i < 4; i++) s += c->u[1][3 * (size_t)ind[i] + d] * a[i * npi + g];
                u[d * npi + g] = s;
                }
            }
        for (int g = 0; g < npi; g++)
            {
            double s = 0;
            for (int i = 0; i < 4; i++) s += c->phi[1][ind[i]] * a[i * npi + g];
            phi[g] = s;
            }
        /* exchangeEnergy tetra.cpp:309-318 */
        double dens_ex = dot3(dudx, dudx) + dot3(dudy, dudy) + dot3(dudz, dudz), s = 0;
        for (int g = 0; g < npi; g++) s += w[g] * dens_ex;
        E[0] += prm->A * s;
        /* demagEnergy tetra.cpp:361-372 */
        s = 0;
        for (int g = 0; g < npi; g++) s += w[g] * ((dudx[0] + dudy[1] + dudz[2]) * phi[g]);
        E[2] += -0.5 * FGO_MU0 * prm->Ms * s;
        /* uniaxialAnisotropyEnergy tetra.cpp:320-328 */
        if (prm->K != 0.0)
            {
            s = 0;
            for (int g = 0; g < npi; g++)
                {
                double ug[3] = {u[g], u[npi + g], u[2 * npi + g]};
                double q = dot3(prm->uk, ug);
                s += w[g] * (q * q);
                }
            E[1] += -prm->K * s;
            }
        /* cubicAnisotropyEnergy tetra.cpp:330-345 */
        if (prm->K3 != 0.0)
            {
            s = 0;
            for (int g = 0; g < npi; g++)
                {
                double ug[3] = {u[g], u[npi + g], u[2 * npi + g]};
                double al0 = dot3(ug, prm->ex), al1 = dot3(ug, prm->ey), al2 = dot3(ug, prm->ez);
                s += w[g] * ((al0 * al1) * (al0 * al1) + (al1 * al2) * (al1 * al2)
                             + (al2 * al0) * (al2 * al0));
                }
            E[1] += prm->K3 * s;
            }
        s = 0;
        if (!space)
            {   /* zeemanEnergy (uniform field) tetra.cpp:374-380 */
            for (int g = 0; g < npi; g++)
                {
                double ug[3] = {u[g], u[npi + g], u[2 * npi + g]};
                s += w[g] * dot3(ug, Hext);
                }
            E[3] += -FGO_MU0 * prm->Ms * s;
            }
        else
            {   /* zeemanEnergy (space field x amplitude) tetra.cpp:382-391, energy.cpp:41-43 */
            const double *sf = c->extSpaceField + (size_t)3 * npi * t; /* [d][g] */
            for (int g = 0; g < npi; g++)
                s += w[g] * (u[g] * sf[g] + u[npi + g] * sf[npi + g] + u[2 * npi + g] * sf[2 * npi + g]);
            E[3] += -FGO_MU0 * prm->Ms * fieldAmp * s;
            }
        }
    const int npt = c->npi_tri;
    const double *at = fgo_tri_a(npt);
    for (int m = 0; m < c->n_magTri; m++)
        {
        const int f = c->magTri[m];
        const int *ind = c->tri_ind + 3 * (size_t)f;
        const fgo_tri_prm *prm = &c->prm_tri[c->tri_reg[f]];
        const double *w = c->tri_w + (size_t)npt * f;
        double surf, nrm[3], wtmp[4];
        fgo_tri_setup(c->p, ind, npt, &surf, nrm, wtmp);
        double s_an = 0, s_dm = 0;
        for (int g = 0; g < npt; g++)
            {
            double ug[3] = {0, 0, 0}, ph = 0;
            for (int i = 0; i < 3; i++)
                {
                for (int d = 0; d < 3; d++) ug[d] += c->u[1][3 * (size_t)ind[i] + d] * at[i * npt + g];
                ph += c->phi[1][ind[i]] * at[i * npt + g];
                }
            double q = dot3(ug, prm->uk);
            s_an += w[g] * (q * q);          /* triangle.cpp:38-43 */
            s_dm += (dot3(ug, nrm) * ph) * w[g]; /* triangle.cpp:80-85 */
            }
        if (prm->Ks != 0.0) E[1] += -prm->Ks * s_an;
        E[2] += 0.5 * FGO_MU0 * c->tri_dMs[f] * s_dm;
        }
    }

void fgo_energy(const fgo_ctx *c, const double Hext[3], double E[4]) { energy_core(c, Hext, 0, 0.0, E); }
void fgo_energy_space(const fgo_ctx *c, double fieldAmp, double E[4]) { energy_core(c, NULL, 1, fieldAmp, E); }

double fgo_total_mag_vol(const fgo_ctx *c)
    {
    double vol = 0; /* mesh.h:81-90 */
    for (int t = 0; t < c->NT; t++)
        if (is_magnetic_tet(c, t))
            {
            double s = 0;
            for (int g = 0; g < c->npi; g++) s += c->tet_w[(size_t)c->npi * t + g];
            vol += s;
            }
    return vol;
    }

/* settings.paramTetra[region].volume, src/mesh.h:81-90 (all tets of the region, magnetic or not) */
double fgo_region_vol(const fgo_ctx *c, int region)
    {
    double vol = 0;
    for (int t = 0; t < c->NT; t++)
        if (c->tet_reg[t] == region)
            {
            double s = 0;
            for (int g = 0; g < c->npi; g++) s += c->tet_w[(size_t)c->npi * t + g];
            vol += s;
            }
    return vol;
    }

void fgo_avg(const fgo_ctx *c, int what, double out[3]) { fgo_avg_region(c, what, -1, out); }

void fgo_avg_region(const fgo_ctx *c, int what, int region, double out[3])
    {
    const int npi = c->npi;
    const double *a = fgo_tet_a(npi);
    const double *field = what == 0 ? c->u[1] : c->v[1];
    double vol = region == -1 ? fgo_total_mag_vol(c) : fgo_region_vol(c, region);
    for (int d = 0; d < 3; d++)
        {
        double sum = 0;
        for (int m = 0; m < c->n_magTet; m++)
            {
            const int t = c->magTet[m];
            if (c->tet_reg[t] != region && region != -1) continue;
            const int *ind = c->tet_ind + 4 * (size_t)t;
            const double *w = c->tet_w + (size_t)npi * t;
            double s = 0;
            for (int g = 0; g < npi; g++)
                {
                double val = 0;
                for (int i = 0; i < 4; i++) val += field[3 * (size_t)ind[i] + d] * a[i * npi + g];
                s += w[g] * val;
                }
            sum += s;
            }
        out[d] = sum / vol;
        }
    }

double fgo_max_angle(const fgo_ctx *c)
    {
    double min_dot = 1.0;
    for (int e = 0; e < c->n_edges; e++)
        {
        double d = dot3(c->u[1] + 3 * (size_t)c->edges[2 * e], c->u[1] + 3 * (size_t)c->edges[2 * e + 1]);
        /* std::min(a, b) with NaN from non-magnetic nodes: (b < a) ? b : a keeps a */
        if (d < min_dot) min_dot = d;
        }
    return acos(min_dot);
    }

/* ------------------------------------------------------------------------------------------ */
/* "next" rows rank 2: magnetic charges feeding the demag solver (src/tetra.cpp:347-359,        */
/* src/triangle.cpp:45-78,87-125, orchestration src/fmm_demag.h:155-223)                        */
/* ------------------------------------------------------------------------------------------ */
/* Tet::charges, src/tetra.cpp:347-359: -Ms * weight * div(vec) */
void fgo_tet_charges(int npi, double Ms, const double da[12], const double *weight,
                     const double *vec_nod /* [i*3+d] */, double *out)
    {
    double sx = 0, sy = 0, sz = 0;
    for (int i = 0; i < 4; i++) sx += vec_nod[3 * i + 0] * da[3 * i + 0];
    for (int i = 0; i < 4; i++) sy += vec_nod[3 * i + 1] * da[3 * i + 1];
    for (int i = 0; i < 4; i++) sz += vec_nod[3 * i + 2] * da[3 * i + 2];
    const double dud_sum = sx + sy + sz;
    for (int g = 0; g < npi; g++) out[g] = -Ms * weight[g] * dud_sum;
    }

/* Tri::charges, src/triangle.cpp:45-59 */
void fgo_tri_charges(int npi, double dMs, const double n[3], const double *weight,
                     const double *vec_nod /* [i*3+d] */, double *out)
    {
    const double *a = fgo_tri_a(npi);
    for (int g = 0; g < npi; g++) out[g] = 0.0;
    if (dMs == 0.0) return;
    for (int g = 0; g < npi; g++)
        {
        double ug[3] = {0, 0, 0};
        for (int d = 0; d < 3; d++)
            for (int i = 0; i < 3; i++) ug[d] += vec_nod[3 * i + d] * a[i * npi + g];
        out[g] = dMs * (weight[g] * dot3(ug, n));
        }
    }

/* Tri::potential, src/triangle.cpp:87-125: analytic potential at local node i of the linear
 * surface charge of the triangle (p: 3 nodes x 3, vec_nod [i*3+d]) */
double fgo_tri_potential(const double *p /* 9 */, const double *vec_nod /* 9 */, double surf,
                         const double n[3], double dMs, int i)
    {
    const int ii = (i + 1) % 3, iii = (i + 2) % 3;
    const double *p1 = p + 3 * i, *p2 = p + 3 * ii, *p3 = p + 3 * iii;
    double p1p2[3], p1p3[3];
    for (int k = 0; k < 3; k++) { p1p2[k] = p2[k] - p1[k]; p1p3[k] = p3[k] - p1[k]; }
    const double b = sqrt(dot3(p1p2, p1p2));
    const double t = dot3(p1p2, p1p3) / b;
    const double _2s = 2. * surf;
    const double h = _2s / b;
    const double c = (t - b) / h;
    const double fth = sqrt(1.0 + (t / h) * (t / h)), fc = sqrt(1.0 + c * c);
    const double r = h * fth;
    const double log_1 = log((c * t + h + fc * r) / (b * (c + fc)));
    const double xi = b * log_1 / fc;
    const double s1 = dot3(vec_nod + 3 * i, n), s2 = dot3(vec_nod + 3 * ii, n), s3 = dot3(vec_nod + 3 * iii, n);
    const double pot = xi * s1
                       + ((xi * (h + c * t) - b * (r - b)) * s2 + b * (r - b - c * xi) * s3) * b
                                 / (_2s * (1 + c * c));
    return 0.5 * dMs * pot;
    }

/* number of sources of fmm::calc_charges: magTet*NPI + magTri*NPI_tri (src/fmm_demag.h:88) */
long long fgo_n_sources(const fgo_ctx *c)
    { return (long long)c->n_magTet * c->npi + (long long)c->n_magTri * c->npi_tri; }

/* source positions = Gauss points (getPtGauss, src/tetra.h:344-352, src/triangle.h:209-218) in the
 * order of insertCharges (src/fmm_demag.h:82-83): tets of magTet, then triangles of magTri */
void fgo_source_positions(const fgo_ctx *c, double *pos /* nsrc x 3 */)
    {
    const double *a = fgo_tet_a(c->npi), *at = fgo_tri_a(c->npi_tri);
    size_t k = 0;
    for (int m = 0; m < c->n_magTet; m++)
        {
        const int *ind = c->tet_ind + 4 * (size_t)c->magTet[m];
        for (int g = 0; g < c->npi; g++, k++)
            for (int d = 0; d < 3; d++)
                {
                double s = 0;
                for (int i = 0; i < 4; i++) s += c->p[3 * (size_t)ind[i] + d] * a[i * c->npi + g];
                pos[3 * k + d] = s;
                }
        }
    for (int m = 0; m < c->n_magTri; m++)
        {
        const int *ind = c->tri_ind + 3 * (size_t)c->magTri[m];
        for (int g = 0; g < c->npi_tri; g++, k++)
            for (int d = 0; d < 3; d++)
                {
                double s = 0;
                for (int i = 0; i < 3; i++) s += c->p[3 * (size_t)ind[i] + d] * at[i * c->npi_tri + g];
                pos[3 * k + d] = s;
                }
        }
    }

/* fmm::calc_charges, src/fmm_demag.h:155-185, on the NEXT state: which = 0 u | 1 v */
void fgo_calc_charges(const fgo_ctx *c, int which, double *srcDen, double *corr)
    {
    const double *field = which == 0 ? c->u[1] : c->v[1];
    size_t nsrc = 0;
    for (int m = 0; m < c->n_magTet; m++)
        {
        const int t = c->magTet[m];
        const int *ind = c->tet_ind + 4 * (size_t)t;
        double vn[12];
        for (int i = 0; i < 4; i++)
            for (int d = 0; d < 3; d++) vn[3 * i + d] = field[3 * (size_t)ind[i] + d];
        fgo_tet_charges(c->npi, c->prm_tet[c->tet_reg[t]].Ms, c->tet_da + 12 * (size_t)t,
                        c->tet_w + (size_t)c->npi * t, vn, srcDen + nsrc);
        nsrc += c->npi;
        }
    for (int i = 0; i < c->NOD; i++) corr[i] = 0.0;
    const double *at = fgo_tri_a(c->npi_tri);
    for (int m = 0; m < c->n_magTri; m++)
        {
        const int f = c->magTri[m];
        const int *ind = c->tri_ind + 3 * (size_t)f;
        double surf, nrm[3], w[4], vn[9], pp[9], q[4];
        fgo_tri_setup(c->p, ind, c->npi_tri, &surf, nrm, w);
        for (int i = 0; i < 3; i++)
            for (int d = 0; d < 3; d++)
                {
                vn[3 * i + d] = field[3 * (size_t)ind[i] + d];
                pp[3 * i + d] = c->p[3 * (size_t)ind[i] + d];
                }
        fgo_tri_charges(c->npi_tri, c->tri_dMs[f], nrm, w, vn, q);
        for (int g = 0; g < c->npi_tri; g++) srcDen[nsrc + g] = q[g];
        nsrc += c->npi_tri;
        /* Tri::correctionCharges, src/triangle.cpp:61-78 */
        for (int i = 0; i < 3; i++)
            {
            for (int g = 0; g < c->npi_tri; g++)
                {
                double gp[3] = {0, 0, 0}, dd[3];
                for (int d = 0; d < 3; d++)
                    for (int k = 0; k < 3; k++) gp[d] += pp[3 * k + d] * at[k * c->npi_tri + g];
                for (int d = 0; d < 3; d++) dd[d] = pp[3 * i + d] - gp[d];
                corr[ind[i]] -= q[g] / sqrt(dot3(dd, dd));
                }
            corr[ind[i]] += fgo_tri_potential(pp, vn, surf, nrm, c->tri_dMs[f], i);
            }
        }
    }

/* All-pairs stand-in for scal_fmm::fmm::demag (src/fmm_demag.h:187-223): the potential the FMM
 * approximates, phi_i = (sum_j q_j / |p_i - x_j| + corr_i) / (4 pi) on the magnetic nodes, written to
 * NEXT phi (which = 0, from u) or phiv (which = 1, from v).  ScalFMM itself (third party, rev
 * 22b9e4f6cf, P = 9) is absent here; this is what it converges to. */
void fgo_demag_direct(fgo_ctx *c, int which)
    {
    const long long ns = fgo_n_sources(c);
    double *src = (double *)malloc(sizeof(double) * (size_t)(ns > 0 ? ns : 1));
    double *pos = (double *)malloc(sizeof(double) * 3 * (size_t)(ns > 0 ? ns : 1));
    double *corr = (double *)malloc(sizeof(double) * (size_t)c->NOD);
    fgo_calc_charges(c, which, src, corr);
    fgo_source_positions(c, pos);
    double *out = which == 0 ? c->phi[1] : c->phiv[1];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->NOD; i++)
        {
        if (!c->magNode[i]) continue;
        const double *pi = c->p + 3 * (size_t)i;
        double s = 0.0;
        for (long long j = 0; j < ns; j++)
            {
            const double dx = pi[0] - pos[3 * j], dy = pi[1] - pos[3 * j + 1], dz = pi[2] - pos[3 * j + 2];
            s += src[j] / sqrt(dx * dx + dy * dy + dz * dz);
            }
        out[i] = (s + corr[i]) / (4 * 3.14159265358979323846);
        }
    free(src);
    free(pos);
    free(corr);
    }

/* scal_fmm::fmm::calc_demag, src/fmm_demag.h:98-103 (FIRST_ORDER off: second call for phiv) */
void fgo_calc_demag_direct(fgo_ctx *c, int second_order)
    {
    fgo_demag_direct(c, 0);
    if (second_order) fgo_demag_direct(c, 1);
    }
