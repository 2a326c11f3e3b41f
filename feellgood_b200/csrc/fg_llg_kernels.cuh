// fg_llg_kernels.cuh — device code of the per-time-step LLG kernels (sm_100a, FP64).
//
//   k_basis          Node::setBasis                  reference src/node.h:73-102
//   k_tet            Tet::integrales                 reference src/tetra.cpp:210-307
//   k_tri            Tri::integrales                 reference src/triangle.cpp:6-36
//   k_assemble_sell  solver::buildMat/buildVect + mask + buildInitGuess + build_diag_precond
//                    reference src/solver.cpp:9-59, src/solver.h:110-143, sparseMat.h:174-183
//   k_update         node update + v_max             reference src/solver.cpp:74-88, node.h:116-122
//
// Structure exploited.  With AE v = E v + a_w (m x v) (src/tetra.cpp:108-148) and
// P depending on the node only, the projected element matrix is
//     Kp[(r,a),(c,b)] = E_ab (e_r,a . e_c,b) + delta_ab a_w,a e_r,a . (m_a x e_c,a),
//     E_ab = prefactor s_dt Abis wsum (grad a_a . grad a_b) + delta_ab sum_g a_a(g) w_g alpha_eff(g)
// so the global K is the projection of a NODxNOD scalar matrix whose off-diagonal part is constant
// per mesh (S, built once) and whose diagonal gains one state-dependent number per node per step
// (Malpha).  The per-step element kernel therefore emits, per (tet, local node), one 32-byte
// record {sum_g a w alpha_eff, BE(3)}; the assembly gathers the records of a node through its
// incidence list (no atomics, fixed order => deterministic), projects the summed BE on the node's
// basis once (L = (eq . BE, ep . BE)) and writes the node's part of the system.
#pragma once
#include "fg_common.cuh"
#include "fg_reduce.cuh"
#include "fg_krylov_state.cuh"

namespace fg
{
// Gauss tables (src/tetra.h:29-81, src/triangle.h:21-65), filled by fg_create
__constant__ double c_tet_a5[20], c_tet_pds5[5], c_tet_a1[4], c_tet_pds1[1];
__constant__ double c_tri_a4[12], c_tri_pds4[4], c_tri_a1[3], c_tri_pds1[1];

// The element math below (tet_core, tet_iso_front, tet_iso_be, alpha_eff) is __host__ __device__ so that
// tests/cpp/device_math_test.cu can run it on the CPU; on the host the Gauss tables come from the same
// tet_tables() that fills the __constant__ copies (fg_setup.cpp).  Device code is unchanged by this.
void tet_tables(int npi, double a[20], double pds[5]);
#ifndef __CUDA_ARCH__
inline double host_tet_table(int npi, bool weights, int idx)
    {
    static double a5[20], p5[5], a1[4], p1[1];
    static const bool init = (tet_tables(5, a5, p5), tet_tables(1, a1, p1), true);
    (void)init;
    return npi == 5 ? (weights ? p5[idx] : a5[idx]) : (weights ? p1[idx] : a1[idx]);
    }
#endif
template <int NPI> __host__ __device__ __forceinline__ double tet_a(int i, int g)
    {
#ifdef __CUDA_ARCH__
    return NPI == 5 ? c_tet_a5[i * 5 + g] : c_tet_a1[i];
#else
    return host_tet_table(NPI, false, NPI == 5 ? i * 5 + g : i);
#endif
    }
template <int NPI> __host__ __device__ __forceinline__ double tet_pds(int g)
    {
#ifdef __CUDA_ARCH__
    return NPI == 5 ? c_tet_pds5[g] : c_tet_pds1[0];
#else
    return host_tet_table(NPI, true, NPI == 5 ? g : 0);
#endif
    }
// one rounding, whatever the compiler contracts around it
__host__ __device__ __forceinline__ double mul_rn(double a, double b)
    {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
    }
void tri_tables(int npi, double a[12], double pds[4]);
#ifndef __CUDA_ARCH__
inline double host_tri_table(int npi, bool weights, int idx)
    {
    static double a4[12], p4[4], a1[3], p1[1];
    static const bool init = (tri_tables(4, a4, p4), tri_tables(1, a1, p1), true);
    (void)init;
    return npi == 4 ? (weights ? p4[idx] : a4[idx]) : (weights ? p1[idx] : a1[idx]);
    }
#endif
template <int NPI> __host__ __device__ __forceinline__ double tri_a(int i, int g)
    {
#ifdef __CUDA_ARCH__
    return NPI == 4 ? c_tri_a4[i * 4 + g] : c_tri_a1[i];
#else
    return host_tri_table(NPI, false, NPI == 4 ? i * 4 + g : i);
#endif
    }
template <int NPI> __host__ __device__ __forceinline__ double tri_pds(int g)
    {
#ifdef __CUDA_ARCH__
    return NPI == 4 ? c_tri_pds4[g] : c_tri_pds1[0];
#else
    return host_tri_table(NPI, true, NPI == 4 ? g : 0);
#endif
    }

__host__ __device__ __forceinline__ double dot3(const double *a, const double *b)
    { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__host__ __device__ __forceinline__ void cross3(const double *a, const double *b, double *r)
    {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
    }
// Eigen normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z)
__host__ __device__ __forceinline__ void normalize3(double *a)
    {
    const double z = dot3(a, a);
    if (z > 0.0)
        {
        const double s = sqrt(z);
        a[0] /= s;
        a[1] /= s;
        a[2] /= s;
        }
    }

__device__ __forceinline__ void load_rec(const NodeRec *p, double u[3], double v[3], double &phi,
                                         double &phiv)
    {  // 64-byte aligned record: two 256-bit requests (LDG.E.ENL2.256)
    const double4 *q = reinterpret_cast<const double4 *>(p);
    const double4 a = ld256_nc(q), b = ld256_nc(q + 1);
    u[0] = a.x; u[1] = a.y; u[2] = a.z;
    v[0] = a.w; v[1] = b.x; v[2] = b.y;
    phi = b.z; phiv = b.w;
    }
__device__ __forceinline__ void load_basis(const Basis *p, double ep[3], double eq[3])
    {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    ep[0] = a.x; ep[1] = a.y; ep[2] = b.x;
    eq[0] = b.y; eq[1] = c.x; eq[2] = c.y;
    }

// ------------------------------------------------------------------------------------------
// Node::setBasis, src/node.h:73-102.  cr = cos(r), sr = sin(r) are evaluated by the host libm so
// that the rotation uses the very numbers the reference would.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void node_set_basis(const double u[3], double cr, double sr,
                                               double ep[3], double eq[3])
    {
    int k = 0;
    double m = fabs(u[0]);
    if (fabs(u[1]) < m) { m = fabs(u[1]); k = 1; }
    if (fabs(u[2]) < m) { k = 2; }
    double e[3] = {0.0, 0.0, 0.0};
    e[k] = 1.0;
    const double d = dot3(e, u);
    e[0] -= d * u[0];
    e[1] -= d * u[1];
    e[2] -= d * u[2];
    normalize3(e);
    double f[3];
    cross3(u, e, f);
#pragma unroll
    for (int c = 0; c < 3; c++)
        {
        ep[c] = cr * e[c] - sr * f[c];
        eq[c] = sr * e[c] + cr * f[c];
        }
    }

// fg_set_state: max over the magnetic nodes of | |u|^2 - 1 | (non-negative doubles order like their bit
// patterns, so an integer atomicMax does it; NaN compares above everything and is caught as well)
__global__ void __launch_bounds__(BLOCK)
k_check_unit(int NOD, const unsigned char *__restrict__ nonmag, const NodeRec *__restrict__ cur,
             unsigned long long *out)
    {
    const int a = blockIdx.x * BLOCK + threadIdx.x;
    double dev = 0.0;
    if (a < NOD && !nonmag[a])
        {
        const double *u = cur[a].u;
        dev = fabs(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] - 1.0);
        if (!(dev == dev)) dev = 1.7976931348623157e308;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    if ((threadIdx.x & 31) == 0 && dev > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(dev));
    }

// Also writes the basis as a unit quaternion (32 B): what the Krylov kernels read (fg_common.cuh).
// COMMIT: mesh::evolution (src/mesh.h:189-193) fused in: `src` is the NEXT record, copied to CURRENT (`dst`)
// on the way (fg_commit is lazy), so the copy costs one 64-byte write per node instead of a pass of its own.
template <bool COMMIT>
__global__ void __launch_bounds__(BLOCK)
k_basis(int NOD, const NodeRec *__restrict__ src, NodeRec *dst, double cr, double sr, Basis *__restrict__ basis,
        double4 *__restrict__ qbasis)
    {
    const int stride = gridDim.x * BLOCK;
    for (int a = blockIdx.x * BLOCK + threadIdx.x; a < NOD; a += stride)
        {
        double u[3], v[3], phi, phiv, ep[3], eq[3];
        if (COMMIT)
            {
            const double4 *q = reinterpret_cast<const double4 *>(src + a);
            const double4 r0 = ld256_nc(q), r1 = ld256_nc(q + 1);
            double4 *o = reinterpret_cast<double4 *>(dst + a);
            st256(o, r0);
            st256(o + 1, r1);
            u[0] = r0.x; u[1] = r0.y; u[2] = r0.z;
            }
        else
            load_rec(src + a, u, v, phi, phiv);
        node_set_basis(u, cr, sr, ep, eq);
        double2 *q = reinterpret_cast<double2 *>(basis + a);
        q[0] = make_double2(ep[0], ep[1]);
        q[1] = make_double2(ep[2], eq[0]);
        q[2] = make_double2(eq[1], eq[2]);
        st256(qbasis + a, basis_to_quat(ep, eq));
        }
    }

// ------------------------------------------------------------------------------------------
// Tet::integrales, src/tetra.cpp:210-307
// ------------------------------------------------------------------------------------------
// src/tetra.cpp:47-75
__host__ __device__ __forceinline__ double alpha_eff(double dt, double alpha, double h)
    {
    const double reduced_dt = FG_GAMMA0 * dt;
    const double r = 0.1;
    const double M = 2. * alpha * r / reduced_dt;
    if (h > 0.)
        return (h > M) ? alpha + reduced_dt / 2. * M : alpha + reduced_dt / 2. * h;
    return (h < -M) ? alpha / (1. + reduced_dt / (2. * alpha) * M)
                    : alpha / (1. - reduced_dt / (2. * alpha) * h);
    }

struct TetIn
    {
    double da[4][3];
    double detJ;
    double u[4][3], v[4][3], phi[4], phiv[4];
    };

struct StepPrm
    {
    double dt, prefactor, Hext[3], A_Hext, Vdrift;
    int idx_dir;
    };

// Element core shared by the production kernel and the Kp/Lp tap.  Hext is [d][g].
// Outputs: contrib[i] = sum_g a_i(g) w_g alpha_eff(g)   (the state-dependent diagonal of E)
//          BE[d][i]                                      (src/tetra.cpp:277-303)
// Organised Gauss-point-outermost so that only the accumulators (contrib, BE), the node values and
// the per-element constants (grad U, Hd, Hv, the exchange products Ex) stay live: the reference's
// per-point tables U, V, H, H_aniso (4 x 3 x NPI doubles) never exist.  Every sum keeps the
// reference's order of accumulation.
// CUBIC / DRIFT = false compile the cubic-anisotropy and the recentring-drift branches out (the caller
// guarantees that no region has K3 / that idx_dir is undefined): same arithmetic on what remains, fewer
// live registers (the general kernel needs 254).
template <int NPI, bool CUBIC = true, bool DRIFT = true>
__host__ __device__ __forceinline__ void tet_core(const TetIn &T, const TetRegion &R, const StepPrm &sp,
                                         const double (&Hext)[3][NPI], double contrib[4],
                                         double BE[3][4])
    {
    const double alpha = R.alpha, Abis = R.Abis;
    const double s_dt = FG_THETA * sp.dt * FG_GAMMA0;
    const double th_dt = s_dt / FG_GAMMA0;  // what calc_aniso_* and Hv receive (tetra.cpp:240,292)

    // interpolation of the element-wise constants, src/tetra.h:183-218
    double dU[3][3], Hd[3], Hv[3];
#pragma unroll
    for (int d = 0; d < 3; d++)
        {
#pragma unroll
        for (int k = 0; k < 3; k++)
            {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) s += T.u[i][d] * T.da[i][k];
            dU[d][k] = s;  // dUd{x,y,z}(d) for k = 0,1,2
            }
        double hd = 0.0, hv = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            hd -= T.phi[i] * T.da[i][d];
            hv -= T.phiv[i] * T.da[i][d];
            }
        Hd[d] = hd;
        Hv[d] = hv;
        }
    // tetra.cpp:232-233
    double gsq = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++)
        gsq += dU[0][k] * dU[0][k] + dU[1][k] * dU[1][k] + dU[2][k] * dU[2][k];
    // exchange products da_i . grad U_d of tetra.cpp:296-297
    double Ex[3][4];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int i = 0; i < 4; i++)
            Ex[d][i] = T.da[i][0] * dU[d][0] + T.da[i][1] * dU[d][1] + T.da[i][2] * dU[d][2];

#pragma unroll
    for (int i = 0; i < 4; i++) contrib[i] = 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int i = 0; i < 4; i++) BE[d][i] = 0.0;

    const bool drift = DRIFT && sp.idx_dir != FG_IDX_UNDEF;
    double dUk[3] = {0, 0, 0}, dVk[3] = {0, 0, 0};
    if (drift)  // add_drift_BE, tetra.cpp:150-169 (accumulated for all points before the fields)
        {
        const int k = sp.idx_dir;
#pragma unroll
        for (int d = 0; d < 3; d++)
            {
            dUk[d] = (k == 2) ? dU[d][2] : (k == 1 ? dU[d][1] : dU[d][0]);
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) s += T.v[i][d] * ((k == 2) ? T.da[i][2] : (k == 1 ? T.da[i][1] : T.da[i][0]));
            dVk[d] = s;
            }
#pragma unroll
        for (int g = 0; g < NPI; g++)
            {
            const double w = T.detJ * tet_pds<NPI>(g);
            double Ug[3], Vg[3];
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                double su = 0.0, sv = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    {
                    su += T.u[i][d] * tet_a<NPI>(i, g);
                    sv += T.v[i][d] * tet_a<NPI>(i, g);
                    }
                Ug[d] = su;
                Vg[d] = sv;
                }
            double c1[3], c2[3], c3[3];
            cross3(Ug, dUk, c1);
            cross3(Ug, dVk, c2);
            cross3(Vg, dUk, c3);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int d = 0; d < 3; d++)
                    {
                    const double interim = tet_a<NPI>(i, g)
                        * (alpha * dUk[d] + c1[d] + s_dt * (alpha * dVk[d] + c2[d] + c3[d]));
                    BE[d][i] += sp.Vdrift * w * interim;
                    }
            }
        }

#pragma unroll
    for (int g = 0; g < NPI; g++)
        {
        const double w = T.detJ * tet_pds<NPI>(g);
        double Ug[3], Vg[3] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < 3; d++)
            {
            double su = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) su += T.u[i][d] * tet_a<NPI>(i, g);
            Ug[d] = su;
            }
        if (R.has_K || (CUBIC && R.has_K3))
            {
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                double sv = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) sv += T.v[i][d] * tet_a<NPI>(i, g);
                Vg[d] = sv;
                }
            }
        double uH = -Abis * gsq;
        double Han[3] = {0.0, 0.0, 0.0};
        if (R.has_K)  // calc_aniso_uniax, tetra.cpp:171-181
            {
            const double t0 = Ug[0] + th_dt * Vg[0], t1 = Ug[1] + th_dt * Vg[1], t2 = Ug[2] + th_dt * Vg[2];
            const double f = R.Kbis * (R.uk[0] * t0 + R.uk[1] * t1 + R.uk[2] * t2);
            Han[0] += f * R.uk[0];
            Han[1] += f * R.uk[1];
            Han[2] += f * R.uk[2];
            const double s = Ug[0] * R.uk[0] + Ug[1] * R.uk[1] + Ug[2] * R.uk[2];
            uH += R.Kbis * (s * s);
            }
        if (CUBIC && R.has_K3)  // calc_aniso_cub, tetra.cpp:183-208 (uk_v.cwiseProduct(ex) kept literally)
            {
            const double uu[3] = {dot3(R.ex, Ug), dot3(R.ey, Ug), dot3(R.ez, Ug)};
            const double uv[3] = {dot3(R.ex, Vg), dot3(R.ey, Vg), dot3(R.ez, Vg)};
            double u3[3];
#pragma unroll
            for (int k = 0; k < 3; k++) u3[k] = uu[k] * (1.0 - uu[k] * uu[k]);
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                const double tmp = uv[d] * R.ex[d];
                const double inner = u3[0] * R.ex[d] + u3[1] * R.ey[d] + u3[2] * R.ez[d]
                                     + th_dt * (tmp * (1.0 - 3 * (uu[d] * uu[d])));
                Han[d] += -R.K3bis * inner;
                }
            uH += -R.K3bis * dot3(uu, u3);
            }
        // tetra.cpp:248-255 (Hst = 0: Tet::extraField is a no-op without spin accumulation)
        double H[3];
            {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                H[d] = Hd[d] + Hext[d][g];
                s += Ug[d] * H[d];
                }
            uH += s;
            }
        // tetra.cpp:257-261 + lumping :114 (the alpha_eff part of the diagonal block)
        const double wa = w * alpha_eff(sp.dt, alpha, uH);
#pragma unroll
        for (int d = 0; d < 3; d++) H[d] += Han[d] + th_dt * Hv[d];
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            contrib[i] += tet_a<NPI>(i, g) * wa;
            const double ai_w = w * tet_a<NPI>(i, g);
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                BE[d][i] -= w * Abis * Ex[d][i];
                BE[d][i] += ai_w * H[d];
                }
            }
        }
    }

// device-resident description of the magnetic tetrahedra (SoA, compact index tm)
struct TetArrays
    {
    int NTm;
    const int4 *ind;        // NTm
    const double *da;       // [12][NTm]
    const double *detJ;     // NTm
    const int *reg;         // NTm
    const TetRegion *regions;
    const double *ext_field;  // [3*npi][NTm] or NULL
    const int4 *slot;       // NTm : record slots of the 4 local nodes (incidence order of the row)
    };

template <int NPI>
__device__ __forceinline__ void tet_load(const TetArrays &A, int tm, const NodeRec *cur, int4 &ind,
                                         TetIn &T)
    {
    ind = __ldg(A.ind + tm);
#pragma unroll
    for (int k = 0; k < 12; k++) T.da[k / 3][k % 3] = __ldcs(A.da + (size_t)k * A.NTm + tm);
    T.detJ = __ldcs(A.detJ + tm);
    const int nd[4] = {ind.x, ind.y, ind.z, ind.w};
#pragma unroll
    for (int i = 0; i < 4; i++) load_rec(cur + nd[i], T.u[i], T.v[i], T.phi[i], T.phiv[i]);
    }

template <int NPI>
__device__ __forceinline__ void tet_field(const TetArrays &A, int tm, const StepPrm &sp, bool space,
                                          double (&Hext)[3][NPI])
    {
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int g = 0; g < NPI; g++)
            Hext[d][g] = space ? sp.A_Hext * __ldcs(A.ext_field + (size_t)(d * NPI + g) * A.NTm + tm)
                               : sp.Hext[d];
    }

constexpr int TET_CTAS_PER_SM = 1;
// one thread per magnetic tetrahedron; emits 4 records {contrib, BE(3)} (32 B), each stored at the
// slot of its node's incidence list so that the row assembly reads them as a stream.  The node bases
// are not needed here: the assembly projects the summed BE of a node once.
template <int NPI, bool SPACE>
__global__ void __launch_bounds__(BLOCK, TET_CTAS_PER_SM)
k_tet(const TetArrays A, const NodeRec *__restrict__ cur, const Basis *__restrict__ basis,
      const StepPrm sp, double4 *__restrict__ rec)
    {
    const int stride = gridDim.x * BLOCK;
    for (int tm = blockIdx.x * BLOCK + threadIdx.x; tm < A.NTm; tm += stride)
        {
        TetIn T;
        int4 ind;
        tet_load<NPI>(A, tm, cur, ind, T);
        double Hext[3][NPI];
        tet_field<NPI>(A, tm, sp, SPACE, Hext);
        const TetRegion R = A.regions[__ldg(A.reg + tm)];
        double contrib[4], BE[3][4];
        tet_core<NPI>(T, R, sp, Hext, contrib, BE);
        const int4 s4 = __ldcs(A.slot + tm);
        const int sl[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            // the record goes where the node's row will stream it from (its incidence slot); the
            // projection Lp = Perm P BE (tetra.cpp:306) is applied once per NODE by the assembly,
            // to the sum of the BE of its tetrahedra (P depends on the node only)
            if (sl[i] < 0) continue;
            st256(rec + sl[i], make_double4(contrib[i], BE[0][i], BE[1][i], BE[2][i]));
            }
        }
    }

// The same element for contexts without cubic anisotropy and without recentring drift (every uniaxial
// material, e.g. the reference's ci-tests/full_test.py): those branches compiled out, the region read where
// it is needed instead of copied into 46 registers, CTAs of 128 threads, three per SM (<= 168 registers: 12
// resident warps per SM instead of 8).
constexpr int TET_LEAN_BLOCK = 128;
template <int NPI, bool SPACE>
__global__ void __launch_bounds__(TET_LEAN_BLOCK, 3)
k_tet_lean(const TetArrays A, const NodeRec *__restrict__ cur, const StepPrm sp, double4 *__restrict__ rec)
    {
    const int stride = gridDim.x * TET_LEAN_BLOCK;
    for (int tm = blockIdx.x * TET_LEAN_BLOCK + threadIdx.x; tm < A.NTm; tm += stride)
        {
        TetIn T;
        int4 ind;
        tet_load<NPI>(A, tm, cur, ind, T);
        double Hext[3][NPI];
        tet_field<NPI>(A, tm, sp, SPACE, Hext);
        const TetRegion &R = A.regions[__ldg(A.reg + tm)];
        double contrib[4], BE[3][4];
        tet_core<NPI, false, false>(T, R, sp, Hext, contrib, BE);
        const int4 s4 = __ldcs(A.slot + tm);
        const int sl[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            if (sl[i] < 0) continue;
            st256(rec + sl[i], make_double4(contrib[i], BE[0][i], BE[1][i], BE[2][i]));
            }
        }
    }

// ---- fast path of Tet::integrales: no magnetocrystalline anisotropy in any region (K = K3 = 0),
// uniform applied field, no recentring drift -- every permalloy-like material, all five BASELINE
// configurations.  Same formulas and the same order of accumulation as tet_core (H_aniso = 0 is
// dropped, not approximated), organised so that the Gauss loop keeps only u, contrib and the field
// live: BE of a node is accumulated when its record is written (it needs only the node's own
// gradient), the velocities are never loaded.  ~120 registers instead of 254: two to three times the
// resident warps for the node gathers.
struct TetIsoIn
    {
    double da[4][3];
    double detJ;
    double u[4][3], phi[4], phiv[4];
    };
struct TetIsoMid
    {
    double dU[3][3];  // grad U
    double H[3];      // (Hd + Hext) + theta dt Hv : the field of tetra.cpp:292-303 with H_aniso = 0
    };

template <int NPI>
__host__ __device__ __forceinline__ void tet_iso_front(const TetIsoIn &T, const TetRegion &R, const StepPrm &sp,
                                              TetIsoMid &M, double contrib[4])
    {
    const double s_dt = FG_THETA * sp.dt * FG_GAMMA0;
    const double th_dt = s_dt / FG_GAMMA0;
    double Heff[3];
#pragma unroll
    for (int d = 0; d < 3; d++)
        {
#pragma unroll
        for (int k = 0; k < 3; k++)
            {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) s += T.u[i][d] * T.da[i][k];
            M.dU[d][k] = s;
            }
        double hd = 0.0, hv = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            hd -= T.phi[i] * T.da[i][d];
            hv -= T.phiv[i] * T.da[i][d];
            }
        Heff[d] = hd + sp.Hext[d];                     // tetra.cpp:248-255, Hst = 0
        M.H[d] = Heff[d] + mul_rn(th_dt, hv);       // :292 with H_aniso = 0 (two roundings, like tet_core)
        }
    double gsq = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++)
        gsq += M.dU[0][k] * M.dU[0][k] + M.dU[1][k] * M.dU[1][k] + M.dU[2][k] * M.dU[2][k];
#pragma unroll
    for (int i = 0; i < 4; i++) contrib[i] = 0.0;
#pragma unroll
    for (int g = 0; g < NPI; g++)
        {
        const double w = T.detJ * tet_pds<NPI>(g);
        double uH = -R.Abis * gsq;
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++)
            {
            double su = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) su += T.u[i][d] * tet_a<NPI>(i, g);
            s += su * Heff[d];
            }
        uH += s;
        const double wa = w * alpha_eff(sp.dt, R.alpha, uH);
#pragma unroll
        for (int i = 0; i < 4; i++) contrib[i] += tet_a<NPI>(i, g) * wa;
        }
    }

// BE(:, i) of local node i (tetra.cpp:294-303), Gauss points in the reference's order
template <int NPI>
__host__ __device__ __forceinline__ void tet_iso_be(const double da_i[3], int i, double detJ, double Abis,
                                           const TetIsoMid &M, double be[3])
    {
    double Ex[3];
#pragma unroll
    for (int d = 0; d < 3; d++)
        {
        Ex[d] = da_i[0] * M.dU[d][0] + da_i[1] * M.dU[d][1] + da_i[2] * M.dU[d][2];
        be[d] = 0.0;
        }
#pragma unroll
    for (int g = 0; g < NPI; g++)
        {
        const double w = detJ * tet_pds<NPI>(g);
        const double ai_w = w * tet_a<NPI>(i, g);
#pragma unroll
        for (int d = 0; d < 3; d++)
            {
            be[d] -= w * Abis * Ex[d];
            be[d] += ai_w * M.H[d];
            }
        }
    }

constexpr int TET_ISO_CTAS_PER_SM = 2;
// The kernel is bound by the latency of its dependent gathers (ncu: long-scoreboard stalls, 16
// resident warps per SM) and by L1 throughput, so (i) the records carry BE itself and the node bases
// are never gathered here (the assembly projects once per node: -192 B of gathers per tetrahedron
// and one dependent stall less), (ii) PIPE: the connectivity and region of a thread's NEXT
// tetrahedron are fetched one iteration ahead.  A deeper variant (connectivity two iterations ahead,
// records of the next tetrahedron prefetched into L2) was slower: more LSU instructions and spills
// (profiles/r01n_kernel_classes_n1_deep_prefetch.txt), and so was cp.async staging of the next
// tetrahedron in shared memory (profiles/experiments/r01p_cp_async_element_staging.md): the kernel is
// co-limited by the L1 tag stage (scattered 32-byte sectors) and DRAM, not by exposed latency alone.
template <int NPI, bool PIPE>
__global__ void __launch_bounds__(BLOCK, TET_ISO_CTAS_PER_SM)
k_tet_iso(const TetArrays A, const NodeRec *__restrict__ cur, const Basis *__restrict__ basis,
          const StepPrm sp, double4 *__restrict__ rec)
    {
    const int stride = gridDim.x * BLOCK;
    int tm = blockIdx.x * BLOCK + threadIdx.x;
    if (tm >= A.NTm) return;
    int4 ind = __ldg(A.ind + tm);
    int reg = __ldg(A.reg + tm);
    for (;;)
        {
        const int tn = tm + stride;
        const bool more = tn < A.NTm;
        int4 ind_n = ind;
        int reg_n = reg;
        if (PIPE && more)
            {
            ind_n = __ldg(A.ind + tn);
            reg_n = __ldg(A.reg + tn);
            }
        TetIsoIn T;
        const int nd[4] = {ind.x, ind.y, ind.z, ind.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            const double4 *q = reinterpret_cast<const double4 *>(cur + nd[i]);
            const double4 a = ld256_nc(q), b = ld256_nc(q + 1);
            T.u[i][0] = a.x; T.u[i][1] = a.y; T.u[i][2] = a.z;
            T.phi[i] = b.z; T.phiv[i] = b.w;
            }
#pragma unroll
        for (int k = 0; k < 12; k++) T.da[k / 3][k % 3] = __ldcs(A.da + (size_t)k * A.NTm + tm);
        T.detJ = __ldcs(A.detJ + tm);
        TetRegion Rl;
        Rl.alpha = A.regions[reg].alpha;
        Rl.Abis = A.regions[reg].Abis;
        TetIsoMid M;
        double contrib[4];
        tet_iso_front<NPI>(T, Rl, sp, M, contrib);
        const int4 s4 = __ldcs(A.slot + tm);
        const int sl[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            if (sl[i] < 0) continue;
            double be[3];
            tet_iso_be<NPI>(T.da[i], i, T.detJ, Rl.Abis, M, be);
            st256(rec + sl[i], make_double4(contrib[i], be[0], be[1], be[2]));
            }
        if (!more) break;
        tm = tn;
        if (PIPE)
            {
            ind = ind_n;
            reg = reg_n;
            }
        else
            {
            ind = __ldg(A.ind + tm);
            reg = __ldg(A.reg + tm);
            }
        }
    }

// ---- k_tet_iso with the node records of a chunk of tetrahedra staged in shared memory ------------------
// The tetrahedra are sorted by their smallest device row, so the 256 tetrahedra of a chunk touch only ~100
// distinct nodes, each of them ~10 times.  Gathering them from global memory costs one L1 tag-stage wavefront
// per lane and 32-byte sector (8 x 32 per warp and tetrahedron: more than half of the kernel's L1 time, ncu
// r01v); here a CTA copies the chunk's distinct node records once into shared memory with cp.async (48 of
// the 64 bytes: u, phi, phiv), double-buffered across chunks, and the tetrahedra read them with 16-bit
// local indices.  Same arithmetic as k_tet_iso (same device functions), so the records are bit-identical.
struct TetChunks
    {
    int nchunk;                // ceil(NTm / 256)
    int cap;                   // largest number of distinct nodes of a chunk
    const int *ptr;            // nchunk + 1
    const int *nodes;          // device rows of the distinct nodes, ascending inside a chunk
    const ushort4 *loc;        // NTm : local indices of the 4 nodes of a tetrahedron inside its chunk's list
    };
constexpr int TET_CHUNK = 256;
// staged record: ISO u0 u1 | u2 v0 | phi phiv (48 B; v0 rides along: 16-byte copies) ; general: the whole NodeRec
__host__ __device__ constexpr int tet_stage_doubles(bool iso) { return iso ? 6 : 8; }

__device__ __forceinline__ void tet_cp_async16(unsigned int dst_smem, const void *src)
    { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory"); }

// ISO: the fast path (tet_iso_front / tet_iso_be, no anisotropy, uniform field, no drift); otherwise the
// general element (tet_core: anisotropies, space-dependent field, recentring drift), whose 254 registers allow
// only 8 resident warps per SM: with the records staged and double-buffered it no longer depends on
// occupancy to hide the gather latency.
template <int NPI, bool ISO, bool SPACE>
__global__ void __launch_bounds__(BLOCK, ISO ? TET_ISO_CTAS_PER_SM : 1)
k_tet_st(const TetArrays A, const TetChunks C, const NodeRec *__restrict__ cur, const StepPrm sp,
         double4 *__restrict__ rec)
    {
    extern __shared__ double tet_stage[];  // 2 buffers of cap staged records
    constexpr int SD = tet_stage_doubles(ISO), NQ = ISO ? 3 : 4;  // doubles and 16-byte pieces per record
    const int tid = threadIdx.x;
    const unsigned int sb0 = (unsigned int)__cvta_generic_to_shared(tet_stage);
    const unsigned int bufbytes = (unsigned int)C.cap * SD * 8u;
    auto issue = [&](int c, int k)
        {
        if (c < C.nchunk)
            {
            const int n0 = __ldg(C.ptr + c), nn = __ldg(C.ptr + c + 1) - n0;
            const unsigned int sb = sb0 + (unsigned int)k * bufbytes;
            for (int i = tid; i < NQ * nn; i += BLOCK)
                {
                const int e = i / NQ, q = i - NQ * e;
                const int node = __ldg(C.nodes + n0 + e);
                const int off = ISO ? (q == 2 ? 48 : 16 * q) : 16 * q;
                tet_cp_async16(sb + (unsigned int)(SD * 8 * e + 16 * q), reinterpret_cast<const char *>(cur + node) + off);
                }
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
        };
    int c = blockIdx.x, k = 0;
    issue(c, 0);
    for (; c < C.nchunk; c += gridDim.x, k ^= 1)
        {
        issue(c + gridDim.x, k ^ 1);  // next chunk of this CTA into the other buffer
        const int tm = c * TET_CHUNK + tid;
        const bool act = tm < A.NTm;
        ushort4 lc = make_ushort4(0, 0, 0, 0);
        int reg = 0;
        int4 s4 = make_int4(-1, -1, -1, -1);
        double da[4][3], detJ = 0.0;
        if (act)
            {  // the coalesced streams of this tetrahedron travel while the node records land
            lc = __ldcs(C.loc + tm);
            reg = __ldg(A.reg + tm);
#pragma unroll
            for (int q = 0; q < 12; q++) da[q / 3][q % 3] = __ldcs(A.da + (size_t)q * A.NTm + tm);
            detJ = __ldcs(A.detJ + tm);
            s4 = __ldcs(A.slot + tm);
            }
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (act)
            {
            const double *buf = tet_stage + (size_t)k * C.cap * SD;
            const int nd[4] = {lc.x, lc.y, lc.z, lc.w};
            const int sl[4] = {s4.x, s4.y, s4.z, s4.w};
            if (ISO)
                {
                TetIsoIn T;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    {
#pragma unroll
                    for (int d = 0; d < 3; d++) T.da[i][d] = da[i][d];
                    const double *r = buf + (size_t)nd[i] * SD;
                    const double2 a = *reinterpret_cast<const double2 *>(r), b = *reinterpret_cast<const double2 *>(r + 4);
                    T.u[i][0] = a.x; T.u[i][1] = a.y; T.u[i][2] = r[2];
                    T.phi[i] = b.x; T.phiv[i] = b.y;
                    }
                T.detJ = detJ;
                TetRegion Rl;
                Rl.alpha = A.regions[reg].alpha;
                Rl.Abis = A.regions[reg].Abis;
                TetIsoMid M;
                double contrib[4];
                tet_iso_front<NPI>(T, Rl, sp, M, contrib);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    {
                    if (sl[i] < 0) continue;
                    double be[3];
                    tet_iso_be<NPI>(T.da[i], i, T.detJ, Rl.Abis, M, be);
                    st256(rec + sl[i], make_double4(contrib[i], be[0], be[1], be[2]));
                    }
                }
            else
                {
                TetIn T;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    {
#pragma unroll
                    for (int d = 0; d < 3; d++) T.da[i][d] = da[i][d];
                    const double2 *r = reinterpret_cast<const double2 *>(buf + (size_t)nd[i] * SD);
                    const double2 a = r[0], b = r[1], cc = r[2], e = r[3];
                    T.u[i][0] = a.x; T.u[i][1] = a.y; T.u[i][2] = b.x;
                    T.v[i][0] = b.y; T.v[i][1] = cc.x; T.v[i][2] = cc.y;
                    T.phi[i] = e.x; T.phiv[i] = e.y;
                    }
                T.detJ = detJ;
                double Hext[3][NPI];
                tet_field<NPI>(A, tm, sp, SPACE, Hext);
                const TetRegion R = A.regions[reg];
                double contrib[4], BE[3][4];
                tet_core<NPI>(T, R, sp, Hext, contrib, BE);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    {
                    if (sl[i] < 0) continue;
                    st256(rec + sl[i], make_double4(contrib[i], BE[0][i], BE[1][i], BE[2][i]));
                    }
                }
            }
        __syncthreads();  // this buffer is the target of the prefetch issued at the top of the next turn
        }
    }

// projection of one node pair: the 2x2 block of K / Kp (SURVEY.md §8a index facts)
__host__ __device__ __forceinline__ void project_block(double E, const double ep_a[3], const double eq_a[3],
                                              const double ep_b[3], const double eq_b[3],
                                              double &k00, double &k01, double &k10, double &k11)
    {
    k00 = E * dot3(eq_a, ep_b);
    k01 = E * dot3(eq_a, eq_b);
    k10 = E * dot3(ep_a, ep_b);
    k11 = E * dot3(ep_a, eq_b);
    }
// gyrotropic part of the diagonal block: a_w e_r . (m x e_c)
__host__ __device__ __forceinline__ void gyro_block(double aw, const double m[3], const double ep[3],
                                           const double eq[3], double &k00, double &k01,
                                           double &k10, double &k11)
    {
    double mp[3], mq[3];
    cross3(m, ep, mp);
    cross3(m, eq, mq);
    k00 += aw * dot3(eq, mp);
    k01 += aw * dot3(eq, mq);
    k10 += aw * dot3(ep, mp);
    k11 += aw * dot3(ep, mq);
    }

// Tap: full element Kp (8x8 row-major) and Lp (8) of the magnetic tets list[0..count), computed
// with the same device functions as the production path (element.h:62,65 layout).
template <int NPI, bool SPACE, bool ISO = false>
__global__ void __launch_bounds__(BLOCK)
k_tet_tap(const TetArrays A, const NodeRec *__restrict__ cur, const Basis *__restrict__ basis,
          const StepPrm sp, const int *__restrict__ list, int count, double *__restrict__ Kp,
          double *__restrict__ Lp)
    {
    const int q = blockIdx.x * BLOCK + threadIdx.x;
    if (q >= count) return;
    const int tm = list[q];
    TetIn T;
    int4 ind;
    tet_load<NPI>(A, tm, cur, ind, T);
    double Hext[3][NPI];
    tet_field<NPI>(A, tm, sp, SPACE, Hext);
    const TetRegion R = A.regions[A.reg[tm]];
    double contrib[4], BE[3][4];
    if (ISO)
        {  // the device functions of the production fast path (k_tet_iso)
        TetIsoIn Ti;
        for (int i = 0; i < 4; i++)
            {
            for (int k = 0; k < 3; k++)
                {
                Ti.da[i][k] = T.da[i][k];
                Ti.u[i][k] = T.u[i][k];
                }
            Ti.phi[i] = T.phi[i];
            Ti.phiv[i] = T.phiv[i];
            }
        Ti.detJ = T.detJ;
        TetIsoMid M;
        tet_iso_front<NPI>(Ti, R, sp, M, contrib);
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            double be[3];
            tet_iso_be<NPI>(Ti.da[i], i, Ti.detJ, R.Abis, M, be);
            BE[0][i] = be[0];
            BE[1][i] = be[1];
            BE[2][i] = be[2];
            }
        }
    else
        tet_core<NPI>(T, R, sp, Hext, contrib, BE);
    const int nd[4] = {ind.x, ind.y, ind.z, ind.w};
    double ep[4][3], eq[4][3];
    for (int i = 0; i < 4; i++) load_basis(basis + nd[i], ep[i], eq[i]);
    const double s_dt = FG_THETA * sp.dt * FG_GAMMA0;
    double wsum = 0.0, aw[4] = {0, 0, 0, 0};
    for (int g = 0; g < NPI; g++)
        {
        const double w = T.detJ * tet_pds<NPI>(g);
        wsum += w;
        for (int i = 0; i < 4; i++) aw[i] += tet_a<NPI>(i, g) * w;
        }
    const double cw = (sp.prefactor * s_dt) * (R.Abis * wsum);
    double *K = Kp + 64 * (size_t)q, *L = Lp + 8 * (size_t)q;
    for (int a = 0; a < 4; a++)
        {
        for (int b = 0; b < 4; b++)
            {
            double E = (T.da[a][0] * T.da[b][0] + T.da[a][1] * T.da[b][1] + T.da[a][2] * T.da[b][2]) * cw;
            if (a == b) E += contrib[a];
            double k00, k01, k10, k11;
            project_block(E, ep[a], eq[a], ep[b], eq[b], k00, k01, k10, k11);
            if (a == b) gyro_block(aw[a], T.u[a], ep[a], eq[a], k00, k01, k10, k11);
            K[(0 * 4 + a) * 8 + (0 * 4 + b)] = k00;
            K[(0 * 4 + a) * 8 + (1 * 4 + b)] = k01;
            K[(1 * 4 + a) * 8 + (0 * 4 + b)] = k10;
            K[(1 * 4 + a) * 8 + (1 * 4 + b)] = k11;
            }
        const double be[3] = {BE[0][a], BE[1][a], BE[2][a]};
        L[a] = dot3(eq[a], be);
        L[4 + a] = dot3(ep[a], be);
        }
    }

// ------------------------------------------------------------------------------------------
// Tri::integrales, src/triangle.cpp:6-36 : surface (Neel) anisotropy, rhs only
// ------------------------------------------------------------------------------------------
struct TriArrays
    {
    int NFa;
    const int *ind;       // [3][NFa]
    const double *surf;   // NFa
    const double *dMs;    // NFa
    const int *reg;       // NFa
    const TriRegion *regions;
    };

template <int NPI>
__host__ __device__ __forceinline__ void tri_core(const TriRegion &R, double surf, double dMs,
                                         const double u[3][3], double BE[3][3])
    {
    const double Kbis = 2.0 * R.Ks / dMs;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) BE[k][i] = 0.0;
#pragma unroll
    for (int g = 0; g < NPI; g++)
        {
        double ug[3];
#pragma unroll
        for (int d = 0; d < 3; d++)
            {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++) s += u[i][d] * tri_a<NPI>(i, g);
            ug[d] = s;
            }
        const double wg = 2.0 * surf * tri_pds<NPI>(g);  // triangle.h:121-122
        const double pf = wg * Kbis * dot3(R.uk, ug);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int k = 0; k < 3; k++) BE[k][i] += pf * tri_a<NPI>(i, g) * R.uk[k];
        }
    }

// one thread per active triangle; emits 3 records {eq.BE, ep.BE}
template <int NPI>
__global__ void __launch_bounds__(BLOCK)
k_tri(const TriArrays A, const NodeRec *__restrict__ cur, const Basis *__restrict__ basis,
      double2 *__restrict__ trec)
    {
    const int stride = gridDim.x * BLOCK;
    for (int fa = blockIdx.x * BLOCK + threadIdx.x; fa < A.NFa; fa += stride)
        {
        int nd[3];
        double u[3][3], v[3], phi, phiv;
#pragma unroll
        for (int i = 0; i < 3; i++)
            {
            nd[i] = A.ind[(size_t)i * A.NFa + fa];
            load_rec(cur + nd[i], u[i], v, phi, phiv);
            }
        const TriRegion R = A.regions[A.reg[fa]];
        double BE[3][3];
        tri_core<NPI>(R, A.surf[fa], A.dMs[fa], u, BE);
#pragma unroll
        for (int i = 0; i < 3; i++)
            {
            double ep[3], eq[3];
            load_basis(basis + nd[i], ep, eq);
            const double be[3] = {BE[0][i], BE[1][i], BE[2][i]};
            trec[3 * (size_t)fa + i] = make_double2(dot3(eq, be), dot3(ep, be));
            }
        }
    }

// ------------------------------------------------------------------------------------------
// Row assembly in SELL-32 order: one warp per slice, one lane per node row.  The lane gathers the
// node's element records through its (SELL-stored, coalesced) incidence list in a fixed order — no
// atomics, bitwise reproducible — then streams its blocks: S and the column index are coalesced
// reads, the neighbour's basis a 48-byte gather, the 2x2 block two coalesced 16-byte stores.  The
// rhs, the initial guess and the Jacobi diagonal leave in the same pass.
// ------------------------------------------------------------------------------------------
struct RowArrays
    {
    int nslice;
    const int *sptr, *scol, *sdeg;     // SELL pattern (fg_common.cuh Operator) + blocks per row
    const double *sS;                  // S in SELL order
    const double *Aw;                  // NODp lumped mass
    const int *iptr;                   // SELL incidence lists: slice extents of the record stream
    const int *itptr, *sinct;          // same for the active triangles
    const unsigned char *nonmag;       // NODp : 1 = node outside the magnetic material (or pad row)
    };

constexpr int ASM_CTAS_PER_SM = 3;
__global__ void __launch_bounds__(BLOCK, ASM_CTAS_PER_SM)
k_assemble_sell(const RowArrays A, const NodeRec *__restrict__ cur, const NodeRec *__restrict__ next,
                const Basis *__restrict__ basis, const double4 *__restrict__ rec,
                const double2 *__restrict__ trec, double cS, double *__restrict__ val,
                double *__restrict__ rhs, double *__restrict__ x0, double *__restrict__ D)
    {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (BLOCK / 32);
    double2 *val2 = reinterpret_cast<double2 *>(val);
    for (int s = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); s < A.nslice; s += nwarps)
        {
        const int row = s * SLICE + lane;
        // ---- gather of the element records -------------------------------------------------
        double Ma = 0.0, L0 = 0.0, L1 = 0.0;  // L: the triangle part, already projected (k_tri)
        double B0 = 0.0, B1 = 0.0, B2 = 0.0;  // sum of BE over the node's tetrahedra
            {
            // the records were written in incidence order: a pure coalesced stream (1 KB per warp
            // request), empty slots hold zeros
            const int i0 = __ldg(A.iptr + s), i1 = __ldg(A.iptr + s + 1);
            const double4 *rp = rec + (size_t)i0 * SLICE + lane;
#pragma unroll 4
            for (int q = i0; q < i1; ++q, rp += SLICE)
                {  // one 256-bit request per record: a warp streams 1 KB contiguous
                const double4 r = ld256_cs(rp);
                Ma += r.x;
                B0 += r.y;
                B1 += r.z;
                B2 += r.w;
                }
            const int t0 = __ldg(A.itptr + s), t1 = __ldg(A.itptr + s + 1);
            const int *tp = A.sinct + (size_t)t0 * SLICE + lane;
            for (int q = t0; q < t1; ++q, tp += SLICE)
                {
                const int idx = __ldcs(tp);
                if (idx >= 0)
                    {
                    const double2 r = trec[idx];
                    L0 += r.x;
                    L1 += r.y;
                    }
                }
            }
        const int p0 = __ldg(A.sptr + s), p1 = __ldg(A.sptr + s + 1);
        const int deg = A.sdeg[row];
        const bool nonmag = A.nonmag[row] != 0;
        const int *cp = A.scol + (size_t)p0 * SLICE + lane;
        const double *sp = A.sS + (size_t)p0 * SLICE + lane;
        double2 *vp = val2 + (size_t)p0 * (2 * SLICE) + lane;
        double m[3] = {0, 0, 0}, vc[3], phi, phiv, ep_a[3] = {0, 0, 0}, eq_a[3] = {0, 0, 0};
        double2 *rhs2 = reinterpret_cast<double2 *>(rhs) + row, *x02 = reinterpret_cast<double2 *>(x0) + row,
                *D2 = reinterpret_cast<double2 *>(D) + row;
        if (nonmag)
            {  // identity rows, zero rhs and guess (src/solver.cpp:46-48, linear_algebra.cpp:13-24)
            *rhs2 = make_double2(0.0, 0.0);
            *x02 = make_double2(0.0, 0.0);
            *D2 = make_double2(0.0, 0.0);
            }
        else
            {
            load_rec(cur + row, m, vc, phi, phiv);
            load_basis(basis + row, ep_a, eq_a);
            double un[3], vn[3];
            load_rec(next + row, un, vn, phi, phiv);
            // buildVect: L[2a] = eq . BE, L[2a+1] = ep . BE (Perm of tetra.cpp:306)
            *rhs2 = make_double2(L0 + (eq_a[0] * B0 + eq_a[1] * B1 + eq_a[2] * B2),
                                 L1 + (ep_a[0] * B0 + ep_a[1] * B1 + ep_a[2] * B2));
            *x02 = make_double2(dot3(vn, ep_a) / FG_GAMMA0, dot3(vn, eq_a) / FG_GAMMA0);
            }
        const double aw = nonmag ? 0.0 : A.Aw[row];
#pragma unroll 2
        for (int j = 0; j < p1 - p0; ++j, cp += SLICE, sp += SLICE, vp += 2 * SLICE)
            {
            const int b = __ldcs(cp);
            const double Sv = __ldcs(sp);
            double k00 = 0.0, k01 = 0.0, k10 = 0.0, k11 = 0.0;
            if (j < deg)
                {
                if (nonmag)
                    {
                    k00 = (b == row) ? 1.0 : 0.0;
                    k11 = k00;
                    }
                else
                    {
                    double E = cS * Sv;
                    if (b == row)
                        {
                        E += Ma;
                        project_block(E, ep_a, eq_a, ep_a, eq_a, k00, k01, k10, k11);
                        gyro_block(aw, m, ep_a, eq_a, k00, k01, k10, k11);
                        *D2 = make_double2(1.0 / k00, 1.0 / k11);
                        }
                    else
                        {
                        double ep_b[3], eq_b[3];
                        load_basis(basis + b, ep_b, eq_b);
                        project_block(E, ep_a, eq_a, ep_b, eq_b, k00, k01, k10, k11);
                        }
                    }
                }
            __stcs(vp, make_double2(k00, k01));
            __stcs(vp + SLICE, make_double2(k10, k11));
            }
        }
    }

// Production assembly for the matrix-free operator (OP_NODE3, fg_common.cuh): K itself is never
// written.  Same record gather as above (one warp per slice, one lane per node row, fixed order, no
// atomics); per node it emits the rhs (buildVect), the initial guess (buildInitGuess) with its
// 3-vector image for the first SpMV, the state-dependent 2x2 node-diagonal part Dg of K
// (alpha_eff mass + gyrotropic term, src/tetra.cpp:108-148) and the Jacobi diagonal 1/K(i,i)
// (build_diag_precond), which also needs the constant diagonal entry S_aa.
struct NodeAsmArrays
    {
    int nslice;
    const double *Sdiag;               // NODp : S_aa
    const double *Aw;                  // NODp lumped mass
    const int *iptr;                   // SELL incidence lists: slice extents of the record stream
    const int *itptr, *sinct;          // same for the active triangles
    const unsigned char *nonmag;       // NODp : 1 = node outside the magnetic material (or pad row)
    };

__global__ void __launch_bounds__(BLOCK, 4)
k_assemble_node(const NodeAsmArrays A, const NodeRec *__restrict__ cur, const NodeRec *__restrict__ next,
                const Basis *__restrict__ basis, const double4 *__restrict__ rec,
                const double2 *__restrict__ trec, double cS, double2 *__restrict__ Dm,
                double *__restrict__ rhs, double *__restrict__ x0, double4 *__restrict__ w0,
                double *__restrict__ D)
    {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (BLOCK / 32);
    for (int s = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); s < A.nslice; s += nwarps)
        {
        const int row = s * SLICE + lane;
        double Ma = 0.0, L0 = 0.0, L1 = 0.0;  // L: the triangle part, already projected (k_tri)
        double B0 = 0.0, B1 = 0.0, B2 = 0.0;  // sum of BE over the node's tetrahedra
            {
            const int i0 = __ldg(A.iptr + s), i1 = __ldg(A.iptr + s + 1);
            const double4 *rp = rec + (size_t)i0 * SLICE + lane;
#pragma unroll 4
            for (int q = i0; q < i1; ++q, rp += SLICE)
                {  // one 256-bit request per record: a warp streams 1 KB contiguous
                const double4 r = ld256_cs(rp);
                Ma += r.x;
                B0 += r.y;
                B1 += r.z;
                B2 += r.w;
                }
            const int t0 = __ldg(A.itptr + s), t1 = __ldg(A.itptr + s + 1);
            const int *tp = A.sinct + (size_t)t0 * SLICE + lane;
            for (int q = t0; q < t1; ++q, tp += SLICE)
                {
                const int idx = __ldcs(tp);
                if (idx >= 0)
                    {
                    const double2 r = trec[idx];
                    L0 += r.x;
                    L1 += r.y;
                    }
                }
            }
        double2 *rhs2 = reinterpret_cast<double2 *>(rhs) + row, *x02 = reinterpret_cast<double2 *>(x0) + row,
                *D2 = reinterpret_cast<double2 *>(D) + row;
        if (A.nonmag[row] != 0)
            {  // identity rows, zero rhs and guess (src/solver.cpp:46-48, linear_algebra.cpp:13-24)
            *rhs2 = make_double2(0.0, 0.0);
            *x02 = make_double2(0.0, 0.0);
            *D2 = make_double2(0.0, 0.0);
            Dm[row] = make_double2(0.0, 0.0);  // identity row: the SpMV tests the nonmag flag
            w0[row] = make_double4(0.0, 0.0, 0.0, 0.0);
            continue;
            }
        double m[3], vc[3], phi, phiv, ep[3], eq[3], un[3], vn[3];
        load_rec(cur + row, m, vc, phi, phiv);
        load_basis(basis + row, ep, eq);
        load_rec(next + row, un, vn, phi, phiv);
        const double g0 = dot3(vn, ep) / FG_GAMMA0, g1 = dot3(vn, eq) / FG_GAMMA0;
        // buildVect: L[2a] = eq . BE, L[2a+1] = ep . BE (Perm of tetra.cpp:306)
        *rhs2 = make_double2(L0 + (eq[0] * B0 + eq[1] * B1 + eq[2] * B2), L1 + (ep[0] * B0 + ep[1] * B1 + ep[2] * B2));
        *x02 = make_double2(g0, g1);
        w0[row] = make_double4(ep[0] * g0 + eq[0] * g1, ep[1] * g0 + eq[1] * g1, ep[2] * g0 + eq[2] * g1, 0.0);
        const double aw = A.Aw[row];
        // node-diagonal part without S, Ma P_a^T P_a + a_w e_r . (m x e_c), in its closed form
        // [[a_w, Ma], [Ma, -a_w]] (fg_common.cuh OP_NODE3): the SpMV needs only the two numbers
        Dm[row] = make_double2(Ma, aw);
        double k00, k01, k10, k11;
        // Jacobi: the full diagonal entries K(2a,2a), K(2a+1,2a+1) as the assembled matrix has them
        project_block(cS * A.Sdiag[row] + Ma, ep, eq, ep, eq, k00, k01, k10, k11);
        gyro_block(aw, m, ep, eq, k00, k01, k10, k11);
        *D2 = make_double2(1.0 / k00, 1.0 / k11);
        }
    }

// ------------------------------------------------------------------------------------------
// Node update (src/solver.cpp:62-88): gated on the device-side outcome of the solve.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
k_update(int NOD, int NOWN, const unsigned char *__restrict__ nonmag, const NodeRec *__restrict__ cur,
         NodeRec *__restrict__ next, const Basis *__restrict__ basis, const double *__restrict__ x,
         double dt, KState *st, const RedBuf red)
    {
    if (!st->done || st->updated) return;
    const bool failed = solve_failed(st);
    double v2max = 0.0;
    if (!failed)
        {
        const int stride = gridDim.x * BLOCK;
        for (int a = blockIdx.x * BLOCK + threadIdx.x; a < NOD; a += stride)
            {
            if (nonmag[a]) continue;
            const double2 xv = reinterpret_cast<const double2 *>(x)[a];
            const double v2 = xv.x * xv.x + xv.y * xv.y;
            if (a < NOWN) v2max = fmax(v2max, v2);  // ghost rows (multi-GPU) belong to another rank
            double u[3], vc[3], phi, phiv, ep[3], eq[3], vn[3], un[3];
            load_rec(cur + a, u, vc, phi, phiv);
            load_basis(basis + a, ep, eq);
            const double vp = xv.x * FG_GAMMA0, vq = xv.y * FG_GAMMA0;  // mesh.h:185-186
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                vn[c] = vp * ep[c] + vq * eq[c];
                un[c] = u[c] + dt * vn[c];
                }
            normalize3(un);
            double2 *q = reinterpret_cast<double2 *>(next + a);
            q[0] = make_double2(un[0], un[1]);
            q[1] = make_double2(un[2], vn[0]);
            q[2] = make_double2(vn[1], vn[2]);
            }
        }
    double tot;
    if (!grid_reduce_max(v2max, red, tot)) return;
    st->failed = failed ? 1 : 0;
    if (!failed)
        {
        st->v2max = tot;
        st->v_max = FG_GAMMA0 * sqrt(tot);
        }
    st->updated = 1;
    }

// multi-GPU: initial guess of the ghost rows (their owner's assembly writes the same numbers there)
__global__ void __launch_bounds__(BLOCK)
k_ghost_guess(int first, int last, const unsigned char *__restrict__ nonmag,
              const NodeRec *__restrict__ next, const Basis *__restrict__ basis, double *__restrict__ x0,
              double4 *__restrict__ w0)
    {
    const int row = first + blockIdx.x * BLOCK + threadIdx.x;
    if (row >= last) return;
    double2 g = make_double2(0.0, 0.0);
    double4 w = make_double4(0.0, 0.0, 0.0, 0.0);
    if (!nonmag[row])
        {
        double un[3], vn[3], phi, phiv, ep[3], eq[3];
        load_rec(next + row, un, vn, phi, phiv);
        load_basis(basis + row, ep, eq);
        g = make_double2(dot3(vn, ep) / FG_GAMMA0, dot3(vn, eq) / FG_GAMMA0);
        w = make_double4(ep[0] * g.x + eq[0] * g.y, ep[1] * g.x + eq[1] * g.y, ep[2] * g.x + eq[2] * g.y, 0.0);
        }
    reinterpret_cast<double2 *>(x0)[row] = g;
    w0[row] = w;
    }

// ------------------------------------------------------------------------------------------
// energies, averages, maximum angle — element-wise reductions over the state
// that is already resident (the reference runs them serially on the host every accepted step).
// ------------------------------------------------------------------------------------------
struct FieldPrm
    {
    double Hext[3];   // uniform field (RtoR3)
    double A_Hext;    // amplitude of the space field (R4toR3)
    };

// Fem::energy, src/energy.cpp:17-47: E[EXCHANGE, ANISOTROPY, DEMAG, ZEEMAN] of the magnetic tets on
// the state `st` (the reference uses NEXT).  One thread per tet; Tet::exchangeEnergy /
// uniaxialAnisotropyEnergy / cubicAnisotropyEnergy / demagEnergy / zeemanEnergy of
// src/tetra.cpp:309-391, each a weight.dot(dens).  NOWN: device rows below it are owned; on a
// slab-partitioned mesh a tet shared by several ranks is counted by each one for the fraction of
// its nodes that rank owns (the fractions are exact binary numbers and sum to one).
template <int NPI, bool SPACE>
__global__ void __launch_bounds__(BLOCK)
k_energy_tet(const TetArrays A, const NodeRec *__restrict__ st, const FieldPrm f, int NOWN,
             double *__restrict__ out, const RedBuf red)
    {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int stride = gridDim.x * BLOCK;
    for (int tm = blockIdx.x * BLOCK + threadIdx.x; tm < A.NTm; tm += stride)
        {
        TetIn T;
        int4 ind;
        tet_load<NPI>(A, tm, st, ind, T);
        const TetRegion R = A.regions[__ldg(A.reg + tm)];
        const double share = 0.25 * ((ind.x < NOWN) + (ind.y < NOWN) + (ind.z < NOWN) + (ind.w < NOWN));
        double dU[3][3];  // dU[d][k] = d u_d / d x_k, src/tetra.h:183-201
#pragma unroll
        for (int d = 0; d < 3; d++)
#pragma unroll
            for (int k = 0; k < 3; k++)
                {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) s += T.u[i][d] * T.da[i][k];
                dU[d][k] = s;
                }
        double dens_ex = 0.0;  // |dudx|^2 + |dudy|^2 + |dudz|^2, tetra.cpp:314-316
#pragma unroll
        for (int k = 0; k < 3; k++) dens_ex += dU[0][k] * dU[0][k] + dU[1][k] * dU[1][k] + dU[2][k] * dU[2][k];
        const double div = dU[0][0] + dU[1][1] + dU[2][2];
        double e_ex = 0.0, e_an1 = 0.0, e_an3 = 0.0, e_dm = 0.0, e_ze = 0.0;
#pragma unroll
        for (int g = 0; g < NPI; g++)
            {
            const double w = T.detJ * tet_pds<NPI>(g);
            double ug[3], phig = 0.0;
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                double su = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) su += T.u[i][d] * tet_a<NPI>(i, g);
                ug[d] = su;
                }
#pragma unroll
            for (int i = 0; i < 4; i++) phig += T.phi[i] * tet_a<NPI>(i, g);
            e_ex += w * dens_ex;
            e_dm += w * (div * phig);                                    // tetra.cpp:368-371
            if (R.has_K)
                {
                const double q = dot3(R.uk, ug);
                e_an1 += w * (q * q);                                    // tetra.cpp:325-327
                }
            if (R.has_K3)
                {
                const double al0 = dot3(ug, R.ex), al1 = dot3(ug, R.ey), al2 = dot3(ug, R.ez);
                e_an3 += w * ((al0 * al1) * (al0 * al1) + (al1 * al2) * (al1 * al2) + (al2 * al0) * (al2 * al0));
                }
            double h[3];
#pragma unroll
            for (int d = 0; d < 3; d++)
                h[d] = SPACE ? __ldcs(A.ext_field + (size_t)(d * NPI + g) * A.NTm + tm) : f.Hext[d];
            e_ze += w * dot3(ug, h);                                     // tetra.cpp:374-391
            }
        acc[0] += share * (R.A * e_ex);
        double e_an = 0.0;
        if (R.has_K) e_an += -R.K * e_an1;
        if (R.has_K3) e_an += R.K3 * e_an3;
        acc[1] += share * e_an;
        acc[2] += share * (-0.5 * FG_MU0 * R.Ms * e_dm);
        acc[3] += share * (SPACE ? -FG_MU0 * R.Ms * f.A_Hext * e_ze : -FG_MU0 * R.Ms * e_ze);
        }
    double tot[4];
    if (grid_reduce<4>(acc, red, tot) != 1) return;
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = tot[k];
    }

// magnetic surface triangles of Fem::energy, src/energy.cpp:49-63 (msh.magTri)
struct MagTriArrays
    {
    int NFm;
    const int *ind;       // [3][NFm] device rows
    const double *surf;   // NFm
    const double *nrm;    // [3][NFm] unit normal (src/triangle.h:206-233)
    const double *dMs;    // NFm
    const int *reg;       // NFm
    const TriRegion *regions;
    };

// Tri::anisotropyEnergy (src/triangle.cpp:38-43) -> out[0], Tri::demagEnergy (:80-85) -> out[1]
template <int NPI>
__global__ void __launch_bounds__(BLOCK)
k_energy_tri(const MagTriArrays A, const NodeRec *__restrict__ st, int NOWN, double *__restrict__ out,
             const RedBuf red)
    {
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * BLOCK;
    for (int fa = blockIdx.x * BLOCK + threadIdx.x; fa < A.NFm; fa += stride)
        {
        double u[3][3], ph[3], v[3], phiv;
        int own = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
            {
            const int nd = A.ind[(size_t)i * A.NFm + fa];
            own += nd < NOWN;
            load_rec(st + nd, u[i], v, ph[i], phiv);
            }
        const double share = own == 3 ? 1.0 : own / 3.0;
        const TriRegion R = A.regions[A.reg[fa]];
        const double surf = A.surf[fa];
        const double n[3] = {A.nrm[fa], A.nrm[(size_t)A.NFm + fa], A.nrm[2 * (size_t)A.NFm + fa]};
        double s_an = 0.0, s_dm = 0.0;
#pragma unroll
        for (int g = 0; g < NPI; g++)
            {
            double ug[3] = {0.0, 0.0, 0.0}, pg = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++)
                {
#pragma unroll
                for (int d = 0; d < 3; d++) ug[d] += u[i][d] * tri_a<NPI>(i, g);
                pg += ph[i] * tri_a<NPI>(i, g);
                }
            const double wg = 2.0 * surf * tri_pds<NPI>(g);
            const double q = dot3(ug, R.uk);
            s_an += wg * (q * q);
            s_dm += (dot3(ug, n) * pg) * wg;
            }
        if (R.Ks != 0.0) acc[0] += share * (-R.Ks * s_an);
        acc[1] += share * (0.5 * FG_MU0 * A.dMs[fa] * s_dm);
        }
    double tot[2];
    if (grid_reduce<2>(acc, red, tot) != 1) return;
    out[0] = tot[0];
    out[1] = tot[1];
    }

// mesh::avg, src/mesh.cpp:89-106: out[0..2] = sum_T weight . interp(component), out[3] = the volume
// of the selected tets (region < 0: all magnetic regions).  what: 0 = u, 1 = v of `st`.
template <int NPI>
__global__ void __launch_bounds__(BLOCK)
k_avg(const TetArrays A, const NodeRec *__restrict__ st, int what, int region, int NOWN,
      double *__restrict__ out, const RedBuf red)
    {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int stride = gridDim.x * BLOCK;
    for (int tm = blockIdx.x * BLOCK + threadIdx.x; tm < A.NTm; tm += stride)
        {
        if (region >= 0 && __ldg(A.reg + tm) != region) continue;
        const int4 ind = __ldg(A.ind + tm);
        const double detJ = __ldcs(A.detJ + tm);
        const int nd[4] = {ind.x, ind.y, ind.z, ind.w};
        double f[4][3];
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            double u[3], v[3], phi, phiv;
            load_rec(st + nd[i], u, v, phi, phiv);
#pragma unroll
            for (int d = 0; d < 3; d++) f[i][d] = what ? v[d] : u[d];
            }
        const double share = 0.25 * ((ind.x < NOWN) + (ind.y < NOWN) + (ind.z < NOWN) + (ind.w < NOWN));
        double s[3] = {0.0, 0.0, 0.0}, vol = 0.0;
#pragma unroll
        for (int g = 0; g < NPI; g++)
            {
            const double w = detJ * tet_pds<NPI>(g);
            vol += w;
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                double val = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) val += f[i][d] * tet_a<NPI>(i, g);
                s[d] += w * val;
                }
            }
#pragma unroll
        for (int d = 0; d < 3; d++) acc[d] += share * s[d];
        acc[3] += share * vol;
        }
    double tot[4];
    if (grid_reduce<4>(acc, red, tot) != 1) return;
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = tot[k];
    }

// mesh::max_angle, src/mesh.h:295-306: min over the mesh edges of u_a . u_b.  The magnetic edges
// are the off-diagonal blocks of the SELL pattern (one warp per slice, as in the SpMV); edges with
// a non-magnetic end (never in the pattern) come as an explicit list.  The kernel reduces
// max(-dot); the host returns acos(min(1, -that)).
__global__ void __launch_bounds__(BLOCK)
k_max_angle(int nslice, const int *__restrict__ sptr, const int *__restrict__ scol,
            const int *__restrict__ sdeg, const NodeRec *__restrict__ st, int n_extra,
            const int2 *__restrict__ extra, double *__restrict__ out, const RedBuf red)
    {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (BLOCK / 32);
    double mx = -1.0;  // = -(initial value 1.0 of the reference's transform_reduce)
    for (int s = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); s < nslice; s += nwarps)
        {
        const int row = s * SLICE + lane;
        const int p0 = __ldg(sptr + s), deg = sdeg[row];
        if (deg <= 1) continue;
        const double2 *q = reinterpret_cast<const double2 *>(st + row);
        const double2 a = __ldg(q), b = __ldg(q + 1);
        const int *cp = scol + (size_t)p0 * SLICE + lane;
        for (int j = 0; j < deg; ++j, cp += SLICE)
            {
            const int c = __ldcs(cp);
            if (c <= row) continue;  // each edge once (first < second); skips the diagonal block
            const double2 *qc = reinterpret_cast<const double2 *>(st + c);
            const double2 ca = __ldg(qc), cb = __ldg(qc + 1);
            const double d = a.x * ca.x + a.y * ca.y + b.x * cb.x;
            mx = fmax(mx, -d);
            }
        }
    const int stride = gridDim.x * BLOCK;
    for (int e = blockIdx.x * BLOCK + threadIdx.x; e < n_extra; e += stride)
        {
        const int2 ed = extra[e];
        const double2 *qa = reinterpret_cast<const double2 *>(st + ed.x), *qb = reinterpret_cast<const double2 *>(st + ed.y);
        const double2 a = __ldg(qa), b = __ldg(qa + 1), ca = __ldg(qb), cb = __ldg(qb + 1);
        const double d = a.x * ca.x + a.y * ca.y + b.x * cb.x;
        mx = fmax(mx, -d);
        }
    double tot;
    if (!grid_reduce_max(mx, red, tot, -1.7976931348623157e308)) return;
    out[0] = tot;
    }

// ------------------------------------------------------------------------------------------
// magnetic charges (what the reference feeds ScalFMM, src/fmm_demag.h:155-185)
// and an all-pairs evaluation of the potential they create (what the FMM approximates).
// Sources are double4 {x, y, z, q}: the Gauss points are per-mesh constants, q is refreshed here.
// ------------------------------------------------------------------------------------------
// Tet::charges, src/tetra.cpp:347-359: q_g = -Ms w_g div(vec); which = 0: u, 1: v of `st`
template <int NPI>
__global__ void __launch_bounds__(BLOCK)
k_charges_tet(const TetArrays A, const NodeRec *__restrict__ st, int which, double4 *__restrict__ src)
    {
    const int stride = gridDim.x * BLOCK;
    for (int tm = blockIdx.x * BLOCK + threadIdx.x; tm < A.NTm; tm += stride)
        {
        TetIn T;
        int4 ind;
        tet_load<NPI>(A, tm, st, ind, T);
        const double Ms = A.regions[__ldg(A.reg + tm)].Ms;
        double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            {
            const double *f = which ? T.v[i] : T.u[i];
            sx += f[0] * T.da[i][0];
            sy += f[1] * T.da[i][1];
            sz += f[2] * T.da[i][2];
            }
        const double dud_sum = sx + sy + sz;
#pragma unroll
        for (int g = 0; g < NPI; g++)
            src[(size_t)tm * NPI + g].w = -Ms * (T.detJ * tet_pds<NPI>(g)) * dud_sum;
        }
    }

// Tri::potential, src/triangle.cpp:87-125: potential at local node i of the linear charge s_k = vec_k . n
__device__ __forceinline__ double tri_potential(const double p[3][3], const double sn[3], double surf,
                                                double dMs, int i)
    {
    const int ii = (i + 1) % 3, iii = (i + 2) % 3;
    double p1p2[3], p1p3[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        {
        p1p2[k] = p[ii][k] - p[i][k];
        p1p3[k] = p[iii][k] - p[i][k];
        }
    const double b = sqrt(dot3(p1p2, p1p2));
    const double t = dot3(p1p2, p1p3) / b;
    const double _2s = 2. * surf;
    const double h = _2s / b;
    const double c = (t - b) / h;
    const double fc = sqrt(1.0 + c * c);
    const double r = h * sqrt(1.0 + (t / h) * (t / h));
    const double log_1 = log((c * t + h + fc * r) / (b * (c + fc)));
    const double xi = b * log_1 / fc;
    const double pot = xi * sn[i]
                       + ((xi * (h + c * t) - b * (r - b)) * sn[ii] + b * (r - b - c * xi) * sn[iii]) * b
                                 / (_2s * (1 + c * c));
    return 0.5 * dMs * pot;
    }

// Tri::charges (src/triangle.cpp:45-59) into the source list and, per (triangle, local node), the
// second-order correction of Tri::correctionCharges (src/triangle.cpp:61-78): minus the Gauss-point
// terms of the triangle itself plus the analytic potential.  The node sums are a separate gather
// (k_corr_gather) so that no atomics are needed and the result is reproducible.
template <int NPI>
__global__ void __launch_bounds__(BLOCK)
k_charges_tri(const MagTriArrays A, const double *__restrict__ pos, const NodeRec *__restrict__ st,
              int which, double4 *__restrict__ src, double *__restrict__ tcorr)
    {
    const int stride = gridDim.x * BLOCK;
    for (int fa = blockIdx.x * BLOCK + threadIdx.x; fa < A.NFm; fa += stride)
        {
        double p[3][3], f[3][3], sn[3];
        const double n[3] = {A.nrm[fa], A.nrm[(size_t)A.NFm + fa], A.nrm[2 * (size_t)A.NFm + fa]};
        const double surf = A.surf[fa], dMs = A.dMs[fa];
#pragma unroll
        for (int i = 0; i < 3; i++)
            {
            const int nd = A.ind[(size_t)i * A.NFm + fa];
            double u[3], v[3], phi, phiv;
            load_rec(st + nd, u, v, phi, phiv);
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                f[i][d] = which ? v[d] : u[d];
                p[i][d] = pos[3 * (size_t)nd + d];
                }
            sn[i] = dot3(f[i], n);
            }
        double q[NPI], gp[NPI][3];
#pragma unroll
        for (int g = 0; g < NPI; g++)
            {
            double ug[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int d = 0; d < 3; d++)
                {
                double sp = 0.0;
#pragma unroll
                for (int i = 0; i < 3; i++)
                    {
                    ug[d] += f[i][d] * tri_a<NPI>(i, g);
                    sp += p[i][d] * tri_a<NPI>(i, g);
                    }
                gp[g][d] = sp;
                }
            q[g] = dMs != 0.0 ? dMs * ((2.0 * surf * tri_pds<NPI>(g)) * dot3(ug, n)) : 0.0;
            src[(size_t)fa * NPI + g].w = q[g];
            }
#pragma unroll
        for (int i = 0; i < 3; i++)
            {
            double cr = 0.0;
#pragma unroll
            for (int g = 0; g < NPI; g++)
                {
                const double dd[3] = {p[i][0] - gp[g][0], p[i][1] - gp[g][1], p[i][2] - gp[g][2]};
                cr -= q[g] / sqrt(dot3(dd, dd));
                }
            cr += tri_potential(p, sn, surf, dMs, i);
            tcorr[3 * (size_t)fa + i] = cr;
            }
        }
    }

// corr[row] = sum of the (triangle, node) corrections of the node, in triangle order
__global__ void __launch_bounds__(BLOCK)
k_corr_gather(int NODt, const int *__restrict__ cptr, const int *__restrict__ cidx,
              const double *__restrict__ tcorr, double *__restrict__ corr)
    {
    const int row = blockIdx.x * BLOCK + threadIdx.x;
    if (row >= NODt) return;
    double s = 0.0;
    for (int k = cptr[row]; k < cptr[row + 1]; k++) s += tcorr[cidx[k]];
    corr[row] = s;
    }

// All-pairs potential, the sum scal_fmm::fmm::demag approximates (src/fmm_demag.h:187-223):
// phi_i = (sum_j q_j / |p_i - x_j| + corr_i) / (4 pi) for the magnetic nodes; one thread per target,
// the sources stream through shared memory in tiles of BLOCK.  which = 0: phi, 1: phiv of `next`.
__global__ void __launch_bounds__(BLOCK)
k_demag_direct(int NODt, const unsigned char *__restrict__ nonmag, const double *__restrict__ pos,
               long long nsrc, const double4 *__restrict__ src, const double *__restrict__ corr,
               int which, NodeRec *__restrict__ next)
    {
    __shared__ double4 tile[BLOCK];
    const int row = blockIdx.x * BLOCK + threadIdx.x;
    const bool active = row < NODt && !nonmag[row];
    double px = 0.0, py = 0.0, pz = 0.0;
    if (active)
        {
        px = pos[3 * (size_t)row];
        py = pos[3 * (size_t)row + 1];
        pz = pos[3 * (size_t)row + 2];
        }
    double s = 0.0;
    for (long long j0 = 0; j0 < nsrc; j0 += BLOCK)
        {
        const long long j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < nsrc ? src[j] : make_double4(0.0, 0.0, 0.0, 0.0);
        __syncthreads();
        const int cnt = nsrc - j0 < BLOCK ? (int)(nsrc - j0) : BLOCK;
        if (active)
            for (int k = 0; k < cnt; k++)
                {
                const double4 q = tile[k];
                const double dx = px - q.x, dy = py - q.y, dz = pz - q.z;
                s += q.w / sqrt(dx * dx + dy * dy + dz * dz);
                }
        __syncthreads();
        }
    if (!active) return;
    const double val = (s + corr[row]) / (4 * 3.14159265358979323846);
    if (which) next[row].phiv = val;
    else next[row].phi = val;
    }

// ------------------------------------------------------------------------------------------
// state packing helpers (host <-> NodeRec)
// ------------------------------------------------------------------------------------------
// which: bit0 u, bit1 v, bit2 phi, bit3 phiv ; staging = [u(3N) | v(3N) | phi(N) | phiv(N)] in the
// caller's node order; the records are in device row order (perm[row] = node, -1 = pad row).
__global__ void __launch_bounds__(BLOCK)
k_pack(int NODp, int NOD, const int *__restrict__ perm, NodeRec *__restrict__ dst,
       const double *__restrict__ stage, int which)
    {
    const int stride = gridDim.x * BLOCK;
    const size_t N = (size_t)NOD;
    for (int row = blockIdx.x * BLOCK + threadIdx.x; row < NODp; row += stride)
        {
        const int a = perm[row];
        if (a < 0) continue;
        NodeRec r = dst[row];
        if (which & 1) { r.u[0] = stage[3 * (size_t)a]; r.u[1] = stage[3 * (size_t)a + 1]; r.u[2] = stage[3 * (size_t)a + 2]; }
        if (which & 2) { r.v[0] = stage[3 * N + 3 * (size_t)a]; r.v[1] = stage[3 * N + 3 * (size_t)a + 1]; r.v[2] = stage[3 * N + 3 * (size_t)a + 2]; }
        if (which & 4) r.phi = stage[6 * N + a];
        if (which & 8) r.phiv = stage[7 * N + a];
        dst[row] = r;
        }
    }

__global__ void __launch_bounds__(BLOCK)
k_unpack(int NODp, int NOD, const int *__restrict__ perm, const NodeRec *__restrict__ src,
         double *__restrict__ stage, int which)
    {
    const int stride = gridDim.x * BLOCK;
    const size_t N = (size_t)NOD;
    for (int row = blockIdx.x * BLOCK + threadIdx.x; row < NODp; row += stride)
        {
        const int a = perm[row];
        if (a < 0) continue;
        const NodeRec r = src[row];
        if (which & 1) { stage[3 * (size_t)a] = r.u[0]; stage[3 * (size_t)a + 1] = r.u[1]; stage[3 * (size_t)a + 2] = r.u[2]; }
        if (which & 2) { stage[3 * N + 3 * (size_t)a] = r.v[0]; stage[3 * N + 3 * (size_t)a + 1] = r.v[1]; stage[3 * N + 3 * (size_t)a + 2] = r.v[2]; }
        if (which & 4) stage[6 * N + a] = r.phi;
        if (which & 8) stage[7 * N + a] = r.phiv;
        }
    }

}  // namespace fg
