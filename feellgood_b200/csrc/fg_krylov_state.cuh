// fg_krylov_state.cuh -- the scalar side of the Krylov solvers: the iteration monitor of the reference
// (src/algebra/iter.h:37-179) and the updates of alpha / omega / rho with their breakdown and exit rules
// (src/algebra/bicg.h:163-234, src/algebra/cg.h:15-121), as executed by the last CTA of each reducing
// kernel (fg_krylov.cu).  Everything here is __host__ __device__: tests/cpp/krylov_state_test.cu drives
// the same functions on the CPU in the order the kernels call them and compares the outcome with the
// reference's own bicg_dir.
#pragma once
#include <math.h>

#include "fg_common.cuh"

namespace fg
{
// ------------------------------------------------------------------------------------------
// iteration monitor, reference src/algebra/iter.h:113-157
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ bool it_finished(KState *st, double nr)
    {
    st->res = fabs(nr);
    if (isnan(st->res))
        {
        st->status = FG_CANNOT_CONVERGE;
        return false;
        }
    if (st->res <= st->rhsn * st->resmax)
        {
        st->status = FG_CONVERGED;
        return true;
        }
    return false;
    }

// `while (!iter.finished(norm(r)))`, then rho_1 and the breakdown test (bicg.h:185-195)
__host__ __device__ __forceinline__ void bicg_top_of_loop(KState *st, double rr, double rho1_new)
    {
    if (it_finished(st, sqrt(fabs(rr))))
        {
        st->done = 1;
        return;
        }
    st->rho1 = rho1_new;
    if (st->nit > 0 && (st->rho2 == 0.0 || st->omega == 0.0))
        {
        st->status = FG_CANNOT_CONVERGE;
        st->done = 1;
        }
    }

// stages of the SpMV with fused epilogues (fg_krylov.cu)
enum
    {
    ST_PLAIN = 0,    // y = A x
    ST_BICG_SETUP,   // r = b - A x (masked); rt = r (p = r implicit); ||b||^2, ||r||^2 (bicg.h:172-183)
    ST_BICG_V,       // v = A phat (masked); (v, rt) -> alpha                    (bicg.h:203-206)
    ST_BICG_T,       // t = A shat (masked); (t,s), (t,t) -> omega               (bicg.h:219-222)
    ST_CG_SETUP,     // r = b - A x (masked); p = D r; ||b||^2, ||r||^2, (Dr,r)  (cg.h:24-34)
    ST_CG_Q,         // q = A p (masked); (q,p) -> a                             (cg.h:45-52)
    ST_RESID         // y = b - A x (masked)  [b -= A xd of the *_dir variants]
    };


__host__ __device__ __forceinline__ double fg_fma(double a, double b, double c)
    {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
    }

// The two vector updates whose results cross GPUs are written with explicit roundings so that the
// owner's value and the copy it pushes to a neighbour are bit-identical whatever the compiler
// contracts elsewhere.   p = r + beta (p - omega v)  (bicg.h:196-201);   s = r - alpha v  (:207-208)
__host__ __device__ __forceinline__ double bicg_p_value(double p, double v, double r, double omega, double beta)
    { return fg_fma(fg_fma(-omega, v, p), beta, r); }
__host__ __device__ __forceinline__ double bicg_s_value(double r, double v, double alpha)
    { return fg_fma(-alpha, v, r); }
__host__ __device__ __forceinline__ double bicg_beta(const KState *st)
    { return (st->rho1 / st->rho2) * (st->alpha / st->omega); }

// finalisation of a reducing SpMV stage by the last CTA (the scalars of bicg.h / cg.h)
template <int STAGE> __host__ __device__ __forceinline__ void spmv_finalize(KState *st, const double (&tot)[RED_NV])
    {
    if (STAGE == ST_BICG_SETUP)
        {
        st->rhsn = sqrt(fabs(tot[0]));
        bicg_top_of_loop(st, tot[1], tot[1]);  // rt == r: (rt, r) = ||r||^2
        }
    else if (STAGE == ST_BICG_V)
        {
        st->alpha = st->rho1 / tot[0];
        khist(st, 0, st->rho1);
        khist(st, 1, tot[0]);
        khist(st, 2, st->alpha);
        }
    else if (STAGE == ST_BICG_T)
        {
        st->omega = tot[0] / tot[1];
        khist(st, 4, tot[0]);
        khist(st, 5, tot[1]);
        khist(st, 6, st->omega);
        }
    else if (STAGE == ST_CG_SETUP)
        {
        st->rhsn = sqrt(fabs(tot[0]));
        st->rho1 = tot[2];  // rho
        if (it_finished(st, sqrt(fabs(tot[1]))) || st->status == FG_CANNOT_CONVERGE) st->done = 1;
        }
    else if (STAGE == ST_CG_Q)
        {
        if (tot[0] == 0.0)
            {
            st->status = FG_CANNOT_CONVERGE;
            st->done = 1;
            }
        else
            st->alpha = st->rho1 / tot[0];
        }
    }

// s = r - alpha v ; shat = D s ; ||s||^2 -> mid-iteration exit test    (bicg.h:207-218)
__host__ __device__ __forceinline__ void bicg_s_finalize(KState *st, double ss)
    {
    khist(st, 3, ss);
    if (it_finished(st, sqrt(fabs(ss))))
        {
        st->final_half = 1;  // x += alpha phat is applied by k_bicg_xr
        st->done = 1;
        }
    else if (st->status == FG_ITER_OVERFLOW || st->status == FG_CANNOT_CONVERGE)
        st->done = 1;
    }


// x += alpha phat + omega shat ; r = s - omega t ; ||r||^2, (rt,r) -> next loop test (bicg.h:223-231 then
// :185-195); fh: the loop ended on ||s|| and only x += alpha phat was applied
__host__ __device__ __forceinline__ void bicg_xr_finalize(KState *st, double rr, double rtr, int fh)
    {
    if (fh)
        st->final_half = 0;
    else
        {
        khist(st, 7, rr);
        st->rho2 = st->rho1;
        st->nit++;
        if (st->nit >= st->maxiter) st->status = FG_ITER_OVERFLOW;  // iter.h:119-124
        bicg_top_of_loop(st, rr, rtr);
        }
    }

// x += a p ; r -= a q ; ||r||^2, (Dr, r) -> loop test of cg (cg.h:45-56)
__host__ __device__ __forceinline__ void cg_xr_finalize(KState *st, double rr, double rho)
    {
    st->rho2 = st->rho1;  // rho_1 = rho
    st->rho1 = rho;
    st->nit++;
    if (st->nit >= st->maxiter) st->status = FG_ITER_OVERFLOW;
    if (it_finished(st, sqrt(fabs(rr))) || st->status == FG_ITER_OVERFLOW
        || st->status == FG_CANNOT_CONVERGE)
        st->done = 1;
    }

// LinAlgebra::solve's failure predicate (src/solver.cpp:62-69): ITER_OVERFLOW, CANNOT_CONVERGE, or an
// ABSOLUTE residual above TOL -- a solve that converged relatively (res <= TOL |b|) is still declared
// failed when |b| > 1 and res > TOL; the time-step controller then halves dt.
__host__ __device__ __forceinline__ bool solve_failed(const KState *st)
    { return st->status == FG_ITER_OVERFLOW || st->status == FG_CANNOT_CONVERGE || st->res > st->resmax; }

// iteration::reset (iter.h:92-98) + the Krylov scalars
__host__ __device__ __forceinline__ void kstate_reset(KState *st, double tol, int maxiter)
    {
    st->rho1 = st->rho2 = st->alpha = st->beta = st->omega = 0.0;
    st->res = 1.7976931348623157e308;  // iteration::reset, iter.h:92-98
    st->rhsn = 1.0;
    st->resmax = tol;
    st->nit = 0;
    st->maxiter = maxiter;
    st->status = FG_UNDEFINED;
    st->done = 0;
    st->final_half = 0;
    st->updated = 0;
    st->failed = 0;
    }

}  // namespace fg
