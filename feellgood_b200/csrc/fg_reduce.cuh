// fg_reduce.cuh — warp-shuffle + deterministic two-stage grid reductions (device code).
// Replaces the serial std::inner_product folds of the reference (src/algebra/algebraCore.h:10-17).
#pragma once
#include "fg_common.cuh"
#include "fg_dist.cuh"

namespace fg
{
__device__ __forceinline__ double warp_sum(double v)
    {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
    }

// Compensated pair (s, e): s + e is the running sum, e collects the rounding errors of every addition
// of the reduction tree (Knuth's TwoSum, error-free).  The dots of BiCGStab cancel massively near
// convergence ((rt, r) ~ 1e-14 |rt||r|): a plain tree ends on the sum of two large opposite halves and
// lands on EXACTLY 0.0 about once per 1e4 reductions, which the reference's `rho_2 == 0` test
// (src/algebra/bicg.h:189) turns into a spurious CANNOT_CONVERGE; its own serial fold, which adds
// one small term last, practically never does.  With the compensation the tree returns the correctly
// rounded sum of the per-thread partials, whatever the grid shape or the number of ranks.
__device__ __forceinline__ void dd_add(double &s, double &e, double bs, double be)
    {
    const double t = s + bs;
    const double bb = t - s;
    const double err = (s - (t - bb)) + (bs - bb);
    s = t;
    e += be + err;
    }
__device__ __forceinline__ void warp_sum_dd(double &s, double &e)
    {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        {
        const double bs = __shfl_xor_sync(0xffffffffu, s, o), be = __shfl_xor_sync(0xffffffffu, e, o);
        dd_add(s, e, bs, be);
        }
    }

// Sum NV per-thread values over the whole grid.  Returns 1 on thread 0 of the last CTA to finish,
// with the totals in out[] (all-reduced over the ranks on a distributed context), 2 on the other
// threads of that CTA, 0 elsewhere.  Deterministic: CTA partials are summed in index order.
// red.partials holds [2 * RED_NV][MAX_GRID] doubles: value and compensation of every CTA partial.
template <int NV>
__device__ int grid_reduce(double (&v)[NV], const RedBuf red, double (&out)[NV])
    {
    __shared__ double sm[NV][BLOCK / 32], sme[NV][BLOCK / 32];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++)
        {
        double s = v[k], e = 0.0;
        warp_sum_dd(s, e);
        if (lane == 0)
            {
            sm[k][wid] = s;
            sme[k][wid] = e;
            }
        }
    __syncthreads();
    if (threadIdx.x == 0)
        {
#pragma unroll
        for (int k = 0; k < NV; k++)
            {
            double s = sm[k][0], e = sme[k][0];
#pragma unroll
            for (int w = 1; w < BLOCK / 32; w++) dd_add(s, e, sm[k][w], sme[k][w]);
            red.partials[k * MAX_GRID + blockIdx.x] = s;
            red.partials[(RED_NV + k) * MAX_GRID + blockIdx.x] = e;
            }
        __threadfence();
        unsigned int t = atomicInc(red.ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
        }
    __syncthreads();
    if (!is_last) return 0;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; k++)
        {
        double s = 0.0, e = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += BLOCK)
            dd_add(s, e, __ldcg(&red.partials[k * MAX_GRID + i]), __ldcg(&red.partials[(RED_NV + k) * MAX_GRID + i]));
        warp_sum_dd(s, e);
        __syncthreads();
        if (lane == 0)
            {
            sm[k][wid] = s;
            sme[k][wid] = e;
            }
        }
    __syncthreads();
    if (wid != 0) return 2;
    __shared__ double tot[2 * NV];  // values, then compensations
    if (lane == 0)
        {
#pragma unroll
        for (int k = 0; k < NV; k++)
            {
            double s = sm[k][0], e = sme[k][0];
#pragma unroll
            for (int w = 1; w < BLOCK / 32; w++) dd_add(s, e, sm[k][w], sme[k][w]);
            tot[k] = s;
            tot[NV + k] = e;
            }
        }
    __syncwarp();
    if (red.dist != nullptr)
        {  // the whole warp takes part.  The all-reduce also publishes halo pushes of this kernel's other CTAs
           // to the peers (k_bicg_s_node), hence system-scope fences around it: release before the mailbox
           // stores, acquire after the spin
        __threadfence_system();
        dist_allreduce_warp(red.dist, tot, NV, false);
        __threadfence_system();
        }
    if (lane != 0) return 2;
#pragma unroll
    for (int k = 0; k < NV; k++) out[k] = tot[k] + tot[NV + k];
    return 1;
    }


// Same protocol for a maximum (the v2max of src/solver.cpp:74-88); `lowest` = identity element.
__device__ __forceinline__ bool grid_reduce_max(double v, const RedBuf red, double &out, double lowest = 0.0)
    {
    __shared__ double smx[BLOCK / 32];
    __shared__ int is_last_mx;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) smx[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0)
        {
        double s = smx[0];
#pragma unroll
        for (int w = 1; w < BLOCK / 32; w++) s = fmax(s, smx[w]);
        red.partials[blockIdx.x] = s;
        __threadfence();
        unsigned int t = atomicInc(red.ticket, gridDim.x - 1);
        is_last_mx = (t == gridDim.x - 1);
        }
    __syncthreads();
    if (!is_last_mx) return false;
    __threadfence();
    double s = lowest;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += BLOCK) s = fmax(s, __ldcg(&red.partials[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    __syncthreads();
    if (lane == 0) smx[wid] = s;
    __syncthreads();
    if (wid != 0) return false;
    __shared__ double totmx[1];
    if (lane == 0)
        {
        s = smx[0];
#pragma unroll
        for (int w = 1; w < BLOCK / 32; w++) s = fmax(s, smx[w]);
        totmx[0] = s;
        }
    __syncwarp();
    if (red.dist != nullptr)
        {
        __threadfence_system();
        dist_allreduce_warp(red.dist, totmx, 1, true);  // max: no compensation
        __threadfence_system();
        }
    if (lane != 0) return false;
    out = totmx[0];
    return true;
    }

}  // namespace fg
