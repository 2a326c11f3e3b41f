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

// Sum NV per-thread values over the whole grid.  Returns 1 on thread 0 of the last CTA to finish,
// with the totals in out[] (all-reduced over the ranks on a distributed context), 2 on the other
// threads of that CTA, 0 elsewhere.  Deterministic: CTA partials are summed in index order.
template <int NV>
__device__ int grid_reduce(double (&v)[NV], const RedBuf red, double (&out)[NV])
    {
    __shared__ double sm[NV][BLOCK / 32];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++)
        {
        double s = warp_sum(v[k]);
        if (lane == 0) sm[k][wid] = s;
        }
    __syncthreads();
    if (threadIdx.x == 0)
        {
#pragma unroll
        for (int k = 0; k < NV; k++)
            {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < BLOCK / 32; w++) s += sm[k][w];
            red.partials[k * MAX_GRID + blockIdx.x] = s;
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
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += BLOCK)
            s += __ldcg(&red.partials[k * MAX_GRID + i]);
        s = warp_sum(s);
        __syncthreads();
        if (lane == 0) sm[k][wid] = s;
        }
    __syncthreads();
    if (wid != 0) return 2;
    __shared__ double tot[NV];
    if (lane == 0)
        {
#pragma unroll
        for (int k = 0; k < NV; k++)
            {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < BLOCK / 32; w++) s += sm[k][w];
            tot[k] = s;
            }
        }
    __syncwarp();
    if (red.dist != nullptr) dist_allreduce_warp(red.dist, tot, NV, false);  // the whole warp takes part
    if (lane != 0) return 2;
#pragma unroll
    for (int k = 0; k < NV; k++) out[k] = tot[k];
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
    if (red.dist != nullptr) dist_allreduce_warp(red.dist, totmx, 1, true);
    if (lane != 0) return false;
    out = totmx[0];
    return true;
    }

}  // namespace fg
