// fg_dist.cuh — multi-GPU plumbing of the row-block (slab) partitioned LLG solve, device side.
//
// One process per GPU; every rank owns a contiguous block of node rows (a slab along the axis the
// reference sorts nodes by, src/mesh.cpp:334-367) plus read-only ghost copies of the neighbours'
// boundary nodes.  The reference is single-process, so nothing here replaces reference code: this
// is the B200 execution model of SURVEY.md §8(e).
//
// Communication never leaves the kernels that produce or consume the data:
//   * all-reduce of the Krylov scalars: warp 0 of the last CTA of a reducing kernel stores its
//     partial sums into EVERY rank's mailbox through CUDA-IPC-mapped peer memory (NVLink P2P
//     stores) as self-validating 8-byte words {32 data bits, 32-bit epoch tag} — an aligned 8-byte
//     store is one NVLink transaction, so no fence / flag round trip is needed (the "LL" idea) —
//     then spins on its own mailbox until every word carries the current tag and sums IN RANK
//     ORDER: every rank obtains bit-identical scalars, hence takes identical convergence decisions
//     without any further agreement protocol.  One-way NVLink latency per all-reduce;
//   * halo exchange: the owner pushes its boundary entries of a vector into the ghost tail of the
//     neighbour's copy of that vector (same peer mapping).  D.p: pushed grid-wide by the first CTAs
//     of k_bicg_p (beta is known when it starts); the last pusher CTA raises a halo epoch flag on
//     which the consuming SpMV spins; D.s: pushed at the start of k_bicg_s, whose own all-reduce
//     (|s|^2) is the barrier that publishes it; x: one k_halo launch per solve.
//
// Why single-buffered ghost tails and double-buffered mailboxes are race-free: an all-reduce is a
// barrier (no rank leaves epoch e before every rank has entered it).  Between two successive
// pushes into the same ghost tail there is always at least one all-reduce that the receiver
// enters only after the kernel that read the previous ghost values (BiCGStab: push(phat) -> SpMV
// [reduce alpha] -> push(shat) -> SpMV [reduce omega] -> update [reduce rho, |r|] -> push(phat)...;
// the final push(x) follows the reduce that set `done`, and the next step's first push follows the
// setup reduce, which a rank enters after its node update consumed x).  A rank can be at most one
// all-reduce ahead of another, so mailbox slot (epoch & 1) is free again when it is rewritten.
//
// All spins are bounded: on timeout the error word is set, the solve is marked CANNOT_CONVERGE and
// the host reports FG_ERR_DIST instead of hanging the device.
#pragma once
#include <cuda_runtime.h>

namespace fg
{
constexpr int DIST_MAX_RANKS = 8;
constexpr int DIST_NV = 4;  // = RED_NV
constexpr unsigned long long DIST_TIMEOUT_NS = 10000000000ull;  // 10 s before a spin gives up

// Control block at the start of every rank's exchange arena (one cudaMalloc, IPC-exported).
struct DistCtrl
    {
    // [epoch parity][source rank][32-bit half of double k]: low 32 bits data, high 32 bits epoch tag;
    // a sum carries 2 doubles per value (the value and its compensation, fg_reduce.cuh)
    unsigned long long ll[2][DIST_MAX_RANKS][4 * DIST_NV];
    unsigned long long hflag[DIST_MAX_RANKS];            // halo epoch written by each source rank
    };

// Per-rank descriptor in local device memory (pointers into the peers' mapped arenas).
struct DistDev
    {
    int rank, world;
    DistCtrl *ctrl[DIST_MAX_RANKS];        // ctrl[q] = rank q's control block (q == rank: local)
    unsigned long long epoch;              // all-reduce epoch (same sequence on every rank)
    unsigned long long hepoch;             // halo epoch
    int error;                             // 1 = a spin timed out
    // halo plan
    int recv_from[DIST_MAX_RANKS];         // 1 if that rank pushes ghosts to me
    int send_ptr[DIST_MAX_RANKS + 1];      // my boundary rows grouped by destination
    const int *send_rows;                  // device rows (node index) to send
    long long send_dst[DIST_MAX_RANKS];    // node offset of my segment in the destination's ghost tail
    // ghost tails of the exchanged vectors on every rank (peer-mapped): the solution x (double2 per
    // node) and the 3-vector images of the two SpMV inputs (double4 per node, fg_common.cuh OP_NODE3)
    double2 *tail[DIST_MAX_RANKS];         // x
    double4 *wtail[2][DIST_MAX_RANKS];     // [0] w of D.p  [1] w of D.s
    };

__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v)
    { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p)
    {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
    }
__device__ __forceinline__ unsigned long long now_ns()
    {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
    }
// spin until *flag >= e; false on timeout
__device__ inline bool wait_flag(const unsigned long long *flag, unsigned long long e)
    {
    if (ld_sys(flag) >= e) return true;
    const unsigned long long t0 = now_ns();
    unsigned int k = 0;
    while (ld_sys(flag) < e)
        if ((++k & 1023u) == 0 && now_ns() - t0 > DIST_TIMEOUT_NS) return false;
    return true;
    }
__device__ __forceinline__ void st_sys_f64(double *p, double v)
    { asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_sys_f64(const double *p)
    {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
    }

// All-reduce of nv <= DIST_NV values over the ranks, executed by ONE FULL WARP per rank (warp 0 of
// the last CTA of a reducing kernel).  v points to shared memory.  op_max = false: v holds nv values
// followed by their nv compensations (fg_reduce.cuh); the ranks' pairs are folded in rank order with
// the same error-free addition, result in place.  op_max = true: v holds nv values, maximum in place.
// `e` = the epoch of this all-reduce (the same number on every rank); `err` = where a timeout is recorded.
__device__ inline void dist_allreduce_warp_e(const DistDev *d, int *err, unsigned long long e, double *v, int nv,
                                             bool op_max);
__device__ inline void dist_allreduce_warp(DistDev *d, double *v, int nv, bool op_max)
    {
    unsigned long long e = 0;
    if ((threadIdx.x & 31) == 0) e = ++d->epoch;
    e = __shfl_sync(0xffffffffu, e, 0);
    dist_allreduce_warp_e(d, &d->error, e, v, nv, op_max);
    }
__device__ inline void dist_allreduce_warp_e(const DistDev *d, int *err, unsigned long long e, double *v, int nv,
                                             bool op_max)
    {
    __shared__ unsigned int rx[DIST_MAX_RANKS][4 * DIST_NV];
    const int lane = threadIdx.x & 31;
    const int nd = op_max ? nv : 2 * nv;  // doubles per rank
    if (*err)
        {  // a previous spin timed out: do not wait again, poison the scalars
        if (lane == 0)
            for (int k = 0; k < nd; k++) v[k] = nan("");
        __syncwarp();
        return;
        }
    const int par = (int)(e & 1ull), nw = 2 * nd, world = d->world, rank = d->rank;
    const unsigned long long tag = (e & 0xffffffffull) << 32;
    const unsigned int *v32 = reinterpret_cast<const unsigned int *>(v);
    for (int idx = lane; idx < world * nw; idx += 32)
        {
        const int q = idx / nw, w = idx - q * nw;
        st_sys(&d->ctrl[q]->ll[par][rank][w], tag | (unsigned long long)v32[w]);
        }
    DistCtrl *me = d->ctrl[rank];
    bool ok = true;
        {  // every lane polls its (up to 4) mailbox words together: independent loads per round, not one
           // L2 round trip per word (96 words on 8 ranks and 3 values)
        constexpr int MAXS = (DIST_MAX_RANKS * 4 * DIST_NV + 31) / 32;
        const int total = world * nw;
        unsigned int need = 0;  // bit q: word lane + 32 q has not arrived yet
#pragma unroll
        for (int q = 0; q < MAXS; q++)
            if (lane + 32 * q < total) need |= 1u << q;
        unsigned long long t0 = 0;
        unsigned int k = 0;
        while (need)
            {
            unsigned long long x[MAXS];
#pragma unroll
            for (int q = 0; q < MAXS; q++)
                {
                const int idx = lane + 32 * q, src = idx / nw;
                x[q] = (need >> q) & 1u ? ld_sys(&me->ll[par][src][idx - src * nw]) : 0ull;
                }
#pragma unroll
            for (int q = 0; q < MAXS; q++)
                if (((need >> q) & 1u) && (x[q] & 0xffffffff00000000ull) == tag)
                    {
                    const int idx = lane + 32 * q, src = idx / nw;
                    rx[src][idx - src * nw] = (unsigned int)x[q];
                    need &= ~(1u << q);
                    }
            if (need && (++k & 1023u) == 0)
                {
                if (t0 == 0)
                    t0 = now_ns();
                else if (now_ns() - t0 > DIST_TIMEOUT_NS)
                    {
                    ok = false;
                    break;
                    }
                }
            }
        }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    auto get = [&](int src, int k) { return __hiloint2double((int)rx[src][2 * k + 1], (int)rx[src][2 * k]); };
    if (!ok)
        {
        if (lane == 0)
            {
            *err = 1;
            for (int k = 0; k < nd; k++) v[k] = nan("");  // NaN residual => CANNOT_CONVERGE (iter.h:147)
            }
        }
    else if (lane < nv)
        {  // lane k folds value k over the ranks, in rank order: identical bits on every rank
        const int k = lane;
        if (op_max)
            {
            double s = get(0, k);
            for (int src = 1; src < world; src++) s = fmax(s, get(src, k));
            v[k] = s;
            }
        else
            {
            double s = get(0, k), c = get(0, nv + k);
            for (int src = 1; src < world; src++)
                {  // error-free fold
                const double bs = get(src, k), be = get(src, nv + k);
                const double t = s + bs, bb = t - s;
                const double er = (s - (t - bb)) + (bs - bb);
                s = t;
                c += be + er;
                }
            v[k] = s;
            v[nv + k] = c;
            }
        }
    __syncwarp();
    }

// Push f(row) for every boundary row of this rank into the neighbours' ghost tails `tails` of one
// exchanged vector; executed by `nthreads` cooperating threads (thread index tid).  Each thread fences its
// own stores at system scope, so a later flag / all-reduce by any thread of the grid that is
// ordered after them (block barrier + device-scope ticket) publishes them to the peers.
// FENCE = false: the caller orders the stores itself (the persistent kernel: block barrier, then one
// system-scope fence per CTA before it arrives at the grid barrier that raises the halo flag -- the pushes
// then overlap the CTA's own rows instead of stalling their threads for an NVLink round trip).
template <bool FENCE = true, class T, class F>
__device__ __forceinline__ void dist_push(const DistDev *d, T *const *tails, int tid, int nthreads, F f)
    {
    const int nsend = d->send_ptr[d->world];
    bool any = false;
    for (int idx = tid; idx < nsend; idx += nthreads)
        {
        int q = 0;
        while (idx >= d->send_ptr[q + 1]) q++;
        const T val = f(d->send_rows[idx]);
        *(tails[q] + d->send_dst[q] + (idx - d->send_ptr[q])) = val;
        any = true;
        }
    if (FENCE && any) __threadfence_system();
    }

// The same push dealt out in chunks of 32 consecutive send-list entries (one coalesced warp store into the
// peer's tail): chunk j goes to CTA j mod G, to its warp nw-1, nw-2, ... -- the warps of a CTA that are the
// last to receive a slice of an incomplete round (SliceIter), so the push rides in otherwise idle issue slots.
// No fence here (see dist_push<false>); returns true when this thread stored something.
template <class T, class F>
__device__ __forceinline__ bool dist_push_warps(const DistDev *d, T *const *tails, int nw, F f)
    {
    const int nsend = d->send_ptr[d->world];
    const int nchunk = (nsend + 31) >> 5, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool any = false;
    for (int j = blockIdx.x + gridDim.x * (nw - 1 - wid); j < nchunk; j += gridDim.x * nw)
        {
        const int idx = 32 * j + lane;
        if (idx >= nsend) continue;
        int q = 0;
        while (idx >= d->send_ptr[q + 1]) q++;
        const T val = f(d->send_rows[idx]);
        *(tails[q] + d->send_dst[q] + (idx - d->send_ptr[q])) = val;
        any = true;
        }
    return any;
    }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
    {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
    }
// spin with acquire loads until *flag >= e (no fence needed afterwards); false on timeout
__device__ inline bool wait_flag_acquire(const unsigned long long *flag, unsigned long long e)
    {
    if (ld_acquire_sys(flag) >= e) return true;
    const unsigned long long t0 = now_ns();
    unsigned int k = 0;
    while (ld_acquire_sys(flag) < e)
        if ((++k & 1023u) == 0 && now_ns() - t0 > DIST_TIMEOUT_NS) return false;
    return true;
    }

// After a dist_push by a whole CTA: one thread raises the halo epoch flag on every destination.
__device__ inline void dist_raise_e(const DistDev *d, unsigned long long e)
    {
    __threadfence_system();
    for (int q = 0; q < d->world; q++)
        if (d->send_ptr[q + 1] > d->send_ptr[q]) st_sys(&d->ctrl[q]->hflag[d->rank], e);
    }
__device__ inline void dist_raise(DistDev *d) { dist_raise_e(d, ++d->hepoch); }

// Consumer side: wait until every source rank has raised the current halo epoch (one thread).
__device__ inline void dist_wait(DistDev *d)
    {
    if (d->error) return;
    const unsigned long long e = d->hepoch;
    DistCtrl *me = d->ctrl[d->rank];
    for (int src = 0; src < d->world; src++)
        if (d->recv_from[src] && !wait_flag(&me->hflag[src], e))
            {
            d->error = 1;
            return;
            }
    __threadfence_system();
    }

}  // namespace fg
