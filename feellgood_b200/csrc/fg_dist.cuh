// fg_dist.cuh — multi-GPU plumbing of the row-block (slab) partitioned LLG solve, device side.
//
// One process per GPU; every rank owns a contiguous block of node rows (a slab along the axis the
// reference sorts nodes by, src/mesh.cpp:334-367) plus read-only ghost copies of the neighbours'
// boundary nodes.  The reference is single-process, so nothing here replaces reference code: this
// is the B200 execution model of SURVEY.md §8(e).
//
// Communication never leaves the kernels that produce or consume the data:
//   * all-reduce of the Krylov scalars: the last CTA of a reducing kernel stores its partial sums
//     into EVERY rank's mailbox through CUDA-IPC-mapped peer memory (NVLink P2P stores), fences at
//     system scope, raises a per-source epoch flag, then waits for the flags of all ranks and sums
//     the mailbox IN RANK ORDER — every rank obtains bit-identical scalars, hence takes identical
//     convergence decisions without any further agreement protocol;
//   * halo exchange: the owner pushes its boundary entries of a vector into the ghost tail of the
//     neighbour's copy of that vector (same peer mapping).  D.p: pushed by the last CTA of the kernel
//     whose reduction fixes beta (setup SpMV, k_bicg_xr) right after that reduction, followed by a
//     halo epoch flag on which the consuming SpMV spins; D.s: pushed at the start of k_bicg_s, whose
//     own all-reduce (|s|^2) is the barrier that publishes it; x: one k_halo launch per solve.
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
    double mailbox[2][DIST_MAX_RANKS][DIST_NV];          // [epoch parity][source rank][value]
    unsigned long long mflag[2][DIST_MAX_RANKS];         // epoch of the mailbox entry
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
    // ghost tails of the three exchanged vectors on every rank (peer-mapped), as double2 per node
    double2 *tail[3][DIST_MAX_RANKS];      // [0] x  [1] phat  [2] shat
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

// All-reduce of nv <= DIST_NV doubles over the ranks; called by ONE thread per rank (thread 0 of the
// last CTA of a reducing kernel).  op_max = false: sum in rank order; true: maximum.
__device__ inline void dist_allreduce(DistDev *d, double *v, int nv, bool op_max)
    {
    const unsigned long long e = ++d->epoch;
    const int par = (int)(e & 1ull);
    if (d->error)
        {  // a previous spin timed out: do not wait again, poison the scalars
        for (int k = 0; k < nv; k++) v[k] = nan("");
        return;
        }
    for (int q = 0; q < d->world; q++)
        for (int k = 0; k < nv; k++) st_sys_f64(&d->ctrl[q]->mailbox[par][d->rank][k], v[k]);
    __threadfence_system();
    for (int q = 0; q < d->world; q++) st_sys(&d->ctrl[q]->mflag[par][d->rank], e);
    DistCtrl *me = d->ctrl[d->rank];
    bool ok = true;
    for (int src = 0; src < d->world && ok; src++) ok = wait_flag(&me->mflag[par][src], e);
    __threadfence_system();
    if (!ok)
        {
        d->error = 1;
        for (int k = 0; k < nv; k++) v[k] = nan("");  // NaN residual => CANNOT_CONVERGE (iter.h:147)
        return;
        }
    for (int k = 0; k < nv; k++)
        {
        double s = ld_sys_f64(&me->mailbox[par][0][k]);
        for (int src = 1; src < d->world; src++)
            {
            const double t = ld_sys_f64(&me->mailbox[par][src][k]);
            s = op_max ? fmax(s, t) : s + t;
            }
        v[k] = s;
        }
    }

// Push f(row) for every boundary row of this rank into the neighbours' ghost tails of vector
// `which`; executed by `nthreads` cooperating threads (thread index tid).  Each thread fences its
// own stores at system scope, so a later flag / all-reduce by any thread of the grid that is
// ordered after them (block barrier + device-scope ticket) publishes them to the peers.
template <class F>
__device__ __forceinline__ void dist_push(const DistDev *d, int which, int tid, int nthreads, F f)
    {
    const int nsend = d->send_ptr[d->world];
    bool any = false;
    for (int idx = tid; idx < nsend; idx += nthreads)
        {
        int q = 0;
        while (idx >= d->send_ptr[q + 1]) q++;
        const double2 val = f(d->send_rows[idx]);
        *(d->tail[which][q] + d->send_dst[q] + (idx - d->send_ptr[q])) = val;
        any = true;
        }
    if (any) __threadfence_system();
    }

// After a dist_push by a whole CTA: one thread raises the halo epoch flag on every destination.
__device__ inline void dist_raise(DistDev *d)
    {
    __threadfence_system();
    const unsigned long long e = ++d->hepoch;
    for (int q = 0; q < d->world; q++)
        if (d->send_ptr[q + 1] > d->send_ptr[q]) st_sys(&d->ctrl[q]->hflag[d->rank], e);
    }

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
