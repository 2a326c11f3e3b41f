// fg_solve_pk.cuh — LinAlgebra::solve's Krylov part (src/solver.cpp:50-88: bicg_dir + node update) as ONE
// persistent cooperative kernel.  Included by fg_krylov.cu after the SpMV slice loop.
//
// The multi-kernel driver (bicgstab_run) pays a launch ramp, a last-CTA reduction tail and a kernel drain
// on each of its 5 kernels per iteration; at the per-rank size of an 8-GPU run (or on the reference's own
// 10^4..10^6-tetrahedron meshes) those fixed costs are larger than the memory time of the kernels.  Here
// the grid is resident for the whole solve:
//   * static row ownership (SliceIter): the grid sweeps the SELL slices as one front, a warp owns the same
//     slices in every phase, the incomplete last round is dealt out per CTA so that all SMs carry the same
//     load to one slice; every Krylov vector entry (r, p, v, s, t, x) is produced and consumed by the same
//     thread; the only data that crosses threads are the 3-vector images w gathered by the SpMV and the
//     scalars;
//   * 5 grid-wide synchronisations per iteration (reference order src/algebra/bicg.h:185-232):
//       A  p = r + beta (p - omega v), w_p = P D p            | barrier (+ halo flag)
//       B  v = K D p, (v, rt)                                 | reduce  -> alpha
//       C  s = r - alpha v, w_s = P D s, |s|^2                | one GPU: reduce -> exit test;
//                                                             | multi-GPU: barrier only (+ halo flag), |s|^2
//                                                             | travels with the next reduction and t is
//                                                             | computed speculatively (discarded on exit)
//       D  t = K D s, (t, s), (t, t)                          | reduce  -> omega
//       E  x += alpha D p + omega D s, r = s - omega t        | reduce  -> rho, loop test
//     i.e. 3 cross-GPU all-reduces per iteration instead of 4;
//   * the barrier is a generation counter in global memory; the LAST arriving CTA sums the CTA partials
//     in index order (error-free, fg_reduce.cuh), runs the cross-GPU all-reduce (fg_dist.cuh), raises the
//     halo flags of the phase it closes and only then releases the generation, so every CTA of every rank
//     reads bit-identical totals and runs the scalar state machine (fg_krylov_state.cuh) redundantly on a
//     shared-memory copy of KState: no broadcast kernel, no host round trip until the solve is over;
//   * multi-GPU: boundary images are pushed into the neighbours' ghost tails by the first CTAs at the
//     start of phases A and C; slices with ghost columns are processed last in B and D, after the halo
//     flag of every source was seen, so interior rows never wait for NVLink;
//   * the node update of src/solver.cpp:62-88 (gated on the failure predicate) is the last phase.
// Memory-model notes: a release/acquire pair at gpu scope (fence + atomic arrive / flag poll + fence by
// thread 0, bar.sync around them) orders every CTA's writes before every other CTA's later reads; the
// acquire fence also drops stale L1 lines.  Images are gathered with plain (coherent) loads issued as
// volatile asm (ld_image<true>), never through ld.global.nc.
#pragma once

namespace fg
{
constexpr int PK_MAX_GRID = 1024;

struct PkSync  // one per context, device memory, zeroed at creation
    {
    unsigned int count;  // arrivals at the open barrier
    unsigned int pad0[31];
    unsigned int gen;    // generation of the last released barrier
    unsigned int pad1[31];
    unsigned long long t_last, t_ar;          // profiling: when the last CTA arrived / the cross-GPU part ended
    // totals of the last two reductions (slot = parity), as self-validating 8-byte words {half of the double,
    // generation}: a CTA that sees the generation in a word has the data with it, so the totals need no
    // second round trip after the release of the barrier (two words per value)
    unsigned long long tot_ll[2][2 * RED_NV];
    double part[2][2 * RED_NV][PK_MAX_GRID];  // CTA partials (values, then compensations)
    };

struct PkArgs
    {
    Operator op;
    int NODp, NODt;
    double *x, *b, *r, *rt, *p, *p2, *v, *s, *t;
    const double *D;
    double4 *w3p, *w3s;
    const unsigned char *mask;
    KState *st;
    PkSync *sync;
    DistDev *dist;
    double tol;
    int maxiter;
    // fused node update (cur == NULL: none)
    const unsigned char *nonmag;
    const NodeRec *cur;
    NodeRec *next;
    const Basis *basis;
    double dt;
    unsigned long long *phase_acc;  // optional [64]: per phase id summed ns [0,16), counts [16,32), ns until the
                                    // last CTA of this GPU arrived at the closing barrier [32,48), ns of the
                                    // cross-GPU part of that barrier (all-reduce over the ranks) [48,64)
    // result mailbox in mapped host memory (NULL: the host copies KState itself)
    KState *h_st;
    unsigned long long *h_seq;
    unsigned long long seq;
    };


template <int BS> struct PkShared
    {
    KState ks;
    double red_s[RED_NV][BS / 32], red_e[RED_NV][BS / 32];
    double tot[2 * RED_NV];
    unsigned int gen;            // barrier generation this CTA waits for next (thread 0)
    unsigned int nred;           // reductions so far (slot parity)
    unsigned long long hepoch;   // halo epoch this CTA expects next
    unsigned long long t_prev;   // time stamp of the previous phase boundary (CTA 0, thread 0)
    int pushed;                  // a thread of this CTA stored into peer memory since the last halo barrier
    // multi-GPU: this CTA's copy of the exchange descriptor (ranks, peer pointers, halo plan: read in every
    // push and all-reduce, each field a dependent L2 round trip when left in global memory); `epoch` and
    // sh.hepoch advance in every CTA alike and go back to global memory when the kernel ends
    DistDev dd;
    int derr;                    // a spin of this CTA timed out
    int last;                    // this CTA arrived last at the open barrier (thread 0 tells its team)
    // halo epochs: the first warp of the CTA that needs epoch e claims it and polls the flags in global
    // memory; the other warps wait for `halo_seen` here (one system-scope poller per CTA, not one per warp)
    unsigned long long halo_claim, halo_seen;
    };

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
    {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
    }
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v)
    { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
    {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
    }
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
    { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_u32(unsigned int *p, unsigned int v)
    { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// arrival at the grid barrier: one acquire-release atomic (fence.acq_rel + ATOMG; `__threadfence()` would be
// the heavier sequentially-consistent MEMBAR.SC on top of it)
__device__ __forceinline__ unsigned int atom_add_acq_rel_u32(unsigned int *p, unsigned int v)
    {
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
    }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

template <int BS> __device__ __forceinline__ void pk_stamp(const PkArgs &a, PkShared<BS> &sh, int id)
    {
    if (a.phase_acc != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
        {  // phase `id` = the time since the previous boundary, as CTA 0 saw it (barrier wait included)
        const unsigned long long t = now_ns();
        if (id != PKP_START)
            {
            a.phase_acc[id] += t - sh.t_prev;
            a.phase_acc[16 + id] += 1ull;
            if (gridDim.x > 1 || a.dist != nullptr)
                {
                const unsigned long long tl = __ldcg(&a.sync->t_last), ta = __ldcg(&a.sync->t_ar);
                if (tl >= sh.t_prev && ta >= tl && t >= ta)
                    {
                    a.phase_acc[32 + id] += tl - sh.t_prev;
                    a.phase_acc[48 + id] += ta - tl;
                    }
                }
            }
        sh.t_prev = t;
        }
    }

// Grid barrier with an optional sum (NV > 0) or maximum (MAXOP) over the grid and the ranks.
// halo: 0 none, 1 raise the halo flags on the neighbours when the barrier closes (every CTA's pushes of
// this phase are then complete and fenced).  On return sh.tot[0..NV) holds the totals (all threads).
template <int BS, int NV, bool MAXOP>
__device__ void pk_sync(const PkArgs &a, PkShared<BS> &sh, const double (&acc)[RED_NV], const int halo)
    {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int NW = (int)blockDim.x >> 5;  // warps really launched (<= BS / 32, see pk_plan)
    if (NV > 0)
        {
#pragma unroll
        for (int k = 0; k < NV; k++)
            {
            double s = acc[k], e = 0.0;
            if (MAXOP)
                {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
                }
            else
                warp_sum_dd(s, e);
            if (lane == 0)
                {
                sh.red_s[k][wid] = s;
                sh.red_e[k][wid] = e;
                }
            }
        }
    __syncthreads();  // every thread's phase work (and its partial) is done
    // Warp k < NV handles value k (the CTA partial and, in the last arriving CTA, the sum over the CTAs: the
    // values are summed side by side, not one after the other); warp 0 does the barrier itself.
    // (a team of one warp synchronising with __syncwarp measured 3 us slower per reduction than two warps on a
    // named barrier -- r02n/r02o -- so a single value still gets a team of two)
    constexpr int NVW = NV > 1 ? NV : (NV == 1 ? 2 : 1);
    auto team_sync = []()
        {
        if (NVW > 1)
            asm volatile("bar.sync 15, %0;" ::"n"(32 * NVW) : "memory");
        else
            __syncwarp();
        };
    if (wid < NVW)
        {
        const bool solo = gridDim.x == 1 && a.dist == nullptr;
        const unsigned int slot = sh.nred & 1u;
        const int k = wid;  // this warp's value
        if (NV > 0 && k < NV)
            {  // CTA partial, fixed tree
            double s = lane < NW ? sh.red_s[k][lane] : 0.0, e = lane < NW ? sh.red_e[k][lane] : 0.0;
            if (MAXOP)
                {
                if (lane >= NW) s = -1.7976931348623157e308;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
                }
            else
                warp_sum_dd(s, e);
            if (lane == 0)
                {
                if (solo)
                    sh.tot[k] = MAXOP ? s : s + e;
                else
                    {
                    a.sync->part[slot][k][blockIdx.x] = s;
                    if (!MAXOP) a.sync->part[slot][RED_NV + k][blockIdx.x] = e;
                    }
                }
            }
        if (!solo)
            {
            team_sync();  // every value's partial is written before thread 0 publishes them
            unsigned int target = 0;
            if (wid == 0 && lane == 0)
                {
                if (NV > 0 && a.dist != nullptr) sh.dd.epoch++;  // the epoch of this all-reduce, in every CTA alike
                target = ++sh.gen;
                // release + acquire in one atomic.  A barrier that raises the halo flags also publishes this
                // CTA's pushes into peer memory (plain stores by any of its threads, ordered before this point
                // by the bar.sync above): one system-scope fence per pushing CTA, issued when the stores have
                // long been on their way
                if (halo && a.dist != nullptr && sh.pushed) fence_acq_rel_sys();
                const unsigned int last = atom_add_acq_rel_u32(&a.sync->count, 1u) == gridDim.x - 1 ? 1u : 0u;
                if (last && a.phase_acc != nullptr) a.sync->t_last = now_ns();
                sh.last = (int)last;
                }
            team_sync();
            const bool last = sh.last != 0;
            if (last && NV > 0 && k < NV)
                {  // value k over the CTAs, in index order; every partial of this lane is fetched before the
                   // first addition (independent L2 round trips)
                constexpr int PF = 5;  // partials per lane held in registers: grids up to 160 CTAs in one go
                double s = MAXOP ? -1.7976931348623157e308 : 0.0, e = 0.0;
                for (int i0 = 0; i0 < (int)gridDim.x; i0 += 32 * PF)
                    {
                    double vs[PF], ve[PF];
#pragma unroll
                    for (int u = 0; u < PF; u++)
                        {
                        const int i = i0 + lane + 32 * u;
                        const bool in = i < (int)gridDim.x;
                        vs[u] = in ? __ldcg(&a.sync->part[slot][k][i]) : (MAXOP ? -1.7976931348623157e308 : 0.0);
                        ve[u] = (in && !MAXOP) ? __ldcg(&a.sync->part[slot][RED_NV + k][i]) : 0.0;
                        }
#pragma unroll
                    for (int u = 0; u < PF; u++)
                        {
                        if (MAXOP)
                            s = fmax(s, vs[u]);
                        else
                            dd_add(s, e, vs[u], ve[u]);
                        }
                    }
                if (MAXOP)
                    {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
                    }
                else
                    warp_sum_dd(s, e);
                if (lane == 0)
                    {
                    sh.tot[k] = s;
                    sh.tot[NV + k] = e;
                    }
                }
            if (last) team_sync();  // all NV sums are in sh.tot (uniform: every team warp read the same sh.last)
            if (wid == 0)
                {
                const unsigned int tgt = __shfl_sync(0xffffffffu, target, 0);
                if (last)
                    {
                    if (NV > 0 && a.dist != nullptr) dist_allreduce_warp_e(&sh.dd, &sh.derr, sh.dd.epoch, sh.tot, NV, MAXOP);
                    if (lane == 0)
                        {
                        a.sync->count = 0;
                        if (a.phase_acc != nullptr) a.sync->t_ar = now_ns();
                        if (NV > 0)
                            {  // release pattern: one fence, then the words the other CTAs spin on
                            fence_acq_rel_gpu();
#pragma unroll
                            for (int q = 0; q < NV; q++)
                                {
                                const double tq = MAXOP ? sh.tot[q] : sh.tot[q] + sh.tot[NV + q];
                                const unsigned long long bits = (unsigned long long)__double_as_longlong(tq);
                                const unsigned long long flag = (unsigned long long)tgt << 32;
                                st_relaxed_u64(&a.sync->tot_ll[slot][2 * q], flag | (bits & 0xffffffffull));
                                st_relaxed_u64(&a.sync->tot_ll[slot][2 * q + 1], flag | (bits >> 32));
                                sh.tot[q] = tq;
                                }
                            st_relaxed_u32(&a.sync->gen, tgt);  // kept current: the next launch starts from it
                            }
                        else
                            st_release_u32(&a.sync->gen, tgt);  // (release store) this GPU's CTAs go on ...
                        // ... while the neighbours learn that every push of the phase is complete and fenced
                        if (halo && a.dist != nullptr) dist_raise_e(&sh.dd, sh.hepoch);
                        }
                    }
                else if (NV > 0)
                    {  // lanes 0 .. 2 NV - 1 each wait for their word of the totals: the acquire load orders
                       // everything that follows after the releasing fence (for these threads; for the rest
                       // of the CTA through the bar.sync below) and drops stale L1 lines
                    unsigned long long wv = 0ull;
                    if (lane < 2 * NV)
                        do
                            wv = ld_acquire_u64(&a.sync->tot_ll[slot][lane]);
                        while ((unsigned int)(wv >> 32) != tgt);
                    const unsigned int half = (unsigned int)wv;
                    const unsigned int lo = __shfl_sync(0xffffffffu, half, (2 * lane) & 31);
                    const unsigned int hi = __shfl_sync(0xffffffffu, half, (2 * lane + 1) & 31);
                    if (lane < NV) sh.tot[lane] = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
                    }
                else if (lane == 0)
                    {
                    while (ld_acquire_u32(&a.sync->gen) != tgt)
                        ;
                    }
                __syncwarp();
                }
            }
        if (wid == 0 && lane == 0)
            {
            if (NV > 0) sh.nred++;
            if (halo)
                {
                sh.hepoch++;
                sh.pushed = 0;
                }
            }
        }
    __syncthreads();
    }

// multi-GPU consumer side: one lane waits until every source rank has raised halo epoch e
__device__ inline void pk_halo_wait(const DistDev *d, int *err, unsigned long long e)
    {
    if (*err) return;
    DistCtrl *me = d->ctrl[d->rank];
    // acquire loads at system scope order the pushed entries before the flag for this thread (and, through
    // the __syncwarp that follows, for its warp); stale L1 copies of the ghost rows cannot exist: they were
    // dropped by the fence of the last grid barrier and nobody reads a ghost row before its flag
    for (int src = 0; src < d->world; src++)
        if (d->recv_from[src] && !wait_flag_acquire(&me->hflag[src], e))
            {
            *err = 1;
            return;
            }
    }

// The same for a whole CTA: the first caller polls, later callers (lane 0 of other warps) wait in shared
// memory.  The poller's acquire drops the SM's L1 (CCTL.IVALL), which is the L1 of every warp of the CTA.
// (Measured against one poller per warp at N = 2, film20m: -2 us of work per product, -0.5 % per solve.)
template <int BS> __device__ __forceinline__ void pk_halo_wait_cta(PkShared<BS> &sh, unsigned long long e)
    {
    volatile unsigned long long *seen = &sh.halo_seen;
    if (*seen < e)
        {
        if (atomicMax(&sh.halo_claim, e) < e)
            {
            pk_halo_wait(&sh.dd, &sh.derr, e);  // on a timeout derr is set and the kernel ends through its error path
            __threadfence_block();
            *seen = e;
            }
        else
            while (*seen < e) {}
        }
    __threadfence_block();
    }

// ---- gather blocks staged in shared memory (fg_setup.hpp, Operator::lcol) --------------------------------
// A scattered 32-byte gather costs one L1 tag-stage wavefront per lane whether it hits or not: 62 M stored
// pairs on the 20 M-tet mesh = 213 us of wavefronts per SM and product, which is what the SpMV measured
// (217-227 us) before this path existed.  Here a group of 4 warps copies the images its block of 256 rows
// needs (the block's own rows, contiguous, plus its halo list) into shared memory once with cp.async, and the
// stored pairs read them with 16-bit local indices through the shared-memory pipe (8 cycles per warp and
// pair instead of 32).  Each warp of the group owns slices w and 7 - w of the block: the slices of a block
// are sorted by width, so every warp gets the same number of pairs.
constexpr int PK_GROUP = 128;          // threads per group
constexpr int PK_SPB = 8;              // slices per gather block
__device__ __forceinline__ void pk_group_sync(int gid)
    { asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(PK_GROUP) : "memory"); }
__device__ __forceinline__ void cp_async16(unsigned int dst_smem, const void *src)
    { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int STAGE>
__device__ __forceinline__ void pk_slice_staged(const Operator &op, const SpmvArgs &a, const int s, const int sl,
                                                const int p0, const int p1, const double4 *stage, const int lane,
                                                double (&acc)[RED_NV])
    {
    const int row = s * SLICE + lane;
    const unsigned short *cp = op.lcol + (size_t)p0 * SLICE + lane;
    const double *sp = op.val + (size_t)p0 * SLICE + lane;
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
        {
        constexpr int U = 8;
        int cn[U];
#pragma unroll
        for (int u = 0; u < U; u++) cn[u] = p0 + u < p1 ? (int)__ldcs(cp + u * SLICE) : 0;
        for (int j = p0; j < p1; j += U)
            {
            int c[U];
            double Sv[U];
#pragma unroll
            for (int u = 0; u < U; u++)
                {
                c[u] = cn[u];
                Sv[u] = j + u < p1 ? __ldcs(sp + u * SLICE) : 0.0;
                }
            cp += U * SLICE;
            sp += U * SLICE;
#pragma unroll
            for (int u = 0; u < U; u++) cn[u] = j + U + u < p1 ? (int)__ldcs(cp + u * SLICE) : 0;
#pragma unroll
            for (int u = 0; u < U; u++)
                {
                const double4 wb = stage[c[u]];
                z0 += Sv[u] * wb.x;
                z1 += Sv[u] * wb.y;
                z2 += Sv[u] * wb.z;
                }
            }
        }
    double ep[3], eq[3];
    load_basis_k(op.qbasis + row, ep, eq);
    double2 xa;
    if (a.x != nullptr)
        xa = reinterpret_cast<const double2 *>(a.x)[row];
    else
        {  // x_a = P_a^T w_a: the node's own image is entry (slice in block, lane) of the staging buffer
        const double4 wr = stage[sl * SLICE + lane];
        xa = make_double2(ep[0] * wr.x + ep[1] * wr.y + ep[2] * wr.z, eq[0] * wr.x + eq[1] * wr.y + eq[2] * wr.z);
        }
    const double2 dm = __ldcs(op.Dm + row);
    double y0 = op.cS * (eq[0] * z0 + eq[1] * z1 + eq[2] * z2) + (dm.y * xa.x + dm.x * xa.y);
    double y1 = op.cS * (ep[0] * z0 + ep[1] * z1 + ep[2] * z2) + (dm.x * xa.x - dm.y * xa.y);
    if (op.nonmag[row] != 0)
        {  // identity row (src/solver.cpp:46-48)
        y0 = xa.x;
        y1 = xa.y;
        }
    spmv_row2<STAGE>(a, row, y0, y1, acc);
    }

// the blocks of this group (pass: 0 all | 1 blocks without ghost rows | 2 blocks with ghost rows)
template <int STAGE>
__device__ __forceinline__ void pk_blocks_staged(const PkArgs &a, const SpmvArgs &sa, const SliceIter &own,
                                                 const int pass, double4 *stage, double (&acc)[RED_NV])
    {
    const Operator &op = a.op;
    const int lane = threadIdx.x & 31, w4 = (threadIdx.x >> 5) & 3, tg = threadIdx.x & (PK_GROUP - 1),
              gid = threadIdx.x / PK_GROUP;
    const unsigned int sbase = (unsigned int)__cvta_generic_to_shared(stage);
    const char *wbytes = reinterpret_cast<const char *>(sa.w);
    for (int b = own.begin(); b < op.nblock; b = own.next(b))
        {
        if (pass != 0 && (op.bghost[b] != 0) != (pass == 2)) continue;
        // slice extents of this warp's two slices, fetched under the shadow of the staging copies
        const int sA = PK_SPB * b + w4, sB = PK_SPB * b + (PK_SPB - 1) - w4;
        int pA0 = 0, pA1 = 0, pB0 = 0, pB1 = 0;
        if (sA < op.nslice)
            {
            pA0 = __ldg(op.ptr + sA);
            pA1 = __ldg(op.ptr + sA + 1);
            }
        if (sB < op.nslice)
            {
            pB0 = __ldg(op.ptr + sB);
            pB1 = __ldg(op.ptr + sB + 1);
            }
        pk_group_sync(gid);  // the group is done with the images of its previous block
        const int row0 = b * (PK_SPB * SLICE);
        const int nown = min(PK_SPB * SLICE, a.NODp - row0);
        const char *src = wbytes + (size_t)row0 * 32;
        for (int i = tg; i < 2 * nown; i += PK_GROUP) cp_async16(sbase + 16 * i, src + 16 * i);
        const int h0 = __ldg(op.bptr + b), nh = __ldg(op.bptr + b + 1) - h0;
        for (int i = tg; i < 2 * nh; i += PK_GROUP)
            {
            const int g = __ldg(op.bhalo + h0 + (i >> 1));
            cp_async16(sbase + PK_SPB * SLICE * 32 + 16 * i, wbytes + (size_t)g * 32 + 16 * (i & 1));
            }
        cp_async_wait_all();
        pk_group_sync(gid);
        if (sA < op.nslice) pk_slice_staged<STAGE>(op, sa, sA, w4, pA0, pA1, stage, lane, acc);
        if (sB < op.nslice) pk_slice_staged<STAGE>(op, sa, sB, (PK_SPB - 1) - w4, pB0, pB1, stage, lane, acc);
        }
    }

template <int STAGE, bool IDX16, bool STAGED, int BS, bool HEAD>
__device__ __forceinline__ void pk_spmv(const PkArgs &a, PkShared<BS> &sh, const SpmvArgs &sa, const SliceIter own,
                                        const int lane, const bool wait_halo, double4 *stage, double (&acc)[RED_NV],
                                        const SpmvHead &head)
    {
    if (STAGED)
        {
        if (a.dist != nullptr && a.op.bghost != nullptr && wait_halo)
            {
            pk_blocks_staged<STAGE>(a, sa, own, 1, stage, acc);
            // blocks whose halo has ghost rows: wait for the neighbours' pushes of this phase (per group:
            // only groups that own such a block wait; the condition is uniform inside a group)
            int b = own.begin();
            while (b < a.op.nblock && a.op.bghost[b] == 0) b = own.next(b);
            if (b < a.op.nblock)
                {
                if (lane == 0) pk_halo_wait_cta(sh, sh.hepoch - 1);
                __syncwarp();
                pk_blocks_staged<STAGE>(a, sa, own, 2, stage, acc);
                }
            }
        else
            pk_blocks_staged<STAGE>(a, sa, own, 0, stage, acc);
        return;
        }
    if (a.dist != nullptr && a.op.sghost != nullptr)
        {  // a partition: slices without ghost columns first (16-bit offsets when they fit), then -- after the
           // halo flags of the phase, if there is a halo to wait for -- the slices with ghost columns, whose
           // far columns need the 32-bit indices unless every offset of the mesh fits
        const bool has_ghost = spmv_node3_slices<STAGE, IDX16, true, HEAD>(a.op, sa, own, lane, 1, acc, head);
        if (has_ghost)
            {  // per warp: only warps that own such a slice wait
            if (wait_halo)
                {
                if (lane == 0) pk_halo_wait_cta(sh, sh.hepoch - 1);
                __syncwarp();
                }
            if (IDX16 && a.op.col16_partial)
                spmv_node3_slices<STAGE, false, true>(a.op, sa, own, lane, 2, acc);
            else
                spmv_node3_slices<STAGE, IDX16, true>(a.op, sa, own, lane, 2, acc);
            }
        }
    else
        spmv_node3_slices<STAGE, IDX16, true, HEAD>(a.op, sa, own, lane, 0, acc, head);
    }

// every slice this warp owns: plain slices, or slices w and 7 - w of the group's gather blocks
template <bool STAGED, class F>
__device__ __forceinline__ void pk_for_slices(const SliceIter &own, const Operator &op, F f)
    {
    if (STAGED)
        {
        const int w4 = (threadIdx.x >> 5) & 3;
        for (int b = own.begin(); b < op.nblock; b = own.next(b))
            {
            const int sA = PK_SPB * b + w4, sB = PK_SPB * b + (PK_SPB - 1) - w4;
            if (sA < op.nslice) f(sA);
            if (sB < op.nslice) f(sB);
            }
        }
    else
        for (int s = own.begin(); s < op.nslice; s = own.next(s)) f(s);
    }

extern __shared__ double4 pk_stage_smem[];  // STAGED: one staging buffer of stage_cap images per thread group

// HEAD: the head of each warp's first slice (extent, first batch of column indices and values: spmv_head) is
// fetched once and kept in registers for the whole solve.  27 registers that the 64 of a full SM of threads do
// not have (ptxas spilled ~600 bytes per thread all over the kernel), so this is the variant of meshes small
// enough for 16 warps per SM to own a slice each: one CTA of 512 threads per SM, up to 128 registers.
template <int BS, bool IDX16, bool STAGED, bool HEAD = false>
__global__ void __launch_bounds__(BS, (BS == 1024 || HEAD) ? 1 : (1024 / BS)) k_llg_solve(const PkArgs a)
    {
    __shared__ PkShared<BS> sh;
    const int lane = threadIdx.x & 31;
    // Row ownership (SliceIter, fg_krylov.cu): a front of W slices per round, the last round dealt out per CTA
    // STAGED: the units are gather blocks of 8 slices, owned by thread groups of 4 warps
    // BS is the LARGEST CTA of this instantiation; the launch may use fewer warps (pk_plan picks the count that
    // fills the last round of slices best), so strides and ownership come from blockDim
    const int nthr = (int)blockDim.x;
    const SliceIter own = STAGED ? slices_balanced(a.op.nblock, nthr / PK_GROUP, threadIdx.x / PK_GROUP)
                                 : slices_balanced(a.op.nslice, nthr / 32, threadIdx.x >> 5);
    double4 *const stage = STAGED ? pk_stage_smem + (size_t)(threadIdx.x / PK_GROUP) * (size_t)((a.op.stage_cap + 3) & ~3)
                                  : nullptr;
    const int gtid = blockIdx.x * nthr + threadIdx.x, gthreads = gridDim.x * nthr;
    const bool spec = a.dist != nullptr;  // speculative second SpMV: one all-reduce less per iteration
    if (a.dist != nullptr)
        {  // the exchange descriptor into shared memory, word by word
        const unsigned int *src = reinterpret_cast<const unsigned int *>(a.dist);
        unsigned int *dst = reinterpret_cast<unsigned int *>(&sh.dd);
        for (int i = threadIdx.x; i < (int)(sizeof(DistDev) / sizeof(unsigned int)); i += nthr) dst[i] = src[i];
        }
    __syncthreads();
    if (threadIdx.x == 0)
        {
        sh.derr = a.dist != nullptr ? sh.dd.error : 0;
        sh.ks = *a.st;
        kstate_reset(&sh.ks, a.tol, a.maxiter);
        if (blockIdx.x != 0) sh.ks.hist = nullptr;  // the history is recorded once
        sh.gen = ld_acquire_u32(&a.sync->gen);
        sh.nred = 0;
        sh.hepoch = a.dist != nullptr ? sh.dd.hepoch + 1 : 1;
        sh.t_prev = 0ull;
        sh.pushed = 0;
        sh.halo_claim = sh.halo_seen = 0ull;
        }
    __syncthreads();
    pk_stamp(a, sh, PKP_START);
    double acc[RED_NV];
    const double2 *D2 = reinterpret_cast<const double2 *>(a.D);
    SpmvArgs sa = {};
    sa.mask = a.mask;

    // ---- r = b - K x0 (masked); rt = r (p = r implicit); |b|^2, |r|^2           (bicg.h:172-183)
#pragma unroll
    for (int k = 0; k < RED_NV; k++) acc[k] = 0.0;
    sa.w = a.w3p;  // image of the initial guess, written by the assembly (ghost rows: k_ghost_guess)
    sa.x = a.x;
    sa.y = a.r;
    sa.a0 = a.b;
    sa.o0 = a.rt;
    SpmvHead head;
    head.s = -1;
    if (HEAD && !STAGED) spmv_head<IDX16>(a.op, own, lane, head);  // constant for the whole solve
    pk_spmv<ST_BICG_SETUP, IDX16, STAGED, BS, HEAD>(a, sh, sa, own, lane, false, stage, acc, head);
    pk_sync<BS, 2, false>(a, sh, acc, 0);
    if (threadIdx.x == 0)
        {
        const double tot[RED_NV] = {sh.tot[0], sh.tot[1], 0.0, 0.0};
        spmv_finalize<ST_BICG_SETUP>(&sh.ks, tot);
        }
    __syncthreads();
    pk_stamp(a, sh, PKP_SETUP);

    int it = 0;
    double *p_new = a.p;  // the direction of the last completed phase A
    while (!sh.ks.done)
        {
        // ---- A: p = r + beta (p - omega v) ; w_p = P D p                        (bicg.h:196-202)
        const double2 *p_old2 = reinterpret_cast<const double2 *>((it & 1) ? a.p2 : a.p);
        p_new = (it & 1) ? a.p : a.p2;
            {
            const bool first = sh.ks.nit == 0;
            const double omega = sh.ks.omega;
            const double beta = first ? 0.0 : bicg_beta(&sh.ks);
            const double2 *r2 = reinterpret_cast<const double2 *>(a.r), *v2 = reinterpret_cast<const double2 *>(a.v);
            auto value = [&](int row, double2 &pi)
                {
                const double2 rr = r2[row], d = D2[row];
                if (first)
                    pi = rr;  // p = r (bicg.h:183)
                else
                    {
                    const double2 pp = p_old2[row], vv = v2[row];
                    pi = make_double2(bicg_p_value(pp.x, vv.x, rr.x, omega, beta), bicg_p_value(pp.y, vv.y, rr.y, omega, beta));
                    }
                return make_double2(d.x * pi.x, d.y * pi.y);
                };
            if (a.dist != nullptr)
                {
                if (dist_push_warps(&sh.dd, sh.dd.wtail[0], nthr / 32, [&](int row)
                        {
                        double2 pi;
                        const double2 ph = value(row, pi);
                        return node_w(a.op.qbasis + row, ph.x, ph.y);
                        }))
                    sh.pushed = 1;
                }
            pk_for_slices<STAGED>(own, a.op, [&](const int s)
                {
                const int row = s * SLICE + lane;
                double2 pi;
                const double2 ph = value(row, pi);
                reinterpret_cast<double2 *>(p_new)[row] = pi;
                st256(a.w3p + row, node_w(a.op.qbasis + row, ph.x, ph.y));
                });
            }
        pk_sync<BS, 0, false>(a, sh, acc, 1);
        pk_stamp(a, sh, PKP_A);

        // ---- B: v = K D p (masked) ; (v, rt) -> alpha                           (bicg.h:203-206)
#pragma unroll
        for (int k = 0; k < RED_NV; k++) acc[k] = 0.0;
        sa.w = a.w3p;
        sa.x = nullptr;
        sa.y = a.v;
        sa.a0 = a.rt;
        sa.o0 = nullptr;
        pk_spmv<ST_BICG_V, IDX16, STAGED, BS, HEAD>(a, sh, sa, own, lane, true, stage, acc, head);
        pk_sync<BS, 1, false>(a, sh, acc, 0);
        if (threadIdx.x == 0)
            {
            const double tot[RED_NV] = {sh.tot[0], 0.0, 0.0, 0.0};
            spmv_finalize<ST_BICG_V>(&sh.ks, tot);
            }
        __syncthreads();
        pk_stamp(a, sh, PKP_B);

        // ---- C: s = r - alpha v ; w_s = P D s ; |s|^2                           (bicg.h:207-218)
        double ss_acc = 0.0;
            {
            const double alpha = sh.ks.alpha;
            const double2 *r2 = reinterpret_cast<const double2 *>(a.r), *v2 = reinterpret_cast<const double2 *>(a.v);
            auto value = [&](int row, double2 &si)
                {
                const double2 rr = r2[row], vv = v2[row], d = D2[row];
                si = make_double2(bicg_s_value(rr.x, vv.x, alpha), bicg_s_value(rr.y, vv.y, alpha));
                return make_double2(d.x * si.x, d.y * si.y);
                };
            if (a.dist != nullptr)
                {
                if (dist_push_warps(&sh.dd, sh.dd.wtail[1], nthr / 32, [&](int row)
                        {
                        double2 si;
                        const double2 sh_ = value(row, si);
                        return node_w(a.op.qbasis + row, sh_.x, sh_.y);
                        }))
                    sh.pushed = 1;
                }
            pk_for_slices<STAGED>(own, a.op, [&](const int s)
                {
                const int row = s * SLICE + lane;
                double2 si;
                const double2 sh_ = value(row, si);
                reinterpret_cast<double2 *>(a.s)[row] = si;
                st256(a.w3s + row, node_w(a.op.qbasis + row, sh_.x, sh_.y));
                ss_acc += si.x * si.x + si.y * si.y;
                });
            }
        if (!spec)
            {
            acc[0] = ss_acc;
            pk_sync<BS, 1, false>(a, sh, acc, 0);
            if (threadIdx.x == 0) bicg_s_finalize(&sh.ks, sh.tot[0]);
            __syncthreads();
            }
        else
            pk_sync<BS, 0, false>(a, sh, acc, 1);
        pk_stamp(a, sh, PKP_C);

        // ---- D: t = K D s (masked) ; (t, s), (t, t) -> omega                    (bicg.h:219-222)
        if (spec || !sh.ks.done)
            {
#pragma unroll
            for (int k = 0; k < RED_NV; k++) acc[k] = 0.0;
            sa.w = a.w3s;
            sa.x = nullptr;
            sa.y = a.t;
            sa.a0 = a.s;
            pk_spmv<ST_BICG_T, IDX16, STAGED, BS, HEAD>(a, sh, sa, own, lane, true, stage, acc, head);
            if (spec)
                {
                acc[2] = ss_acc;
                pk_sync<BS, 3, false>(a, sh, acc, 0);
                if (threadIdx.x == 0)
                    {
                    bicg_s_finalize(&sh.ks, sh.tot[2]);
                    if (!sh.ks.done)
                        {
                        const double tot[RED_NV] = {sh.tot[0], sh.tot[1], 0.0, 0.0};
                        spmv_finalize<ST_BICG_T>(&sh.ks, tot);
                        }
                    }
                }
            else
                {
                pk_sync<BS, 2, false>(a, sh, acc, 0);
                if (threadIdx.x == 0)
                    {
                    const double tot[RED_NV] = {sh.tot[0], sh.tot[1], 0.0, 0.0};
                    spmv_finalize<ST_BICG_T>(&sh.ks, tot);
                    }
                }
            __syncthreads();
            }
        pk_stamp(a, sh, PKP_D);

        // ---- E: x += alpha D p + omega D s ; r = s - omega t ; |r|^2, (rt, r)   (bicg.h:223-231, :185-195)
        //      (loop ended on |s|: only x += alpha D p)
            {
            const int fh = sh.ks.final_half;
            const bool skip = sh.ks.done && !fh;  // overflow / breakdown seen at the |s| test: x stays
            const double alpha = sh.ks.alpha, omega = sh.ks.omega;
            double2 *x2 = reinterpret_cast<double2 *>(a.x), *r2 = reinterpret_cast<double2 *>(a.r);
            const double2 *p2 = reinterpret_cast<const double2 *>(p_new), *s2 = reinterpret_cast<const double2 *>(a.s),
                          *t2 = reinterpret_cast<const double2 *>(a.t), *rt2 = reinterpret_cast<const double2 *>(a.rt);
#pragma unroll
            for (int k = 0; k < RED_NV; k++) acc[k] = 0.0;
            if (skip)
                ;
            else if (fh)
                {
                pk_for_slices<STAGED>(own, a.op, [&](const int s)
                    {
                    const int row = s * SLICE + lane;
                    const double2 d = D2[row], pp = p2[row];
                    double2 xv = x2[row];
                    xv.x += alpha * __dmul_rn(d.x, pp.x);
                    xv.y += alpha * __dmul_rn(d.y, pp.y);
                    x2[row] = xv;
                    });
                }
            else
                {
                pk_for_slices<STAGED>(own, a.op, [&](const int s)
                    {
                    const int row = s * SLICE + lane;
                    const double2 d = D2[row], pp = p2[row], sv = s2[row], tv = t2[row], rtv = rt2[row];
                    double2 xv = x2[row];
                    xv.x = (xv.x + alpha * __dmul_rn(d.x, pp.x)) + omega * __dmul_rn(d.x, sv.x);
                    xv.y = (xv.y + alpha * __dmul_rn(d.y, pp.y)) + omega * __dmul_rn(d.y, sv.y);
                    x2[row] = xv;
                    const double2 ri = make_double2(sv.x - omega * tv.x, sv.y - omega * tv.y);
                    r2[row] = ri;
                    acc[0] += ri.x * ri.x + ri.y * ri.y;
                    acc[1] += rtv.x * ri.x + rtv.y * ri.y;
                    });
                }
            if (!sh.ks.done)
                {
                pk_sync<BS, 2, false>(a, sh, acc, 0);
                if (threadIdx.x == 0) bicg_xr_finalize(&sh.ks, sh.tot[0], sh.tot[1], 0);
                }
            else
                {  // the loop is over.  Multi-GPU: the x halo below reads rows of other CTAs
                if (a.dist != nullptr)
                    pk_sync<BS, 0, false>(a, sh, acc, 0);
                else
                    __syncthreads();  // every warp has read final_half
                if (threadIdx.x == 0 && fh) bicg_xr_finalize(&sh.ks, 0.0, 0.0, 1);  // clears final_half
                }
            __syncthreads();
            }
        pk_stamp(a, sh, PKP_E);
        it++;
        }

    // ---- node update (src/solver.cpp:62-88), gated on the failure predicate -----------------------
    if (a.cur != nullptr)
        {
        const bool failed = solve_failed(&sh.ks);
        const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
        if (a.dist != nullptr)
            {  // the solution of my boundary rows goes into the neighbours' ghost tails of x
            // x of rows owned by other CTAs was written before the last grid barrier of the loop
            if (dist_push_warps(&sh.dd, sh.dd.tail, nthr / 32, [&](int row) { return x2[row]; })) sh.pushed = 1;
            pk_sync<BS, 0, false>(a, sh, acc, 1);
            pk_stamp(a, sh, PKP_HALO_X);
            }
        double v2max = 0.0;
        auto update_row = [&](int row, bool owned)
            {
            if (a.nonmag[row]) return;
            const double2 xv = x2[row];
            if (owned) v2max = fmax(v2max, xv.x * xv.x + xv.y * xv.y);
            const double4 *q = reinterpret_cast<const double4 *>(a.cur + row);
            const double4 ca = ld256_nc(q);
            const double2 *bq = reinterpret_cast<const double2 *>(a.basis + row);
            const double2 b0 = __ldg(bq), b1 = __ldg(bq + 1), b2 = __ldg(bq + 2);
            const double ep[3] = {b0.x, b0.y, b1.x}, eq[3] = {b1.y, b2.x, b2.y};
            const double u[3] = {ca.x, ca.y, ca.z};
            const double vp = xv.x * FG_GAMMA0, vq = xv.y * FG_GAMMA0;  // mesh.h:185-186
            double vn[3], un[3];
#pragma unroll
            for (int c = 0; c < 3; c++)
                {
                vn[c] = vp * ep[c] + vq * eq[c];
                un[c] = u[c] + a.dt * vn[c];
                }
            const double z = un[0] * un[0] + un[1] * un[1] + un[2] * un[2];  // Eigen normalize()
            if (z > 0.0)
                {
                const double sq = sqrt(z);
                un[0] /= sq;
                un[1] /= sq;
                un[2] /= sq;
                }
            double2 *o = reinterpret_cast<double2 *>(a.next + row);
            o[0] = make_double2(un[0], un[1]);
            o[1] = make_double2(un[2], vn[0]);
            o[2] = make_double2(vn[1], vn[2]);
            };
        if (!failed)
            {
            pk_for_slices<STAGED>(own, a.op, [&](const int s) { update_row(s * SLICE + lane, true); });
            if (a.NODt > a.NODp)
                {  // ghost rows: their solution was pushed by the owners
                if (gtid - lane < a.NODt - a.NODp)
                    {  // warp-uniform: one lane waits for the owners' flags, then plain loads see the pushes
                    if (lane == 0) pk_halo_wait_cta(sh, sh.hepoch - 1);
                    __syncwarp();
                    for (int row = a.NODp + gtid; row < a.NODt; row += gthreads) update_row(row, false);
                    }
                }
            }
        acc[0] = v2max;
        pk_sync<BS, 1, true>(a, sh, acc, 0);
        if (threadIdx.x == 0)
            {
            sh.ks.failed = failed ? 1 : 0;
            if (!failed)
                {
                sh.ks.v2max = sh.tot[0];
                sh.ks.v_max = FG_GAMMA0 * sqrt(sh.tot[0]);
                }
            sh.ks.updated = 1;
            }
        pk_stamp(a, sh, PKP_UPDATE);
        }
    if (a.dist != nullptr && threadIdx.x == 0 && sh.derr) a.dist->error = 1;  // a spin timed out: the host reports it
    if (blockIdx.x == 0 && threadIdx.x == 0)
        {
        *a.st = sh.ks;
        if (a.dist != nullptr)
            {  // the epochs the next kernel (persistent or not) continues from
            a.dist->epoch = sh.dd.epoch;
            a.dist->hepoch = sh.hepoch - 1;
            }
        if (a.h_st != nullptr)
            {  // the host is spinning on h_seq: outcome first, then the number (system-scope order)
            *a.h_st = sh.ks;
            __threadfence_system();
            st_sys(a.h_seq, a.seq);
            }
        }
    }

}  // namespace fg
