// fg_krylov.cu — sparse algebra of the LLG hot path on B200: SpMV, fused BLAS-1, BiCGStab, CG.
//
// Replaces the reference's src/algebra/{sparseMat,algebra,algebraCore,bicg,cg,iter}.h.  Same
// algorithm, same stopping rules and status codes, different execution model:
//   * the Krylov scalars and the iteration monitor live on the device (KState); every kernel reads
//     them at entry and the last CTA of each reducing kernel updates them, so a whole batch of
//     iterations is enqueued without a host round trip;
//   * every dot / norm is fused into the kernel that produces its operand and reduced with warp
//     shuffles + a deterministic two-stage grid reduction (fixed grid => bitwise reproducible);
//   * 5 kernels per BiCGStab iteration (reference: 2 parallel SpMV + 17 serial vector passes).
// All kernels are HBM-bandwidth bound: persistent grids of 148 SMs x 8 CTAs, coalesced streams.
#include <limits.h>
#include <math.h>
#include <stdio.h>

#include <type_traits>

#include "fg_common.cuh"
#include "fg_reduce.cuh"
#include "fg_krylov_state.cuh"
#include "fg_slice_iter.cuh"

namespace fg
{
__device__ __forceinline__ void load_basis_k(const double4 *qb, double ep[3], double eq[3])
    { quat_to_basis(ld256_nc(qb), ep, eq); }
// w = ep x0 + eq x1: the 3-vector image of the two tangent-plane unknowns of a node (element.h:81-96)
__device__ __forceinline__ double4 node_w(const double4 *b, double x0, double x1)
    {
    double ep[3], eq[3];
    load_basis_k(b, ep, eq);
    return make_double4(ep[0] * x0 + eq[0] * x1, ep[1] * x0 + eq[1] * x1, ep[2] * x0 + eq[2] * x1, 0.0);
    }

// ------------------------------------------------------------------------------------------
// SpMV with fused epilogues
// ------------------------------------------------------------------------------------------
struct SpmvArgs
    {
    const double4 *w;  // OP_NODE3: 3-vector image of x (gathered); x itself is read for the node-diagonal part
    const double *x;
    double *y;
    const double *a0;  // b | rt | s | p
    const double *a1;  // D (CG_SETUP)
    double *o0, *o1;   // rt (BICG_SETUP) ; p (CG_SETUP)
    const unsigned char *mask;
    KState *st;
    RedBuf red;
    DistDev *dist;     // multi-GPU: halo protocol of the SpMV input (fg_dist.cuh)
    };

// per-row epilogue shared by both layouts: applies the mask, writes y (and the fused outputs) and
// accumulates the fused dot products.
template <int STAGE>
__device__ __forceinline__ void spmv_row(const SpmvArgs &a, int i, double y, double (&acc)[RED_NV])
    {
    const bool m = a.mask != nullptr && a.mask[i];
    if (STAGE == ST_PLAIN)
        a.y[i] = m ? 0.0 : y;
    else if (STAGE == ST_RESID)
        a.y[i] = m ? 0.0 : a.a0[i] - y;
    else if (STAGE == ST_BICG_SETUP)
        {
        const double b = a.a0[i];
        const double r = m ? 0.0 : b - y;
        a.y[i] = r;
        a.o0[i] = r;
        acc[0] += b * b;
        acc[1] += r * r;
        }
    else if (STAGE == ST_CG_SETUP)
        {
        const double b = a.a0[i];
        const double r = m ? 0.0 : b - y;
        const double z = a.a1[i] * r;
        a.y[i] = r;
        a.o1[i] = z;
        acc[0] += b * b;
        acc[1] += r * r;
        acc[2] += z * r;
        }
    else if (STAGE == ST_BICG_V || STAGE == ST_CG_Q)
        {
        y = m ? 0.0 : y;
        a.y[i] = y;
        acc[0] += y * a.a0[i];
        }
    else if (STAGE == ST_BICG_T)
        {
        y = m ? 0.0 : y;
        a.y[i] = y;
        acc[0] += y * a.a0[i];
        acc[1] += y * y;
        }
    }

// same, for the two rows of a node at once (16-byte accesses)
template <int STAGE>
__device__ __forceinline__ void spmv_row2(const SpmvArgs &a, int row, double y0, double y1,
                                          double (&acc)[RED_NV])
    {
    bool m0 = false, m1 = false;
    if (a.mask != nullptr)
        {
        const uchar2 mm = reinterpret_cast<const uchar2 *>(a.mask)[row];
        m0 = mm.x != 0;
        m1 = mm.y != 0;
        }
    double2 *Y = reinterpret_cast<double2 *>(a.y) + row;
    if (STAGE == ST_PLAIN)
        *Y = make_double2(m0 ? 0.0 : y0, m1 ? 0.0 : y1);
    else if (STAGE == ST_RESID)
        {
        const double2 b = reinterpret_cast<const double2 *>(a.a0)[row];
        *Y = make_double2(m0 ? 0.0 : b.x - y0, m1 ? 0.0 : b.y - y1);
        }
    else if (STAGE == ST_BICG_SETUP || STAGE == ST_CG_SETUP)
        {
        const double2 b = reinterpret_cast<const double2 *>(a.a0)[row];
        const double2 r = make_double2(m0 ? 0.0 : b.x - y0, m1 ? 0.0 : b.y - y1);
        *Y = r;
        acc[0] += b.x * b.x + b.y * b.y;
        acc[1] += r.x * r.x + r.y * r.y;
        if (STAGE == ST_BICG_SETUP)
            reinterpret_cast<double2 *>(a.o0)[row] = r;  // rt = r; p = r is implicit (k_bicg_p, nit == 0)
        else
            {
            const double2 d = reinterpret_cast<const double2 *>(a.a1)[row];
            const double2 z = make_double2(d.x * r.x, d.y * r.y);
            reinterpret_cast<double2 *>(a.o1)[row] = z;
            acc[2] += z.x * r.x + z.y * r.y;
            }
        }
    else
        {
        y0 = m0 ? 0.0 : y0;
        y1 = m1 ? 0.0 : y1;
        *Y = make_double2(y0, y1);
        const double2 q = reinterpret_cast<const double2 *>(a.a0)[row];
        acc[0] += y0 * q.x + y1 * q.y;
        if (STAGE == ST_BICG_T) acc[1] += y0 * y0 + y1 * y1;
        }
    }

__host__ __device__ constexpr bool stage_gated(int STAGE)
    { return STAGE == ST_BICG_V || STAGE == ST_BICG_T || STAGE == ST_CG_Q; }
__host__ __device__ constexpr bool stage_reduces(int STAGE) { return STAGE != ST_PLAIN && STAGE != ST_RESID; }

// ---- SELL-32 2x2-block SpMV: one warp per slice, one lane per node row -------------------------
// Every request of the streaming part is a full-warp coalesced access (128 B of indices, 512 B of
// values); the only gather is x (16 B per block, L2-resident: the rows of a slice are neighbours
// in the mesh).  UNROLL independent block-columns are in flight per lane.
constexpr int SPMV_UNROLL = 4;
constexpr int SPMV_CTAS_PER_SM = 4;
template <int STAGE>
__global__ void __launch_bounds__(BLOCK, SPMV_CTAS_PER_SM) k_spmv_sell(const Operator op, const SpmvArgs a)
    {
    if (stage_gated(STAGE))
        if (a.st->done) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (BLOCK / 32);
    const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
    const double2 *val2 = reinterpret_cast<const double2 *>(op.val);
    double acc[RED_NV] = {0.0, 0.0, 0.0, 0.0};
    int s = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    int p0 = 0, p1 = 0;
    if (s < op.nslice)
        {
        p0 = __ldg(op.ptr + s);
        p1 = __ldg(op.ptr + s + 1);
        }
    while (s < op.nslice)
        {
        const int sn = s + nwarps;
        int q0 = 0, q1 = 0;
        if (sn < op.nslice)
            {  // next slice's extent, fetched under the shadow of this slice's stream
            q0 = __ldg(op.ptr + sn);
            q1 = __ldg(op.ptr + sn + 1);
            }
        const int *cp = op.col + (size_t)p0 * SLICE + lane;
        const double2 *vp = val2 + (size_t)p0 * (2 * SLICE) + lane;
        double y0 = 0.0, y1 = 0.0;
#pragma unroll SPMV_UNROLL
        for (int j = p0; j < p1; ++j)
            {
            const int c = __ldcs(cp);
            const double2 k0 = __ldcs(vp);
            const double2 k1 = __ldcs(vp + SLICE);
            const double2 xv = x2[c];
            y0 += k0.x * xv.x;
            y0 += k0.y * xv.y;
            y1 += k1.x * xv.x;
            y1 += k1.y * xv.y;
            cp += SLICE;
            vp += 2 * SLICE;
            }
        spmv_row2<STAGE>(a, s * SLICE + lane, y0, y1, acc);
        s = sn;
        p0 = q0;
        p1 = q1;
        }
    if (!stage_reduces(STAGE)) return;
    double tot[RED_NV];
    const int role = grid_reduce<RED_NV>(acc, a.red, tot);
    if (role == 0) return;
    if (role == 1) spmv_finalize<STAGE>(a.st, tot);
    }

// ---- matrix-free LLG operator (OP_NODE3): K = cS P^T (S x I3) P + Dg, one warp per slice ---------
// Streams 12 B per stored node pair (S, column) instead of 36 B per 2x2 block -- 10 B when the column
// offsets of the mesh fit 16 bits (IDX16); the gather is ONE 256-bit request for the aligned 32-byte
// image w_b (LDG.E.ENL2.256); the row epilogue projects z onto the node's (eq, ep) and adds the 2x2
// node-diagonal part.  Inside the BiCGStab loop the preconditioned input vector itself is never
// stored: the node's own unknowns are recovered from its image, x_a = (ep_a . w_a, eq_a . w_a)
// (a.x == NULL); the taps and the setup stage pass x.  Same SELL-32 traversal, same fused epilogues
// as k_spmv_sell.
// The slice loop is a device function shared by the stand-alone kernel below and by the persistent solve
// kernel (fg_solve_pk.cuh).  COH = false: the gathered images are read through the non-coherent path
// (ld.global.nc: legal only when nothing writes them during the kernel, i.e. one GPU, one kernel per
// phase); COH = true: plain coherent loads issued as volatile asm, for images that peers (multi-GPU ghost
// tails) or other CTAs of the same persistent kernel write while it runs.  `pass` selects the slices of a
// partitioned operator: 0 all, 1 only slices without ghost columns, 2 only slices with ghost columns
// (op.sghost), so that the interior rows never wait for the halo.
template <bool COH> __device__ __forceinline__ double4 ld_image(const double4 *p)
    {
    if (COH)
        {
        double4 r;
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
        return r;
        }
    return ld256_nc(p);
    }

// Row ownership (SliceIter, slices_strided, slices_balanced_at): fg_slice_iter.cuh
__device__ __forceinline__ SliceIter slices_balanced(int n, int per_cta, int id)
    { return slices_balanced_at(n, per_cta, id, (int)blockIdx.x, (int)gridDim.x); }

// Returns true when the warp met a slice that belongs to the other pass (pass 1: a slice with ghost columns).
// The ghost flag of a slice travels with its extent, fetched one slice ahead: deciding on a flag loaded in
// the same turn would stall the (in-order) warp for an L2 round trip per slice -- measured as 12 % of the
// product on a 2-GPU partition.
// The head of a warp's first slice: extent, ghost flag, and the column indices and values of the first batch
// of stored pairs.  The matrix is constant during a solve and so is the ownership of slices, so the
// persistent kernel fetches it once (HEAD variant, where registers allow: k_llg_solve) and every product
// starts with its gathers instead of behind two dependent loads (extent -> indices).  That is most of what
// a product costs when a warp owns a single slice.
constexpr int SPMV_U = 8;  // stored pairs per batch
struct SpmvHead
    {
    int s;  // the slice, or -1
    int p0, p1;
    unsigned char g;
    int cn[SPMV_U];
    double sv[SPMV_U];
    };
template <bool IDX16>
__device__ __forceinline__ void spmv_head(const Operator &op, const SliceIter it, const int lane, SpmvHead &h)
    {
    typedef typename std::conditional<IDX16, short, int>::type idx_t;
    const idx_t *colbase = IDX16 ? reinterpret_cast<const idx_t *>(op.col16) : reinterpret_cast<const idx_t *>(op.col);
    h.s = it.begin();
    if (h.s >= op.nslice)
        {
        h.s = -1;
        return;
        }
    h.p0 = __ldg(op.ptr + h.s);
    h.p1 = __ldg(op.ptr + h.s + 1);
    h.g = op.sghost != nullptr ? __ldg(op.sghost + h.s) : 0;
    const idx_t *cp = colbase + (size_t)h.p0 * SLICE + lane;
    const double *sp = op.val + (size_t)h.p0 * SLICE + lane;
#pragma unroll
    for (int u = 0; u < SPMV_U; u++)
        {
        const bool in = h.p0 + u < h.p1;
        h.cn[u] = in ? (int)__ldcs(cp + u * SLICE) : 0;
        h.sv[u] = in ? __ldcs(sp + u * SLICE) : 0.0;
        }
    }

template <int STAGE, bool IDX16, bool COH, bool HEAD>
__device__ __forceinline__ bool spmv_node3_slices(const Operator &op, const SpmvArgs &a, const SliceIter it,
                                                  const int lane, const int pass, double (&acc)[RED_NV],
                                                  const SpmvHead &head)
    {
    const int s_end = op.nslice;
    typedef typename std::conditional<IDX16, short, int>::type idx_t;
    const idx_t *colbase = IDX16 ? reinterpret_cast<const idx_t *>(op.col16) : reinterpret_cast<const idx_t *>(op.col);
    const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
    bool other = false;
    int s = it.begin();
    int p0 = 0, p1 = 0;
    unsigned char g = 0;
    constexpr int U = SPMV_U;
    int cn[U];
    bool have_cn = false;  // cn holds the first batch of slice s (from the head fetched before the barrier)
    if (HEAD && head.s == s)
        {
        p0 = head.p0;
        p1 = head.p1;
        g = head.g;
#pragma unroll
        for (int u = 0; u < U; u++) cn[u] = head.cn[u];
        have_cn = true;
        }
    else if (s < s_end)
        {
        p0 = __ldg(op.ptr + s);
        p1 = __ldg(op.ptr + s + 1);
        if (pass != 0) g = __ldg(op.sghost + s);
        }
    while (s < s_end)
        {
        const int sn = it.next(s);
        int q0 = 0, q1 = 0;
        unsigned char gn = 0;
        if (sn < s_end)
            {
            q0 = __ldg(op.ptr + sn);
            q1 = __ldg(op.ptr + sn + 1);
            if (pass != 0) gn = __ldg(op.sghost + sn);
            }
        if (pass != 0 && (g != 0) != (pass == 2))
            {  // this slice is for the other pass
            other = true;
            s = sn;
            p0 = q0;
            p1 = q1;
            g = gn;
            have_cn = false;
            continue;
            }
        const int row = s * SLICE + lane;
        if (op.prefetch)
            {  // a slice is ~4 dependent memory phases (indices -> gathers, twice, -> row operands): pull
               // the row-epilogue operands into L2 now, no registers held (measured 235/217 -> 231/212 us;
               // also prefetching the first batch of the next slice was slower, 242/231 us:
               // profiles/experiments/r01x_spmv_prefetch.md)
            prefetch_l2(op.qbasis + row);
            prefetch_l2(op.Dm + row);
            if (a.a0 != nullptr) prefetch_l2(reinterpret_cast<const double2 *>(a.a0) + row);
            }
        const int cbase = IDX16 ? row : 0;  // 16-bit columns are offsets from the lane's own row
        const idx_t *cp = colbase + (size_t)p0 * SLICE + lane;
        const double *sp = op.val + (size_t)p0 * SLICE + lane;
        double z0 = 0.0, z1 = 0.0, z2 = 0.0;
            {  // batches of U pairs; the column indices of the next batch are fetched while this batch
               // gathers, so the index -> gather dependency is paid once per slice, not once per batch
               // (measured 291 us against 317 us for the plain unrolled loop, film20m)
            const bool from_head = HEAD && have_cn;
            if (!have_cn)
                {
#pragma unroll
                for (int u = 0; u < U; u++) cn[u] = p0 + u < p1 ? (int)__ldcs(cp + u * SLICE) : 0;
                }
            have_cn = false;
            for (int j = p0; j < p1; j += U)
                {
                int c[U];
                double Sv[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    c[u] = cbase + cn[u];
                    if (HEAD && from_head && j == p0)
                        Sv[u] = head.sv[u];
                    else
                        Sv[u] = j + u < p1 ? __ldcs(sp + u * SLICE) : 0.0;
                    }
                cp += U * SLICE;
                sp += U * SLICE;
#pragma unroll
                for (int u = 0; u < U; u++) cn[u] = j + U + u < p1 ? (int)__ldcs(cp + u * SLICE) : 0;
#pragma unroll
                for (int u = 0; u < U; u++)
                    {
                    const double4 wb = ld_image<COH>(a.w + c[u]);
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
            xa = x2[row];
        else
            {  // x_a = P_a^T w_a (ep, eq orthonormal): the line is in L1/L2, the diagonal pair gathered it
            const double4 wr = ld_image<COH>(a.w + row);
            xa = make_double2(ep[0] * wr.x + ep[1] * wr.y + ep[2] * wr.z, eq[0] * wr.x + eq[1] * wr.y + eq[2] * wr.z);
            }
        // node-diagonal part in closed form: Dg = [[a_w, Ma], [Ma, -a_w]] (fg_common.cuh, OP_NODE3)
        const double2 dm = __ldcs(op.Dm + row);
        double y0 = op.cS * (eq[0] * z0 + eq[1] * z1 + eq[2] * z2) + (dm.y * xa.x + dm.x * xa.y);
        double y1 = op.cS * (ep[0] * z0 + ep[1] * z1 + ep[2] * z2) + (dm.x * xa.x - dm.y * xa.y);
        if (op.nonmag[row] != 0)
            {  // identity row (src/solver.cpp:46-48)
            y0 = xa.x;
            y1 = xa.y;
            }
        spmv_row2<STAGE>(a, row, y0, y1, acc);
        s = sn;
        p0 = q0;
        p1 = q1;
        g = gn;
        }
    return other;
    }

template <int STAGE, bool IDX16, bool COH>
__device__ __forceinline__ bool spmv_node3_slices(const Operator &op, const SpmvArgs &a, const SliceIter it,
                                                  const int lane, const int pass, double (&acc)[RED_NV])
    {
    SpmvHead none;
    none.s = -1;
    return spmv_node3_slices<STAGE, IDX16, COH, false>(op, a, it, lane, pass, acc, none);
    }

template <int STAGE, bool IDX16, bool COH>
__global__ void __launch_bounds__(BLOCK, SPMV_CTAS_PER_SM) k_spmv_node3(const Operator op, const SpmvArgs a)
    {
    if (stage_gated(STAGE))
        if (a.st->done) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (BLOCK / 32);
    const int s0 = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    double acc[RED_NV] = {0.0, 0.0, 0.0, 0.0};
    if (COH && STAGE == ST_BICG_V && a.dist != nullptr && op.sghost != nullptr)
        {  // multi-GPU: the ghost entries of w(D.p) are pushed by the neighbours' k_bicg_p while this kernel
           // runs.  Rows without ghost columns do not wait; the others are done after the halo flag of
           // every source rank was seen, and they read the images with coherent loads (COH).
        spmv_node3_slices<STAGE, IDX16, COH>(op, a, slices_strided(s0, nwarps, op.nslice), lane, 1, acc);
        if (threadIdx.x == 0) dist_wait(a.dist);
        __syncthreads();
        spmv_node3_slices<STAGE, IDX16, COH>(op, a, slices_strided(s0, nwarps, op.nslice), lane, 2, acc);
        }
    else
        {
        if (STAGE == ST_BICG_V && a.dist != nullptr)
            {
            if (threadIdx.x == 0) dist_wait(a.dist);
            __syncthreads();
            }
        spmv_node3_slices<STAGE, IDX16, COH>(op, a, slices_strided(s0, nwarps, op.nslice), lane, 0, acc);
        }
    if (!stage_reduces(STAGE)) return;
    double tot[RED_NV];
    const int role = grid_reduce<RED_NV>(acc, a.red, tot);
    if (role == 1) spmv_finalize<STAGE>(a.st, tot);
    }

}  // namespace fg
#include "fg_solve_pk.cuh"
namespace fg
{
// ---- plain CSR (algebra::SparseMatrix of the side solvers): G lanes per row ---------------------
template <int STAGE>
__global__ void __launch_bounds__(BLOCK) k_spmv_csr(const Operator op, const SpmvArgs a)
    {
    if (stage_gated(STAGE))
        if (a.st->done) return;
    const int G = op.lanes;
    const int lane_in_group = threadIdx.x & (G - 1);
    const int groups_per_cta = BLOCK / G;
    const long long total_groups = (long long)gridDim.x * groups_per_cta;
    double acc[RED_NV] = {0.0, 0.0, 0.0, 0.0};
    for (long long u0 = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G;; u0 += total_groups)
        {
        const bool active = u0 < op.n;
        if (!__any_sync(0xffffffffu, active)) break;
        const int u = (int)u0;
        double y0 = 0.0;
        if (active)
            {
            const int beg = op.ptr[u], end = op.ptr[u + 1];
            for (int j = beg + lane_in_group; j < end; j += G)
                y0 += __ldcs(op.val + j) * a.x[__ldg(op.col + j)];
            }
        for (int o = G >> 1; o > 0; o >>= 1) y0 += __shfl_xor_sync(0xffffffffu, y0, o);
        if (active && lane_in_group == 0) spmv_row<STAGE>(a, u, y0, acc);
        }
    if (!stage_reduces(STAGE)) return;
    double tot[RED_NV];
    if (grid_reduce<RED_NV>(acc, a.red, tot) != 1) return;
    spmv_finalize<STAGE>(a.st, tot);
    }

int grid_for(long long work_items, int items_per_cta)
    {
    long long g = (work_items + items_per_cta - 1) / items_per_cta;
    if (g < 1) g = 1;
    if (g > MAX_GRID) g = MAX_GRID;
    return (int)g;
    }

// persistent grids are sized from the occupancy the compiled kernel really gets (a grid larger
// than one resident wave would serialise a second, mostly idle wave)
template <class K> static int resident_grid(K kernel)
    {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    int dev = 0, sms = NUM_SMS;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int g = per_sm * sms;
    return g > MAX_GRID ? MAX_GRID : g;
    }

template <int STAGE>
static int launch_spmv(const Operator &op, const KrylovWork &w, const SpmvArgs &a)
    {
    const int cls = STAGE == ST_BICG_V ? KC_SPMV_V : (STAGE == ST_BICG_T ? KC_SPMV_T : (STAGE == ST_BICG_SETUP ? KC_SPMV_SETUP : KC_OTHER));
    const bool prof = prof_begin(w.prof, w.stream, cls);
    if (op.kind == OP_NODE3)
        {
        static int wave = 0, wave16 = 0;
        if (!wave)
            {
            wave = resident_grid(k_spmv_node3<STAGE, false, false>);
            wave16 = resident_grid(k_spmv_node3<STAGE, true, false>);
            }
        const int need = (op.nslice + BLOCK / 32 - 1) / (BLOCK / 32);
        const int wv = op.col16 ? wave16 : wave;
        const int grid = need < wv ? (need > 0 ? need : 1) : wv;
        // a distributed context reads the images with coherent loads (peers write the ghost tails while
        // the consuming kernel runs); one GPU keeps the non-coherent path
        if (a.dist != nullptr)
            {
            if (op.col16)
                k_spmv_node3<STAGE, true, true><<<grid, BLOCK, 0, w.stream>>>(op, a);
            else
                k_spmv_node3<STAGE, false, true><<<grid, BLOCK, 0, w.stream>>>(op, a);
            }
        else if (op.col16)
            k_spmv_node3<STAGE, true, false><<<grid, BLOCK, 0, w.stream>>>(op, a);
        else
            k_spmv_node3<STAGE, false, false><<<grid, BLOCK, 0, w.stream>>>(op, a);
        }
    else if (op.kind == OP_SELL2)
        {
        static int wave = 0;
        if (!wave) wave = resident_grid(k_spmv_sell<STAGE>);
        const int need = (op.nslice + BLOCK / 32 - 1) / (BLOCK / 32);
        const int grid = need < wave ? (need > 0 ? need : 1) : wave;
        k_spmv_sell<STAGE><<<grid, BLOCK, 0, w.stream>>>(op, a);
        }
    else
        {
        static int wave = 0;
        if (!wave) wave = resident_grid(k_spmv_csr<STAGE>);
        int grid = grid_for(op.n, BLOCK / op.lanes);
        if (grid > wave) grid = wave;
        k_spmv_csr<STAGE><<<grid, BLOCK, 0, w.stream>>>(op, a);
        }
    if (prof) prof_end(w.prof, w.stream);
    if (w.launches) ++*w.launches;
    FG_CUDA(cudaGetLastError());
    return FG_OK;
    }

// plain y = A x; for OP_NODE3 the 3-vector image of x is built first (taps and microbenchmarks only:
// the solver's own producers write it in the kernel that produces x)
__global__ void __launch_bounds__(BLOCK)
k_make_w(int nnode, const double *__restrict__ x, const double4 *__restrict__ basis, double4 *__restrict__ w3)
    {
    const int stride = gridDim.x * BLOCK;
    for (int a = blockIdx.x * BLOCK + threadIdx.x; a < nnode; a += stride)
        {
        const double2 xv = reinterpret_cast<const double2 *>(x)[a];
        w3[a] = node_w(basis + a, xv.x, xv.y);
        }
    }

int spmv(const Operator &op, const KrylovWork &w, const double *x, double *y, bool masked, bool make_w)
    {
    SpmvArgs a = {};
    if (op.kind == OP_NODE3 && !make_w)
        a.w = w.w3s;
    else if (op.kind == OP_NODE3)
        {
        k_make_w<<<grid_for(w.nx / 2, BLOCK), BLOCK, 0, w.stream>>>(w.nx / 2, x, op.qbasis, w.w3s);
        if (w.launches) ++*w.launches;
        FG_CUDA(cudaGetLastError());
        a.w = w.w3s;
        }
    a.x = x;
    a.y = y;
    a.mask = masked ? w.mask : nullptr;
    a.st = w.st;
    a.red = w.red;
    return launch_spmv<ST_PLAIN>(op, w, a);
    }

// ------------------------------------------------------------------------------------------
// fused vector kernels of BiCGStab (reference src/algebra/bicg.h:185-232)
// ------------------------------------------------------------------------------------------
// p = r + beta (p - omega v) ; phat = D p                       (bicg.h:196-202)
// p is ping-ponged (read p, write pn): see the node-wise variant below.
__global__ void __launch_bounds__(BLOCK)
k_bicg_p(int n, const double *__restrict__ r, const double *__restrict__ p, double *__restrict__ pn,
         const double *__restrict__ v, const double *__restrict__ D, double *__restrict__ phat,
         const KState *st)
    {
    if (st->done) return;
    const bool first = st->nit == 0;
    const double omega = st->omega;
    const double beta = first ? 0.0 : bicg_beta(st);
    const int stride = gridDim.x * BLOCK;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride)
        {
        double pi;
        if (first)
            pi = r[i];  // p = r (bicg.h:183)
        else
            pi = bicg_p_value(p[i], v[i], r[i], omega, beta);
        pn[i] = pi;
        phat[i] = D[i] * pi;
        }
    }

// Node-wise variant for the matrix-free LLG operator: it writes w = P phat, the 3-vector image the
// next SpMV gathers; phat itself is stored only when asked for (phat != NULL): the solver recomputes
// D p where it needs it (k_bicg_xr_node), which saves one vector write here and one read in the SpMV.  Multi-GPU: beta is known when this kernel starts, so its first
// CTAs push w of the boundary rows into the neighbours' ghost tails before doing their share of the
// update; the last of those CTAs raises the halo flag the consuming SpMV waits on (fg_dist.cuh).
// p is ping-ponged (read p, write pn) so that the pushers can read the old direction of rows
// another CTA is updating.
__global__ void __launch_bounds__(BLOCK)
k_bicg_p_node(int nnode, const double *__restrict__ r, const double *__restrict__ p, double *__restrict__ pn,
              const double *__restrict__ v, const double *__restrict__ D, double *__restrict__ phat,
              const double4 *__restrict__ basis, double4 *__restrict__ w3, const KState *st, DistDev *dist,
              unsigned int *ticket, int npush)
    {
    if (st->done) return;
    const bool first = st->nit == 0;
    const double omega = st->omega;
    const double beta = first ? 0.0 : bicg_beta(st);
    const int stride = gridDim.x * BLOCK;
    const double2 *p2 = reinterpret_cast<const double2 *>(p), *v2 = reinterpret_cast<const double2 *>(v),
                  *r2 = reinterpret_cast<const double2 *>(r), *D2 = reinterpret_cast<const double2 *>(D);
    auto value = [&](int row, double2 &pi)
        {
        const double2 rr = r2[row], d = D2[row];
        if (first)
            pi = rr;  // p = r (bicg.h:183)
        else
            {
            const double2 pp = p2[row], vv = v2[row];
            pi = make_double2(bicg_p_value(pp.x, vv.x, rr.x, omega, beta), bicg_p_value(pp.y, vv.y, rr.y, omega, beta));
            }
        return make_double2(d.x * pi.x, d.y * pi.y);
        };
    if (dist != nullptr && (int)blockIdx.x < npush)
        {
        dist_push(dist, dist->wtail[0], blockIdx.x * BLOCK + threadIdx.x, npush * BLOCK, [&](int row)
            {
            double2 pi;
            const double2 ph = value(row, pi);
            return node_w(basis + row, ph.x, ph.y);
            });
        __syncthreads();
        if (threadIdx.x == 0)
            {
            const unsigned int t = atomicInc(ticket, (unsigned int)npush - 1);
            if (t == (unsigned int)npush - 1) dist_raise(dist);
            }
        }
    for (int a = blockIdx.x * BLOCK + threadIdx.x; a < nnode; a += stride)
        {
        double2 pi;
        const double2 ph = value(a, pi);
        reinterpret_cast<double2 *>(pn)[a] = pi;
        if (phat != nullptr) reinterpret_cast<double2 *>(phat)[a] = ph;
        st256(w3 + a, node_w(basis + a, ph.x, ph.y));
        }
    }

// s = r - alpha v ; shat = D s ; ||s||^2 -> mid-iteration exit test    (bicg.h:207-218)
__global__ void __launch_bounds__(BLOCK)
k_bicg_s(int n, const double *__restrict__ r, const double *__restrict__ v,
         const double *__restrict__ D, double *__restrict__ s, double *__restrict__ shat, KState *st,
         const RedBuf red)
    {
    if (st->done) return;
    const double alpha = st->alpha;
    double acc[1] = {0.0};
    const int stride = gridDim.x * BLOCK;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride)
        {
        const double si = bicg_s_value(r[i], v[i], alpha);
        s[i] = si;
        shat[i] = D[i] * si;
        acc[0] += si * si;
        }
    double tot[1];
    if (grid_reduce<1>(acc, red, tot) != 1) return;
    bicg_s_finalize(st, tot[0]);
    }

// Node-wise variant (matrix-free LLG operator): also writes w = P shat.  Multi-GPU: w of the
// boundary rows goes to the neighbours first; the all-reduce of |s|^2 below is the barrier that
// publishes it (fg_dist.cuh).
__global__ void __launch_bounds__(BLOCK)
k_bicg_s_node(int nnode, const double *__restrict__ r, const double *__restrict__ v,
              const double *__restrict__ D, double *__restrict__ s, double *__restrict__ shat,
              const double4 *__restrict__ basis, double4 *__restrict__ w3, KState *st, const RedBuf red)
    {
    if (st->done) return;
    const double alpha = st->alpha;
    double acc[1] = {0.0};
    const int stride = gridDim.x * BLOCK;
    const double2 *r2 = reinterpret_cast<const double2 *>(r), *v2 = reinterpret_cast<const double2 *>(v),
                  *D2 = reinterpret_cast<const double2 *>(D);
    auto value = [&](int row, double2 &si)
        {
        const double2 rr = r2[row], vv = v2[row], d = D2[row];
        si = make_double2(bicg_s_value(rr.x, vv.x, alpha), bicg_s_value(rr.y, vv.y, alpha));
        return make_double2(d.x * si.x, d.y * si.y);
        };
    if (red.dist != nullptr)
        dist_push(red.dist, red.dist->wtail[1], blockIdx.x * BLOCK + threadIdx.x, stride, [&](int row)
            {
            double2 si;
            const double2 sh = value(row, si);
            return node_w(basis + row, sh.x, sh.y);
            });
    for (int a = blockIdx.x * BLOCK + threadIdx.x; a < nnode; a += stride)
        {
        double2 si;
        const double2 sh = value(a, si);
        reinterpret_cast<double2 *>(s)[a] = si;
        if (shat != nullptr) reinterpret_cast<double2 *>(shat)[a] = sh;
        st256(w3 + a, node_w(basis + a, sh.x, sh.y));
        acc[0] += si.x * si.x + si.y * si.y;
        }
    double tot[1];
    if (grid_reduce<1>(acc, red, tot) != 1) return;
    bicg_s_finalize(st, tot[0]);
    }

// x += alpha phat + omega shat ; r = s - omega t ; ||r||^2, (rt,r) -> next loop test
// (bicg.h:223-231 then :185-195).  When the loop ended on ||s|| only x += alpha phat is applied.
__global__ void __launch_bounds__(BLOCK)
k_bicg_xr(int n, double *__restrict__ x, const double *__restrict__ phat,
          const double *__restrict__ shat, const double *__restrict__ s,
          const double *__restrict__ t, const double *__restrict__ rt, double *__restrict__ r,
          KState *st, const RedBuf red)
    {
    const int fh = st->final_half;
    if (st->done && !fh) return;
    const double alpha = st->alpha, omega = st->omega;
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * BLOCK;
    if (fh)
        {
        for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride) x[i] += alpha * phat[i];
        }
    else
        {
        for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride)
            {
            x[i] = (x[i] + alpha * phat[i]) + omega * shat[i];
            const double ri = s[i] - omega * t[i];
            r[i] = ri;
            acc[0] += ri * ri;
            acc[1] += rt[i] * ri;
            }
        }
    double tot[2];
    const int role = grid_reduce<2>(acc, red, tot);
    if (role == 0) return;
    if (role == 1) bicg_xr_finalize(st, tot[0], tot[1], fh);
    }

// Node-wise variant for the matrix-free LLG operator: phat = D p and shat = D s are not stored by
// their producers; they are recomputed here with the same single rounding (D p, then the same
// update expression), so x is bit-identical to the stored-vector formulation.
__global__ void __launch_bounds__(BLOCK)
k_bicg_xr_node(int nnode, double *__restrict__ x, const double *__restrict__ p, const double *__restrict__ D,
               const double *__restrict__ s, const double *__restrict__ t, const double *__restrict__ rt,
               double *__restrict__ r, KState *st, const RedBuf red)
    {
    const int fh = st->final_half;
    if (st->done && !fh) return;
    const double alpha = st->alpha, omega = st->omega;
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * BLOCK;
    double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
    const double2 *p2 = reinterpret_cast<const double2 *>(p), *D2 = reinterpret_cast<const double2 *>(D),
                  *s2 = reinterpret_cast<const double2 *>(s), *t2 = reinterpret_cast<const double2 *>(t),
                  *rt2 = reinterpret_cast<const double2 *>(rt);
    if (fh)
        {
        for (int a = blockIdx.x * BLOCK + threadIdx.x; a < nnode; a += stride)
            {
            const double2 d = D2[a], pp = p2[a];
            double2 xv = x2[a];
            xv.x += alpha * __dmul_rn(d.x, pp.x);
            xv.y += alpha * __dmul_rn(d.y, pp.y);
            x2[a] = xv;
            }
        }
    else
        {
        for (int a = blockIdx.x * BLOCK + threadIdx.x; a < nnode; a += stride)
            {
            const double2 d = D2[a], pp = p2[a], sv = s2[a], tv = t2[a], rtv = rt2[a];
            double2 xv = x2[a];
            xv.x = (xv.x + alpha * __dmul_rn(d.x, pp.x)) + omega * __dmul_rn(d.x, sv.x);
            xv.y = (xv.y + alpha * __dmul_rn(d.y, pp.y)) + omega * __dmul_rn(d.y, sv.y);
            x2[a] = xv;
            const double2 ri = make_double2(sv.x - omega * tv.x, sv.y - omega * tv.y);
            r2[a] = ri;
            acc[0] += ri.x * ri.x + ri.y * ri.y;
            acc[1] += rtv.x * ri.x + rtv.y * ri.y;
            }
        }
    double tot[2];
    const int role = grid_reduce<2>(acc, red, tot);
    if (role == 0) return;
    if (role == 1) bicg_xr_finalize(st, tot[0], tot[1], fh);
    }

// ------------------------------------------------------------------------------------------
// fused vector kernels of CG (reference src/algebra/cg.h:36-56)
// ------------------------------------------------------------------------------------------
// p = (rho/rho_1) p + D r   (nit > 0)
__global__ void __launch_bounds__(BLOCK)
k_cg_p(int n, const double *__restrict__ r, const double *__restrict__ D, double *__restrict__ p,
       const KState *st)
    {
    if (st->done || st->nit == 0) return;
    const double f = st->rho1 / st->rho2;
    const int stride = gridDim.x * BLOCK;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride)
        p[i] = p[i] * f + D[i] * r[i];
    }

// x += a p ; r -= a q ; ||r||^2, (Dr, r) -> loop test
__global__ void __launch_bounds__(BLOCK)
k_cg_xr(int n, double *__restrict__ x, const double *__restrict__ p, const double *__restrict__ q,
        const double *__restrict__ D, double *__restrict__ r, KState *st, const RedBuf red)
    {
    if (st->done) return;
    const double a = st->alpha;
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * BLOCK;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride)
        {
        x[i] += a * p[i];
        const double ri = r[i] - a * q[i];
        r[i] = ri;
        acc[0] += ri * ri;
        acc[1] += (D[i] * ri) * ri;
        }
    double tot[2];
    if (grid_reduce<2>(acc, red, tot) != 1) return;
    cg_xr_finalize(st, tot[0], tot[1]);
    }

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
k_init_state(KState *st, double tol, int maxiter)
    {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    kstate_reset(st, tol, maxiter);
    }

__global__ void __launch_bounds__(BLOCK)
k_mask(int n, const unsigned char *__restrict__ mask, double *__restrict__ x)
    {
    const int stride = gridDim.x * BLOCK;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride)
        if (mask[i]) x[i] = 0.0;
    }

__global__ void __launch_bounds__(BLOCK)
k_axpy(int n, double a, const double *__restrict__ x, double *__restrict__ y)
    {
    const int stride = gridDim.x * BLOCK;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride) y[i] += a * x[i];
    }

// D[i] = 1 / A(i,i) (0 when masked), src/algebra/sparseMat.h:174-183
__global__ void __launch_bounds__(BLOCK)
k_diag_precond(const Operator op, const unsigned char *__restrict__ mask, double *__restrict__ D)
    {
    const int stride = gridDim.x * BLOCK;
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < op.n; i += stride)
        {
        double c = 0.0;
        int lo = op.ptr[i], hi = op.ptr[i + 1];
        const int end = hi;
        while (lo < hi)
            {
            const int mid = lo + (hi - lo) / 2;
            if (op.col[mid] < i)
                lo = mid + 1;
            else
                hi = mid;
            }
        if (lo < end && op.col[lo] == i) c = op.val[lo];
        D[i] = (mask != nullptr && mask[i]) ? 0.0 : 1.0 / c;
        }
    }

#define FG_LAUNCH_C(w, cls, kernel, grid, ...)                     \
    do                                                             \
        {                                                          \
        const bool prof_ = prof_begin((w).prof, (w).stream, cls);  \
        kernel<<<(grid), BLOCK, 0, (w).stream>>>(__VA_ARGS__);     \
        if (prof_) prof_end((w).prof, (w).stream);                 \
        if ((w).launches) ++*(w).launches;                         \
        FG_CUDA(cudaGetLastError());                               \
        } while (0)
#define FG_LAUNCH(w, kernel, grid, ...) FG_LAUNCH_C(w, KC_OTHER, kernel, grid, __VA_ARGS__)

int vec_mask(const KrylovWork &w, double *x)
    {
    if (!w.mask) return FG_OK;
    FG_LAUNCH(w, k_mask, grid_for(w.n, BLOCK * 4), w.n, w.mask, x);
    return FG_OK;
    }

int vec_axpy(const KrylovWork &w, double a, const double *x, double *y)
    {
    FG_LAUNCH(w, k_axpy, grid_for(w.n, BLOCK * 4), w.n, a, x, y);
    return FG_OK;
    }

int build_diag_precond_csr(const Operator &A, const KrylovWork &w)
    {
    FG_LAUNCH(w, k_diag_precond, grid_for(A.n, BLOCK), A, w.mask, w.D);
    return FG_OK;
    }

int krylov_alloc(KrylovWork &w, int n, int n_ghost, cudaStream_t stream, long long *launch_counter,
                 bool node3, double *const ext[3])
    {
    w = KrylovWork();
    w.n = n;
    w.nx = n + n_ghost;
    w.stream = stream;
    w.launches = launch_counter;
    const size_t nb = sizeof(double) * (size_t)(w.nx > 0 ? w.nx : 1);
    double **vecs[] = {&w.x, &w.b, &w.r, &w.rt, &w.p, &w.p2, &w.v, &w.s, &w.t, &w.phat, &w.shat, &w.D};
    if (ext)
        {  // multi-GPU: the exchanged vectors live in the IPC arena
        w.x = ext[0];
        w.w3p = reinterpret_cast<double4 *>(ext[1]);
        w.w3s = reinterpret_cast<double4 *>(ext[2]);
        w.arena = ext[0];
        }
    else if (node3)
        {
        const size_t wb = sizeof(double4) * (size_t)(w.nx / 2 > 0 ? w.nx / 2 : 1);
        FG_CUDA(cudaMalloc(&w.w3p, wb));
        FG_CUDA(cudaMalloc(&w.w3s, wb));
        FG_CUDA(cudaMemsetAsync(w.w3p, 0, wb, stream));
        FG_CUDA(cudaMemsetAsync(w.w3s, 0, wb, stream));
        }
    for (double **v : vecs)
        {
        if (!*v) FG_CUDA(cudaMalloc(v, nb));
        FG_CUDA(cudaMemsetAsync(*v, 0, nb, stream));
        }
    FG_CUDA(cudaMalloc(&w.st, sizeof(KState)));
    FG_CUDA(cudaMemsetAsync(w.st, 0, sizeof(KState), stream));
    FG_CUDA(cudaHostAlloc(&w.h_st, sizeof(KState), cudaHostAllocMapped));
    memset(w.h_st, 0, sizeof(KState));
    FG_CUDA(cudaHostAlloc(&w.h_seq, sizeof(unsigned long long), cudaHostAllocMapped));
    *w.h_seq = 0ull;
    w.seq = 0ull;
    FG_CUDA(cudaHostGetDevicePointer(&w.d_h_st, w.h_st, 0));
    FG_CUDA(cudaHostGetDevicePointer(&w.d_h_seq, w.h_seq, 0));
    FG_CUDA(cudaMalloc(&w.red.partials, sizeof(double) * 2 * RED_NV * MAX_GRID));  // values + compensations
    FG_CUDA(cudaMalloc(&w.red.ticket, sizeof(unsigned int)));
    FG_CUDA(cudaMemsetAsync(w.red.ticket, 0, sizeof(unsigned int), stream));
    w.red.dist = nullptr;
    FG_CUDA(cudaEventCreateWithFlags(&w.ev_poll, cudaEventDisableTiming));
    w.last_iters = 0;
    if (node3)
        {
        FG_CUDA(cudaMalloc(&w.pk, sizeof(PkSync)));
        FG_CUDA(cudaMemsetAsync(w.pk, 0, sizeof(PkSync), stream));
        FG_CUDA(cudaMalloc(&w.pk_phase_acc, sizeof(unsigned long long) * 64));
        FG_CUDA(cudaMemsetAsync(w.pk_phase_acc, 0, sizeof(unsigned long long) * 64, stream));
        }
    return FG_OK;
    }

void krylov_free(KrylovWork &w)
    {
    if (w.arena)
        {  // owned by the exchange arena
        w.x = nullptr;
        w.w3p = w.w3s = nullptr;
        }
    if (w.w3p) cudaFree(w.w3p);
    if (w.w3s) cudaFree(w.w3s);
    double *vecs[] = {w.x, w.b, w.r, w.rt, w.p, w.p2, w.v, w.s, w.t, w.phat, w.shat, w.D};
    for (double *v : vecs)
        if (v) cudaFree(v);
    if (w.st) cudaFree(w.st);
    if (w.h_st) cudaFreeHost(w.h_st);
    if (w.h_seq) cudaFreeHost(w.h_seq);
    if (w.red.partials) cudaFree(w.red.partials);
    if (w.red.ticket) cudaFree(w.red.ticket);
    if (w.ev_poll) cudaEventDestroy(w.ev_poll);
    if (w.pk) cudaFree(w.pk);
    if (w.pk_phase_acc) cudaFree(w.pk_phase_acc);
    w = KrylovWork();
    }

// ------------------------------------------------------------------------------------------
// multi-GPU halo exchange (fg_dist.cuh): push my boundary entries into the neighbours' ghost tails
// over NVLink peer memory, raise the halo epoch flag, wait for my own sources.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
k_halo(DistDev *d, const double *__restrict__ vec, int gate, const KState *st, unsigned int *ticket)
    {
    if (gate == 1 && st->done) return;
    if (gate == 2 && (!st->done || st->updated)) return;
    __shared__ int is_last;
    const double2 *v2 = reinterpret_cast<const double2 *>(vec);
    const int nsend = d->send_ptr[d->world];
    const int stride = gridDim.x * BLOCK;
    for (int idx = blockIdx.x * BLOCK + threadIdx.x; idx < nsend; idx += stride)
        {
        int q = 0;
        while (idx >= d->send_ptr[q + 1]) q++;
        const double2 val = v2[d->send_rows[idx]];
        double2 *dst = d->tail[q] + d->send_dst[q] + (idx - d->send_ptr[q]);
        *dst = val;
        }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        {
        const unsigned int t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
        }
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    __threadfence_system();
    const unsigned long long e = ++d->hepoch;
    for (int q = 0; q < d->world; q++)
        if (d->send_ptr[q + 1] > d->send_ptr[q]) st_sys(&d->ctrl[q]->hflag[d->rank], e);
    if (d->error) return;
    DistCtrl *me = d->ctrl[d->rank];
    for (int src = 0; src < d->world; src++)
        if (d->recv_from[src] && !wait_flag(&me->hflag[src], e))
            {
            d->error = 1;
            break;
            }
    __threadfence_system();
    }

int halo_exchange(const KrylovWork &w, int gate)
    {
    if (!w.dist) return FG_OK;
    const bool prof = prof_begin(w.prof, w.stream, KC_HALO);
    k_halo<<<w.halo_grid, BLOCK, 0, w.stream>>>(w.dist, w.x, gate, w.st, w.red.ticket);
    if (prof) prof_end(w.prof, w.stream);
    if (w.launches) ++*w.launches;
    FG_CUDA(cudaGetLastError());
    return FG_OK;
    }

static int poll_state(KrylovWork &w)
    {
    FG_CUDA(cudaMemcpyAsync(w.h_st, w.st, sizeof(KState), cudaMemcpyDeviceToHost, w.stream));
    FG_CUDA(cudaEventRecord(w.ev_poll, w.stream));
    FG_CUDA(cudaEventSynchronize(w.ev_poll));
    return FG_OK;
    }

// ------------------------------------------------------------------------------------------
// BiCGStab driver
// ------------------------------------------------------------------------------------------
int bicgstab_run(const Operator &op, KrylovWork &w, double tol, int maxiter, post_batch_fn post,
                 void *user)
    {
    const int n = w.n;
    // one resident wave per vector kernel (their register counts differ)
    const bool node3 = op.kind == OP_NODE3;
    static int wave_pn = 0, wave_sn = 0, wave_ps = 0, wave_ss = 0, wave_xr = 0, wave_xrn = 0;
    if (!wave_xr)
        {
        wave_xrn = resident_grid(k_bicg_xr_node);
        wave_pn = resident_grid(k_bicg_p_node);
        wave_sn = resident_grid(k_bicg_s_node);
        wave_ps = resident_grid(k_bicg_p);
        wave_ss = resident_grid(k_bicg_s);
        wave_xr = resident_grid(k_bicg_xr);
        }
    const int wave_p = node3 ? wave_pn : wave_ps, wave_s = node3 ? wave_sn : wave_ss;
    if (w.dist && !node3)
        {
        set_error("bicgstab_run: the multi-GPU path needs the matrix-free LLG operator");
        return FG_ERR_STATE;
        }
    const int gneed = grid_for(n, BLOCK * 2);
    const int wave_x = node3 ? wave_xrn : wave_xr;
    const int gp = gneed < wave_p ? gneed : wave_p, gs = gneed < wave_s ? gneed : wave_s,
              gxr = gneed < wave_x ? gneed : wave_x;
    // multi-GPU: the first npush CTAs of k_bicg_p push the boundary rows of D.p (>= 1 so that the
    // halo epoch advances on every rank, even one without neighbours)
    int npush = w.dist ? (w.nsend + BLOCK - 1) / BLOCK : 0;
    if (w.dist && npush < 1) npush = 1;
    if (npush > gp) npush = gp;
    FG_LAUNCH(w, k_init_state, 1, w.st, tol, maxiter);
    // r = b - A x0 (masked); rt = p = r; rhsn, first loop test
        {
        SpmvArgs a = {};
        a.x = w.x;
        a.y = w.r;
        a.a0 = w.b;
        a.w = w.w3p;  // image of the initial guess, written by the assembly
        a.o0 = w.rt;
        a.mask = w.mask;
        a.st = w.st;
        a.red = w.red;
        a.dist = w.dist;
        FG_TRY(launch_spmv<ST_BICG_SETUP>(op, w, a));
        }
    int enq = 0;
    // first batch sized by the previous solve of this context, then short batches
    int batch = w.last_iters > 0 ? w.last_iters + 1 : 8;
    for (;;)
        {
        if (batch > maxiter + 1 - enq) batch = maxiter + 1 - enq;
        if (batch < 1) batch = 1;
        for (int k = 0; k < batch; k++)
            {
            // iteration `enq + k` of this solve reads p from one buffer and writes the other one
            double *p_old = ((enq + k) & 1) ? w.p2 : w.p, *p_new = ((enq + k) & 1) ? w.p : w.p2;
            if (node3)
                FG_LAUNCH_C(w, KC_BICG_P, k_bicg_p_node, gp, n / 2, w.r, p_old, p_new, w.v, w.D,
                            static_cast<double *>(nullptr), w.qbasis, w.w3p, w.st, w.dist, w.red.ticket, npush);
            else
                FG_LAUNCH_C(w, KC_BICG_P, k_bicg_p, gp, n, w.r, p_old, p_new, w.v, w.D, w.phat, w.st);
            SpmvArgs a = {};
            a.dist = w.dist;
            a.w = w.w3p;
            a.x = node3 ? nullptr : w.phat;  // matrix-free: x_a is recovered from its image w_a
            a.y = w.v;
            a.a0 = w.rt;
            a.mask = w.mask;
            a.st = w.st;
            a.red = w.red;
            FG_TRY(launch_spmv<ST_BICG_V>(op, w, a));
            if (node3)
                FG_LAUNCH_C(w, KC_BICG_S, k_bicg_s_node, gs, n / 2, w.r, w.v, w.D, w.s,
                            static_cast<double *>(nullptr), w.qbasis, w.w3s, w.st, w.red);
            else
                FG_LAUNCH_C(w, KC_BICG_S, k_bicg_s, gs, n, w.r, w.v, w.D, w.s, w.shat, w.st, w.red);
            a.w = w.w3s;
            a.x = node3 ? nullptr : w.shat;
            a.y = w.t;
            a.a0 = w.s;
            FG_TRY(launch_spmv<ST_BICG_T>(op, w, a));
            if (node3)
                FG_LAUNCH_C(w, KC_BICG_XR, k_bicg_xr_node, gxr, n / 2, w.x, p_new, w.D, w.s, w.t, w.rt, w.r, w.st,
                            w.red);
            else
                FG_LAUNCH_C(w, KC_BICG_XR, k_bicg_xr, gxr, n, w.x, w.phat, w.shat, w.s, w.t, w.rt, w.r, w.st, w.red);
            }
        enq += batch;
        if (post) FG_TRY(post(user));
        FG_TRY(poll_state(w));
        if (w.h_st->done) break;
        if (enq > maxiter + 1)
            {
            set_error("bicgstab_run: device loop did not terminate after %d iterations", enq);
            return FG_ERR_STATE;
            }
        batch = 4;
        }
    w.last_iters = w.h_st->nit;
    return FG_OK;
    }

// ------------------------------------------------------------------------------------------
// BiCGStab as one persistent cooperative kernel (fg_solve_pk.cuh)
// ------------------------------------------------------------------------------------------
// the instantiation for a block size, column width and staging mode
static const void *pk_kernel(int bs, bool i16, bool staged, bool head)
    {
    if (staged)
        return bs == 1024 ? (const void *)k_llg_solve<1024, true, true> : (const void *)k_llg_solve<256, true, true>;
    if (bs == 1024) return i16 ? (const void *)k_llg_solve<1024, true, false> : (const void *)k_llg_solve<1024, false, false>;
    if (bs == 512)
        return i16 ? (const void *)k_llg_solve<512, true, false, true> : (const void *)k_llg_solve<512, false, false, true>;
    return i16 ? (const void *)k_llg_solve<256, true, false> : (const void *)k_llg_solve<256, false, false>;
    }

// Launch shape of the persistent kernel for an operator: CTA size, and whether the gathered images are
// staged in shared memory (returns true) with `smem` bytes of dynamic shared memory per CTA.
static bool pk_plan_sms(const Operator &op, int sms, int *bs_out, size_t *smem_out, bool *head_out)
    {
    // CTA size: 1024 threads (one CTA per SM, 148 arrivals per barrier) once every warp of such a grid has a
    // slice of its own; 256 threads below that, so that small meshes still spread over all SMs.  When 16
    // warps per SM own at most a slice each: the 512-thread variant that keeps the head of that slice in
    // registers (k_llg_solve<.., HEAD>), one CTA per SM (sp4, 1483 slices: 7.9 / 7.1 us per product
    // against 8.9 / 7.5, +3.4 % steps/s, r02z).  Not for a mesh that fits one 256-thread CTA, which has no
    // grid barrier at all (ellipsoid, 6 slices: the same variant with 256 threads measured 3.0 us per
    // product against 2.4).  FG_PK_BLOCK=512 forces the variant on any mesh (film20m: 16 warps per SM
    // with 128 registers are 33 % slower per product than 32 with 64).
    static const int forced = getenv("FG_PK_BLOCK") ? atoi(getenv("FG_PK_BLOCK")) : 0;
    static const bool no_head = getenv("FG_PK_NOHEAD") != nullptr && atoi(getenv("FG_PK_NOHEAD")) != 0;
    int bs = op.nslice >= sms * 32 ? 1024 : ((op.nslice > 8 && op.nslice <= sms * 16) ? 512 : 256);
    if (forced == 256 || forced == 512 || forced == 1024) bs = forced;
    const bool head = !no_head && bs == 512;
    if (bs == 512 && !head) bs = 256;  // 512 threads exist only as the HEAD variant
    if (head_out) *head_out = head;
    // gathered images staged in shared memory when the mesh has gather blocks, a buffer per thread group fits,
    // and there are enough blocks to keep every group of a full grid busy (a group walks 8 slices per block:
    // below that size one slice per warp through L1 has the shorter critical path)
    static const int min_blocks = getenv("FG_STAGE_MIN_BLOCKS") ? atoi(getenv("FG_STAGE_MIN_BLOCKS")) : -1;
    const int cap = (op.stage_cap + 3) & ~3;
    size_t smem = (size_t)(bs / 128) * (size_t)cap * sizeof(double4);
    const int need = min_blocks >= 0 ? min_blocks : sms * 8;
    const bool staged = !head && op.lcol != nullptr && op.stage_cap > 0 && smem <= (size_t)200 * 1024 && op.nblock >= need;
    if (bs_out) *bs_out = bs;
    if (smem_out) *smem_out = staged ? smem : 0;
    return staged;
    }
bool pk_plan(const Operator &op, int *bs_out, size_t *smem_out, bool *head_out)
    {
    int dev = 0, sms = NUM_SMS;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return pk_plan_sms(op, sms, bs_out, smem_out, head_out);
    }

// Warps per CTA of a 1024-thread launch on `grid` CTAs.  A product takes one round of the slice front per
// slice of the busiest warp, and a round costs about the same with 24 warps per SM as with 32 (latency, not
// issue slots: 6.7 us against 7.3 us on a 4-GPU partition of the 20 M-tet mesh), so: the fewest rounds first
// -- ceil(slices / warps of the grid) -- and among the counts in [24, 32] that reach it the smallest one,
// which leaves the last round fullest (8-GPU partition, 4.1 slices per warp at 32: 27 warps make the fifth
// round 90 % full, -1.7 us per product).  (Round 2 first picked the fullest last round outright: 24 warps
// and 11 rounds instead of 30 and 9 on the interior ranks at N = 4, 73 us per product against 66.)
static int pk_warps(int nslice, int grid)
    {
    static const int forced = getenv("FG_PK_WARPS") ? atoi(getenv("FG_PK_WARPS")) : 0;
    if (forced >= 1 && forced <= 32) return forced;
    if (nslice >= 16 * grid * 32) return 32;  // many rounds: the tail does not matter
    auto rounds = [&](int nw) { return (nslice + grid * nw - 1) / (grid * nw); };
    int best = 32;
    for (int nw = 31; nw >= 24; nw--)
        if (rounds(nw) <= rounds(best)) best = nw;
    return best;
    }

// What bicgstab_run_pk launches for an unstaged operator of `nslice` slices on a device with `sms` SMs,
// computed without a device (fg_solver_launch_shape; the residency per SM follows the launch bounds of
// k_llg_solve): out = {CTA size of the instantiation, warps per CTA launched, CTAs, 1 = HEAD variant}.
void pk_launch_shape(int nslice, int sms, int out[4])
    {
    Operator op = {};
    op.nslice = nslice;
    int bs = 0;
    bool head = false;
    pk_plan_sms(op, sms, &bs, nullptr, &head);
    const int per_sm = (bs == 1024 || head) ? 1 : 1024 / bs;
    int wave = per_sm * sms;
    if (wave > PK_MAX_GRID) wave = PK_MAX_GRID;
    int grid = (nslice + bs / 32 - 1) / (bs / 32);
    if (grid > wave) grid = wave;
    if (grid < 1) grid = 1;
    out[0] = bs;
    out[1] = bs == 1024 ? pk_warps(nslice, grid) : bs / 32;
    out[2] = grid;
    out[3] = head ? 1 : 0;
    }

int bicgstab_run_pk(const Operator &op, KrylovWork &w, double tol, int maxiter, const PkUpdate *upd)
    {
    if (op.kind != OP_NODE3 || !w.pk)
        {
        set_error("bicgstab_run_pk: needs the matrix-free LLG operator");
        return FG_ERR_STATE;
        }
    int bs = 0;
    size_t smem = 0;
    bool head = false;
    const bool staged = pk_plan(op, &bs, &smem, &head);
    const bool i16 = op.col16 != nullptr;
    int dev = 0, sms = NUM_SMS;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const void *fn = pk_kernel(bs, i16, staged, head);
    if (staged && smem > 48 * 1024)
        FG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, bs, smem) != cudaSuccess || per_sm < 1)
        {
        set_error("bicgstab_run_pk: the solve kernel does not fit an SM");
        return FG_ERR_CUDA;
        }
    int wave = per_sm * sms;
    if (wave > PK_MAX_GRID) wave = PK_MAX_GRID;
    // units of ownership: gather blocks per thread group (staged) or slices per warp
    const int units = staged ? op.nblock : op.nslice, per_cta = staged ? bs / 128 : bs / 32;
    int grid = (units + per_cta - 1) / per_cta;
    if (grid > wave) grid = wave;
    if (grid < 1) grid = 1;
    PkArgs a = {};
    a.op = op;
    a.NODp = w.n / 2;
    a.NODt = w.nx / 2;
    a.x = w.x; a.b = w.b; a.r = w.r; a.rt = w.rt; a.p = w.p; a.p2 = w.p2; a.v = w.v; a.s = w.s; a.t = w.t;
    a.D = w.D;
    a.w3p = w.w3p;
    a.w3s = w.w3s;
    a.mask = w.mask;
    a.st = w.st;
    a.sync = w.pk;
    a.dist = w.dist;
    a.tol = tol;
    a.maxiter = maxiter;
    if (upd)
        {
        a.nonmag = upd->nonmag;
        a.cur = upd->cur;
        a.next = upd->next;
        a.basis = upd->basis;
        a.dt = upd->dt;
        a.NODp = upd->NODp;
        a.NODt = upd->NODt;
        }
    a.phase_acc = w.pk_stamps_on ? w.pk_phase_acc : nullptr;
    static const bool mailbox = getenv("FG_PK_MAILBOX") == nullptr || atoi(getenv("FG_PK_MAILBOX")) != 0;
    if (mailbox)
        {
        a.h_st = w.d_h_st;
        a.h_seq = w.d_h_seq;
        a.seq = ++w.seq;
        }
    void *args[] = {&a};
    const bool prof = prof_begin(w.prof, w.stream, KC_SOLVE);
    // the instantiation bounds the CTA at `bs` threads; a 1024-thread unstaged launch may use fewer warps
    const int threads = (bs == 1024 && !staged) ? 32 * pk_warps(op.nslice, grid) : bs;
    FG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, w.stream));
    if (prof) prof_end(w.prof, w.stream);
    if (w.launches) ++*w.launches;
    if (mailbox)
        {  // spin on the sequence number the kernel's last thread writes (a few us after it, against ~20 us
           // for copy + event); the stream is queried now and then so that a failed launch cannot hang us
        volatile unsigned long long *seq = w.h_seq;
        unsigned int spins = 0;
        while (*seq != a.seq)
            {
            if ((++spins & 0xfffffu) == 0)
                {
                const cudaError_t q = cudaStreamQuery(w.stream);
                if (q != cudaErrorNotReady && *seq != a.seq)
                    {
                    if (q != cudaSuccess)
                        {
                        set_error("bicgstab_run_pk: %s", cudaGetErrorString(q));
                        return FG_ERR_CUDA;
                        }
                    break;  // finished without writing the mailbox: fall back to the copy
                    }
                }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            }
        if (*seq != a.seq) FG_TRY(poll_state(w));
        __sync_synchronize();
        }
    else
        FG_TRY(poll_state(w));
    if (!w.h_st->done)
        {
        set_error("bicgstab_run_pk: the solve kernel returned without finishing");
        return FG_ERR_STATE;
        }
    w.last_iters = w.h_st->nit;
    return FG_OK;
    }

// ------------------------------------------------------------------------------------------
// CG driver
// ------------------------------------------------------------------------------------------
int cg_run(const Operator &op, KrylovWork &w, double tol, int maxiter)
    {
    const int n = w.n;
    const int gv = grid_for(n, BLOCK * 4);
    FG_LAUNCH(w, k_init_state, 1, w.st, tol, maxiter);
        {
        SpmvArgs a = {};
        a.x = w.x;
        a.y = w.r;
        a.a0 = w.b;
        a.a1 = w.D;
        a.o1 = w.p;
        a.mask = w.mask;
        a.st = w.st;
        a.red = w.red;
        FG_TRY(launch_spmv<ST_CG_SETUP>(op, w, a));
        }
    int enq = 0;
    int batch = 16;
    for (;;)
        {
        if (batch > maxiter + 1 - enq) batch = maxiter + 1 - enq;
        if (batch < 1) batch = 1;
        for (int k = 0; k < batch; k++)
            {
            FG_LAUNCH(w, k_cg_p, gv, n, w.r, w.D, w.p, w.st);
            SpmvArgs a = {};
            a.x = w.p;
            a.y = w.t;  // q
            a.a0 = w.p;
            a.mask = w.mask;
            a.st = w.st;
            a.red = w.red;
            FG_TRY(launch_spmv<ST_CG_Q>(op, w, a));
            FG_LAUNCH(w, k_cg_xr, gv, n, w.x, w.p, w.t, w.D, w.r, w.st, w.red);
            }
        enq += batch;
        FG_TRY(poll_state(w));
        if (w.h_st->done) break;
        if (enq > maxiter + 1)
            {
            set_error("cg_run: device loop did not terminate after %d iterations", enq);
            return FG_ERR_STATE;
            }
        }
    return FG_OK;
    }

// b <- b - A xd (masked afterwards by the caller), the prelude of the *_dir(xd) variants
int resid_into(const Operator &op, const KrylovWork &w, const double *xd, const double *b,
               double *out)
    {
    SpmvArgs a = {};
    a.x = xd;
    a.y = out;
    a.a0 = b;
    a.mask = nullptr;
    a.st = w.st;
    a.red = w.red;
    return launch_spmv<ST_RESID>(op, w, a);
    }

}  // namespace fg
