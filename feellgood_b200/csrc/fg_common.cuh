// fg_common.cuh — shared declarations of the B200 (sm_100a) LLG hot-path library.
// FP64, HBM-bound kernels: no tensor cores anywhere (nothing here is a dense contraction).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/feellgood_b200.h"

// physical constants and scheme parameters, reference src/config.h.in:31-32,52
#define FG_MU0 1.25663706127e-6
#define FG_GAMMA0 (1.76085962784e11 * FG_MU0)
#define FG_THETA 0.5
#define FG_EPSILON 1e-40

namespace fg
{
void set_error(const char *fmt, ...);

#define FG_CUDA(call)                                                                          \
    do                                                                                         \
        {                                                                                      \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            {                                                                                  \
            fg::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return FG_ERR_CUDA;                                                                \
            }                                                                                  \
        } while (0)

#define FG_TRY(call)               \
    do                             \
        {                          \
        int rc_ = (call);          \
        if (rc_ != FG_OK) return rc_; \
        } while (0)

constexpr int NUM_SMS = 148;          // B200
constexpr int BLOCK = 256;            // every kernel of the library uses 256-thread CTAs
constexpr int CTAS_PER_SM = 8;        // 2048 resident threads per SM
constexpr int MAX_GRID = NUM_SMS * CTAS_PER_SM;  // persistent grid-stride kernels: one full wave
constexpr int RED_NV = 4;             // max simultaneous reductions per kernel
constexpr int SLICE = 32;             // SELL slice height = warp size

// Node record = the reference's Nodes::dataNode {u, v, phi, phiv} (src/node.h:47-53): 64 B, so a
// gather of one node by a tetrahedron is exactly two aligned 32-byte sectors.
struct __align__(16) NodeRec
    {
    double u[3];
    double v[3];
    double phi;
    double phiv;
    };
static_assert(sizeof(NodeRec) == 64, "NodeRec must be 64 bytes");

// tangent-plane basis of a node (src/node.h:62-64): 48 B
struct __align__(16) Basis
    {
    double ep[3];
    double eq[3];
    };
static_assert(sizeof(Basis) == 48, "Basis must be 48 bytes");

// per-region constants, pre-digested from fg_tet_prm (src/tetra.cpp:216-218,239,244)
struct TetRegion
    {
    double alpha, Abis, Kbis, K3bis;
    double A, K, K3, Ms;  // raw constants, read by the energy kernels (src/tetra.cpp:309-391)
    double uk[3], ex[3], ey[3], ez[3];
    int has_K, has_K3;
    };
struct TriRegion
    {
    double Ks;
    double uk[3];
    };

// iteration<T> (src/algebra/iter.h:37-179) + the Krylov scalars, resident on the device.
struct KState
    {
    double rho1, rho2, alpha, beta, omega;
    double res, rhsn, resmax;
    double v2max, v_max;
    int nit, maxiter, status;
    int done;        // loop finished (any reason)
    int final_half;  // converged on ||s||: x += alpha*phat still to apply (bicg.h:211-215)
    int updated;     // node update applied for this solve
    int failed;      // LinAlgebra::solve return value (src/solver.cpp:62-69)
    int hist_cap;    // rows of hist (0 = off)
    double *hist;    // optional per-iteration record [hist_cap][8]: rho_1, (v,rt), alpha, |s|^2, (t,s), (t,t),
                     // omega, |r|^2 — diagnosis of breakdowns (fg_get_krylov_history)
    };
__host__ __device__ __forceinline__ void khist(KState *st, int col, double v)
    {
    if (st->hist != nullptr && st->nit < st->hist_cap) st->hist[8 * (size_t)st->nit + col] = v;
    }

// buffers of the deterministic two-stage grid reduction
struct DistDev;
struct RedBuf
    {
    double *partials;     // [2 * RED_NV][MAX_GRID]: CTA partial sums and their compensations
    unsigned int *ticket; // one counter, self-resetting (atomicInc wrap)
    DistDev *dist;        // multi-GPU: the grid totals are all-reduced over the ranks (fg_dist.cuh)
    };

// The operator y = A x.
//  OP_SELL2: the LLG matrix K (src/solver.h:75-104) in SELL-32-sigma with 2x2 blocks, device row
//            order.  Slice s holds 32 node rows and sptr[s+1]-sptr[s] block-columns;
//            block-column j of the slice stores, for lane l = row s*32+l,
//              scol[(sptr[s]+j)*32 + l]                  device row of the column node
//              val2[((sptr[s]+j)*2 + 0)*32 + l]          (K[2r,2c], K[2r,2c+1])     as double2
//              val2[((sptr[s]+j)*2 + 1)*32 + l]          (K[2r+1,2c], K[2r+1,2c+1]) as double2
//            so one warp streams a slice with fully coalesced 128 B / 512 B requests and every
//            lane folds its own row left to right (no shuffles).  4 B of index per 32 B of values.
//  OP_CSR:   any algebra::SparseMatrix (src/algebra/sparseMat.h), plain CSR.
//  OP_NODE3: the same K, never materialised: K = cS P^T (S x I3) P + Dg, where S is
//            the per-mesh scalar stiffness in the same SELL-32 layout (8 B value + 4 B index per
//            node pair instead of 36 B per 2x2 block), P maps the 2 tangent-plane unknowns of a node
//            to the 3-vector w = ep x0 + eq x1, and Dg is the state-dependent 2x2 node-diagonal part
//            (alpha_eff mass + gyrotropic term).  The producer of an SpMV input writes w (double4
//            per node: one aligned 32-byte sector per gather); the SpMV folds z_a = sum_b S_ab w_b and
//            projects: y = cS (eq_a.z, ep_a.z) + Dg x_a.
//            Inside the Krylov kernels (ep, eq) are decoded from a unit quaternion (32 B per node instead
//            of 48) and Dg is applied in its closed form: with m x ep = eq, m x eq = -ep for the
//            orthonormal right-handed triad (ep, eq, m) of Node::setBasis,
//                Dg = [[a_w, Ma], [Ma, -a_w]]   (Ma = sum of the alpha_eff records, a_w = lumped mass),
//            16 B per node instead of 32.  The explicit triple products of the reference
//            (src/tetra.cpp:108-148) differ from this form by rounding only (~1e-16 relative, against a
//            parity target of 1e-12); they are kept in the assembled-block operator and the taps.
enum { OP_SELL2 = 0, OP_CSR = 1, OP_NODE3 = 2 };
struct Operator
    {
    int kind;
    int n;       // rows (OP_SELL2 / OP_NODE3: 2 * padded node count)
    int lanes;   // OP_CSR: lanes cooperating on one row: 2..32
    const int *ptr;     // OP_CSR rowptr (n+1) | OP_SELL2 / OP_NODE3 sptr (nslice+1)
    const int *col;     // OP_CSR col | OP_SELL2 / OP_NODE3 scol
    const short *col16; // OP_NODE3, optional: scol as 16-bit offsets from the lane's own row (2 B instead of
                        // 4 B per stored pair); NULL when an offset of the mesh does not fit
    int col16_partial;  // 1: col16 is valid only in the slices WITHOUT ghost columns (sghost == 0); the others
                        // read `col` (persistent kernel on a partition, where the ghost rows are out of reach)
    const double *val;  // OP_SELL2: 2x2 blocks | OP_NODE3: S (one double per stored node pair)
    int nslice;         // OP_SELL2 / OP_NODE3
    // OP_NODE3
    const double4 *qbasis;       // tangent-plane basis of every node as a unit quaternion (32 B, one 256-bit
                                 // request instead of 48 B): see basis_to_quat / quat_to_basis below
    const double2 *Dm;           // state-dependent node-diagonal part: (Ma, a_w) per node row, see OP_NODE3 above
    const unsigned char *nonmag; // 1 = identity row (node outside the magnetic material, pad row)
    double cS;                   // prefactor * s_dt (src/tetra.cpp:261)
    int prefetch;                // 1: pull the row-epilogue operands and the next slice's indices into L2 early
    const unsigned char *sghost; // partitioned operator (multi-GPU): 1 = the slice has a ghost column, i.e. it
                                 // must wait for the halo of its SpMV input; NULL on one GPU
    // gather blocks (fg_setup.hpp): 256 rows = 8 slices whose gathered images (own rows + halo) are staged in
    // shared memory by a thread group of the persistent kernel and addressed with 16-bit local indices
    const unsigned short *lcol;  // local column index per stored pair, same layout as col / col16; NULL = none
    const int *bptr;             // nblock + 1
    const int *bhalo;            // device rows of the halo images
    const unsigned char *bghost; // 1 = the halo of the block has a ghost row (multi-GPU), NULL on one GPU
    int nblock;
    int stage_cap;               // images a staging buffer must hold (256 + largest halo)
    };

// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): one request per 32-byte node image
// instead of a 128-bit + 64-bit pair -- halves the L1 wavefronts of the SpMV gather.  p must be 32-byte
// aligned (the images are double4 arrays in 256-byte-aligned allocations).
#ifdef __CUDACC__
__device__ __forceinline__ double4 ld256_nc(const double4 *p)
    {
    double4 r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
    }
// Tangent-plane basis <-> unit quaternion.  R = [ep | eq | ep x eq] is a rotation (Node::setBasis builds
// an orthonormal right-handed triad), so 4 doubles carry it; the decoded (ep, eq) are orthonormal to
// rounding and differ from the stored ones by a few 1e-16.  Every Krylov kernel decodes with this one
// function, so the operator they apply is consistent.
__host__ __device__ __forceinline__ double4 basis_to_quat(const double ep[3], const double eq[3])
    {
    const double n0 = ep[1] * eq[2] - ep[2] * eq[1], n1 = ep[2] * eq[0] - ep[0] * eq[2],
                 n2 = ep[0] * eq[1] - ep[1] * eq[0];
    const double r00 = ep[0], r10 = ep[1], r20 = ep[2], r01 = eq[0], r11 = eq[1], r21 = eq[2], r02 = n0,
                 r12 = n1, r22 = n2;
    const double tr = r00 + r11 + r22;
    double w, x, y, z;
    if (tr > 0.0)  // Shepperd's method: always divide by the largest component
        {
        const double s = 2.0 * sqrt(tr + 1.0);
        w = 0.25 * s; x = (r21 - r12) / s; y = (r02 - r20) / s; z = (r10 - r01) / s;
        }
    else if (r00 > r11 && r00 > r22)
        {
        const double s = 2.0 * sqrt(fmax(1.0 + r00 - r11 - r22, 0.0));
        w = (r21 - r12) / s; x = 0.25 * s; y = (r01 + r10) / s; z = (r02 + r20) / s;
        }
    else if (r11 > r22)
        {
        const double s = 2.0 * sqrt(fmax(1.0 + r11 - r00 - r22, 0.0));
        w = (r02 - r20) / s; x = (r01 + r10) / s; y = 0.25 * s; z = (r12 + r21) / s;
        }
    else
        {
        const double s = 2.0 * sqrt(fmax(1.0 + r22 - r00 - r11, 0.0));
        w = (r10 - r01) / s; x = (r02 + r20) / s; y = (r12 + r21) / s; z = 0.25 * s;
        }
    const double nn = w * w + x * x + y * y + z * z;
    if (!(nn > 1e-300) || !(nn < 1e300)) return make_double4(0.0, 0.0, 0.0, 1.0);  // degenerate (u = 0): identity
    const double inv = 1.0 / sqrt(nn);
    return make_double4(x * inv, y * inv, z * inv, w * inv);
    }
__host__ __device__ __forceinline__ void quat_to_basis(const double4 q, double ep[3], double eq[3])
    {
    const double x = q.x, y = q.y, z = q.z, w = q.w;
    ep[0] = 1.0 - 2.0 * (y * y + z * z);
    ep[1] = 2.0 * (x * y + z * w);
    ep[2] = 2.0 * (x * z - y * w);
    eq[0] = 2.0 * (x * y - z * w);
    eq[1] = 1.0 - 2.0 * (x * x + z * z);
    eq[2] = 2.0 * (y * z + x * w);
    }

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// streaming variant (evict-first): data read exactly once per step (the element records)
__device__ __forceinline__ double4 ld256_cs(const double4 *p)
    {
    double4 r;
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
    return r;
    }
__device__ __forceinline__ void st256(double4 *p, const double4 v)
    {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    }
#endif

// optional CUDA-event pairs around kernel launches: fg_set_profiling(ctx, 2) brackets every SpMV
// launch, fg_set_profiling(ctx, 3) every kernel of the step (classes below)
enum
    {
    KC_BASIS = 0, KC_TET, KC_TRI, KC_ASSEMBLE, KC_SPMV_SETUP, KC_BICG_P, KC_SPMV_V, KC_BICG_S,
    KC_SPMV_T, KC_BICG_XR, KC_HALO, KC_UPDATE, KC_SOLVE, KC_OTHER, KC_COUNT
    };
struct SpmvProf
    {
    cudaEvent_t *ev;  // 2*cap events
    int n, cap;       // pairs recorded / capacity
    int *cls;         // kernel class of each pair
    int mode;         // 2: SpMV launches only | 3: every kernel
    };
inline bool prof_is_spmv(int cls) { return cls == KC_SPMV_SETUP || cls == KC_SPMV_V || cls == KC_SPMV_T; }
inline bool prof_mode2(int cls) { return prof_is_spmv(cls) || cls == KC_SOLVE; }
inline bool prof_begin(SpmvProf *p, cudaStream_t s, int cls)
    {
    if (!p || p->n >= p->cap || (p->mode == 2 && !prof_mode2(cls))) return false;
    p->cls[p->n] = cls;
    cudaEventRecord(p->ev[2 * p->n], s);
    return true;
    }
inline void prof_end(SpmvProf *p, cudaStream_t s)
    {
    cudaEventRecord(p->ev[2 * p->n + 1], s);
    p->n++;
    }

// Krylov workspace (all device pointers, length n [+ ghosts for vectors that feed an SpMV])
struct KrylovWork
    {
    int n;       // owned rows
    int nx;      // length of vectors that are SpMV inputs (n + ghost entries)
    double *x, *b, *r, *rt, *p, *p2, *v, *s, *t, *phat, *shat, *D;  // p2: ping-pong partner of p
    // OP_NODE3: the 3-vector images (double4 per node, length nx/2) of the SpMV inputs x0 | phat
    // (shared buffer) and shat, and the basis that maps between the two representations
    double4 *w3p, *w3s;
    const double4 *qbasis;
    const unsigned char *mask; // n : 1 = Dirichlet dof (lvd), may be NULL
    KState *st;                // device
    KState *h_st;              // pinned host mirror
    RedBuf red;
    cudaStream_t stream;
    long long *launches;       // host counter of kernel launches
    cudaEvent_t ev_poll;
    int last_iters;            // iterations of the previous solve (sizes the first batch)
    SpmvProf *prof;            // NULL unless SpMV profiling is on
    // row-block multi-GPU path (NULL on one GPU): x, phat, shat then live in the IPC arena
    DistDev *dist;
    void *arena;
    int halo_grid;
    int nsend;                 // boundary rows this rank pushes to its neighbours
    // persistent solve kernel (fg_solve_pk.cuh)
    struct PkSync *pk;         // barrier / reduction scratch (NULL: not allocated)
    unsigned long long *pk_phase_acc;  // device [32]: in-kernel phase times (ns) and counts, by PKP_* id
    int pk_stamps_on;
    // result mailbox of the persistent kernel: its last thread writes KState and a sequence number straight
    // into mapped pinned host memory; the host spins on the number instead of waiting for a copy + event
    unsigned long long *h_seq;     // pinned, mapped
    unsigned long long seq;        // last sequence number handed to a kernel
    KState *d_h_st;                // device views of h_st / h_seq
    unsigned long long *d_h_seq;
    };

// phases of the persistent solve kernel (ids of its in-kernel time stamps)
enum { PKP_START = 1, PKP_SETUP, PKP_A, PKP_B, PKP_C, PKP_D, PKP_E, PKP_HALO_X, PKP_UPDATE, PKP_END };

// what the persistent solve kernel needs for the fused node update (src/solver.cpp:62-88)
struct PkUpdate
    {
    const unsigned char *nonmag;
    const NodeRec *cur;
    NodeRec *next;
    const Basis *basis;
    double dt;
    int NODp, NODt;
    };


// ---- fg_krylov.cu ----
// node3: also allocate the 3-vector images of the SpMV inputs (matrix-free LLG operator);
// ext (optional): externally owned storage for x, w3p, w3s (the multi-GPU exchange arena)
int krylov_alloc(KrylovWork &w, int n, int n_ghost, cudaStream_t stream, long long *launch_counter,
                 bool node3 = false, double *const ext[3] = nullptr);
void krylov_free(KrylovWork &w);
int grid_for(long long work_items, int items_per_cta);
// y = A x (optionally masked).  OP_NODE3: make_w builds the 3-vector image of x in w.w3s first;
// pass false when w.w3s already holds it (repeated products with the same x)
int spmv(const Operator &op, const KrylovWork &w, const double *x, double *y, bool masked, bool make_w = true);
// BiCGStab with Jacobi preconditioner and Dirichlet mask, reference src/algebra/bicg.h:163-234.
// w.x holds the initial guess, w.b the (already masked) rhs, w.D the (masked) inverse diagonal.
// post_batch, if not NULL, is called after each enqueued batch of iterations (same stream) so the
// caller can append gated work (the node update) before the host polls the state.
typedef int (*post_batch_fn)(void *user);
int bicgstab_run(const Operator &op, KrylovWork &w, double tol, int maxiter, post_batch_fn post,
                 void *user);
// The same solve (OP_NODE3 only) as ONE persistent cooperative kernel, fused with the node update when
// upd != NULL (fg_solve_pk.cuh); one host synchronisation per solve.
int bicgstab_run_pk(const Operator &op, KrylovWork &w, double tol, int maxiter, const PkUpdate *upd);
bool pk_plan(const Operator &op, int *bs_out, size_t *smem_out, bool *head_out = nullptr);
void pk_launch_shape(int nslice, int sms, int out[4]);
// Jacobi-preconditioned CG, reference src/algebra/cg.h:15-58,68-121 (same conventions)
int cg_run(const Operator &op, KrylovWork &w, double tol, int maxiter);
// D = 1/diag(A) for a plain CSR operator (src/algebra/sparseMat.h:174-183), then masked
int build_diag_precond_csr(const Operator &A, const KrylovWork &w);
int vec_mask(const KrylovWork &w, double *x);                  // x[lvd] = 0
int vec_axpy(const KrylovWork &w, double a, const double *x, double *y);  // y += a x
// multi-GPU halo exchange of the solution x (fg_dist.cuh);
// gate = 1: skipped once the solve is done | 2: only when done and the node update is pending
int halo_exchange(const KrylovWork &w, int gate);

}  // namespace fg
