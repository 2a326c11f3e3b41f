// fg_llg.cu — the LLG step context and its C ABI (include/feellgood_b200.h).
//
// One fg_ctx = one LinAlgebra object of the reference (src/linear_algebra.h) bound to one B200 and
// one CUDA stream.  All mesh tables, the node state (CURRENT and NEXT), the sparse system and the
// Krylov workspace stay resident in HBM between steps; a step moves no bulk data across PCIe.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stddef.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "fg_common.cuh"
#include "fg_llg_kernels.cuh"
#include "fg_setup.hpp"

namespace fg
{
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
    {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    }
int resid_into(const Operator &op, const KrylovWork &w, const double *xd, const double *b,
               double *out);
}  // namespace fg

using namespace fg;

struct fg_ctx
    {
    int device = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    HostSetup h;
    int NOD = 0, NODp = 0, NODt = 0, NTm = 0, NFa = 0, n = 0, np = 0, nnzb = 0;
    // multi-GPU (fg_dist_*): world == 1 on a single device
    int rank = 0, world = 1;
    void *arena = nullptr;
    size_t arena_bytes = 0, tail_off[3] = {0, 0, 0};  // ghost tails of x, w3p, w3s in the arena
    DistDev h_dist = {};
    DistDev *d_dist = nullptr;
    int *send_rows = nullptr;
    void *peer_base[DIST_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool connected = false;
    long long nblk = 0;          // stored blocks incl. SELL padding
    double tol = 1e-6;
    int maxiter = 700;
    // node state
    NodeRec *cur = nullptr, *next = nullptr;
    Basis *basis = nullptr;
    unsigned char *nonmag = nullptr, *dofmask = nullptr;
    double *stage = nullptr;     // 8*NOD doubles, device
    double *h_stage = nullptr;   // 8*NOD doubles, pinned host
    // tets
    int4 *tet_ind = nullptr, *tet_slot = nullptr;
    // chunks of 256 tetrahedra and their distinct nodes (k_tet_iso_st)
    int *tch_ptr = nullptr, *tch_nodes = nullptr;
    ushort4 *tet_loc = nullptr;
    int tch_n = 0, tch_cap = 0;
    double *tet_da = nullptr, *tet_detJ = nullptr, *ext_field = nullptr;
    int *tet_reg = nullptr;
    TetRegion *reg_tet = nullptr;
    double4 *rec = nullptr;
    // tris
    int *tri_ind = nullptr, *tri_reg = nullptr;
    double *tri_surf = nullptr, *tri_dMs = nullptr;
    TriRegion *reg_tri = nullptr;
    double2 *trec = nullptr;
    std::vector<TriRegion> h_reg_tri;
    // pattern and per-mesh constants
    short *scol16 = nullptr;  // scol as 16-bit offsets from the row (NULL when one does not fit)
    short *scol16_ng = nullptr;  // multi-GPU: the same, valid in the slices without ghost columns only
    int *perm = nullptr, *sptr = nullptr, *scol = nullptr, *sdeg = nullptr, *iptr = nullptr,
        *itptr = nullptr, *sinct = nullptr;
    double *sS = nullptr, *Aw = nullptr, *Sdiag = nullptr;
    double2 *Dm = nullptr;     // (Ma, a_w) per node row: the node-diagonal part of the matrix-free operator
    double4 *qbasis = nullptr; // the basis of every node (owned + ghost) as a unit quaternion
    double *val = nullptr;       // assembled 2x2 blocks of K: only the parity tap materialises them
    Operator op_K = {};          // the assembled-K operator of the tap (OP_SELL2)
    bool use_blocks = false;     // fg_set_operator(ctx, 1): solve with the assembled 2x2 blocks (A/B checks)
    int solver_kind = 0;         // fg_set_solver: 0 persistent kernel, 1 one kernel per phase
    unsigned char *sghost = nullptr;  // multi-GPU: slices with a ghost column
    unsigned long long *d_unit = nullptr;  // fg_set_state: max | |u|^2 - 1 | over the magnetic nodes (bits of a double)
    unsigned short *lcol = nullptr;   // gather blocks of the persistent SpMV (fg_setup.hpp)
    int *bptr = nullptr, *bhalo = nullptr;
    unsigned char *bghost = nullptr;
    KrylovWork kw;
    Operator op;
    // energies / averages / max angle: tables built on first use
    bool obs_ready = false;
    int NFm = 0, n_extra = 0, nreg_tet = 0;
    int *mtri_ind = nullptr, *mtri_reg = nullptr;
    double *mtri_surf = nullptr, *mtri_nrm = nullptr, *mtri_dMs = nullptr;
    int2 *extra_edges = nullptr;
    double *d_scal = nullptr, *h_scal = nullptr;  // 8 doubles: reduction results (device / pinned)
    // charges + all-pairs demag: tables built on first use
    bool demag_ready = false;
    long long nsrc = 0;
    double *node_pos = nullptr, *corr = nullptr, *tcorr = nullptr;
    double4 *src = nullptr;
    int *cptr = nullptr, *cidx = nullptr;
    // step bookkeeping
    StepPrm sp = {};
    bool have_basis = false, prepared = false, space_field = false, assembled = false;
    bool commit_pending = false;  // fg_commit is lazy: the next k_basis copies NEXT -> CURRENT on its way
    bool iso_regions = true;   // no region has K or K3: the element fast path applies (k_tet_iso)
    bool cubic_regions = false;  // a magnetic region has K3: the lean general kernel does not apply
    double v_max = 0.0;
    // profiling
    int profiling = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    SpmvProf prof = {nullptr, 0, 0, nullptr, 2};
    };

namespace
{
template <class T> int dev_upload(T **dst, const std::vector<T> &src, cudaStream_t s, size_t min_elems = 1)
    {
    const size_t nel = src.size() > min_elems ? src.size() : min_elems;
    FG_CUDA(cudaMalloc(dst, sizeof(T) * nel));
    if (!src.empty())
        FG_CUDA(cudaMemcpyAsync(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, s));
    return FG_OK;
    }

#define CTX_LAUNCH_C(c, cls, kernel, grid, ...)                    \
    do                                                             \
        {                                                          \
        const bool prof_ = prof_begin((c)->kw.prof, (c)->stream, cls); \
        kernel<<<(grid), BLOCK, 0, (c)->stream>>>(__VA_ARGS__);    \
        if (prof_) prof_end((c)->kw.prof, (c)->stream);            \
        ++(c)->launches;                                           \
        FG_CUDA(cudaGetLastError());                               \
        } while (0)
#define CTX_LAUNCH(c, kernel, grid, ...) CTX_LAUNCH_C(c, KC_OTHER, kernel, grid, __VA_ARGS__)

int check_ctx(const fg_ctx *c)
    {
    if (!c)
        {
        set_error("null context");
        return FG_ERR_INVALID;
        }
    FG_CUDA(cudaSetDevice(c->device));
    return FG_OK;
    }

// mesh::evolution is executed lazily: fg_commit only marks it, the next base_projection copies NEXT ->
// CURRENT inside k_basis (which reads every node record anyway); any other entry point that could see
// CURRENT first performs the copy here.
int flush_commit(fg_ctx *c)
    {
    if (!c->commit_pending) return FG_OK;
    FG_CUDA(cudaMemcpyAsync(c->cur, c->next, sizeof(NodeRec) * (size_t)c->NODt, cudaMemcpyDeviceToDevice, c->stream));
    c->commit_pending = false;
    return FG_OK;
    }

TetArrays tet_arrays(const fg_ctx *c)
    {
    TetArrays A;
    A.NTm = c->NTm;
    A.ind = c->tet_ind;
    A.da = c->tet_da;
    A.detJ = c->tet_detJ;
    A.reg = c->tet_reg;
    A.regions = c->reg_tet;
    A.ext_field = c->ext_field;
    A.slot = c->tet_slot;
    return A;
    }

int launch_basis(fg_ctx *c, double angle)
    {
    if (c->commit_pending)
        {  // evolution fused: read NEXT, write it to CURRENT, build the basis from it
        CTX_LAUNCH_C(c, KC_BASIS, k_basis<true>, grid_for(c->NODt, BLOCK), c->NODt, c->next, c->cur, cos(angle), sin(angle),
                     c->basis, c->qbasis);
        c->commit_pending = false;
        }
    else
        CTX_LAUNCH_C(c, KC_BASIS, k_basis<false>, grid_for(c->NODt, BLOCK), c->NODt, c->cur, c->cur, cos(angle), sin(angle),
                     c->basis, c->qbasis);
    c->have_basis = true;
    c->prepared = false;
    c->assembled = false;
    return FG_OK;
    }

// the element fast path (k_tet_iso): no anisotropy anywhere, uniform field, no recentring drift
static bool use_iso(const fg_ctx *c)
    {
    static const bool off = getenv("FG_NO_ISO") != nullptr;
    return !off && c->iso_regions && !c->space_field && c->sp.idx_dir == FG_IDX_UNDEF;
    }

int launch_elements(fg_ctx *c)
    {
    if (!c->have_basis)
        {
        set_error("prepareElements before base_projection");
        return FG_ERR_STATE;
        }
    if (c->NTm > 0)
        {
        const TetArrays A = tet_arrays(c);
        const int grid = grid_for(c->NTm, BLOCK);
        // A/B (FG_TET_STAGE=1): node records of a chunk of tetrahedra staged in shared memory (k_tet_st).  Measured
        // slower than the free-running gather kernels on the 20 M-tet film (r02g: 1.51 against 1.35 ms isotropic,
        // 2.92 against 2.10 ms with anisotropy): the element kernel is co-limited by DRAM (0.87 ms of 1.34), the
        // L1 pipe (0.88 ms) and FP64 issue (38 % busy), and the barriers of a staged chunk loop overlap the three
        // worse than 16 independent warps do.  Off by default.
        static const bool tet_stage = getenv("FG_TET_STAGE") != nullptr && atoi(getenv("FG_TET_STAGE")) != 0;
        const bool iso = use_iso(c);
        const size_t st_smem = 2 * (size_t)c->tch_cap * tet_stage_doubles(iso) * sizeof(double);
        if (tet_stage && c->tch_n > 0 && st_smem <= 160 * 1024)
            {  // the chunk's distinct node records staged in shared memory (k_tet_st)
            TetChunks C;
            C.nchunk = c->tch_n;
            C.cap = c->tch_cap;
            C.ptr = c->tch_ptr;
            C.nodes = c->tch_nodes;
            C.loc = c->tet_loc;
            typedef void (*kfn)(const TetArrays, const TetChunks, const NodeRec *, const StepPrm, double4 *);
            const int five = c->h.npi_tet == 5 ? 1 : 0, space = c->space_field ? 1 : 0;
            static const kfn table[2][3] = {{k_tet_st<1, true, false>, k_tet_st<1, false, false>, k_tet_st<1, false, true>},
                                            {k_tet_st<5, true, false>, k_tet_st<5, false, false>, k_tet_st<5, false, true>}};
            const int variant = iso ? 0 : 1 + space;
            const kfn fn = table[five][variant];
            static int waves[2][3] = {{0, 0, 0}, {0, 0, 0}};
            static size_t wave_smem[2][3] = {{0, 0, 0}, {0, 0, 0}};
            int &wave = waves[five][variant];
            if (!wave || wave_smem[five][variant] != st_smem)
                {
                if (st_smem > 48 * 1024)
                    FG_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem));
                int per_sm = 1, sms = NUM_SMS;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, BLOCK, st_smem) != cudaSuccess || per_sm < 1)
                    per_sm = 1;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
                wave = per_sm * sms;
                wave_smem[five][variant] = st_smem;
                }
            const int g = c->tch_n < wave ? c->tch_n : wave;
            const bool prof_ = prof_begin(c->kw.prof, c->stream, KC_TET);
            fn<<<g, BLOCK, st_smem, c->stream>>>(A, C, c->cur, c->sp, c->rec);
            if (prof_) prof_end(c->kw.prof, c->stream);
            ++c->launches;
            FG_CUDA(cudaGetLastError());
            }
        else if (use_iso(c))
            {
            static const bool pipe = getenv("FG_TET_NOPIPE") == nullptr;  // A/B switch of the prefetch pipeline
            if (c->h.npi_tet == 5 && pipe)
                CTX_LAUNCH_C(c, KC_TET, (k_tet_iso<5, true>), grid, A, c->cur, c->basis, c->sp, c->rec);
            else if (c->h.npi_tet == 5)
                CTX_LAUNCH_C(c, KC_TET, (k_tet_iso<5, false>), grid, A, c->cur, c->basis, c->sp, c->rec);
            else if (pipe)
                CTX_LAUNCH_C(c, KC_TET, (k_tet_iso<1, true>), grid, A, c->cur, c->basis, c->sp, c->rec);
            else
                CTX_LAUNCH_C(c, KC_TET, (k_tet_iso<1, false>), grid, A, c->cur, c->basis, c->sp, c->rec);
            }
        else if (!c->cubic_regions && c->sp.idx_dir == FG_IDX_UNDEF && getenv("FG_TET_NOLEAN") == nullptr)
            {  // uniaxial anisotropy only, no drift: the lean general kernel (k_tet_lean)
            const int gl = grid_for(c->NTm, TET_LEAN_BLOCK);
            const bool prof_ = prof_begin(c->kw.prof, c->stream, KC_TET);
            if (c->h.npi_tet == 5 && c->space_field)
                k_tet_lean<5, true><<<gl, TET_LEAN_BLOCK, 0, c->stream>>>(A, c->cur, c->sp, c->rec);
            else if (c->h.npi_tet == 5)
                k_tet_lean<5, false><<<gl, TET_LEAN_BLOCK, 0, c->stream>>>(A, c->cur, c->sp, c->rec);
            else if (c->space_field)
                k_tet_lean<1, true><<<gl, TET_LEAN_BLOCK, 0, c->stream>>>(A, c->cur, c->sp, c->rec);
            else
                k_tet_lean<1, false><<<gl, TET_LEAN_BLOCK, 0, c->stream>>>(A, c->cur, c->sp, c->rec);
            if (prof_) prof_end(c->kw.prof, c->stream);
            ++c->launches;
            FG_CUDA(cudaGetLastError());
            }
        else if (c->h.npi_tet == 5)
            {
            if (c->space_field)
                CTX_LAUNCH_C(c, KC_TET, (k_tet<5, true>), grid, A, c->cur, c->basis, c->sp, c->rec);
            else
                CTX_LAUNCH_C(c, KC_TET, (k_tet<5, false>), grid, A, c->cur, c->basis, c->sp, c->rec);
            }
        else
            {
            if (c->space_field)
                CTX_LAUNCH_C(c, KC_TET, (k_tet<1, true>), grid, A, c->cur, c->basis, c->sp, c->rec);
            else
                CTX_LAUNCH_C(c, KC_TET, (k_tet<1, false>), grid, A, c->cur, c->basis, c->sp, c->rec);
            }
        }
    if (c->NFa > 0)
        {
        TriArrays F;
        F.NFa = c->NFa;
        F.ind = c->tri_ind;
        F.surf = c->tri_surf;
        F.dMs = c->tri_dMs;
        F.reg = c->tri_reg;
        F.regions = c->reg_tri;
        const int grid = grid_for(c->NFa, BLOCK);
        if (c->h.npi_tri == 4)
            CTX_LAUNCH_C(c, KC_TRI, k_tri<4>, grid, F, c->cur, c->basis, c->trec);
        else
            CTX_LAUNCH_C(c, KC_TRI, k_tri<1>, grid, F, c->cur, c->basis, c->trec);
        }
    c->prepared = true;
    c->assembled = false;
    return FG_OK;
    }

int launch_assemble_K(fg_ctx *c);

int launch_assemble(fg_ctx *c, double dt)
    {
    if (!c->prepared)
        {
        set_error("solve before prepareElements");
        return FG_ERR_STATE;
        }
    if (dt != c->sp.dt)
        {
        set_error("solve(dt=%g) does not match prepareElements(dt=%g)", dt, c->sp.dt);
        return FG_ERR_STATE;
        }
    NodeAsmArrays R;
    R.nslice = c->h.nslice;
    R.Sdiag = c->Sdiag;
    R.Aw = c->Aw;
    R.iptr = c->iptr;
    R.itptr = c->itptr;
    R.sinct = c->sinct;
    R.nonmag = c->nonmag;
    const double s_dt = FG_THETA * dt * FG_GAMMA0;
    const double cS = c->sp.prefactor * s_dt;  // tetra.cpp:261: lumping(a_eff, prefactor*s_dt*Abis)
    static int wave = 0;
    if (!wave)
        {
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_node, BLOCK, 0) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        int sms = NUM_SMS;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        wave = per_sm * sms;
        }
    int grid = (c->h.nslice + BLOCK / 32 - 1) / (BLOCK / 32);
    if (grid > wave) grid = wave;
    if (grid < 1) grid = 1;
    c->op.cS = cS;
    CTX_LAUNCH_C(c, KC_ASSEMBLE, k_assemble_node, grid, R, c->cur, c->next, c->basis, c->rec, c->trec, cS,
                 c->Dm, c->kw.b, c->kw.x, c->kw.w3p, c->kw.D);
    if (c->NODt > c->NODp)
        CTX_LAUNCH(c, k_ghost_guess, (c->NODt - c->NODp + BLOCK - 1) / BLOCK, c->NODp, c->NODt, c->nonmag,
                   c->next, c->basis, c->kw.x, c->kw.w3p);
    if (c->use_blocks) FG_TRY(launch_assemble_K(c));
    c->assembled = true;
    return FG_OK;
    }

// Parity tap only: materialise the 2x2 blocks of K (solver::buildMat, src/solver.h:110-129) in the
// SELL layout with the row-assembly kernel; rhs / guess / diagonal go to Krylov scratch vectors.
int launch_assemble_K(fg_ctx *c)
    {
    const size_t nval = 4 * (size_t)(c->nblk > 0 ? c->nblk : 1);
    if (!c->val)
        {
        FG_CUDA(cudaMalloc(&c->val, sizeof(double) * nval));
        c->op_K = c->op;
        c->op_K.kind = OP_SELL2;
        c->op_K.val = c->val;
        }
    RowArrays R;
    R.nslice = c->h.nslice;
    R.sptr = c->sptr;
    R.scol = c->scol;
    R.sdeg = c->sdeg;
    R.sS = c->sS;
    R.Aw = c->Aw;
    R.iptr = c->iptr;
    R.itptr = c->itptr;
    R.sinct = c->sinct;
    R.nonmag = c->nonmag;
    int grid = (c->h.nslice + BLOCK / 32 - 1) / (BLOCK / 32);
    if (grid > MAX_GRID) grid = MAX_GRID;
    if (grid < 1) grid = 1;
    CTX_LAUNCH(c, k_assemble_sell, grid, R, c->cur, c->next, c->basis, c->rec, c->trec, c->op.cS, c->val,
               c->kw.t, c->kw.v, c->kw.s);
    return FG_OK;
    }

int post_update(void *user)
    {
    fg_ctx *c = static_cast<fg_ctx *>(user);
    FG_TRY(halo_exchange(c->kw, 2));  // multi-GPU: the solution of the ghost rows
    CTX_LAUNCH_C(c, KC_UPDATE, k_update, grid_for(c->NODt, BLOCK), c->NODt, c->NODp, c->nonmag, c->cur, c->next, c->basis,
               c->kw.x, c->sp.dt, c->kw.st, c->kw.red);
    return FG_OK;
    }

int run_solve(fg_ctx *c, double dt, fg_step_result *out)
    {
    if (c->arena && !c->connected)
        {
        set_error("solve: distributed context not connected (fg_dist_connect)");
        return FG_ERR_DIST;
        }
    if (c->profiling) FG_CUDA(cudaEventRecord(c->ev[2], c->stream));
    FG_TRY(launch_assemble(c, dt));
    if (c->profiling) FG_CUDA(cudaEventRecord(c->ev[3], c->stream));
    if (c->solver_kind == 0 && !c->use_blocks)
        {
        PkUpdate upd;
        upd.nonmag = c->nonmag;
        upd.cur = c->cur;
        upd.next = c->next;
        upd.basis = c->basis;
        upd.dt = c->sp.dt;
        upd.NODp = c->NODp;
        upd.NODt = c->NODt;
        Operator opk = c->op;
        if (!opk.col16 && c->scol16_ng)
            {  // 16-bit offsets in the slices without ghost columns, 32-bit columns in the others
            opk.col16 = c->scol16_ng;
            opk.col16_partial = 1;
            }
        FG_TRY(bicgstab_run_pk(opk, c->kw, c->tol, c->maxiter, &upd));
        }
    else
        FG_TRY(bicgstab_run(c->use_blocks ? c->op_K : c->op, c->kw, c->tol, c->maxiter, post_update, c));
    if (c->profiling)
        {
        FG_CUDA(cudaEventRecord(c->ev[4], c->stream));
        FG_CUDA(cudaEventSynchronize(c->ev[4]));
        float ms;
        for (int k = 0; k < 4; k++)
            {
            FG_CUDA(cudaEventElapsedTime(&ms, c->ev[k], c->ev[k + 1]));
            c->phase_ms[k] = ms;
            }
        }
    const KState &st = *c->kw.h_st;
    if (c->d_dist && st.status == FG_CANNOT_CONVERGE)
        {
        int derr = 0;
        FG_CUDA(cudaMemcpy(&derr, reinterpret_cast<char *>(c->d_dist) + offsetof(DistDev, error), sizeof(int),
                           cudaMemcpyDeviceToHost));
        if (derr)
            {
            set_error("solve: a peer did not answer within %g s (multi-GPU exchange timed out)",
                      (double)DIST_TIMEOUT_NS * 1e-9);
            return FG_ERR_DIST;
            }
        }
    if (!st.updated)
        {
        set_error("solve: node update did not run (done=%d)", st.done);
        return FG_ERR_STATE;
        }
    if (!st.failed) c->v_max = st.v_max;
    c->prepared = false;  // Kp/Lp are consumed; the reference would reuse them, we require a new prepare
    if (out)
        {
        out->failed = st.failed;
        out->status = st.status;
        out->iters = st.nit;
        out->pad_ = 0;
        out->res = st.res;
        out->rhsnorm = st.rhsn;
        out->v_max = c->v_max;
        }
    return FG_OK;
    }

// host array (pageable or pinned) -> staging -> NodeRec fields
int push_fields(fg_ctx *c, NodeRec *dst, const double *u, const double *v, const double *phi,
                const double *phiv, bool zero_missing)
    {
    const size_t N = (size_t)c->NOD;
    int which = 0;
    struct Part { const double *src; size_t off, len; int bit; };
    const Part parts[4] = {{u, 0, 3 * N, 1}, {v, 3 * N, 3 * N, 2}, {phi, 6 * N, N, 4}, {phiv, 7 * N, N, 8}};
    for (const Part &p : parts)
        {
        if (p.src)
            FG_CUDA(cudaMemcpyAsync(c->stage + p.off, p.src, sizeof(double) * p.len,
                                    cudaMemcpyHostToDevice, c->stream));
        else if (zero_missing)
            FG_CUDA(cudaMemsetAsync(c->stage + p.off, 0, sizeof(double) * p.len, c->stream));
        else
            continue;
        which |= p.bit;
        }
    if (which) CTX_LAUNCH(c, k_pack, grid_for(c->NODt, BLOCK), c->NODt, c->NOD, c->perm, dst, c->stage, which);
    // the caller's buffer may be pageable and reused right away
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FG_OK;
    }

// device row order <-> caller's dof order, for the taps (host side)
void dofs_to_caller(const fg_ctx *c, const std::vector<double> &dev, double *out)
    {
    for (int a = 0; a < c->NOD; a++)
        {
        const size_t r = (size_t)c->h.iperm[a];
        out[2 * (size_t)a] = dev[2 * r];
        out[2 * (size_t)a + 1] = dev[2 * r + 1];
        }
    }
void dofs_to_device(const fg_ctx *c, const double *in, std::vector<double> &dev)
    {
    dev.assign(2 * (size_t)c->NODt, 0.0);
    for (int a = 0; a < c->NOD; a++)
        {
        const size_t r = (size_t)c->h.iperm[a];
        dev[2 * r] = in[2 * (size_t)a];
        dev[2 * r + 1] = in[2 * (size_t)a + 1];
        }
    }
}  // namespace

extern "C" {

const char *fg_last_error(void) { return g_err; }
int fg_version(void) { return 100; }

static int create_ctx(const fg_mesh *mesh, const fg_params *prm, int device, const fg_dist_desc *dd,
                      fg_ctx **out)
    {
    if (!mesh || !prm || !out)
        {
        set_error("fg_create: null argument");
        return FG_ERR_INVALID;
        }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        {
        set_error("fg_create: no CUDA device available (this library has no CPU fallback)");
        return FG_ERR_CUDA;
        }
    if (device < 0 || device >= ndev)
        {
        set_error("fg_create: device %d out of range (%d devices)", device, ndev);
        return FG_ERR_INVALID;
        }
    FG_CUDA(cudaSetDevice(device));
    fg_ctx *c = new fg_ctx();
    c->device = device;
    std::string err;
    int rc = host_setup(*mesh, *prm, dd ? dd->n_owned : -1, c->h, err);
    if (rc != FG_OK)
        {
        set_error("%s", err.c_str());
        delete c;
        return rc;
        }
    HostSetup &h = c->h;
    c->NOD = h.NOD;
    c->NODp = h.NODp;
    c->NODt = h.NODt;
    if (dd)
        {
        c->rank = dd->rank;
        c->world = dd->world;
        }
    c->NTm = (int)h.magTet.size();
    c->NFa = (int)h.actTri.size();
    c->n = 2 * h.NOD;
    c->np = 2 * h.NODp;
    c->nnzb = h.nptr[h.NOD];
    c->nblk = (long long)h.sptr[h.nslice] * SLICE;
    c->tol = prm->tol;
    c->maxiter = prm->maxiter;
    c->nreg_tet = prm->nreg_tet;

#define CK(x)                      \
    do                             \
        {                          \
        int rc_ = (x);             \
        if (rc_ != FG_OK)          \
            {                      \
            fg_destroy(c);         \
            return rc_;            \
            }                      \
        } while (0)
#define CKCUDA(call)                                                                             \
    do                                                                                           \
        {                                                                                        \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            {                                                                                    \
            set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));       \
            fg_destroy(c);                                                                       \
            return FG_ERR_CUDA;                                                                  \
            }                                                                                    \
        } while (0)

    CKCUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    cudaStream_t s = c->stream;
    // Gauss tables
        {
        double a[20], p[5];
        tet_tables(5, a, p);
        CKCUDA(cudaMemcpyToSymbol(c_tet_a5, a, sizeof(double) * 20));
        CKCUDA(cudaMemcpyToSymbol(c_tet_pds5, p, sizeof(double) * 5));
        tet_tables(1, a, p);
        CKCUDA(cudaMemcpyToSymbol(c_tet_a1, a, sizeof(double) * 4));
        CKCUDA(cudaMemcpyToSymbol(c_tet_pds1, p, sizeof(double) * 1));
        tri_tables(4, a, p);
        CKCUDA(cudaMemcpyToSymbol(c_tri_a4, a, sizeof(double) * 12));
        CKCUDA(cudaMemcpyToSymbol(c_tri_pds4, p, sizeof(double) * 4));
        tri_tables(1, a, p);
        CKCUDA(cudaMemcpyToSymbol(c_tri_a1, a, sizeof(double) * 3));
        CKCUDA(cudaMemcpyToSymbol(c_tri_pds1, p, sizeof(double) * 1));
        }
    const size_t N = (size_t)c->NODt;   // node arrays are in device row order, padded to 32 (+ ghosts)
    CKCUDA(cudaMalloc(&c->cur, sizeof(NodeRec) * N));
    CKCUDA(cudaMalloc(&c->next, sizeof(NodeRec) * N));
    CKCUDA(cudaMalloc(&c->basis, sizeof(Basis) * N));
    CKCUDA(cudaMemsetAsync(c->cur, 0, sizeof(NodeRec) * N, s));
    CKCUDA(cudaMemsetAsync(c->next, 0, sizeof(NodeRec) * N, s));
    CKCUDA(cudaMemsetAsync(c->basis, 0, sizeof(Basis) * N, s));
    CKCUDA(cudaMalloc(&c->stage, sizeof(double) * 8 * N));
    CKCUDA(cudaMallocHost(&c->h_stage, sizeof(double) * 8 * N));
        {
        std::vector<unsigned char> nonmag(N, 1), dofmask(2 * N, 1);
        std::vector<double> aw(N, 0.0);
        for (size_t r = 0; r < N; r++)
            {
            const int a = h.perm[r];
            if (a < 0) continue;
            nonmag[r] = h.magNode[a] ? 0 : 1;
            dofmask[2 * r] = dofmask[2 * r + 1] = nonmag[r];
            aw[r] = h.Aw[a];
            }
        CK(dev_upload(&c->nonmag, nonmag, s));
        CK(dev_upload(&c->dofmask, dofmask, s));
        CK(dev_upload(&c->Aw, aw, s));
        CK(dev_upload(&c->perm, h.perm, s));
        CKCUDA(cudaStreamSynchronize(s));
        }
    // magnetic tets, SoA
        {
        const size_t M = (size_t)c->NTm;
        std::vector<int4> ind(M);
        std::vector<double> da(12 * M), dj(M);
        std::vector<int> reg(M);
        for (size_t tm = 0; tm < M; tm++)
            {
            const size_t t = (size_t)h.magTet[tm];
            ind[tm] = make_int4(h.tet_dev_ind[4 * tm], h.tet_dev_ind[4 * tm + 1], h.tet_dev_ind[4 * tm + 2], h.tet_dev_ind[4 * tm + 3]);
            for (int k = 0; k < 12; k++) da[(size_t)k * M + tm] = h.tet_da[12 * t + k];
            dj[tm] = h.tet_detJ[t];
            reg[tm] = h.tet_reg[t];
            }
        CK(dev_upload(&c->tet_ind, ind, s));
            {  // chunks of TET_CHUNK consecutive tetrahedra (they are sorted by their smallest device row):
               // the distinct nodes of a chunk and the 16-bit local indices of every tetrahedron in that list
            const int nch = (int)((M + TET_CHUNK - 1) / TET_CHUNK);
            std::vector<int> cptr((size_t)nch + 1, 0);
            std::vector<std::vector<int>> lists((size_t)nch);
            std::vector<ushort4> loc(M);
#pragma omp parallel for schedule(dynamic, 64)
            for (int ch = 0; ch < nch; ch++)
                {
                const size_t t0 = (size_t)ch * TET_CHUNK, t1 = std::min(M, t0 + TET_CHUNK);
                std::vector<int> &L = lists[(size_t)ch];
                L.assign(h.tet_dev_ind.begin() + 4 * t0, h.tet_dev_ind.begin() + 4 * t1);
                std::sort(L.begin(), L.end());
                L.erase(std::unique(L.begin(), L.end()), L.end());
                for (size_t tm = t0; tm < t1; tm++)
                    {
                    unsigned short q[4];
                    for (int i = 0; i < 4; i++)
                        q[i] = (unsigned short)(std::lower_bound(L.begin(), L.end(), h.tet_dev_ind[4 * tm + i]) - L.begin());
                    loc[tm] = make_ushort4(q[0], q[1], q[2], q[3]);
                    }
                }
            int cap = 1;
            for (int ch = 0; ch < nch; ch++)
                {
                cptr[(size_t)ch + 1] = cptr[(size_t)ch] + (int)lists[(size_t)ch].size();
                cap = std::max(cap, (int)lists[(size_t)ch].size());
                }
            std::vector<int> nodes((size_t)cptr[(size_t)nch]);
            for (int ch = 0; ch < nch; ch++)
                std::copy(lists[(size_t)ch].begin(), lists[(size_t)ch].end(), nodes.begin() + cptr[(size_t)ch]);
            c->tch_n = nch;
            c->tch_cap = cap;
            CK(dev_upload(&c->tch_ptr, cptr, s));
            CK(dev_upload(&c->tch_nodes, nodes, s));
            CK(dev_upload(&c->tet_loc, loc, s));
            CKCUDA(cudaStreamSynchronize(s));
            }
            {
            std::vector<int4> slot(M);
            for (size_t tm = 0; tm < M; tm++)
                slot[tm] = make_int4(h.tet_slot[4 * tm], h.tet_slot[4 * tm + 1], h.tet_slot[4 * tm + 2], h.tet_slot[4 * tm + 3]);
            CK(dev_upload(&c->tet_slot, slot, s));
            CKCUDA(cudaStreamSynchronize(s));
            }
        CK(dev_upload(&c->tet_da, da, s));
        CK(dev_upload(&c->tet_detJ, dj, s));
        CK(dev_upload(&c->tet_reg, reg, s));
        std::vector<TetRegion> regs((size_t)prm->nreg_tet);
        for (int r = 0; r < prm->nreg_tet; r++)
            {
            const fg_tet_prm &p = prm->prm_tet[r];
            TetRegion &R = regs[r];
            memset(&R, 0, sizeof R);
            R.alpha = p.alpha_LLG;
            R.A = p.A;
            R.K = p.K;
            R.K3 = p.K3;
            R.Ms = p.Ms;
            if (p.Ms > 0)
                {
                R.Abis = 2.0 * p.A / (FG_MU0 * p.Ms);   // tetra.cpp:217
                R.Kbis = 2.0 * p.K / (FG_MU0 * p.Ms);   // tetra.cpp:239
                R.K3bis = 2.0 * p.K3 / (FG_MU0 * p.Ms); // tetra.cpp:244
                }
            R.has_K = p.K != 0;
            R.has_K3 = p.K3 != 0;
            if (p.Ms > 0 && (R.has_K || R.has_K3)) c->iso_regions = false;
            if (p.Ms > 0 && R.has_K3) c->cubic_regions = true;
            for (int k = 0; k < 3; k++)
                {
                R.uk[k] = p.uk[k];
                R.ex[k] = p.ex[k];
                R.ey[k] = p.ey[k];
                R.ez[k] = p.ez[k];
                }
            }
        CK(dev_upload(&c->reg_tet, regs, s));
        // records live in the SELL incidence layout (slots without a tet stay zero for ever)
        const size_t nrec = (size_t)h.iptr[h.nslice] * SLICE;
        CKCUDA(cudaMalloc(&c->rec, sizeof(double4) * (nrec > 0 ? nrec : 1)));
        CKCUDA(cudaMemsetAsync(c->rec, 0, sizeof(double4) * (nrec > 0 ? nrec : 1), s));
        CKCUDA(cudaStreamSynchronize(s));
        }
    // active triangles
        {
        const size_t M = (size_t)c->NFa;
        std::vector<int> ind(3 * M), reg(M);
        std::vector<double> surf(M), dMs(M);
        for (size_t fa = 0; fa < M; fa++)
            {
            const size_t f = (size_t)h.actTri[fa];
            for (int i = 0; i < 3; i++) ind[(size_t)i * M + fa] = h.iperm[h.tri_ind[3 * f + i]];
            reg[fa] = h.tri_reg[f];
            surf[fa] = h.tri_surf[f];
            dMs[fa] = h.tri_dMs[f];
            }
        CK(dev_upload(&c->tri_ind, ind, s));
        CK(dev_upload(&c->tri_reg, reg, s));
        CK(dev_upload(&c->tri_surf, surf, s));
        CK(dev_upload(&c->tri_dMs, dMs, s));
        c->h_reg_tri.resize((size_t)(prm->nreg_tri > 0 ? prm->nreg_tri : 1));
        for (int r = 0; r < prm->nreg_tri; r++)
            {
            c->h_reg_tri[r].Ks = prm->prm_tri[r].Ks;
            for (int k = 0; k < 3; k++) c->h_reg_tri[r].uk[k] = prm->prm_tri[r].uk[k];
            }
        CK(dev_upload(&c->reg_tri, c->h_reg_tri, s));
        CKCUDA(cudaMalloc(&c->trec, sizeof(double2) * 3 * (M > 0 ? M : 1)));
        CKCUDA(cudaStreamSynchronize(s));
        }
    CK(dev_upload(&c->sptr, h.sptr, s));
    CK(dev_upload(&c->scol, h.scol, s));
    std::vector<unsigned char> sg;  // multi-GPU: slices that gather a ghost entry
    if (dd)
        {  // they wait for the halo, the others do not (fg_solve_pk.cuh)
        sg.assign((size_t)h.nslice, 0);
#pragma omp parallel for schedule(static)
        for (int sl = 0; sl < h.nslice; sl++)
            {
            unsigned char g = 0;
            for (size_t pos = (size_t)h.sptr[sl] * SLICE; pos < (size_t)h.sptr[sl + 1] * SLICE && !g; pos++)
                if (h.scol[pos] >= h.NODp) g = 1;
            sg[(size_t)sl] = g;
            }
        CK(dev_upload(&c->sghost, sg, s));
        }
        {  // 16-bit column offsets for the matrix-free SpMV (2 B instead of 4 B per stored pair) when every
           // neighbour of every row lies within +-32767 device rows (single-GPU meshes in the reference's
           // sorted node order).  On a partition the ghost rows sit behind the owned rows, out of reach for the
           // rows at the start of a slab: there the offsets are used for the slices WITHOUT ghost columns and
           // the few slices with ghost columns (processed in a pass of their own anyway) read the 32-bit
           // columns (scol16_ng, persistent kernel only).
        const size_t ne = h.scol.size();
        const bool want = getenv("FG_NO_COL16") == nullptr;
        std::vector<short> c16;
        if (want) c16.assign(ne, 0);
        long long bad_all = 0, bad_ng = 0;
        if (want)
            {
#pragma omp parallel for schedule(static) reduction(+ : bad_all, bad_ng)
            for (int sl = 0; sl < h.nslice; sl++)
                for (int j = h.sptr[sl]; j < h.sptr[sl + 1]; j++)
                    for (int l = 0; l < SLICE; l++)
                        {
                        const size_t pos = (size_t)j * SLICE + l;
                        const int d = h.scol[pos] - (sl * SLICE + l);
                        if (d < -32767 || d > 32767)
                            {
                            bad_all++;
                            if (sg.empty() || !sg[(size_t)sl]) bad_ng++;
                            }
                        else
                            c16[pos] = (short)d;
                        }
            }
        if (want && ne > 0 && bad_all == 0)
            CK(dev_upload(&c->scol16, c16, s));
        else if (want && ne > 0 && dd && bad_ng == 0)
            CK(dev_upload(&c->scol16_ng, c16, s));
        }
    if (h.stage_cap > 0)
        {
        CK(dev_upload(&c->lcol, h.lcol, s));
        CK(dev_upload(&c->bptr, h.bptr, s));
        CK(dev_upload(&c->bhalo, h.bhalo, s));
        if (dd) CK(dev_upload(&c->bghost, h.bghost, s));
        }
    CK(dev_upload(&c->sdeg, h.sdeg, s));
    CK(dev_upload(&c->sS, h.sS, s));
    CK(dev_upload(&c->iptr, h.iptr, s));
    CK(dev_upload(&c->itptr, h.itptr, s));
    CK(dev_upload(&c->sinct, h.sinct, s));
        {  // S_aa in device row order (Jacobi diagonal of the never-assembled K)
        std::vector<double> sd((size_t)c->NODp, 0.0);
        for (int a = 0; a < h.NOD; a++)
            {
            if (h.iperm[a] >= c->NODp) continue;  // ghost node (multi-GPU): no row here
            for (int j = h.nptr[a]; j < h.nptr[a + 1]; j++)
                if (h.ncol[j] == a) sd[(size_t)h.iperm[a]] = h.S[j];
            }
        CK(dev_upload(&c->Sdiag, sd, s));
        CKCUDA(cudaMalloc(&c->Dm, sizeof(double2) * (size_t)c->NODp));
        CKCUDA(cudaMemsetAsync(c->Dm, 0, sizeof(double2) * (size_t)c->NODp, s));
        CKCUDA(cudaMalloc(&c->qbasis, sizeof(double4) * (size_t)c->NODt));
        CKCUDA(cudaMemsetAsync(c->qbasis, 0, sizeof(double4) * (size_t)c->NODt, s));
        CKCUDA(cudaStreamSynchronize(s));
        }
    if (dd)
        {  // exchange arena: control block + the three halo-exchanged vectors, one IPC-exportable block
        const size_t vb = ((sizeof(double) * 2 * (size_t)c->NODt + 255) / 256) * 256;   // x
        const size_t wb = ((sizeof(double4) * (size_t)c->NODt + 255) / 256) * 256;      // w3p, w3s
        const size_t cb = ((sizeof(DistCtrl) + 255) / 256) * 256;
        c->arena_bytes = cb + vb + 2 * wb;
        CKCUDA(cudaMalloc(&c->arena, c->arena_bytes));
        CKCUDA(cudaMemsetAsync(c->arena, 0, c->arena_bytes, s));
        char *base = static_cast<char *>(c->arena);
        double *ext[3];
        ext[0] = reinterpret_cast<double *>(base + cb);
        ext[1] = reinterpret_cast<double *>(base + cb + vb);
        ext[2] = reinterpret_cast<double *>(base + cb + vb + wb);
        c->tail_off[0] = cb + sizeof(double) * 2 * (size_t)c->NODp;
        c->tail_off[1] = cb + vb + sizeof(double4) * (size_t)c->NODp;
        c->tail_off[2] = cb + vb + wb + sizeof(double4) * (size_t)c->NODp;
        CK(krylov_alloc(c->kw, c->np, 2 * (c->NODt - c->NODp), s, &c->launches, true, ext));
        c->kw.arena = c->arena;
        // halo plan: boundary rows in device order, grouped by destination rank
        DistDev &D = c->h_dist;
        D.rank = dd->rank;
        D.world = dd->world;
        std::vector<int> rows((size_t)dd->send_ptr[dd->world]);
        for (int q = 0; q <= dd->world; q++) D.send_ptr[q] = dd->send_ptr[q];
        for (int q = 0; q < dd->world; q++)
            {
            D.send_dst[q] = dd->send_dst[q];
            D.recv_from[q] = dd->recv_from[q];
            }
        for (size_t k = 0; k < rows.size(); k++)
            {
            const int a = dd->send_nodes[k];
            if (a < 0 || a >= h.n_owned)
                {
                set_error("fg_dist_create: send node %d is not an owned node", a);
                fg_destroy(c);
                return FG_ERR_DIST;
                }
            rows[k] = h.iperm[a];
            }
        CK(dev_upload(&c->send_rows, rows, s));
        D.send_rows = c->send_rows;
        CKCUDA(cudaMalloc(&c->d_dist, sizeof(DistDev)));
        int ng = ((int)rows.size() + BLOCK - 1) / BLOCK;
        c->kw.halo_grid = ng < 1 ? 1 : (ng > 64 ? 64 : ng);
        c->kw.nsend = (int)rows.size();
        }
    else
        CK(krylov_alloc(c->kw, c->np, 0, s, &c->launches, true));
    c->kw.mask = c->dofmask;
    c->kw.qbasis = c->qbasis;
    c->op.kind = OP_NODE3;   // K is never materialised on the step path
    c->op.n = c->np;
    c->op.lanes = 32;
    c->op.ptr = c->sptr;
    c->op.col = c->scol;
    c->op.col16 = c->scol16;
    c->op.val = c->sS;
    c->op.nslice = h.nslice;
    c->op.qbasis = c->qbasis;
    c->op.Dm = c->Dm;
    c->op.nonmag = c->nonmag;
    c->op.cS = 0.0;
    // early L2 prefetch of the SpMV row-epilogue operands; FG_SPMV_PF=0 switches it off (A/B)
    c->op.prefetch = getenv("FG_SPMV_PF") ? atoi(getenv("FG_SPMV_PF")) : 1;
    c->op.sghost = c->sghost;
    c->op.lcol = getenv("FG_NO_STAGE") ? nullptr : c->lcol;
    c->op.bptr = c->bptr;
    c->op.bhalo = c->bhalo;
    c->op.bghost = c->bghost;
    c->op.nblock = h.nblock;
    c->op.stage_cap = h.stage_cap;
        {
        const char *sv = getenv("FG_SOLVER");
        c->solver_kind = (sv && (!strcmp(sv, "multi") || !strcmp(sv, "1"))) ? 1 : 0;
        }
    for (int k = 0; k < 5; k++) CKCUDA(cudaEventCreate(&c->ev[k]));
    CKCUDA(cudaStreamSynchronize(s));
    // the per-mesh host tables no longer needed are released (their SELL images are on the device)
    std::vector<double>().swap(h.S);
    std::vector<double>().swap(h.sS);
    std::vector<int>().swap(h.inc);
    std::vector<int>().swap(h.inc_tri);
    std::vector<int>().swap(h.sinc);
    std::vector<int>().swap(h.tet_slot);
    std::vector<int>().swap(h.sinct);
    std::vector<int>().swap(h.scol);
    std::vector<unsigned short>().swap(h.lcol);
    std::vector<int>().swap(h.bhalo);
    c->sp.idx_dir = FG_IDX_UNDEF;
    *out = c;
    return FG_OK;
    }

int fg_create(const fg_mesh *mesh, const fg_params *prm, int device, fg_ctx **out)
    { return create_ctx(mesh, prm, device, nullptr, out); }

// ---- multi-GPU: one rank of the slab-partitioned solve (fg_dist.cuh) ----
int fg_dist_create(const fg_mesh *local_mesh, const fg_params *prm, int device, const fg_dist_desc *dd,
                   fg_ctx **out)
    {
    if (!dd || dd->world < 1 || dd->world > DIST_MAX_RANKS || dd->rank < 0 || dd->rank >= dd->world
        || !dd->send_ptr || !dd->send_dst || !dd->recv_from || (dd->send_ptr[dd->world] > 0 && !dd->send_nodes)
        || !local_mesh || dd->n_owned < 0 || dd->n_owned > local_mesh->NOD)
        {
        set_error("fg_dist_create: bad distribution descriptor (world 1..%d)", DIST_MAX_RANKS);
        return FG_ERR_DIST;
        }
    return create_ctx(local_mesh, prm, device, dd, out);
    }

namespace
{
struct DistBlob  // FG_DIST_BLOB_BYTES
    {
    cudaIpcMemHandle_t handle;       // 64 B
    unsigned long long tail_off[3];  // byte offsets of the ghost tails of x, w3p, w3s in the arena
    int rank, n_ghost;
    char pad_[FG_DIST_BLOB_BYTES - 64 - 24 - 8];
    };
static_assert(sizeof(DistBlob) == FG_DIST_BLOB_BYTES, "blob size");
}  // namespace

int fg_dist_export(fg_ctx *c, void *blob)
    {
    FG_TRY(check_ctx(c));
    if (!c->arena || !blob)
        {
        set_error("fg_dist_export: not a distributed context");
        return FG_ERR_DIST;
        }
    DistBlob b;
    memset(&b, 0, sizeof b);
    FG_CUDA(cudaIpcGetMemHandle(&b.handle, c->arena));
    for (int k = 0; k < 3; k++) b.tail_off[k] = c->tail_off[k];
    b.rank = c->rank;
    b.n_ghost = c->NODt - c->NODp;
    memcpy(blob, &b, sizeof b);
    return FG_OK;
    }

int fg_dist_connect(fg_ctx *c, const void *blobs)
    {
    FG_TRY(check_ctx(c));
    if (!c->arena || !blobs)
        {
        set_error("fg_dist_connect: not a distributed context");
        return FG_ERR_DIST;
        }
    const DistBlob *B = static_cast<const DistBlob *>(blobs);
    DistDev &D = c->h_dist;
    for (int q = 0; q < c->world; q++)
        {
        if (B[q].rank != q)
            {
            set_error("fg_dist_connect: blob %d belongs to rank %d", q, B[q].rank);
            return FG_ERR_DIST;
            }
        void *base = c->arena;
        if (q != c->rank)
            {
            if (!c->peer_base[q])  // every peer is mapped: the all-reduce mailboxes go to all of them
                FG_CUDA(cudaIpcOpenMemHandle(&c->peer_base[q], B[q].handle, cudaIpcMemLazyEnablePeerAccess));
            base = c->peer_base[q];
            }
        else
            c->peer_base[q] = c->arena;
        D.ctrl[q] = static_cast<DistCtrl *>(base);
        D.tail[q] = reinterpret_cast<double2 *>(static_cast<char *>(base) + B[q].tail_off[0]);
        D.wtail[0][q] = reinterpret_cast<double4 *>(static_cast<char *>(base) + B[q].tail_off[1]);
        D.wtail[1][q] = reinterpret_cast<double4 *>(static_cast<char *>(base) + B[q].tail_off[2]);
        if (D.send_ptr[q + 1] - D.send_ptr[q] + D.send_dst[q] > B[q].n_ghost)
            {
            set_error("fg_dist_connect: segment for rank %d exceeds its ghost range", q);
            return FG_ERR_DIST;
            }
        }
    D.epoch = D.hepoch = 0;
    D.error = 0;
    FG_CUDA(cudaMemcpyAsync(c->d_dist, &D, sizeof(DistDev), cudaMemcpyHostToDevice, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    c->kw.dist = c->d_dist;
    c->kw.red.dist = c->d_dist;
    c->connected = true;
    return FG_OK;
    }

void fg_destroy(fg_ctx *c)
    {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void *ptrs[] = {c->cur, c->next, c->basis, c->nonmag, c->dofmask, c->stage, c->tet_ind, c->tet_da,
                    c->tet_detJ, c->ext_field, c->tet_reg, c->reg_tet, c->rec, c->tri_ind,
                    c->tri_reg, c->tri_surf, c->tri_dMs, c->reg_tri, c->trec, c->perm, c->sptr,
                    c->scol, c->sdeg, c->iptr, c->tch_ptr, c->tch_nodes, c->tet_loc, c->tet_slot, c->itptr, c->sinct, c->sS, c->Aw, c->val, c->Sdiag, c->Dm, c->qbasis,
                    c->scol16, c->scol16_ng, c->sghost, c->d_unit, c->lcol, c->bptr, c->bhalo, c->bghost, c->mtri_ind, c->mtri_reg, c->mtri_surf, c->mtri_nrm, c->mtri_dMs, c->extra_edges,
                    c->d_scal, c->node_pos, c->corr, c->tcorr, c->src, c->cptr, c->cidx};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int q = 0; q < DIST_MAX_RANKS; q++)
        if (c->peer_base[q] && q != c->rank) cudaIpcCloseMemHandle(c->peer_base[q]);
    if (c->d_dist) cudaFree(c->d_dist);
    if (c->send_rows) cudaFree(c->send_rows);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->prof.ev)
        {
        for (int k = 0; k < 2 * c->prof.cap; k++) cudaEventDestroy(c->prof.ev[k]);
        delete[] c->prof.ev;
        delete[] c->prof.cls;
        }
    if (c->kw.st)
        {
        KState hs;
        if (cudaMemcpy(&hs, c->kw.st, sizeof(KState), cudaMemcpyDeviceToHost) == cudaSuccess && hs.hist) cudaFree(hs.hist);
        }
    krylov_free(c->kw);
    if (c->arena) cudaFree(c->arena);
    for (int k = 0; k < 5; k++)
        if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    }

int fg_get_sizes(const fg_ctx *c, long long out[10])
    {
    if (!c || !out)
        {
        set_error("fg_get_sizes: null argument");
        return FG_ERR_INVALID;
        }
    out[0] = c->h.NOD;
    out[1] = c->h.NT;
    out[2] = c->h.NF;
    out[3] = (long long)c->h.magTet.size();
    out[4] = (long long)c->h.magTri.size();
    out[5] = c->h.n_edges;
    out[6] = c->h.n_edges_mag;
    out[7] = c->n;
    out[8] = 4LL * c->nnzb;
    out[9] = (long long)c->h.lvd.size();
    return FG_OK;
    }

int fg_get_layout(const fg_ctx *c, long long out[4])
    {
    if (!c || !out)
        {
        set_error("fg_get_layout: null argument");
        return FG_ERR_INVALID;
        }
    // the persistent solver addresses staged images with 16-bit local indices whatever the global distance
    out[0] = (c->solver_kind == 0 && (pk_plan(c->op, nullptr, nullptr) || c->scol16_ng)) || c->scol16 ? 2 : 4;
    out[1] = c->nblk;
    out[2] = c->iso_regions ? 1 : 0;
    out[3] = c->NODp;
    return FG_OK;
    }

int fg_solver_launch_shape(int nslice, int n_sm, int out[4])
    {
    if (nslice < 1 || !out)
        {
        set_error("fg_solver_launch_shape: nslice >= 1 and out are required");
        return FG_ERR_INVALID;
        }
    pk_launch_shape(nslice, n_sm > 0 ? n_sm : NUM_SMS, out);
    return FG_OK;
    }

int fg_set_state(fg_ctx *c, const double *u, const double *v, const double *phi, const double *phiv)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (!u)
        {
        set_error("fg_set_state: u is required");
        return FG_ERR_INVALID;
        }
    FG_TRY(push_fields(c, c->cur, u, v, phi, phiv, true));
        {  // |u| = 1 on the magnetic nodes is a precondition (include/feellgood_b200.h): the matrix-free operator
           // applies the node-diagonal block in the closed form that holds for an orthonormal triad (ep, eq, u)
        if (!c->d_unit) FG_CUDA(cudaMalloc(&c->d_unit, sizeof(unsigned long long)));
        FG_CUDA(cudaMemsetAsync(c->d_unit, 0, sizeof(unsigned long long), c->stream));
        CTX_LAUNCH(c, k_check_unit, grid_for(c->NODt, BLOCK), c->NODt, c->nonmag, c->cur, c->d_unit);
        unsigned long long bits = 0;
        FG_CUDA(cudaMemcpyAsync(&bits, c->d_unit, sizeof bits, cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        double dev = 0.0;
        memcpy(&dev, &bits, sizeof dev);
        if (!(dev <= 1e-6))
            {
            set_error("fg_set_state: the magnetisation of a magnetic node is not a unit vector (max | |u|^2 - 1 | = %g)", dev);
            return FG_ERR_INVALID;
            }
        }
    FG_CUDA(cudaMemcpyAsync(c->next, c->cur, sizeof(NodeRec) * (size_t)c->NODt, cudaMemcpyDeviceToDevice, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    c->have_basis = c->prepared = c->assembled = false;
    return FG_OK;
    }

int fg_set_next_v(fg_ctx *c, const double *v)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (!v)
        {
        set_error("fg_set_next_v: null v");
        return FG_ERR_INVALID;
        }
    return push_fields(c, c->next, nullptr, v, nullptr, nullptr, false);
    }

int fg_set_potentials(fg_ctx *c, const double *phi, const double *phiv)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (!phi || !phiv)
        {
        set_error("fg_set_potentials: null argument");
        return FG_ERR_INVALID;
        }
    return push_fields(c, c->next, nullptr, nullptr, phi, phiv, false);
    }

int fg_get_state(fg_ctx *c, int step, double *u, double *v, double *phi, double *phiv)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (step != 0 && step != 1)
        {
        set_error("fg_get_state: step must be 0 (CURRENT) or 1 (NEXT)");
        return FG_ERR_INVALID;
        }
    const size_t N = (size_t)c->NOD;
    const int which = (u ? 1 : 0) | (v ? 2 : 0) | (phi ? 4 : 0) | (phiv ? 8 : 0);
    if (!which) return FG_OK;
    CTX_LAUNCH(c, k_unpack, grid_for(c->NODt, BLOCK), c->NODt, c->NOD, c->perm, step ? c->next : c->cur, c->stage, which);
    struct Part { double *dst; size_t off, len; };
    const Part parts[4] = {{u, 0, 3 * N}, {v, 3 * N, 3 * N}, {phi, 6 * N, N}, {phiv, 7 * N, N}};
    for (const Part &p : parts)
        if (p.dst)
            FG_CUDA(cudaMemcpyAsync(p.dst, c->stage + p.off, sizeof(double) * p.len,
                                    cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FG_OK;
    }

int fg_commit(fg_ctx *c)
    {
    FG_TRY(check_ctx(c));
    static const bool eager = getenv("FG_EAGER_COMMIT") != nullptr;  // A/B: copy now instead of in k_basis
    c->commit_pending = true;
    if (eager) FG_TRY(flush_commit(c));
    c->have_basis = c->prepared = c->assembled = false;
    return FG_OK;
    }

int fg_set_ext_space_field(fg_ctx *c, const double *field)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (!field)
        {
        set_error("fg_set_ext_space_field: null field");
        return FG_ERR_INVALID;
        }
    const int npi = c->h.npi_tet;
    const size_t M = (size_t)c->NTm, K = 3 * (size_t)npi;
    std::vector<double> soa(K * (M > 0 ? M : 1));
    for (size_t tm = 0; tm < M; tm++)
        for (size_t k = 0; k < K; k++) soa[k * M + tm] = field[K * (size_t)c->h.magTet[tm] + k];
    if (!c->ext_field) FG_CUDA(cudaMalloc(&c->ext_field, sizeof(double) * soa.size()));
    FG_CUDA(cudaMemcpyAsync(c->ext_field, soa.data(), sizeof(double) * soa.size(), cudaMemcpyHostToDevice, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FG_OK;
    }

int fg_base_projection(fg_ctx *c, double angle)
    {
    FG_TRY(check_ctx(c));
    if (c->profiling) FG_CUDA(cudaEventRecord(c->ev[0], c->stream));
    return launch_basis(c, angle);
    }

static int set_step_prm(fg_ctx *c, const double Hext[3], double A_Hext, double dt, double prefactor,
                        int idx_dir, double Vdrift, bool space)
    {
    if (!(dt > 0.0))
        {
        set_error("prepareElements: dt must be positive");
        return FG_ERR_INVALID;
        }
    if (idx_dir < FG_IDX_UNDEF || idx_dir > FG_IDX_Z)
        {
        set_error("prepareElements: bad idx_dir %d", idx_dir);
        return FG_ERR_INVALID;
        }
    if (space && !c->ext_field)
        {
        set_error("prepareElements(A_Hext): no space field set (fg_set_ext_space_field)");
        return FG_ERR_STATE;
        }
    c->sp.dt = dt;
    c->sp.prefactor = prefactor;
    for (int k = 0; k < 3; k++) c->sp.Hext[k] = Hext ? Hext[k] : 0.0;
    c->sp.A_Hext = A_Hext;
    c->sp.idx_dir = idx_dir;
    c->sp.Vdrift = Vdrift;
    c->space_field = space;
    return FG_OK;
    }

int fg_prepare_elements(fg_ctx *c, const double Hext[3], double dt, double prefactor, int idx_dir,
                        double Vdrift)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (!Hext)
        {
        set_error("fg_prepare_elements: null Hext");
        return FG_ERR_INVALID;
        }
    FG_TRY(set_step_prm(c, Hext, 0.0, dt, prefactor, idx_dir, Vdrift, false));
    if (c->profiling) FG_CUDA(cudaEventRecord(c->ev[1], c->stream));
    return launch_elements(c);
    }

int fg_prepare_elements_space(fg_ctx *c, double A_Hext, double dt, double prefactor, int idx_dir,
                              double Vdrift)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    FG_TRY(set_step_prm(c, nullptr, A_Hext, dt, prefactor, idx_dir, Vdrift, true));
    if (c->profiling) FG_CUDA(cudaEventRecord(c->ev[1], c->stream));
    return launch_elements(c);
    }

int fg_solve(fg_ctx *c, double dt, fg_step_result *out)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    return run_solve(c, dt, out);
    }

int fg_step(fg_ctx *c, double angle, const double Hext[3], double dt, double prefactor, int idx_dir,
            double Vdrift, fg_step_result *out)
    {
    FG_TRY(check_ctx(c));
    if (!Hext)
        {
        set_error("fg_step: null Hext");
        return FG_ERR_INVALID;
        }
    FG_TRY(set_step_prm(c, Hext, 0.0, dt, prefactor, idx_dir, Vdrift, false));
    if (c->profiling) FG_CUDA(cudaEventRecord(c->ev[0], c->stream));
    FG_TRY(launch_basis(c, angle));
    if (c->profiling) FG_CUDA(cudaEventRecord(c->ev[1], c->stream));
    FG_TRY(launch_elements(c));
    return run_solve(c, dt, out);
    }

// ---- energies, averages, maximum angle ----
namespace
{
// device tables of the magnetic surface triangles (msh.magTri) and of the mesh edges that have a
// non-magnetic end, built the first time an observable is asked for
int ensure_obs_tables(fg_ctx *c)
    {
    if (c->obs_ready) return FG_OK;
    const HostSetup &h = c->h;
    const size_t M = h.magTri.size();
    std::vector<int> ind(3 * M), reg(M);
    std::vector<double> surf(M), nrm(3 * M), dMs(M);
    for (size_t k = 0; k < M; k++)
        {
        const size_t f = (size_t)h.magTri[k];
        for (int i = 0; i < 3; i++)
            {
            ind[(size_t)i * M + k] = h.iperm[h.tri_ind[3 * f + i]];
            nrm[(size_t)i * M + k] = h.tri_nrm[3 * f + i];
            }
        reg[k] = h.tri_reg[f];
        surf[k] = h.tri_surf[f];
        dMs[k] = h.tri_dMs[f];
        }
    FG_TRY(dev_upload(&c->mtri_ind, ind, c->stream));
    FG_TRY(dev_upload(&c->mtri_reg, reg, c->stream));
    FG_TRY(dev_upload(&c->mtri_surf, surf, c->stream));
    FG_TRY(dev_upload(&c->mtri_nrm, nrm, c->stream));
    FG_TRY(dev_upload(&c->mtri_dMs, dMs, c->stream));
    c->NFm = (int)M;
    std::vector<int2> ex(h.extra_edges.size() / 2);
    for (size_t k = 0; k < ex.size(); k++)
        ex[k] = make_int2(h.iperm[h.extra_edges[2 * k]], h.iperm[h.extra_edges[2 * k + 1]]);
    FG_TRY(dev_upload(&c->extra_edges, ex, c->stream));
    c->n_extra = (int)ex.size();
    FG_CUDA(cudaMalloc(&c->d_scal, sizeof(double) * 8));
    FG_CUDA(cudaMallocHost(&c->h_scal, sizeof(double) * 8));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    c->obs_ready = true;
    return FG_OK;
    }

int fetch_scalars(fg_ctx *c, int count)
    {
    FG_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    return FG_OK;
    }

int obs_guard(fg_ctx *c, const char *who)
    {
    FG_TRY(check_ctx(c));
    if (c->arena && !c->connected)
        {
        set_error("%s: distributed context not connected (fg_dist_connect)", who);
        return FG_ERR_DIST;
        }
    return ensure_obs_tables(c);
    }

int run_energy(fg_ctx *c, const FieldPrm &f, bool space, double E[4])
    {
    if (!E)
        {
        set_error("fg_energy: null output");
        return FG_ERR_INVALID;
        }
    if (space && !c->ext_field)
        {
        set_error("fg_energy_space: no space field set (fg_set_ext_space_field)");
        return FG_ERR_STATE;
        }
    FG_CUDA(cudaMemsetAsync(c->d_scal, 0, sizeof(double) * 8, c->stream));
    const TetArrays A = tet_arrays(c);
    const int gt = grid_for(c->NTm, BLOCK);
    // NEXT state, like Fem::energy (src/energy.cpp:26-27)
    if (c->h.npi_tet == 5)
        {
        if (space) CTX_LAUNCH(c, (k_energy_tet<5, true>), gt, A, c->next, f, c->NODp, c->d_scal, c->kw.red);
        else CTX_LAUNCH(c, (k_energy_tet<5, false>), gt, A, c->next, f, c->NODp, c->d_scal, c->kw.red);
        }
    else
        {
        if (space) CTX_LAUNCH(c, (k_energy_tet<1, true>), gt, A, c->next, f, c->NODp, c->d_scal, c->kw.red);
        else CTX_LAUNCH(c, (k_energy_tet<1, false>), gt, A, c->next, f, c->NODp, c->d_scal, c->kw.red);
        }
    MagTriArrays F;
    F.NFm = c->NFm;
    F.ind = c->mtri_ind;
    F.surf = c->mtri_surf;
    F.nrm = c->mtri_nrm;
    F.dMs = c->mtri_dMs;
    F.reg = c->mtri_reg;
    F.regions = c->reg_tri;
    const int gf = grid_for(c->NFm, BLOCK);
    if (c->h.npi_tri == 4) CTX_LAUNCH(c, k_energy_tri<4>, gf, F, c->next, c->NODp, c->d_scal + 4, c->kw.red);
    else CTX_LAUNCH(c, k_energy_tri<1>, gf, F, c->next, c->NODp, c->d_scal + 4, c->kw.red);
    FG_TRY(fetch_scalars(c, 6));
    const double *r = c->h_scal;
    E[0] = r[0];
    E[1] = r[1] + r[4];  // volume + surface anisotropy
    E[2] = r[2] + r[5];  // volume + surface charges
    E[3] = r[3];
    return FG_OK;
    }
}  // namespace

int fg_energy(fg_ctx *c, const double Hext[3], double E[4])
    {
    FG_TRY(obs_guard(c, "fg_energy"));
    if (!Hext)
        {
        set_error("fg_energy: null Hext");
        return FG_ERR_INVALID;
        }
    FieldPrm f = {{Hext[0], Hext[1], Hext[2]}, 0.0};
    return run_energy(c, f, false, E);
    }

int fg_energy_space(fg_ctx *c, double A_Hext, double E[4])
    {
    FG_TRY(obs_guard(c, "fg_energy_space"));
    FieldPrm f = {{0.0, 0.0, 0.0}, A_Hext};
    return run_energy(c, f, true, E);
    }

int fg_avg(fg_ctx *c, int what, int region, double out[3])
    {
    FG_TRY(obs_guard(c, "fg_avg"));
    if (!out || (what != 0 && what != 1) || region < -1 || region >= c->nreg_tet)
        {
        set_error("fg_avg: bad argument (what=%d, region=%d of %d)", what, region, c->nreg_tet);
        return FG_ERR_INVALID;
        }
    const TetArrays A = tet_arrays(c);
    const int gt = grid_for(c->NTm, BLOCK);
    FG_CUDA(cudaMemsetAsync(c->d_scal, 0, sizeof(double) * 4, c->stream));
    if (c->h.npi_tet == 5) CTX_LAUNCH(c, k_avg<5>, gt, A, c->next, what, region, c->NODp, c->d_scal, c->kw.red);
    else CTX_LAUNCH(c, k_avg<1>, gt, A, c->next, what, region, c->NODp, c->d_scal, c->kw.red);
    FG_TRY(fetch_scalars(c, 4));
    // sum / volume (src/mesh.cpp:104-105).  The sum runs over magTet, the volume is the region's
    // (src/mesh.h:81-90): a non-magnetic region gives 0 / vol = 0, a region without any tet 0/0 = NaN.
    double vol = c->h_scal[3];
    if (vol == 0.0 && region >= 0)
        for (size_t t = 0; t < c->h.tet_reg.size() && vol == 0.0; t++)
            if (c->h.tet_reg[t] == region) vol = c->h.tet_detJ[t];  // any positive number: 0 / vol
    for (int d = 0; d < 3; d++) out[d] = c->h_scal[d] / vol;
    return FG_OK;
    }

int fg_max_angle(fg_ctx *c, double *angle)
    {
    FG_TRY(obs_guard(c, "fg_max_angle"));
    if (!angle)
        {
        set_error("fg_max_angle: null output");
        return FG_ERR_INVALID;
        }
    int grid = (c->h.nslice + BLOCK / 32 - 1) / (BLOCK / 32);
    if (grid > MAX_GRID) grid = MAX_GRID;
    if (grid < 1) grid = 1;
    CTX_LAUNCH(c, k_max_angle, grid, c->h.nslice, c->sptr, c->scol, c->sdeg, c->next, c->n_extra,
               c->extra_edges, c->d_scal, c->kw.red);
    FG_TRY(fetch_scalars(c, 1));
    const double min_dot = -c->h_scal[0];
    *angle = acos(min_dot);  // src/mesh.h:305
    return FG_OK;
    }

// ---- magnetic charges and the all-pairs demag potential ----
namespace
{
int ensure_demag_tables(fg_ctx *c)
    {
    FG_TRY(ensure_obs_tables(c));
    if (c->demag_ready) return FG_OK;
    const HostSetup &h = c->h;
    const size_t N = (size_t)c->NODt;
    std::vector<double> pos(3 * N, 0.0);
    for (size_t r = 0; r < N; r++)
        {
        const int a = h.perm[r];
        if (a < 0) continue;
        for (int d = 0; d < 3; d++) pos[3 * r + d] = h.node_p[3 * (size_t)a + d];
        }
    FG_TRY(dev_upload(&c->node_pos, pos, c->stream));
    // source positions = Gauss points (getPtGauss): tets in device order, then msh.magTri
    const int npi = h.npi_tet, npt = h.npi_tri;
    double a[20], pds[5], at[12], pt[4];
    tet_tables(npi, a, pds);
    tri_tables(npt, at, pt);
    const size_t M = (size_t)c->NTm, F = (size_t)c->NFm;
    c->nsrc = (long long)(M * npi + F * npt);
    std::vector<double4> src((size_t)c->nsrc);
    for (size_t tm = 0; tm < M; tm++)
        {
        const int *ind = &h.tet_ind[4 * (size_t)h.magTet[tm]];
        for (int g = 0; g < npi; g++)
            {
            double x[3] = {0, 0, 0};
            for (int d = 0; d < 3; d++)
                for (int i = 0; i < 4; i++) x[d] += h.node_p[3 * (size_t)ind[i] + d] * a[i * npi + g];
            src[tm * npi + g] = make_double4(x[0], x[1], x[2], 0.0);
            }
        }
    std::vector<int> cnt(N + 1, 0);
    for (size_t k = 0; k < F; k++)
        {
        const int *ind = &h.tri_ind[3 * (size_t)h.magTri[k]];
        for (int g = 0; g < npt; g++)
            {
            double x[3] = {0, 0, 0};
            for (int d = 0; d < 3; d++)
                for (int i = 0; i < 3; i++) x[d] += h.node_p[3 * (size_t)ind[i] + d] * at[i * npt + g];
            src[M * npi + k * npt + g] = make_double4(x[0], x[1], x[2], 0.0);
            }
        for (int i = 0; i < 3; i++) cnt[(size_t)h.iperm[ind[i]] + 1]++;
        }
    FG_TRY(dev_upload(&c->src, src, c->stream));
    // node -> (triangle, local node) incidences of msh.magTri, in triangle order
    for (size_t r = 0; r < N; r++) cnt[r + 1] += cnt[r];
    std::vector<int> cidx(3 * F), fill(cnt.begin(), cnt.end() - 1);
    for (size_t k = 0; k < F; k++)
        {
        const int *ind = &h.tri_ind[3 * (size_t)h.magTri[k]];
        for (int i = 0; i < 3; i++) cidx[(size_t)fill[h.iperm[ind[i]]]++] = (int)(3 * k + i);
        }
    FG_TRY(dev_upload(&c->cptr, cnt, c->stream));
    FG_TRY(dev_upload(&c->cidx, cidx, c->stream));
    FG_CUDA(cudaMalloc(&c->tcorr, sizeof(double) * (3 * F > 0 ? 3 * F : 1)));
    FG_CUDA(cudaMalloc(&c->corr, sizeof(double) * N));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    c->demag_ready = true;
    return FG_OK;
    }

MagTriArrays mag_tri_arrays(const fg_ctx *c)
    {
    MagTriArrays F;
    F.NFm = c->NFm;
    F.ind = c->mtri_ind;
    F.surf = c->mtri_surf;
    F.nrm = c->mtri_nrm;
    F.dMs = c->mtri_dMs;
    F.reg = c->mtri_reg;
    F.regions = c->reg_tri;
    return F;
    }

// fmm::calc_charges (src/fmm_demag.h:155-185) on the NEXT state
int launch_charges(fg_ctx *c, int which)
    {
    const TetArrays A = tet_arrays(c);
    const int gt = grid_for(c->NTm, BLOCK);
    if (c->NTm > 0)
        {
        if (c->h.npi_tet == 5) CTX_LAUNCH(c, k_charges_tet<5>, gt, A, c->next, which, c->src);
        else CTX_LAUNCH(c, k_charges_tet<1>, gt, A, c->next, which, c->src);
        }
    double4 *src_tri = c->src + (size_t)c->NTm * c->h.npi_tet;
    if (c->NFm > 0)
        {
        const MagTriArrays F = mag_tri_arrays(c);
        const int gf = grid_for(c->NFm, BLOCK);
        if (c->h.npi_tri == 4) CTX_LAUNCH(c, k_charges_tri<4>, gf, F, c->node_pos, c->next, which, src_tri, c->tcorr);
        else CTX_LAUNCH(c, k_charges_tri<1>, gf, F, c->node_pos, c->next, which, src_tri, c->tcorr);
        }
    CTX_LAUNCH(c, k_corr_gather, (c->NODt + BLOCK - 1) / BLOCK, c->NODt, c->cptr, c->cidx, c->tcorr, c->corr);
    return FG_OK;
    }

int demag_guard(fg_ctx *c, const char *who)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));  // fg_demag_direct writes the potentials of NEXT
    if (c->arena)
        {
        set_error("%s: not available on a distributed context (single-GPU stand-in for the FMM)", who);
        return FG_ERR_STATE;
        }
    return ensure_demag_tables(c);
    }
}  // namespace

int fg_calc_charges(fg_ctx *c, int which, double *srcDen, double *corr)
    {
    FG_TRY(demag_guard(c, "fg_calc_charges"));
    if (which != 0 && which != 1)
        {
        set_error("fg_calc_charges: which must be 0 (u) or 1 (v)");
        return FG_ERR_INVALID;
        }
    FG_TRY(launch_charges(c, which));
    const HostSetup &h = c->h;
    if (srcDen)
        {
        std::vector<double4> src((size_t)c->nsrc);
        FG_CUDA(cudaMemcpyAsync(src.data(), c->src, sizeof(double4) * src.size(), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        // reference order (src/fmm_demag.h:160-170): magTet in ascending tet index, then magTri
        const int npi = h.npi_tet;
        std::vector<int> rank((size_t)h.NT, -1);
        int r = 0;
        for (int t = 0; t < h.NT; t++)
            if (h.tet_to_mag[t] >= 0) rank[t] = r++;
        for (size_t tm = 0; tm < (size_t)c->NTm; tm++)
            for (int g = 0; g < npi; g++) srcDen[(size_t)rank[h.magTet[tm]] * npi + g] = src[tm * npi + g].w;
        for (size_t k = (size_t)c->NTm * npi; k < src.size(); k++) srcDen[k] = src[k].w;
        }
    if (corr)
        {
        std::vector<double> tmp((size_t)c->NODt);
        FG_CUDA(cudaMemcpyAsync(tmp.data(), c->corr, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        for (int a = 0; a < c->NOD; a++) corr[a] = tmp[(size_t)h.iperm[a]];
        }
    return FG_OK;
    }

int fg_demag_direct(fg_ctx *c, int second_order)
    {
    FG_TRY(demag_guard(c, "fg_demag_direct"));
    for (int which = 0; which < (second_order ? 2 : 1); which++)
        {
        FG_TRY(launch_charges(c, which));
        CTX_LAUNCH(c, k_demag_direct, (c->NODt + BLOCK - 1) / BLOCK, c->NODt, c->nonmag, c->node_pos, c->nsrc,
                   c->src, c->corr, which, c->next);
        }
    return FG_OK;
    }

// ---- taps ----
int fg_get_basis(fg_ctx *c, double *ep, double *eq)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    const size_t N = (size_t)c->NODp;
    std::vector<Basis> b(N);
    FG_CUDA(cudaMemcpyAsync(b.data(), c->basis, sizeof(Basis) * N, cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    for (int a = 0; a < c->NOD; a++)
        {
        const Basis &q = b[(size_t)c->h.iperm[a]];
        for (int k = 0; k < 3; k++)
            {
            if (ep) ep[3 * (size_t)a + k] = q.ep[k];
            if (eq) eq[3 * (size_t)a + k] = q.eq[k];
            }
        }
    return FG_OK;
    }

int fg_get_elements(fg_ctx *c, int first, int count, double *Kp, double *Lp)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (first < 0 || count < 0 || first + count > c->h.NT || !Kp || !Lp)
        {
        set_error("fg_get_elements: bad range [%d,%d) of %d", first, first + count, c->h.NT);
        return FG_ERR_INVALID;
        }
    if (!c->prepared && !c->assembled)
        {
        set_error("fg_get_elements: no prepareElements since the last state change");
        return FG_ERR_STATE;
        }
    memset(Kp, 0, sizeof(double) * 64 * (size_t)count);
    memset(Lp, 0, sizeof(double) * 8 * (size_t)count);
    std::vector<int> list, where;
    for (int t = first; t < first + count; t++)
        if (c->h.tet_to_mag[t] >= 0)
            {
            list.push_back(c->h.tet_to_mag[t]);
            where.push_back(t - first);
            }
    const int cnt = (int)list.size();
    if (cnt == 0) return FG_OK;
    double *dK = nullptr, *dL = nullptr;
    int *dlist = nullptr;
    FG_CUDA(cudaMalloc(&dK, sizeof(double) * 64 * (size_t)cnt));
    FG_CUDA(cudaMalloc(&dL, sizeof(double) * 8 * (size_t)cnt));
    FG_CUDA(cudaMalloc(&dlist, sizeof(int) * (size_t)cnt));
    FG_CUDA(cudaMemcpyAsync(dlist, list.data(), sizeof(int) * (size_t)cnt, cudaMemcpyHostToDevice, c->stream));
    const TetArrays A = tet_arrays(c);
    const int grid = (cnt + BLOCK - 1) / BLOCK;
    if (use_iso(c))
        {
        if (c->h.npi_tet == 5) k_tet_tap<5, false, true><<<grid, BLOCK, 0, c->stream>>>(A, c->cur, c->basis, c->sp, dlist, cnt, dK, dL);
        else k_tet_tap<1, false, true><<<grid, BLOCK, 0, c->stream>>>(A, c->cur, c->basis, c->sp, dlist, cnt, dK, dL);
        }
    else if (c->h.npi_tet == 5)
        {
        if (c->space_field) k_tet_tap<5, true><<<grid, BLOCK, 0, c->stream>>>(A, c->cur, c->basis, c->sp, dlist, cnt, dK, dL);
        else k_tet_tap<5, false><<<grid, BLOCK, 0, c->stream>>>(A, c->cur, c->basis, c->sp, dlist, cnt, dK, dL);
        }
    else
        {
        if (c->space_field) k_tet_tap<1, true><<<grid, BLOCK, 0, c->stream>>>(A, c->cur, c->basis, c->sp, dlist, cnt, dK, dL);
        else k_tet_tap<1, false><<<grid, BLOCK, 0, c->stream>>>(A, c->cur, c->basis, c->sp, dlist, cnt, dK, dL);
        }
    ++c->launches;
    std::vector<double> hK(64 * (size_t)cnt), hL(8 * (size_t)cnt);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(hK.data(), dK, sizeof(double) * hK.size(), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hL.data(), dL, sizeof(double) * hL.size(), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dK);
    cudaFree(dL);
    cudaFree(dlist);
    if (e != cudaSuccess)
        {
        set_error("fg_get_elements: %s", cudaGetErrorString(e));
        return FG_ERR_CUDA;
        }
    for (int q = 0; q < cnt; q++)
        {
        memcpy(Kp + 64 * (size_t)where[q], &hK[64 * (size_t)q], sizeof(double) * 64);
        memcpy(Lp + 8 * (size_t)where[q], &hL[8 * (size_t)q], sizeof(double) * 8);
        }
    return FG_OK;
    }

int fg_get_records(fg_ctx *c, int first, int count, double *rec_out)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (first < 0 || count < 0 || first + count > c->h.NT || !rec_out)
        {
        set_error("fg_get_records: bad range [%d,%d) of %d", first, first + count, c->h.NT);
        return FG_ERR_INVALID;
        }
    if (!c->prepared && !c->assembled)
        {
        set_error("fg_get_records: no prepareElements since the last state change");
        return FG_ERR_STATE;
        }
    memset(rec_out, 0, sizeof(double) * 16 * (size_t)count);
    FG_CUDA(cudaStreamSynchronize(c->stream));
    for (int t = first; t < first + count; t++)
        {
        const int tm = c->h.tet_to_mag[t];
        if (tm < 0) continue;
        int4 sl;
        FG_CUDA(cudaMemcpy(&sl, c->tet_slot + tm, sizeof(int4), cudaMemcpyDeviceToHost));
        const int slot[4] = {sl.x, sl.y, sl.z, sl.w};
        for (int i = 0; i < 4; i++)
            if (slot[i] >= 0)
                FG_CUDA(cudaMemcpy(rec_out + 16 * (size_t)(t - first) + 4 * i, c->rec + slot[i], sizeof(double4),
                                   cudaMemcpyDeviceToHost));
        }
    return FG_OK;
    }

int fg_get_tri_elements(fg_ctx *c, int first, int count, double *Lp)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (first < 0 || count < 0 || first + count > c->h.NF || !Lp)
        {
        set_error("fg_get_tri_elements: bad range");
        return FG_ERR_INVALID;
        }
    if (!c->have_basis)
        {
        set_error("fg_get_tri_elements: no basis");
        return FG_ERR_STATE;
        }
    memset(Lp, 0, sizeof(double) * 6 * (size_t)count);
    // Tri::integrales runs for magnetic triangles with Ks != 0 (src/linear_algebra.cpp:45-51),
    // whether or not their Lp later reaches the rhs
    std::vector<int> sel;
    for (int f = first; f < first + count; f++)
        {
        const int *ind = &c->h.tri_ind[3 * (size_t)f];
        const bool mag = c->h.magNode[ind[0]] && c->h.magNode[ind[1]] && c->h.magNode[ind[2]];
        if (mag && c->h_reg_tri[c->h.tri_reg[f]].Ks != 0) sel.push_back(f);
        }
    if (sel.empty()) return FG_OK;
    const size_t M = sel.size();
    std::vector<int> ind(3 * M), reg(M);
    std::vector<double> surf(M), dMs(M);
    for (size_t k = 0; k < M; k++)
        {
        const size_t f = (size_t)sel[k];
        for (int i = 0; i < 3; i++) ind[(size_t)i * M + k] = c->h.iperm[c->h.tri_ind[3 * f + i]];
        reg[k] = c->h.tri_reg[f];
        surf[k] = c->h.tri_surf[f];
        dMs[k] = c->h.tri_dMs[f];
        }
    int *d_ind = nullptr, *d_reg = nullptr;
    double *d_surf = nullptr, *d_dMs = nullptr;
    double2 *d_rec = nullptr;
    int rc = FG_OK;
    std::vector<double2> hrec(3 * M);
    do
        {
        if ((rc = dev_upload(&d_ind, ind, c->stream)) != FG_OK) break;
        if ((rc = dev_upload(&d_reg, reg, c->stream)) != FG_OK) break;
        if ((rc = dev_upload(&d_surf, surf, c->stream)) != FG_OK) break;
        if ((rc = dev_upload(&d_dMs, dMs, c->stream)) != FG_OK) break;
        if (cudaMalloc(&d_rec, sizeof(double2) * 3 * M) != cudaSuccess) { rc = FG_ERR_CUDA; break; }
        TriArrays F;
        F.NFa = (int)M;
        F.ind = d_ind;
        F.surf = d_surf;
        F.dMs = d_dMs;
        F.reg = d_reg;
        F.regions = c->reg_tri;
        const int grid = grid_for((long long)M, BLOCK);
        if (c->h.npi_tri == 4) k_tri<4><<<grid, BLOCK, 0, c->stream>>>(F, c->cur, c->basis, d_rec);
        else k_tri<1><<<grid, BLOCK, 0, c->stream>>>(F, c->cur, c->basis, d_rec);
        ++c->launches;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(hrec.data(), d_rec, sizeof(double2) * 3 * M, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess)
            {
            set_error("fg_get_tri_elements: %s", cudaGetErrorString(e));
            rc = FG_ERR_CUDA;
            }
        } while (0);
    cudaFree(d_ind);
    cudaFree(d_reg);
    cudaFree(d_surf);
    cudaFree(d_dMs);
    cudaFree(d_rec);
    if (rc != FG_OK) return rc;
    for (size_t k = 0; k < M; k++)
        for (int i = 0; i < 3; i++)
            {  // Perm = {3,4,5,0,1,2} (triangle.cpp:30-34): rows 0..2 eq-projections, 3..5 ep
            Lp[6 * (size_t)(sel[k] - first) + i] = hrec[3 * k + i].x;
            Lp[6 * (size_t)(sel[k] - first) + 3 + i] = hrec[3 * k + i].y;
            }
    return FG_OK;
    }

int fg_get_csr_pattern(const fg_ctx *c, int *rowptr, int *col)
    {
    if (!c || !rowptr || !col)
        {
        set_error("fg_get_csr_pattern: null argument");
        return FG_ERR_INVALID;
        }
    const HostSetup &h = c->h;
    rowptr[0] = 0;
    for (int a = 0; a < h.NOD; a++)
        {
        const int beg = h.nptr[a], deg = h.nptr[a + 1] - beg;
        for (int k = 0; k < 2; k++)
            {
            const int base = 4 * beg + 2 * deg * k;
            for (int j = 0; j < deg; j++)
                {
                col[base + 2 * j] = 2 * h.ncol[beg + j];
                col[base + 2 * j + 1] = 2 * h.ncol[beg + j] + 1;
                }
            rowptr[2 * a + k + 1] = base + 2 * deg;
            }
        }
    return FG_OK;
    }

int fg_get_system(fg_ctx *c, double dt, double *val, double *rhs, double *x0)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (c->arena)
        {
        set_error("fg_get_system: not available on a distributed context");
        return FG_ERR_STATE;
        }
    // before the solve: assemble now; after it: K and L_rhs are still resident (x0 then holds Xw)
    if (c->prepared || !c->assembled) FG_TRY(launch_assemble(c, dt));
    const HostSetup &h = c->h;
    if (val)
        {  // SELL-32 2x2 blocks -> the reference's CSR order (rows 2a, 2a+1 of node a, sorted columns)
        if (!c->use_blocks) FG_TRY(launch_assemble_K(c));
        std::vector<double> sv(4 * (size_t)c->nblk);
        FG_CUDA(cudaMemcpyAsync(sv.data(), c->val, sizeof(double) * sv.size(), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        for (int a = 0; a < h.NOD; a++)
            {
            const int r = h.iperm[a], sl = r / SLICE, lane = r % SLICE;
            const int beg = h.nptr[a], deg = h.nptr[a + 1] - beg;
            double *row0 = val + 4 * (size_t)beg, *row1 = row0 + 2 * (size_t)deg;
            for (int j = 0; j < deg; j++)
                {
                const size_t p0 = (((size_t)h.sptr[sl] + j) * 2) * SLICE + lane;  // double2 index
                row0[2 * j] = sv[2 * p0];
                row0[2 * j + 1] = sv[2 * p0 + 1];
                row1[2 * j] = sv[2 * (p0 + SLICE)];
                row1[2 * j + 1] = sv[2 * (p0 + SLICE) + 1];
                }
            }
        }
    std::vector<double> tmp(2 * (size_t)c->NODt);
    if (rhs)
        {
        FG_CUDA(cudaMemcpyAsync(tmp.data(), c->kw.b, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        dofs_to_caller(c, tmp, rhs);
        }
    if (x0)
        {
        FG_CUDA(cudaMemcpyAsync(tmp.data(), c->kw.x, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
        FG_CUDA(cudaStreamSynchronize(c->stream));
        dofs_to_caller(c, tmp, x0);
        }
    return FG_OK;
    }

int fg_get_precond(fg_ctx *c, double dt, double *D)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (c->arena || !D)
        {
        set_error("fg_get_precond: null argument or distributed context");
        return FG_ERR_STATE;
        }
    if (c->prepared || !c->assembled) FG_TRY(launch_assemble(c, dt));
    std::vector<double> tmp(2 * (size_t)c->NODt);
    FG_CUDA(cudaMemcpyAsync(tmp.data(), c->kw.D, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    dofs_to_caller(c, tmp, D);
    return FG_OK;
    }

int fg_apply_operator(fg_ctx *c, const double *x, double *y)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (!x || !y)
        {
        set_error("fg_apply_operator: null argument");
        return FG_ERR_INVALID;
        }
    if (!c->assembled || c->arena)
        {
        set_error("fg_apply_operator: no assembled system (or distributed context)");
        return FG_ERR_STATE;
        }
    std::vector<double> tmp;
    dofs_to_device(c, x, tmp);
    FG_CUDA(cudaMemcpyAsync(c->kw.phat, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice, c->stream));
    FG_TRY(spmv(c->op, c->kw, c->kw.phat, c->kw.v, false));
    FG_CUDA(cudaMemcpyAsync(tmp.data(), c->kw.v, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    dofs_to_caller(c, tmp, y);
    return FG_OK;
    }

int fg_get_solution(fg_ctx *c, double *Xw)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (!Xw)
        {
        set_error("fg_get_solution: null argument");
        return FG_ERR_INVALID;
        }
    std::vector<double> tmp(2 * (size_t)c->NODt);
    FG_CUDA(cudaMemcpyAsync(tmp.data(), c->kw.x, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
    FG_CUDA(cudaStreamSynchronize(c->stream));
    dofs_to_caller(c, tmp, Xw);
    return FG_OK;
    }

int fg_get_tet_tables(const fg_ctx *c, int *ind, double *da, double *weight)
    {
    if (!c)
        {
        set_error("fg_get_tet_tables: null context");
        return FG_ERR_INVALID;
        }
    const HostSetup &h = c->h;
    if (ind) memcpy(ind, h.tet_ind.data(), sizeof(int) * h.tet_ind.size());
    if (da) memcpy(da, h.tet_da.data(), sizeof(double) * h.tet_da.size());
    if (weight)
        {
        double a[20], p[5];
        tet_tables(h.npi_tet, a, p);
        for (size_t t = 0; t < (size_t)h.NT; t++)
            for (int g = 0; g < h.npi_tet; g++) weight[(size_t)h.npi_tet * t + g] = h.tet_detJ[t] * p[g];
        }
    return FG_OK;
    }

int fg_host_plan(const fg_mesh *mesh, const fg_params *prm, long long out[12])
    {
    if (!mesh || !prm || !out)
        {
        set_error("fg_host_plan: null argument");
        return FG_ERR_INVALID;
        }
    HostSetup h;
    std::string err;
    const int rc = host_setup(*mesh, *prm, -1, h, err);
    if (rc != FG_OK)
        {
        set_error("%s", err.c_str());
        return rc;
        }
    // the device layout must be a faithful image of the reference pattern
    long long bad = 0;
    std::vector<char> seen((size_t)h.NODp, 0);
    for (int r = 0; r < h.NODp; r++)
        {
        const int a = h.perm[r];
        if (a < 0) continue;
        if (a >= h.NOD || h.iperm[a] != r || seen[a]) bad++;
        else seen[a] = 1;
        if (r / SELL_WINDOW != a / SELL_WINDOW) bad++;  // rows only move inside their window
        }
    for (int a = 0; a < h.NOD; a++)
        if (!seen[a]) bad++;
    int wmax = 0;
    for (int s = 0; s < h.nslice; s++)
        {
        const int w = h.sptr[s + 1] - h.sptr[s];
        wmax = w > wmax ? w : wmax;
        for (int l = 0; l < SELL_C; l++)
            {
            const int r = s * SELL_C + l, a = h.perm[r];
            const int deg = a < 0 ? 0 : h.nptr[a + 1] - h.nptr[a];
            if (deg > w || h.sdeg[r] != deg) bad++;
            for (int j = 0; j < w; j++)
                {
                const size_t pos = ((size_t)h.sptr[s] + j) * SELL_C + l;
                if (j < deg)
                    {
                    if (h.perm[h.scol[pos]] != h.ncol[(size_t)h.nptr[a] + j]) bad++;
                    if (h.sS[pos] != h.S[(size_t)h.nptr[a] + j]) bad++;
                    }
                else if (h.scol[pos] != r || h.sS[pos] != 0.0)
                    bad++;
                }
            const int wi = h.iptr[s + 1] - h.iptr[s];
            const int ni = a < 0 ? 0 : h.inc_ptr[a + 1] - h.inc_ptr[a];
            for (int q = 0; q < wi; q++)
                {
                const int v = h.sinc[((size_t)h.iptr[s] + q) * SELL_C + l];
                if (q < ni ? v != h.inc[(size_t)h.inc_ptr[a] + q] : v != -1) bad++;
                }
            }
        }
    for (size_t tm = 0; tm < h.magTet.size(); tm++)
        for (int i = 0; i < 4; i++)
            {
            if (h.perm[h.tet_dev_ind[4 * tm + i]] != h.tet_ind[4 * (size_t)h.magTet[tm] + i]) bad++;
            const int sl = h.tet_slot[4 * tm + i];
            if (sl < 0 || h.sinc[(size_t)sl] != (int)(4 * tm + i)) bad++;
            else if (sl % SELL_C != h.tet_dev_ind[4 * tm + i] % SELL_C) bad++;
            }
    if (bad)
        {
        set_error("fg_host_plan: device layout inconsistent with the reference pattern (%lld defects)", bad);
        return FG_ERR_STATE;
        }
    out[0] = h.NOD;
    out[1] = h.NODp;
    out[2] = h.nslice;
    out[3] = h.nptr[h.NOD];
    out[4] = (long long)h.sptr[h.nslice] * SELL_C;
    out[5] = (long long)h.magTet.size();
    out[6] = wmax;
    out[7] = (long long)h.iptr[h.nslice] * SELL_C;
    out[8] = 4LL * (long long)h.magTet.size();
    out[9] = (long long)h.lvd.size();
    out[10] = h.n_edges;
    out[11] = h.n_edges_mag;
    return FG_OK;
    }

// ---- instrumentation ----
long long fg_kernel_launches(const fg_ctx *c) { return c ? c->launches : 0; }
void *fg_stream(const fg_ctx *c) { return c ? (void *)c->stream : nullptr; }

int fg_set_profiling(fg_ctx *c, int on)
    {
    FG_TRY(check_ctx(c));
    c->profiling = on == 1 ? 1 : 0;
    if (on == 2 || on == 3)
        {
        if (!c->prof.ev)
            {
            c->prof.cap = 8192;
            c->prof.ev = new cudaEvent_t[2 * c->prof.cap];
            c->prof.cls = new int[c->prof.cap];
            for (int k = 0; k < 2 * c->prof.cap; k++) FG_CUDA(cudaEventCreate(&c->prof.ev[k]));
            }
        c->prof.n = 0;
        c->prof.mode = on;
        c->kw.prof = &c->prof;
        c->kw.pk_stamps_on = 1;
        if (c->kw.pk_phase_acc)
            FG_CUDA(cudaMemsetAsync(c->kw.pk_phase_acc, 0, sizeof(unsigned long long) * 64, c->stream));
        }
    else
        {
        c->kw.prof = nullptr;
        c->kw.pk_stamps_on = 0;
        }
    return FG_OK;
    }

int fg_set_solver(fg_ctx *c, int kind)
    {
    FG_TRY(check_ctx(c));
    if (kind != 0 && kind != 1)
        {
        set_error("fg_set_solver: kind must be 0 (persistent kernel) or 1 (one kernel per phase)");
        return FG_ERR_INVALID;
        }
    c->solver_kind = kind;
    return FG_OK;
    }

int fg_get_solve_times(fg_ctx *c, double ms[27], long long count[9])
    {
    FG_TRY(check_ctx(c));
    if (!ms || !count)
        {
        set_error("fg_get_solve_times: null argument");
        return FG_ERR_INVALID;
        }
    FG_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 27; k++) ms[k] = 0.0;
    for (int k = 0; k < 9; k++) count[k] = 0;
    for (int k = 0; k < c->prof.n; k++)
        {
        if (c->prof.cls[k] != KC_SOLVE) continue;
        float t = 0.f;
        FG_CUDA(cudaEventElapsedTime(&t, c->prof.ev[2 * k], c->prof.ev[2 * k + 1]));
        ms[0] += t;
        count[0]++;
        }
    if (c->kw.pk_phase_acc)
        {
        unsigned long long acc[64];
        FG_CUDA(cudaMemcpy(acc, c->kw.pk_phase_acc, sizeof acc, cudaMemcpyDeviceToHost));
        for (int id = PKP_SETUP; id <= PKP_UPDATE; id++)
            {
            ms[id - PKP_SETUP + 1] = 1e-6 * (double)acc[id];
            count[id - PKP_SETUP + 1] = (long long)acc[16 + id];
            ms[9 + id - PKP_SETUP + 1] = 1e-6 * (double)acc[32 + id];   // until the last CTA arrived
            ms[18 + id - PKP_SETUP + 1] = 1e-6 * (double)acc[48 + id];  // cross-GPU part of the barrier
            }
        }
    return FG_OK;
    }

int fg_get_spmv_times(fg_ctx *c, double *total_ms, int *launches)
    {
    FG_TRY(check_ctx(c));
    if (!total_ms || !launches)
        {
        set_error("fg_get_spmv_times: null argument");
        return FG_ERR_INVALID;
        }
    FG_CUDA(cudaStreamSynchronize(c->stream));
    double sum = 0.0;
    int cnt = 0;
    for (int k = 0; k < c->prof.n; k++)
        {
        if (!prof_is_spmv(c->prof.cls[k])) continue;
        float ms = 0.f;
        FG_CUDA(cudaEventElapsedTime(&ms, c->prof.ev[2 * k], c->prof.ev[2 * k + 1]));
        sum += ms;
        cnt++;
        }
    *total_ms = sum;
    *launches = cnt;
    c->prof.n = 0;
    return FG_OK;
    }

int fg_get_kernel_times(fg_ctx *c, double ms_out[FG_KERNEL_CLASSES], int launches_out[FG_KERNEL_CLASSES])
    {
    FG_TRY(check_ctx(c));
    if (!ms_out || !launches_out)
        {
        set_error("fg_get_kernel_times: null argument");
        return FG_ERR_INVALID;
        }
    static_assert(FG_KERNEL_CLASSES == KC_COUNT + 1, "kernel class table");
    FG_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < FG_KERNEL_CLASSES; k++)
        {
        ms_out[k] = 0.0;
        launches_out[k] = 0;
        }
    for (int k = 0; k < c->prof.n; k++)
        {
        float ms = 0.f;
        FG_CUDA(cudaEventElapsedTime(&ms, c->prof.ev[2 * k], c->prof.ev[2 * k + 1]));
        ms_out[c->prof.cls[k]] += ms;
        launches_out[c->prof.cls[k]]++;
        if (k + 1 < c->prof.n)
            {  // time between the end of this kernel and the start of the next one in the record
            FG_CUDA(cudaEventElapsedTime(&ms, c->prof.ev[2 * k + 1], c->prof.ev[2 * k + 2]));
            if (ms < 1.0f)  // skip the host-side pauses between steps
                {
                ms_out[KC_COUNT] += ms;
                launches_out[KC_COUNT]++;
                }
            }
        }
    return FG_OK;
    }

int fg_set_operator(fg_ctx *c, int kind)
    {
    FG_TRY(check_ctx(c));
    if (kind != 0 && kind != 1)
        {
        set_error("fg_set_operator: kind must be 0 (matrix-free) or 1 (assembled 2x2 blocks)");
        return FG_ERR_INVALID;
        }
    if (kind == 1 && c->arena)
        {
        set_error("fg_set_operator: the assembled-block operator is single-GPU only");
        return FG_ERR_STATE;
        }
    c->use_blocks = kind == 1;
    c->assembled = false;
    return FG_OK;
    }

int fg_get_krylov_history(fg_ctx *c, int rows, double *out)
    {
    FG_TRY(check_ctx(c));
    if (rows < 0 || (rows > 0 && !out))
        {
        set_error("fg_get_krylov_history: bad argument");
        return FG_ERR_INVALID;
        }
    static const int CAP = 1024;
    KState hs;
    FG_CUDA(cudaMemcpy(&hs, c->kw.st, sizeof(KState), cudaMemcpyDeviceToHost));
    if (!hs.hist)
        {  // first call switches the recording on; the NEXT solves are recorded
        double *buf = nullptr;
        FG_CUDA(cudaMalloc(&buf, sizeof(double) * 8 * CAP));
        FG_CUDA(cudaMemset(buf, 0, sizeof(double) * 8 * CAP));
        hs.hist = buf;
        hs.hist_cap = CAP;
        FG_CUDA(cudaMemcpy(c->kw.st, &hs, sizeof(KState), cudaMemcpyHostToDevice));
        if (rows > 0) memset(out, 0, sizeof(double) * 8 * (size_t)rows);
        return FG_OK;
        }
    const int nr = rows < CAP ? rows : CAP;
    if (nr > 0) FG_CUDA(cudaMemcpy(out, hs.hist, sizeof(double) * 8 * (size_t)nr, cudaMemcpyDeviceToHost));
    return FG_OK;
    }

int fg_get_krylov_state(const fg_ctx *c, double out[8])
    {
    if (!c || !out)
        {
        set_error("fg_get_krylov_state: null argument");
        return FG_ERR_INVALID;
        }
    const KState &st = *c->kw.h_st;
    out[0] = st.rho1;
    out[1] = st.rho2;
    out[2] = st.alpha;
    out[3] = st.omega;
    out[4] = st.res;
    out[5] = st.rhsn;
    out[6] = st.nit;
    out[7] = st.status;
    return FG_OK;
    }

int fg_get_phase_times(const fg_ctx *c, double out[8])
    {
    if (!c || !out)
        {
        set_error("fg_get_phase_times: null argument");
        return FG_ERR_INVALID;
        }
    for (int k = 0; k < 8; k++) out[k] = c->phase_ms[k];
    return FG_OK;
    }

int fg_bench_spmv(fg_ctx *c, int reps, double *ms_per_launch)
    {
    FG_TRY(check_ctx(c));
    FG_TRY(flush_commit(c));
    if (reps <= 0 || !ms_per_launch)
        {
        set_error("fg_bench_spmv: bad argument");
        return FG_ERR_INVALID;
        }
    if (!c->assembled)
        {
        set_error("fg_bench_spmv: no assembled system");
        return FG_ERR_STATE;
        }
    for (int k = 0; k < 3; k++) FG_TRY(spmv(c->op, c->kw, c->kw.x, c->kw.t, true, k == 0));
    FG_CUDA(cudaEventRecord(c->ev[0], c->stream));
    for (int k = 0; k < reps; k++) FG_TRY(spmv(c->op, c->kw, c->kw.x, c->kw.t, true, false));
    FG_CUDA(cudaEventRecord(c->ev[1], c->stream));
    FG_CUDA(cudaEventSynchronize(c->ev[1]));
    float ms = 0.f;
    FG_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    *ms_per_launch = (double)ms / reps;
    return FG_OK;
    }

}  // extern "C"
