/* feellgood_b200.h — C ABI of the B200-native LLG hot path (libfeellgood_b200.so).
 *
 * This is the drop-in boundary for FeeLLGood's per-time-step path: every entry point replaces
 * one call of the reference's C++ class surface (file:line are relative to the reference tree,
 * feellgood/FeeLLGood).  The reference has no FFI of its own; the seam is what
 * Fem::time_integration calls on LinAlgebra (src/time_integration.cpp:193-206) plus the algebra::
 * free functions (src/algebra/{bicg,cg}.h) — see SURVEY.md §8(b) and INTEGRATION.md for the C++
 * shim (feellgood_b200/host/) that keeps the reference's class names on top of these symbols.
 *
 * Conventions
 *   - plain C: pointers and sizes only.  All array arguments are HOST pointers (pageable or
 *     pinned; pinned memory makes the copies asynchronous) unless the name ends in _dev.
 *   - every function returns 0 on success, a negative FG_ERR_* otherwise; fg_last_error() gives
 *     the message.  No exceptions cross the boundary, nothing calls exit().
 *   - a context is single-threaded and bound to one CUDA device / one stream, like the
 *     reference's LinAlgebra object is driven from the main thread only.
 *   - there is NO CPU fallback: if no CUDA device is usable fg_create fails.
 *   - FP64 throughout, 32-bit indices.  Node vectors are (NOD x 3) row-major.
 */
#ifndef FEELLGOOD_B200_H
#define FEELLGOOD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FG_OK 0
#define FG_ERR_INVALID -1   /* bad argument */
#define FG_ERR_CUDA -2      /* CUDA runtime error (message in fg_last_error) */
#define FG_ERR_MESH -3      /* singular tetrahedron, index out of range */
#define FG_ERR_STATE -4     /* call sequence error (e.g. solve before prepareElements) */
#define FG_ERR_DIST -5      /* multi-GPU setup error */

/* algebra::algoStatus, src/algebra/iter.h:24-30 */
#define FG_UNDEFINED -1
#define FG_CONVERGED 0
#define FG_ITER_OVERFLOW 1
#define FG_CANNOT_CONVERGE 2

/* Nodes::index, src/node.h:35-41 (recentring direction of add_drift_BE) */
#define FG_IDX_UNDEF -1
#define FG_IDX_X 0
#define FG_IDX_Y 1
#define FG_IDX_Z 2

/* Tetra::prm fields read by the hot path, src/tetra.h:86-112 */
typedef struct fg_tet_prm
    {
    double alpha_LLG, A, Ms, K;
    double uk[3];
    double K3;
    double ex[3], ey[3], ez[3];
    } fg_tet_prm;

/* Triangle::prm fields read by the hot path, src/triangle.h:69-82 */
typedef struct fg_tri_prm
    {
    double Ks;
    double uk[3];
    int suppress_charges;
    int pad_;
    } fg_tri_prm;

/* What Mesh::mesh hands to LinAlgebra (src/mesh.h:38-135): nodes already scaled and sorted
 * (src/mesh.cpp:334-367), zero-based connectivity, per-element region index into the parameter
 * arrays, dMs per surface triangle (src/mesh.cpp:227-242). */
typedef struct fg_mesh
    {
    int NOD;
    const double *node_p; /* NOD x 3 */
    int NT;
    const int *tet_ind; /* NT x 4 ; re-oriented internally like Tet::orientate */
    const int *tet_reg; /* NT */
    int NF;
    const int *tri_ind; /* NF x 3 */
    const int *tri_reg; /* NF */
    const double *tri_dMs; /* NF */
    } fg_mesh;

/* Settings fields that cross into the hot path (src/settings.h; default-settings.yml:229-232) */
typedef struct fg_params
    {
    int nreg_tet;
    const fg_tet_prm *prm_tet;
    int nreg_tri;
    const fg_tri_prm *prm_tri;
    int npi_tet;  /* 5, or 1 = ONE_GAUSS_POINT (src/tetra.h:29-81) */
    int npi_tri;  /* 4, or 1 (src/triangle.h:21-65) */
    double tol;   /* finite_element_solver.tolerance */
    int maxiter;  /* finite_element_solver.max(iter) */
    } fg_params;

/* what LinAlgebra::solve + get_v_max + the iteration monitor report (src/solver.cpp:6-90) */
typedef struct fg_step_result
    {
    int failed;     /* LinAlgebra::solve return value: 1 = failure */
    int status;     /* iter.status */
    int iters;      /* iter.get_iteration() */
    int pad_;
    double res;     /* iter.get_res() */
    double rhsnorm; /* iter.get_rhsnorm() */
    double v_max;   /* LinAlgebra::get_v_max() (kept from the previous solve when failed) */
    } fg_step_result;

typedef struct fg_ctx fg_ctx;

const char *fg_last_error(void);
int fg_version(void);

/* ---- construction: LinAlgebra::LinAlgebra(Settings&, Mesh::mesh&), src/linear_algebra.h:40-64 +
 *      solver<2> ctor src/solver.h:26-41 + Tet/Tri ctors src/tetra.h:140-163, src/triangle.h:113-127 */
int fg_create(const fg_mesh *mesh, const fg_params *prm, int device, fg_ctx **out);
void fg_destroy(fg_ctx *ctx);
/* out[0..9] = NOD, NT, NF, n_magTet, n_magTri, E, E_mag, n, nnz, nlvd */
int fg_get_sizes(const fg_ctx *ctx, long long out[10]);
/* device layout chosen for this mesh (no reference counterpart; bench.py's byte model reads it):
 * out[0] = bytes of column index per stored node pair of the matrix-free operator (2 or 4),
 * out[1] = stored node pairs including SELL padding, out[2] = 1 when the element fast path
 * (no anisotropy in any region) is available, out[3] = padded node rows */
int fg_get_layout(const fg_ctx *ctx, long long out[4]);
/* launch shape of the persistent solve kernel (fg_set_solver 0) for a mesh of `nslice` SELL slices (32 node
 * rows each) on a device with `n_sm` SMs (<= 0: 148), computed on the host, no device needed (no reference
 * counterpart; tests pin the heuristics with it): out = {threads per CTA of the kernel variant, warps per CTA
 * actually launched, CTAs, 1 when the variant keeps the head of each warp's first slice in registers} */
int fg_solver_launch_shape(int nslice, int n_sm, int out[4]);

/* ---- node state (Nodes::Node::d[CURRENT|NEXT], src/node.h:47-70) ---- */
/* mesh::init_distrib + Fem ctor: set CURRENT u, v, phi, phiv and copy to NEXT; NULL = zeros.
 * Precondition: |u| = 1 on every magnetic node, as mesh::init_distrib normalises it (src/mesh.cpp) and
 * Node::make_evol keeps it (src/node.h:116-122); the matrix-free operator relies on the triad (ep, eq, u) of
 * Node::setBasis being orthonormal.  A state that violates it (beyond 1e-6) is rejected with FG_ERR_INVALID. */
int fg_set_state(fg_ctx *ctx, const double *u, const double *v, const double *phi,
                 const double *phiv);
/* d[NEXT].v only (what buildInitGuess reads, src/linear_algebra.cpp:13-24) */
int fg_set_next_v(fg_ctx *ctx, const double *v);
/* Nodes::set_phi / set_phiv on NEXT, written by the demag solver each step (src/fmm_demag.h) */
int fg_set_potentials(fg_ctx *ctx, const double *phi, const double *phiv);
/* step: 0 = CURRENT, 1 = NEXT; any output pointer may be NULL */
int fg_get_state(fg_ctx *ctx, int step, double *u, double *v, double *phi, double *phiv);
/* mesh::evolution, src/mesh.h:189-193: NEXT -> CURRENT */
int fg_commit(fg_ctx *ctx);
/* mesh::extSpaceField (src/mesh.h:366): NT x 3 x npi_tet, [tet][xyz][gauss] */
int fg_set_ext_space_field(fg_ctx *ctx, const double *field);

/* ---- the three per-step calls (src/time_integration.cpp:193-206) ---- */
/* LinAlgebra::base_projection, src/linear_algebra.cpp:3-11: `angle` is M_2_PI * U(0,1) drawn by
 * the caller with the reference's own rand()/mt19937 sequence. */
int fg_base_projection(fg_ctx *ctx, double angle);
/* LinAlgebra::prepareElements(Hext, t_prm), src/linear_algebra.cpp:26-52.  dt and prefactor are
 * timing::get_dt() and timing::prefactor (src/time_integration.h:31-42); idx_dir / Vdrift are
 * LinAlgebra::idx_dir and DW_vz (src/linear_algebra.h:113-118). */
int fg_prepare_elements(fg_ctx *ctx, const double Hext[3], double dt, double prefactor,
                        int idx_dir, double Vdrift);
/* LinAlgebra::prepareElements(A_Hext, t_prm), src/linear_algebra.cpp:54-81 */
int fg_prepare_elements_space(fg_ctx *ctx, double A_Hext, double dt, double prefactor,
                              int idx_dir, double Vdrift);
/* LinAlgebra::solve(t_prm), src/solver.cpp:6-90 (dt must equal the one given to prepareElements) */
int fg_solve(fg_ctx *ctx, double dt, fg_step_result *out);
/* the three calls above in one enqueue (one host synchronisation) */
int fg_step(fg_ctx *ctx, double angle, const double Hext[3], double dt, double prefactor,
            int idx_dir, double Vdrift, fg_step_result *out);

/* ---- energies, averages, maximum angle: what Fem::compute_all / saver evaluate after every accepted
 *      step.  Element-wise reductions over the resident NEXT state; on a
 *      distributed context they are collective and return the global value on every rank. ---- */
/* Fem::energy, src/energy.cpp:5-68 with a uniform applied field (RtoR3): E[0..3] = exchange,
 * anisotropy (volume + surface), demag (volume + surface charges), zeeman — the ENERGY_TYPE order of
 * src/fem.h:30-36.  Tet terms src/tetra.cpp:309-391, Tri terms src/triangle.cpp:38-43,80-85. */
int fg_energy(fg_ctx *ctx, const double Hext[3], double E[4]);
/* same with the space-dependent field times an amplitude (R4toR3, src/energy.cpp:41-43) */
int fg_energy_space(fg_ctx *ctx, double A_Hext, double E[4]);
/* mesh::avg, src/mesh.cpp:89-106: what = 0 Nodes::get_u_comp | 1 Nodes::get_v_comp (NEXT state), all
 * three components at once; region = -1 for all magnetic regions, else a volume-region index. */
int fg_avg(fg_ctx *ctx, int what, int region, double out[3]);
/* mesh::max_angle, src/mesh.h:295-306 (radians; over ALL mesh edges, like the reference) */
int fg_max_angle(fg_ctx *ctx, double *angle);

/* ---- magnetic charges and an all-pairs demag potential (single GPU) ---- */
/* fmm::calc_charges, src/fmm_demag.h:155-185, on the NEXT state: which = 0 u | 1 v.  srcDen holds
 * n_magTet*npi_tet volume charges (Tet::charges, src/tetra.cpp:347-359) then n_magTri*npi_tri surface
 * charges (Tri::charges, src/triangle.cpp:45-59) in the reference's order; corr (NOD) the second-order
 * corrections of Tri::correctionCharges / Tri::potential (src/triangle.cpp:61-78,87-125).  Either
 * output may be NULL. */
int fg_calc_charges(fg_ctx *ctx, int which, double *srcDen, double *corr);
/* Stand-in for scal_fmm::fmm::calc_demag (src/fmm_demag.h:98-103,187-223) on meshes small enough for
 * an all-pairs sum: phi = (sum_j q_j / |p_i - x_j| + corr_i) / 4pi from u, and (second_order != 0, the
 * reference's default FIRST_ORDER=OFF) phiv from v, written to the NEXT state of the magnetic nodes.
 * ScalFMM itself stays outside the path (BASELINE north star); this is the sum it approximates. */
int fg_demag_direct(fg_ctx *ctx, int second_order);

/* ---- taps used by the parity tests ---- */
int fg_get_basis(fg_ctx *ctx, double *ep, double *eq);              /* NOD x 3 each */
/* element<N,NPI>::Kp / Lp (src/element.h:62,65) of tets [first, first+count): count x 64, x 8 */
int fg_get_elements(fg_ctx *ctx, int first, int count, double *Kp, double *Lp);
int fg_get_tri_elements(fg_ctx *ctx, int first, int count, double *Lp); /* count x 6 */
/* What the PRODUCTION element kernels (k_tet_iso / k_tet_lean / k_tet) wrote for tets [first, first+count) by
 * the last prepareElements: per local node {sum_g a_i w_g alpha_eff(g), BE(x,y,z)} (count x 4 x 4; zeros for
 * non-magnetic tets and for nodes without a row on this device).  With the node bases they give Lp
 * (src/tetra.cpp:306) and the state-dependent diagonal of Kp (src/tetra.cpp:108-131); fg_get_elements
 * recomputes Kp/Lp with a separate tap kernel, this one reads the records the assembly consumes. */
int fg_get_records(fg_ctx *ctx, int first, int count, double *rec);
/* solver<2>::K shape (src/solver.h:75-104): rowptr n+1, col nnz */
int fg_get_csr_pattern(const fg_ctx *ctx, int *rowptr, int *col);
/* K values (nnz), L_rhs (n) and the initial guess Xw (n) as LinAlgebra::solve builds them
 * (src/solver.cpp:9-59), for the system prepared by the last prepareElements.  Called after
 * fg_solve it returns the K and L_rhs that solve used, and x0 = the solution Xw. */
int fg_get_system(fg_ctx *ctx, double dt, double *val, double *rhs, double *x0);
/* SparseMatrix::build_diag_precond (src/algebra/sparseMat.h:174-183) as the solver uses it: D[i] = 1 / K(i,i),
 * 0 on the masked dofs (n doubles), for the system prepared by the last prepareElements (or solved last). */
int fg_get_precond(fg_ctx *ctx, double dt, double *D);
/* y = K x with the device SpMV the solver itself uses, on the system assembled last (n each) */
int fg_apply_operator(fg_ctx *ctx, const double *x, double *y);
int fg_get_solution(fg_ctx *ctx, double *Xw); /* n */
/* Tet::orientate result and geometry tables: NT x 4, NT x 12, NT x npi_tet (NULL = skip) */
int fg_get_tet_tables(const fg_ctx *ctx, int *ind, double *da, double *weight);

/* ---- generic sparse algebra: algebra::SparseMatrix + solvers (src/algebra) ---- */
typedef struct fg_matrix fg_matrix;
typedef struct fg_iter_result
    {
    int status, iters;
    double res, rhsnorm;
    } fg_iter_result;

/* SparseMatrix(const MatrixShape&), src/algebra/sparseMat.h:79-90: CSR pattern, sorted columns */
int fg_matrix_create(int n, const int *rowptr, const int *col, int device, fg_matrix **out);
void fg_matrix_destroy(fg_matrix *m);
/* values in CSR order (what clear()/add()/set() produced on the host side) */
int fg_matrix_set_values(fg_matrix *m, const double *val);
/* SparseMatrix::mult, src/algebra/sparseMat.h:158-170 */
int fg_matrix_mult(fg_matrix *m, const double *x, double *y);
/* algebra::bicg, src/algebra/bicg.h:14-72 */
int fg_bicg(fg_matrix *m, double *x, const double *rhs, double tol, int maxiter,
            fg_iter_result *out);
/* algebra::bicg_dir, src/algebra/bicg.h:83-154 (xd != NULL) and :163-234 (xd == NULL) */
int fg_bicg_dir(fg_matrix *m, double *x, const double *rhs, const double *xd, const int *ld,
                int nld, double tol, int maxiter, fg_iter_result *out);
/* algebra::cg, src/algebra/cg.h:15-58 */
int fg_cg(fg_matrix *m, double *x, const double *rhs, double tol, int maxiter,
          fg_iter_result *out);
/* algebra::cg_dir, src/algebra/cg.h:68-121 */
int fg_cg_dir(fg_matrix *m, double *x, const double *rhs, const double *xd, const int *ld, int nld,
              double tol, int maxiter, fg_iter_result *out);

/* ---- multi-GPU: row-block (slab) partition over the GPUs of one box ----
 * No reference equivalent (the reference is single-process); SURVEY.md §8(e), DESIGN.md §8.  One
 * process per GPU.  Each rank passes its LOCAL mesh: nodes [0, n_owned) are the rows it owns (a
 * contiguous range of the reference's sorted node order), nodes [n_owned, NOD) are ghosts (the
 * other nodes of the tetrahedra touching an owned node), tets/tris in local numbering.  The
 * ranks exchange the opaque blobs of fg_dist_export out of band (torch.distributed / MPI / files)
 * and hand all of them to fg_dist_connect, which maps the peers' exchange arenas (CUDA IPC over
 * NVLink).  After that every per-step call of this header works unchanged and must be made by
 * all ranks; fg_step_result is identical on every rank.  State arrays are local (owned + ghosts);
 * the taps that return K / L / solution vectors are not available on a distributed context. */
#define FG_DIST_BLOB_BYTES 128
typedef struct fg_dist_desc
    {
    int rank, world;        /* world <= 8 */
    int n_owned;            /* local nodes [0, n_owned) are owned, the rest are ghosts */
    const int *send_ptr;    /* world+1 : send_nodes grouped by destination rank */
    const int *send_nodes;  /* local ids of owned nodes, in the order the destination stores its ghosts */
    const int *send_dst;    /* world : offset (in nodes) of my segment inside the destination's ghost range */
    const int *recv_from;   /* world : 1 if that rank sends ghosts to me */
    } fg_dist_desc;
int fg_dist_create(const fg_mesh *local_mesh, const fg_params *prm, int device,
                   const fg_dist_desc *dist, fg_ctx **out);
int fg_dist_export(fg_ctx *ctx, void *blob /* FG_DIST_BLOB_BYTES */);
int fg_dist_connect(fg_ctx *ctx, const void *blobs /* world x FG_DIST_BLOB_BYTES, rank order */);

/* ---- host-side planning (no GPU needed) ---- */
/* Runs the once-per-mesh preprocessing of fg_create (orientation, geometry tables, sparsity of
 * solver<2>::build_shape, device row ordering and the SELL-32 layout, incidence lists) on the host
 * only, verifies that the device layout is a faithful image of the reference's CSR pattern, and
 * reports its sizes: out[0..11] = NOD, padded NOD, slices, 2x2 blocks of the pattern (nnz/4),
 * stored blocks incl. SELL padding, magnetic tets, widest slice, stored incidence slots,
 * incidences (4 x magnetic tets), masked dofs, E, E_mag. */
int fg_host_plan(const fg_mesh *mesh, const fg_params *prm, long long out[12]);

/* ---- instrumentation ---- */
/* number of kernels this library launched on the context's stream since creation */
long long fg_kernel_launches(const fg_ctx *ctx);
/* the CUDA stream (cudaStream_t) the context launches on, for CUDA-event timing by the caller */
void *fg_stream(const fg_ctx *ctx);
/* Phase timing with CUDA events on the context's stream (adds host synchronisations: for
 * analysis, not for throughput runs).  After a step, out[] holds milliseconds of
 * out[0]=base_projection out[1]=prepareElements out[2]=assembly of K, L, D, x0
 * out[3]=BiCGStab (setup + iterations + node update) ; out[4..7] reserved (0).
 * on = 0 off, 1 phase timers, 2 per-SpMV event pairs (see fg_get_spmv_times). */
int fg_set_profiling(fg_ctx *ctx, int on);
int fg_get_phase_times(const fg_ctx *ctx, double out[8]);
/* Per-iteration record of the BiCGStab scalars (diagnosis): the first call switches the recording on
 * (up to 1024 iterations per solve, overwritten by each solve); later calls copy the first `rows`
 * rows of 8 doubles: rho_1, (v,rt), alpha, |s|^2, (t,s), (t,t), omega, |r|^2 (src/algebra/bicg.h:185-231). */
int fg_get_krylov_history(fg_ctx *ctx, int rows, double *out);
/* Which form of K the solver applies: 0 (default) the matrix-free operator K = cS P^T (S x I3) P + Dg,
 * 1 the assembled 2x2 blocks (what solver::buildMat produces; single GPU) — same algorithm, kept for
 * A/B checks of parity and bandwidth. */
int fg_set_operator(fg_ctx *ctx, int kind);
/* The Krylov scalars of the last solve as the device left them (diagnosis of breakdowns, src/algebra/
 * bicg.h:189-194): out = rho_1, rho_2, alpha, omega, res, rhsnorm, iterations, status. */
int fg_get_krylov_state(const fg_ctx *ctx, double out[8]);
/* fg_set_profiling(ctx, 2): instead of the phase timers, bracket every SpMV launch of the solver
 * with a CUDA-event pair on the context's stream (no host synchronisation is added; at most 4096
 * launches are kept).  fg_get_spmv_times returns their summed device time and count, and resets. */
int fg_get_spmv_times(fg_ctx *ctx, double *total_ms, int *launches);
/* fg_set_profiling(ctx, 3): the same bracketing around EVERY kernel of the step.  Returns the summed
 * device milliseconds and launch counts per kernel class since the mode was set (and keeps them):
 * 0 basis 1 tet 2 tri 3 assemble 4 spmv(setup) 5 bicg_p 6 spmv(v) 7 bicg_s 8 spmv(t) 9 bicg_xr
 * 10 halo 11 update 12 solve (the persistent kernel: setup + all iterations + node update) 13 other
 * 14 gaps between consecutive kernels (idle stream time). */
#define FG_KERNEL_CLASSES 15
int fg_get_kernel_times(fg_ctx *ctx, double ms[FG_KERNEL_CLASSES], int launches[FG_KERNEL_CLASSES]);
/* How LinAlgebra::solve's Krylov part (src/solver.cpp:50-88) is executed: 0 (default) ONE persistent
 * cooperative kernel per solve (grid barriers between the phases of src/algebra/bicg.h:185-232, node update
 * fused), 1 one kernel per phase (5 per iteration) with the host polling the device-side iteration monitor.
 * Same algorithm, same stopping rules; kept for A/B checks.  The environment variable FG_SOLVER=multi
 * selects 1 for every context of the process. */
int fg_set_solver(fg_ctx *ctx, int kind);
/* fg_set_profiling(ctx, 2 or 3) with the persistent solve kernel: ms[0] / count[0] = summed device time
 * (CUDA events) and number of solve-kernel launches since the mode was set; ms[1..8] / count[1..8] = time
 * spent in the phases of the kernel, from %globaltimer stamps taken by its CTA 0 after the grid barrier
 * that closes each phase: 1 setup (r = b - K x0), 2 A (p, image of D p), 3 B (v = K D p, alpha),
 * 4 C (s, image of D s), 5 D (t = K D s, omega), 6 E (x, r, rho), 7 halo of x (multi-GPU), 8 node update.
 * ms[9 + k] = of phase k, the time until the LAST CTA of this GPU arrived at the closing barrier (its work);
 * ms[18 + k] = the cross-GPU part of that barrier (all-reduce over the ranks: latency + waiting for the
 * slowest rank); the remainder of ms[k] is the release of the barrier. */
int fg_get_solve_times(fg_ctx *ctx, double ms[27], long long count[9]);
/* Microbenchmark of the solver's SpMV on the assembled K: runs `reps` launches back to back and
 * returns the mean milliseconds per launch (CUDA events on the context's stream). */
int fg_bench_spmv(fg_ctx *ctx, int reps, double *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif
