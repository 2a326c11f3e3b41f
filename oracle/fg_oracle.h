/* fg_oracle.h — CPU ORACLE for the FeeLLGood per-time-step LLG hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a plain-C restatement of the
 * reference's algorithm, written line by line from the reference sources cited at each
 * function (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product library
 * (feellgood_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity pinning: the reference has no stored numeric golden vectors for this path; its
 * unit tests pin each formula against "ref code" (SURVEY.md §4, §8c).  Those ref-code blocks
 * are ported in tests/cpp/ut_oracle.cpp and run against this oracle with the reference's own
 * fixtures, seed (mt19937(5489)) and tolerance (UT_TOL = 5e-16, relaxed as in the reference).
 * The sparse algebra (SparseMatrix::mult, bicg, bicg_dir, cg, cg_dir) is additionally checked
 * against the reference's own unmodified src/algebra headers compiled into oracle/_ref/ (see
 * oracle/Makefile, oracle/ref_algebra_wrap.cpp).  The composition (Tet::integrales end to end,
 * scatter, LinAlgebra::solve) is pinned by no reference test (SURVEY.md §4 "gaps"): for it this
 * restatement is cross-checked against an independent numpy dense restatement and the
 * structured closed form (tests/test_oracle_*.py).
 */
#ifndef FG_ORACLE_H
#define FG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* src/config.h.in:31-32,52 */
#define FGO_MU0 1.25663706127e-6
#define FGO_GAMMA0 (1.76085962784e11 * FGO_MU0)
#define FGO_THETA 0.5
#define FGO_EPSILON 1e-40

#define FGO_TET_N 4
#define FGO_TRI_N 3
#define FGO_NPI_TET_MAX 5
#define FGO_NPI_TRI_MAX 4

/* src/algebra/iter.h:24-30 */
enum { FGO_UNDEFINED = -1, FGO_CONVERGED = 0, FGO_ITER_OVERFLOW = 1, FGO_CANNOT_CONVERGE = 2 };

/* src/node.h:35-41 */
enum { FGO_IDX_UNDEF = -1, FGO_IDX_X = 0, FGO_IDX_Y = 1, FGO_IDX_Z = 2 };

/* Volume-region material constants actually read by the hot path (src/tetra.h:86-112). */
typedef struct fgo_tet_prm
    {
    double alpha_LLG, A, Ms, K;
    double uk[3];
    double K3;
    double ex[3], ey[3], ez[3];
    } fgo_tet_prm;

/* Surface-region constants read by the hot path (src/triangle.h:69-82). */
typedef struct fgo_tri_prm
    {
    double Ks;
    double uk[3];
    int suppress_charges;
    int pad_;
    } fgo_tri_prm;

/* ---------------- scalar / single-element functions (unit-test granularity) -------------- */

/* timing ctor + set_dt, src/time_integration.h:11-15,37-42 (abs() there resolves to fabs). */
double fgo_timing_dt0(double dtmin, double dtmax);
double fgo_timing_prefactor(double dt, double dtmax);

/* Gauss tables, src/tetra.h:29-81 and src/triangle.h:21-65; npi = 5|1 (tet), 4|1 (tri). */
const double *fgo_tet_a(int npi);   /* a[i*npi + g] */
const double *fgo_tet_pds(int npi);
const double *fgo_tri_a(int npi);
const double *fgo_tri_pds(int npi);

/* Node::setBasis, src/node.h:73-102 (PARANOID_ORTHONORMALIZATION false). */
void fgo_node_set_basis(const double u[3], double r, double ep[3], double eq[3]);
/* Node::make_evol, src/node.h:116-122 (vp, vq already multiplied by gamma0 by the caller). */
void fgo_node_make_evol(const double u0[3], const double ep[3], const double eq[3], double vp,
                        double vq, double dt, double u1[3], double v1[3]);

/* Tet::orientate src/tetra.cpp:410-424: swaps ind[2],ind[3] when the mixed product is negative.
 * returns 0 ok, 1 swapped, -1 singular. p is indexed by the (possibly swapped) ind afterwards. */
int fgo_tet_orientate(const double *node_p, int ind[4]);
/* Tet ctor src/tetra.h:140-163 + Jacobian src/tetra.cpp:393-408: da (4x3 row-major), weights. */
double fgo_tet_setup(const double *node_p, const int ind[4], int npi, double da[12], double *weight);
/* Tri ctor src/triangle.h:113-127,206-233. */
void fgo_tri_setup(const double *node_p, const int ind[3], int npi, double *surf, double n[3],
                   double *weight);

/* src/tetra.cpp:47-75 */
void fgo_calc_alpha_eff(int npi, double dt, double alpha, const double *uHeff, double *a_eff);
/* src/tetra.cpp:108-148 ; AE is 12x12 row-major, accumulated into (+=). u_nod[i*3+d]. */
void fgo_tet_lumping(int npi, const double da[12], const double *weight, const double *u_nod,
                     const double *alpha_eff, double prefactor, double AE[144]);
/* src/tetra.cpp:171-181 ; U,V,H_aniso are [d*npi+g]; returns contribution to uHeff in out. */
void fgo_calc_aniso_uniax(int npi, const double uk[3], double Kbis, double s_dt, const double *U,
                          const double *V, double *H_aniso, double *out);
/* src/tetra.cpp:183-208 */
void fgo_calc_aniso_cub(int npi, const double ex[3], const double ey[3], const double ez[3],
                        double K3bis, double s_dt, const double *U, const double *V,
                        double *H_aniso, double *out);
/* element::buildMatP src/element.h:81-96 ; P is (2N)x(3N) row-major, ep/eq are [i*3+d]. */
void fgo_build_matP(int N, const double *ep, const double *eq, double *P);

/* Tet::integrales src/tetra.cpp:210-307.  Node inputs are [i*3+d] / [i]; Hext is [d*npi+g]
 * (what calc_Hext() returns).  Kp is 8x8 row-major, Lp 8. */
void fgo_tet_integrales(int npi, const fgo_tet_prm *prm, double dt, double prefactor,
                        const double da[12], const double *weight, const double *u_nod,
                        const double *v_nod, const double *phi_nod, const double *phiv_nod,
                        const double *ep_nod, const double *eq_nod, const double *Hext, int idx_dir,
                        double Vdrift, double Kp[64], double Lp[8]);
/* Tri::integrales src/triangle.cpp:6-36 ; Lp 6. */
void fgo_tri_integrales(int npi, const fgo_tri_prm *prm, double dMs, const double *weight,
                        const double *u_nod, const double *ep_nod, const double *eq_nod,
                        double Lp[6]);

/* ---------------- sparse algebra (src/algebra) on CSR arrays ---------------------------- */

/* SparseMatrix::mult src/algebra/sparseMat.h:158-170 */
void fgo_spmv(int n, const int *rowptr, const int *col, const double *val, const double *x,
              double *y);

typedef struct fgo_iter
    {
    double resmax;  /* in  */
    int maxiter;    /* in  */
    int status;     /* out */
    int nit;        /* out */
    double res;     /* out */
    double rhsn;    /* out */
    } fgo_iter;

/* src/algebra/bicg.h:14-72 */
void fgo_bicg(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val, double *x,
              const double *rhs);
/* src/algebra/bicg.h:163-234 (mask variant, Dirichlet values zero) ; returns res/rhsn */
double fgo_bicg_dir(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                    double *x, const double *rhs, const int *ld, int nld);
/* src/algebra/bicg.h:83-154 (Dirichlet values xd) */
void fgo_bicg_dir_xd(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                     double *x, const double *rhs, const double *xd, const int *ld, int nld);
/* src/algebra/cg.h:15-58 */
void fgo_cg(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val, double *x,
            const double *rhs);
/* src/algebra/cg.h:68-121 */
void fgo_cg_dir(fgo_iter *it, int n, const int *rowptr, const int *col, const double *val,
                double *x, const double *rhs, const double *xd, const int *ld, int nld);

/* ---------------- mesh-level context: the LinAlgebra call surface ----------------------- */

typedef struct fgo_ctx fgo_ctx;

/* Builds everything Mesh::mesh (src/mesh.h:38-135: edges, magNode, magTet, magTri) and the
 * solver<2>/LinAlgebra constructors (src/solver.h:39,75-104, src/linear_algebra.h:40-64) build.
 * node_p: NOD x 3 (already scaled and sorted by the caller), tet_ind: NT x 4 zero-based,
 * tet_reg / tri_reg: region index per element, tri_dMs: per triangle (src/mesh.cpp:227-242).
 * Elements are taken in the order given; tets are re-oriented internally like the Tet ctor. */
fgo_ctx *fgo_create(int NOD, const double *node_p, int NT, const int *tet_ind, const int *tet_reg,
                    int NF, const int *tri_ind, const int *tri_reg, const double *tri_dMs,
                    int nreg_tet, const fgo_tet_prm *prm_tet, int nreg_tri,
                    const fgo_tri_prm *prm_tri, int npi_tet, int npi_tri, double tol, int maxiter);
void fgo_destroy(fgo_ctx *c);
void fgo_set_num_threads(fgo_ctx *c, int nthreads); /* OpenMP stand-in for EXEC_POL (TBB) */
/* if non-NULL, solve() drives the reference's own SparseMatrix + bicg_dir through the
 * oracle/_ref wrapper library at so_path; returns 0 on success. */
int fgo_use_reference_algebra(fgo_ctx *c, const char *so_path);

/* sizes: out[0]=NOD out[1]=NT out[2]=NF out[3]=n_magTet out[4]=n_magTri out[5]=E(all edges)
 * out[6]=E_mag out[7]=n(=2NOD) out[8]=nnz out[9]=nlvd */
void fgo_sizes(const fgo_ctx *c, long long out[10]);

/* state: CURRENT u,v,phi,phiv (NOD x 3, NOD x 3, NOD, NOD); also copied to NEXT like
 * mesh::init_distrib + evolution leave it. */
void fgo_set_state(fgo_ctx *c, const double *u, const double *v, const double *phi,
                   const double *phiv);
void fgo_set_next_v(fgo_ctx *c, const double *v); /* d[NEXT].v, read by buildInitGuess */
void fgo_set_potentials_next(fgo_ctx *c, const double *phi, const double *phiv);
void fgo_get_state(const fgo_ctx *c, int step /*0 CURRENT,1 NEXT*/, double *u, double *v,
                   double *phi, double *phiv);
void fgo_get_basis(const fgo_ctx *c, double *ep, double *eq);
void fgo_evolution(fgo_ctx *c); /* src/mesh.h:189-193 */
void fgo_set_ext_space_field(fgo_ctx *c, const double *field /* NT x 3 x npi, [t][d][g] */);

/* LinAlgebra::base_projection with the angle r = M_2_PI * U(0,1) drawn by the caller
 * (src/linear_algebra.cpp:3-11, src/mesh.h:178-182). */
void fgo_base_projection(fgo_ctx *c, double r);
/* LinAlgebra::prepareElements (uniform field) src/linear_algebra.cpp:26-52 */
void fgo_prepare_elements(fgo_ctx *c, const double Hext[3], double dt, double prefactor,
                          int idx_dir, double Vdrift);
/* LinAlgebra::prepareElements (space field × amplitude) src/linear_algebra.cpp:54-81 */
void fgo_prepare_elements_space(fgo_ctx *c, double A_Hext, double dt, double prefactor,
                                int idx_dir, double Vdrift);
/* LinAlgebra::solve src/solver.cpp:6-90 ; returns 1 on failure like the reference's bool. */
int fgo_solve(fgo_ctx *c, double dt);
/* only the assembly half of solve() (src/solver.cpp:9-48 + buildInitGuess :59), for taps */
void fgo_assemble(fgo_ctx *c);
double fgo_get_v_max(const fgo_ctx *c);
void fgo_get_iter(const fgo_ctx *c, fgo_iter *out);

/* taps */
void fgo_get_tet_ind(const fgo_ctx *c, int *ind /* NT x 4 after orientation */);
void fgo_get_tet_geom(const fgo_ctx *c, double *da /* NT x 12 */, double *weight /* NT x npi */);
void fgo_get_element(const fgo_ctx *c, int tet, double Kp[64], double Lp[8]);
void fgo_get_tri_element(const fgo_ctx *c, int tri, double Lp[6]);
void fgo_get_csr(const fgo_ctx *c, int *rowptr /* n+1 */, int *col /* nnz */);
void fgo_get_system(const fgo_ctx *c, double *val /* nnz */, double *rhs /* n */,
                    double *x /* n: Xw */);
void fgo_get_masks(const fgo_ctx *c, unsigned char *magNode /* NOD */, int *lvd /* nlvd */);
void fgo_get_edges(const fgo_ctx *c, int *edges /* E x 2 */);

/* ---------------- "next" rows: energies and averages ------------------ */
/* Fem::energy src/energy.cpp:5-68 on NEXT state; E[4] = exchange, anisotropy, demag, zeeman */
void fgo_energy(const fgo_ctx *c, const double Hext[3], double E[4]);
/* same with the space-dependent field x amplitude (R4toR3), src/energy.cpp:41-43, tetra.cpp:382-391 */
void fgo_energy_space(const fgo_ctx *c, double fieldAmp, double E[4]);
/* mesh::avg src/mesh.cpp:89-106 for u (what=0) or v (what=1) NEXT, all magnetic regions */
void fgo_avg(const fgo_ctx *c, int what, double out[3]);
/* same restricted to one volume region (region = -1: all magnetic regions) */
void fgo_avg_region(const fgo_ctx *c, int what, int region, double out[3]);
double fgo_region_vol(const fgo_ctx *c, int region);
double fgo_total_mag_vol(const fgo_ctx *c);
/* mesh::max_angle src/mesh.h:295-306 */
double fgo_max_angle(const fgo_ctx *c);

/* ---------------- "next" rows rank 2: magnetic charges and the demag potential ------------- */
/* Tet::charges src/tetra.cpp:347-359 ; Tri::charges src/triangle.cpp:45-59 ; Tri::potential
 * src/triangle.cpp:87-125 (vec_nod is [i*3+d]) */
void fgo_tet_charges(int npi, double Ms, const double da[12], const double *weight,
                     const double *vec_nod, double *out);
void fgo_tri_charges(int npi, double dMs, const double n[3], const double *weight,
                     const double *vec_nod, double *out);
double fgo_tri_potential(const double *p, const double *vec_nod, double surf, const double n[3],
                         double dMs, int i);
long long fgo_n_sources(const fgo_ctx *c);
void fgo_source_positions(const fgo_ctx *c, double *pos);
/* fmm::calc_charges src/fmm_demag.h:155-185 on NEXT: which = 0 u | 1 v ; corr has NOD entries */
void fgo_calc_charges(const fgo_ctx *c, int which, double *srcDen, double *corr);
/* all-pairs stand-in for fmm::demag src/fmm_demag.h:187-223 (ScalFMM is absent): writes NEXT phi
 * (which = 0) or phiv (which = 1) of the magnetic nodes */
void fgo_demag_direct(fgo_ctx *c, int which);
void fgo_calc_demag_direct(fgo_ctx *c, int second_order);

#ifdef __cplusplus
}
#endif
#endif
