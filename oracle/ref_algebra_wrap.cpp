// ref_algebra_wrap.cpp — C-ABI wrapper that compiles the REFERENCE's OWN, UNMODIFIED sparse
// algebra headers (/root/reference/src/algebra/{sparseMat,algebra,algebraCore,iter,bicg,cg}.h and
// src/time_integration.h) into oracle/_ref/libfgref_algebra.so.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Nothing from the reference is copied into this
// repository: the headers are #included from where they lie under $(REF) at build time
// (oracle/Makefile), and the only generated file (config.h, from src/config.h.in) is written to
// the git-ignored oracle/_ref/.  The library is used (a) to validate oracle/fg_oracle.c's
// restatement of SparseMatrix::mult / bicg / bicg_dir / cg / cg_dir bit for bit, and (b) as the
// solver of the CPU baseline ("the reference's own SparseMatrix + bicg_dir").
//
// Note: TBB is absent in this image, so std::execution::par inside SparseMatrix::mult runs on
// libstdc++'s serial backend — exactly the "as shipped in this container" figure of BASELINE.md.

// The reference's time_integration.h calls unqualified abs(log(x)); in the reference build the
// double overload is visible because Eigen's headers include <stdlib.h>/<cmath> first
// (SURVEY.md §8a quirk a1).  Reproduce that include environment.
#include <stdlib.h>
#include <cmath>
#include <math.h>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "algebra/algebra.h"
#include "algebra/bicg.h"
#include "algebra/cg.h"
#include "time_integration.h"
// the reference's own TimeStepper (cut out of src/time_integration.cpp:11-38 by oracle/Makefile at
// build time, never stored in this repository) and LogStats (src/log-stats.h, STL-only)
#include <cfloat>
#include <algorithm>
#include "timestepper_extract.h"
#include "log-stats.h"

namespace
    {
struct RefMatrix
    {
    int n;
    std::vector<int> rowptr, col;
    algebra::SparseMatrix *A;
    };

algebra::MatrixShape make_shape(int n, const int *rowptr, const int *col)
    {
    algebra::MatrixShape shape(n);
    for (int i = 0; i < n; i++)
        for (int k = rowptr[i]; k < rowptr[i + 1]; k++) shape[i].insert(col[k]);
    return shape;
    }

void put_iter(const algebra::iteration<double> &it, int *status, int *nit, double *res,
              double *rhsn)
    {
    if (status) *status = (int)it.status;
    if (nit) *nit = it.get_iteration();
    if (res) *res = it.get_res();
    if (rhsn) *rhsn = it.get_rhsnorm();
    }
    }  // namespace

extern "C"
    {
void *fgref_matrix_create(int n, const int *rowptr, const int *col)
    {
    RefMatrix *m = new RefMatrix;
    m->n = n;
    m->rowptr.assign(rowptr, rowptr + n + 1);
    m->col.assign(col, col + rowptr[n]);
    m->A = new algebra::SparseMatrix(make_shape(n, rowptr, col));
    return m;
    }

void fgref_matrix_destroy(void *h)
    {
    RefMatrix *m = (RefMatrix *)h;
    delete m->A;
    delete m;
    }

void fgref_matrix_clear(void *h) { ((RefMatrix *)h)->A->clear(); }
void fgref_matrix_add(void *h, int i, int j, double v) { ((RefMatrix *)h)->A->add(i, j, v); }
void fgref_matrix_set(void *h, int i, int j, double v) { ((RefMatrix *)h)->A->set(i, j, v); }

void fgref_matrix_set_values(void *h, const double *val)
    {
    RefMatrix *m = (RefMatrix *)h;
    for (int i = 0; i < m->n; i++)
        for (int k = m->rowptr[i]; k < m->rowptr[i + 1]; k++) m->A->set(i, m->col[k], val[k]);
    }

void fgref_matrix_get_values(void *h, double *val)
    {
    RefMatrix *m = (RefMatrix *)h;
    for (int i = 0; i < m->n; i++)
        for (int k = m->rowptr[i]; k < m->rowptr[i + 1]; k++) val[k] = (*m->A)(i, m->col[k]);
    }

void fgref_matrix_mult(void *h, const double *x, double *y)
    {
    RefMatrix *m = (RefMatrix *)h;
    std::vector<double> X(x, x + m->n), Y(m->n);
    algebra::mult(*m->A, X, Y);
    for (int i = 0; i < m->n; i++) y[i] = Y[i];
    }

void fgref_build_diag_precond(void *h, double *D)
    {
    RefMatrix *m = (RefMatrix *)h;
    std::vector<double> d(m->n);
    m->A->build_diag_precond<double>(d);
    for (int i = 0; i < m->n; i++) D[i] = d[i];
    }

void fgref_bicg(void *h, double *x, const double *rhs, int n, double tol, int maxiter, int *status,
                int *nit, double *res, double *rhsn)
    {
    RefMatrix *m = (RefMatrix *)h;
    algebra::iteration<double> it("bicg", tol, false, maxiter);
    std::vector<double> X(x, x + n), B(rhs, rhs + n);
    algebra::bicg<double>(it, *m->A, X, B);
    for (int i = 0; i < n; i++) x[i] = X[i];
    put_iter(it, status, nit, res, rhsn);
    }

double fgref_bicg_dir(void *h, double *x, const double *rhs, int n, const int *ld, int nld,
                      double tol, int maxiter, int *status, int *nit, double *res, double *rhsn)
    {
    RefMatrix *m = (RefMatrix *)h;
    algebra::iteration<double> it("bicg_dir", tol, false, maxiter);
    std::vector<double> X(x, x + n), B(rhs, rhs + n);
    std::vector<int> LD(ld, ld + nld);
    double r = algebra::bicg_dir<double>(it, *m->A, X, B, LD);
    for (int i = 0; i < n; i++) x[i] = X[i];
    put_iter(it, status, nit, res, rhsn);
    return r;
    }

void fgref_bicg_dir_xd(void *h, double *x, const double *rhs, const double *xd, int n,
                       const int *ld, int nld, double tol, int maxiter, int *status, int *nit,
                       double *res, double *rhsn)
    {
    RefMatrix *m = (RefMatrix *)h;
    algebra::iteration<double> it("bicg_dir", tol, false, maxiter);
    std::vector<double> X(x, x + n), B(rhs, rhs + n), XD(xd, xd + n);
    std::vector<int> LD(ld, ld + nld);
    algebra::bicg_dir<double>(it, *m->A, X, B, XD, LD);
    for (int i = 0; i < n; i++) x[i] = X[i];
    put_iter(it, status, nit, res, rhsn);
    }

void fgref_cg(void *h, double *x, const double *rhs, int n, double tol, int maxiter, int *status,
              int *nit, double *res, double *rhsn)
    {
    RefMatrix *m = (RefMatrix *)h;
    algebra::iteration<double> it("cg", tol, false, maxiter);
    std::vector<double> X(x, x + n), B(rhs, rhs + n);
    algebra::cg<double>(it, *m->A, X, B);
    for (int i = 0; i < n; i++) x[i] = X[i];
    put_iter(it, status, nit, res, rhsn);
    }

void fgref_cg_dir(void *h, double *x, const double *rhs, const double *xd, int n, const int *ld,
                  int nld, double tol, int maxiter, int *status, int *nit, double *res,
                  double *rhsn)
    {
    RefMatrix *m = (RefMatrix *)h;
    algebra::iteration<double> it("cg_dir", tol, false, maxiter);
    std::vector<double> X(x, x + n), B(rhs, rhs + n), XD(xd, xd + n);
    std::vector<int> LD(ld, ld + nld);
    algebra::cg_dir<double>(it, *m->A, X, B, XD, LD);
    for (int i = 0; i < n; i++) x[i] = X[i];
    put_iter(it, status, nit, res, rhsn);
    }

// class timing, src/time_integration.h:6-59 : out = {dt0, prefactor(dt0), prefactor(dt)}
void fgref_timing(double tf, double dtmin, double dtmax, double dt, double out[3])
    {
    timing t(tf, dtmin, dtmax);
    out[0] = t.get_dt();
    out[1] = t.prefactor;
    t.set_dt(dt);
    out[2] = t.prefactor;
    }

// BLAS-1 helpers, src/algebra/algebra.h:35-77, algebraCore.h:10-17 (for the ut_algebra port)
double fgref_dot(const double *x, const double *y, int n)
    {
    std::vector<double> X(x, x + n), Y(y, y + n);
    return algebra::dot(X, Y);
    }
double fgref_norm(const double *x, int n)
    {
    std::vector<double> X(x, x + n);
    return algebra::norm(X);
    }
    /* ---- TimeStepper (src/time_integration.cpp:11-38) and LogStats (src/log-stats.h) ---- */
void *fgref_ts_new(double initial, double dtmin, double dtmax) { return new TimeStepper(initial, dtmin, dtmax); }
void fgref_ts_free(void *h) { delete (TimeStepper *)h; }
void fgref_ts_set_soft_limit(void *h, double mx) { ((TimeStepper *)h)->set_soft_limit(mx); }
double fgref_ts_step(void *h, double stride) { return (*(TimeStepper *)h)(stride); }
void *fgref_ls_new(void) { return new LogStats; }
void fgref_ls_free(void *h) { delete (LogStats *)h; }
void fgref_ls_add(void *h, double x) { ((LogStats *)h)->add(x); }
void fgref_ls_get(void *h, double out[3])
    {
    LogStats *l = (LogStats *)h;
    out[0] = (double)l->count();
    out[1] = l->mean();
    out[2] = l->stddev();
    }
}
