// fg_matrix.cu — C ABI of the generic sparse algebra: algebra::SparseMatrix + bicg / bicg_dir /
// cg / cg_dir (reference src/algebra/sparseMat.h, bicg.h, cg.h) on one B200.
// Used by the reference's side solvers (electrostatSolver: cg_dir, spinAccumulationSolver:
// bicg_dir) and by the ported algebra unit tests; the LLG step itself goes through fg_llg.cu.
#include <string.h>

#include <vector>

#include "fg_common.cuh"

namespace fg
{
int resid_into(const Operator &op, const KrylovWork &w, const double *xd, const double *b,
               double *out);
}
using namespace fg;

struct fg_matrix
    {
    int device = 0, n = 0, nnz = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    int *rowptr = nullptr, *col = nullptr;
    double *val = nullptr, *xd = nullptr;
    unsigned char *mask = nullptr;
    KrylovWork kw;
    Operator op;
    };

namespace
{
int check(fg_matrix *m)
    {
    if (!m)
        {
        set_error("null matrix");
        return FG_ERR_INVALID;
        }
    FG_CUDA(cudaSetDevice(m->device));
    return FG_OK;
    }

// upload the Dirichlet list as a byte mask (NULL list = no masking)
int set_mask(fg_matrix *m, const int *ld, int nld)
    {
    if (!ld)
        {
        m->kw.mask = nullptr;
        return FG_OK;
        }
    std::vector<unsigned char> h((size_t)m->n, 0);
    for (int k = 0; k < nld; k++)
        {
        if (ld[k] < 0 || ld[k] >= m->n)
            {
            set_error("Dirichlet index %d out of range", ld[k]);
            return FG_ERR_INVALID;
            }
        h[ld[k]] = 1;
        }
    FG_CUDA(cudaMemcpyAsync(m->mask, h.data(), h.size(), cudaMemcpyHostToDevice, m->stream));
    FG_CUDA(cudaStreamSynchronize(m->stream));
    m->kw.mask = m->mask;
    return FG_OK;
    }

// common body of the four solvers (bicg.h:14-72,83-154,163-234 ; cg.h:15-58,68-121)
int run(fg_matrix *m, bool bicg, double *x, const double *rhs, const double *xd, const int *ld,
        int nld, double tol, int maxiter, fg_iter_result *out)
    {
    FG_TRY(check(m));
    if (!x || !rhs)
        {
        set_error("solver: null x or rhs");
        return FG_ERR_INVALID;
        }
    const size_t nb = sizeof(double) * (size_t)m->n;
    KrylovWork &w = m->kw;
    FG_TRY(set_mask(m, ld, nld));
    FG_CUDA(cudaMemcpyAsync(w.x, x, nb, cudaMemcpyHostToDevice, m->stream));
    FG_CUDA(cudaMemcpyAsync(w.b, rhs, nb, cudaMemcpyHostToDevice, m->stream));
    FG_TRY(build_diag_precond_csr(m->op, w));  // D, masked
    if (xd)
        {  // b -= A xd
        FG_CUDA(cudaMemcpyAsync(m->xd, xd, nb, cudaMemcpyHostToDevice, m->stream));
        FG_TRY(resid_into(m->op, w, m->xd, w.b, w.t));
        FG_CUDA(cudaMemcpyAsync(w.b, w.t, nb, cudaMemcpyDeviceToDevice, m->stream));
        }
    FG_TRY(vec_mask(w, w.b));
    if (bicg)
        FG_TRY(bicgstab_run(m->op, w, tol, maxiter, nullptr, nullptr));
    else
        FG_TRY(cg_run(m->op, w, tol, maxiter));
    if (xd) FG_TRY(vec_axpy(w, 1.0, m->xd, w.x));  // x += xd
    FG_CUDA(cudaMemcpyAsync(x, w.x, nb, cudaMemcpyDeviceToHost, m->stream));
    FG_CUDA(cudaStreamSynchronize(m->stream));
    if (out)
        {
        out->status = w.h_st->status;
        out->iters = w.h_st->nit;
        out->res = w.h_st->res;
        out->rhsnorm = w.h_st->rhsn;
        }
    return FG_OK;
    }
}  // namespace

extern "C" {

int fg_matrix_create(int n, const int *rowptr, const int *col, int device, fg_matrix **out)
    {
    if (!out || !rowptr || n <= 0 || rowptr[0] != 0 || (rowptr[n] > 0 && !col))
        {
        set_error("fg_matrix_create: bad argument");
        return FG_ERR_INVALID;
        }
    *out = nullptr;
    for (int i = 0; i < n; i++)
        {
        if (rowptr[i + 1] < rowptr[i])
            {
            set_error("fg_matrix_create: rowptr not monotone at row %d", i);
            return FG_ERR_INVALID;
            }
        for (int k = rowptr[i]; k < rowptr[i + 1]; k++)
            if (col[k] < 0 || col[k] >= n || (k > rowptr[i] && col[k] <= col[k - 1]))
                {
                set_error("fg_matrix_create: row %d: columns must be sorted, unique and in range", i);
                return FG_ERR_INVALID;
                }
        }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        {
        set_error("fg_matrix_create: no CUDA device available (this library has no CPU fallback)");
        return FG_ERR_CUDA;
        }
    if (device < 0 || device >= ndev)
        {
        set_error("fg_matrix_create: device %d out of range", device);
        return FG_ERR_INVALID;
        }
    FG_CUDA(cudaSetDevice(device));
    fg_matrix *m = new fg_matrix();
    m->device = device;
    m->n = n;
    m->nnz = rowptr[n];
    int rc = FG_OK;
    do
        {
        if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = FG_ERR_CUDA; break; }
        const size_t nz = (size_t)(m->nnz > 0 ? m->nnz : 1);
        if (cudaMalloc(&m->rowptr, sizeof(int) * ((size_t)n + 1)) != cudaSuccess
            || cudaMalloc(&m->col, sizeof(int) * nz) != cudaSuccess
            || cudaMalloc(&m->val, sizeof(double) * nz) != cudaSuccess
            || cudaMalloc(&m->xd, sizeof(double) * (size_t)n) != cudaSuccess
            || cudaMalloc(&m->mask, (size_t)n) != cudaSuccess)
            { rc = FG_ERR_CUDA; break; }
        cudaMemcpyAsync(m->rowptr, rowptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, m->stream);
        if (m->nnz > 0) cudaMemcpyAsync(m->col, col, sizeof(int) * (size_t)m->nnz, cudaMemcpyHostToDevice, m->stream);
        cudaMemsetAsync(m->val, 0, sizeof(double) * nz, m->stream);  // SparseMatrix ctor: zeros
        if ((rc = krylov_alloc(m->kw, n, 0, m->stream, &m->launches)) != FG_OK) break;
        if (cudaStreamSynchronize(m->stream) != cudaSuccess) { rc = FG_ERR_CUDA; break; }
        } while (0);
    if (rc != FG_OK)
        {
        if (rc == FG_ERR_CUDA) set_error("fg_matrix_create: %s", cudaGetErrorString(cudaGetLastError()));
        fg_matrix_destroy(m);
        return rc;
        }
    const double mean = (double)m->nnz / n;
    m->op.kind = OP_CSR;
    m->op.n = n;
    m->op.lanes = mean > 24 ? 32 : mean > 12 ? 16 : mean > 6 ? 8 : mean > 3 ? 4 : 2;
    m->op.ptr = m->rowptr;
    m->op.col = m->col;
    m->op.val = m->val;
    *out = m;
    return FG_OK;
    }

void fg_matrix_destroy(fg_matrix *m)
    {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    void *ptrs[] = {m->rowptr, m->col, m->val, m->xd, m->mask};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    krylov_free(m->kw);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    }

int fg_matrix_set_values(fg_matrix *m, const double *val)
    {
    FG_TRY(check(m));
    if (!val && m->nnz > 0)
        {
        set_error("fg_matrix_set_values: null values");
        return FG_ERR_INVALID;
        }
    if (m->nnz > 0)
        FG_CUDA(cudaMemcpyAsync(m->val, val, sizeof(double) * (size_t)m->nnz, cudaMemcpyHostToDevice, m->stream));
    FG_CUDA(cudaStreamSynchronize(m->stream));
    return FG_OK;
    }

int fg_matrix_mult(fg_matrix *m, const double *x, double *y)
    {
    FG_TRY(check(m));
    if (!x || !y)
        {
        set_error("fg_matrix_mult: null argument");
        return FG_ERR_INVALID;
        }
    const size_t nb = sizeof(double) * (size_t)m->n;
    FG_CUDA(cudaMemcpyAsync(m->kw.phat, x, nb, cudaMemcpyHostToDevice, m->stream));
    FG_TRY(spmv(m->op, m->kw, m->kw.phat, m->kw.v, false));
    FG_CUDA(cudaMemcpyAsync(y, m->kw.v, nb, cudaMemcpyDeviceToHost, m->stream));
    FG_CUDA(cudaStreamSynchronize(m->stream));
    return FG_OK;
    }

int fg_bicg(fg_matrix *m, double *x, const double *rhs, double tol, int maxiter, fg_iter_result *out)
    { return run(m, true, x, rhs, nullptr, nullptr, 0, tol, maxiter, out); }

int fg_bicg_dir(fg_matrix *m, double *x, const double *rhs, const double *xd, const int *ld, int nld,
                double tol, int maxiter, fg_iter_result *out)
    {
    static const int none = 0;
    return run(m, true, x, rhs, xd, ld ? ld : &none, ld ? nld : 0, tol, maxiter, out);
    }

int fg_cg(fg_matrix *m, double *x, const double *rhs, double tol, int maxiter, fg_iter_result *out)
    { return run(m, false, x, rhs, nullptr, nullptr, 0, tol, maxiter, out); }

int fg_cg_dir(fg_matrix *m, double *x, const double *rhs, const double *xd, const int *ld, int nld,
              double tol, int maxiter, fg_iter_result *out)
    {
    static const int none = 0;
    return run(m, false, x, rhs, xd, ld ? ld : &none, ld ? nld : 0, tol, maxiter, out);
    }

}  // extern "C"
