// feellgood_b200.hpp — C++17 drop-in for FeeLLGood's per-time-step call surface, on top of the C ABI
// (include/feellgood_b200.h -> libfeellgood_b200.so).  Header-only; needs no Eigen.
//
// It keeps the names, argument meaning and error behaviour of the reference classes that
// Fem::time_integration drives (reference src/time_integration.cpp:193-206):
//
//   class timing                         src/time_integration.h:6-59       (verbatim semantics)
//   class LinAlgebra                     src/linear_algebra.h:33-121       (base_projection,
//                                        prepareElements x2, solve, get_v_max, set_DW_vz,
//                                        buildInitGuess)
//   namespace algebra:                   src/algebra/{iter,sparseMat,bicg,cg}.h
//       algoStatus, iteration<T>, MatrixShape, SparseMatrix{clear,set,add,operator(),mult,
//       build_diag_precond}, bicg, bicg_dir (both overloads), cg, cg_dir
//
// Differences a maintainer has to know:
//   * the node state (u, v, phi, phiv CURRENT/NEXT) that the reference keeps inside Mesh::mesh lives
//     in HBM; LinAlgebra::set_state / set_potentials / get_state / evolution are the accessors the
//     loop uses where the reference touches msh directly (mesh::evolution, Nodes::set_phi ...);
//   * Eigen::Vector3d arguments are accepted as anything with .data() -> const double* (Eigen's
//     Vector3d qualifies) or as fgb200::Vec3;
//   * fatal errors throw std::runtime_error instead of calling exit(1) (the reference's -DLIBRARY
//     behaviour, src/config.h.in:14-29); solve() still reports solver failure by returning true.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/feellgood_b200.h"

#ifndef M_2_PI
#define M_2_PI 0.63661977236758134308 /* 2/pi: what the reference multiplies by (sic) */
#endif

// ---------------------------------------------------------------------------------------------
// timing — reference src/time_integration.h:6-59.  std::fabs spelled out: the reference's
// unqualified abs() only resolves to the double overload through Eigen's includes.
// ---------------------------------------------------------------------------------------------
class timing
    {
public:
    inline timing(const double _tf, const double _dtmin, const double _dtmax)
        : tf(_tf), DTMIN(_dtmin), DTMAX(_dtmax), TAUR(100. * DTMAX), t(0)
        { set_dt(std::sqrt(DTMIN * DTMAX)); }
    const double tf, DTMIN, DTMAX, TAUR;
    double prefactor;
    inline double get_dt() const { return dt; }
    inline void set_dt(const double _dt)
        {
        dt = _dt;
        const double t_tilde = _dt / TAUR;
        prefactor = (1. + t_tilde * std::fabs(std::log(t_tilde)));
        }
    inline bool is_dt_TooSmall() const { return (dt < DTMIN); }
    inline void inc_t() { t += dt; }
    inline double get_t() const { return t; }
    inline void set_t(const double _t) { t = _t; }

private:
    double t, dt;
    };

namespace fgb200
{
struct Vec3
    {
    double v[3];
    const double *data() const { return v; }
    };

inline void check(int rc, const char *what)
    {
    if (rc != FG_OK)
        throw std::runtime_error(std::string(what) + ": " + fg_last_error());
    }

// The Settings fields LinAlgebra reads (reference src/settings.h; linear_algebra.h:40-53)
struct Settings
    {
    std::vector<fg_tet_prm> paramTetra;     // index 0 = "__default__" like the reference
    std::vector<fg_tri_prm> paramTriangle;
    double TOL = 1e-6;                      // default-settings.yml:232
    int MAXITER = 700;                      // default-settings.yml:229
    int verbose = 0;
    bool recenter = false;
    int recentering_direction = FG_IDX_Z;
    int npi_tet = 5, npi_tri = 4;           // ONE_GAUSS_POINT -> 1, 1
    };

// What Mesh::mesh hands to the solver (src/mesh.h:38-135), as plain arrays
struct MeshView
    {
    std::vector<double> node_p;             // NOD x 3, scaled and sorted (mesh::sortNodes)
    std::vector<int> tet_ind, tet_reg;      // NT x 4, NT
    std::vector<int> tri_ind, tri_reg;      // NF x 3, NF
    std::vector<double> tri_dMs;            // NF
    int NOD() const { return (int)(node_p.size() / 3); }
    int NT() const { return (int)tet_reg.size(); }
    int NF() const { return (int)tri_reg.size(); }
    };
}  // namespace fgb200

// ---------------------------------------------------------------------------------------------
// namespace algebra — reference src/algebra
// ---------------------------------------------------------------------------------------------
namespace algebra
{
enum algoStatus { UNDEFINED = -1, CONVERGED = 0, ITER_OVERFLOW = 1, CANNOT_CONVERGE = 2 };

// iteration<T> (src/algebra/iter.h:37-179): same public surface; the loop itself runs on the GPU
// and fills the monitor when it returns.
template <typename T> class iteration
    {
protected:
    T rhsn;
    const int maxiter;
    const bool noise;
    int nit;
    T res;

private:
    const std::string solver_name;

public:
    iteration(const std::string &name, T r, const bool _noise, const int _maxiter)
        : rhsn(1.0), maxiter(_maxiter), noise(_noise), nit(0), res(std::numeric_limits<T>::max()),
          solver_name(name), status(UNDEFINED), resmax(r) {}
    algoStatus status;
    const T resmax;
    void reset(void)
        {
        rhsn = 1.0;
        nit = 0;
        res = std::numeric_limits<T>::max();
        status = UNDEFINED;
        }
    std::string infos(void) const
        {
        static const char *names[] = {"UNDEFINED", "CONVERGED", "ITER_OVERFLOW", "CANNOT_CONVERGE"};
        std::stringstream s;
        s << solver_name << " status " << names[(int)status + 1] << " after " << nit
          << " iterations, residu= " << res;
        return s.str();
        }
    T get_res() const { return res; }
    int get_iteration() const { return nit; }
    T get_rhsnorm() const { return rhsn; }
    void set_rhsnorm(T r) { rhsn = r; }
    int get_maxiter() const { return maxiter; }
    // filled by the device solvers
    void assign(int st, int iters, T r, T rn)
        {
        status = (algoStatus)st;
        nit = iters;
        res = r;
        rhsn = rn;
        if (noise) std::cout << infos() << "\n";
        }
    };

using MatrixShape = std::vector<std::set<int>>;

// SparseMatrix (src/algebra/sparseMat.h:45-191).  Values are kept on the host in CSR order for
// clear/set/add/operator() (assembly by the caller, as in the reference) and mirrored to the GPU
// lazily for mult and the Krylov solvers.
class SparseMatrix
    {
public:
    SparseMatrix(const MatrixShape &shape) : rowptr(shape.size() + 1, 0)
        {
        for (size_t i = 0; i < shape.size(); ++i) rowptr[i + 1] = rowptr[i] + (int)shape[i].size();
        col.reserve((size_t)rowptr.back());
        for (const auto &row : shape) col.insert(col.end(), row.begin(), row.end());
        val.assign(col.size(), 0.0);
        }
    SparseMatrix(const SparseMatrix &) = delete;
    SparseMatrix &operator=(const SparseMatrix &) = delete;
    ~SparseMatrix() { if (dev) fg_matrix_destroy(dev); }
    void clear() { std::fill(val.begin(), val.end(), 0.0); dirty = true; }
    void set(const int i, const int j, const double v) { val[(size_t)find(i, j, true)] = v; dirty = true; }
    // not thread-safe across rows sharing a cache line is fine; per-entry races are the caller's
    void add(const int i, const int j, const double v) { val[(size_t)find(i, j, true)] += v; dirty = true; }
    double operator()(const int i, const int j) const
        {
        const int k = find(i, j, false);
        return k < 0 ? 0.0 : val[(size_t)k];
        }
    void print(std::ostream &flux = std::cout) const
        {
        flux << "[\n";
        for (size_t i = 0; i + 1 < rowptr.size(); ++i)
            {
            flux << "  {";
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
                flux << col[(size_t)k] << ": " << val[(size_t)k] << (k < rowptr[i + 1] - 1 ? ", " : "\n");
            flux << "}\n";
            }
        flux << ']';
        }
    template <typename T> void mult(const std::vector<T> &X, std::vector<T> &Y)
        {
        static_assert(std::is_same<T, double>::value, "the device path is FP64");
        fgb200::check(fg_matrix_mult(device(), X.data(), Y.data()), "SparseMatrix::mult");
        }
    template <typename T> void build_diag_precond(std::vector<T> &D) const
        {
        for (size_t i = 0; i + 1 < rowptr.size(); ++i) D[i] = 1.0 / (*this)((int)i, (int)i);
        }
    size_t size() const { return rowptr.size() - 1; }
    // the device mirror (uploaded on demand), for the solvers below
    fg_matrix *device(int dev_id = 0)
        {
        if (!dev)
            fgb200::check(fg_matrix_create((int)size(), rowptr.data(), col.data(), dev_id, &dev),
                          "SparseMatrix: fg_matrix_create");
        if (dirty)
            {
            fgb200::check(fg_matrix_set_values(dev, val.data()), "SparseMatrix: fg_matrix_set_values");
            dirty = false;
            }
        return dev;
        }

private:
    int find(int i, int j, bool must) const
        {
        const auto b = col.begin() + rowptr[(size_t)i], e = col.begin() + rowptr[(size_t)i + 1];
        const auto it = std::lower_bound(b, e, j);
        if (it == e || *it != j)
            {
            if (must) throw std::out_of_range("SparseMatrix: (i,j) outside the shape");
            return -1;
            }
        return (int)(it - col.begin());
        }
    std::vector<int> rowptr, col;
    std::vector<double> val;
    fg_matrix *dev = nullptr;
    bool dirty = true;
    };

template <typename T> void mult(SparseMatrix &A, const std::vector<T> &X, std::vector<T> &Y) { A.mult(X, Y); }

namespace detail
{
template <typename T> void store(iteration<T> &iter, const fg_iter_result &r)
    { iter.assign(r.status, r.iters, r.res, r.rhsnorm); }
}

// src/algebra/bicg.h:14-72
template <typename T>
void bicg(iteration<T> &iter, SparseMatrix &A, std::vector<T> &x, const std::vector<T> &rhs)
    {
    fg_iter_result r;
    fgb200::check(fg_bicg(A.device(), x.data(), rhs.data(), iter.resmax, iter.get_maxiter(), &r), "bicg");
    detail::store(iter, r);
    }
// src/algebra/bicg.h:83-154
template <typename T>
void bicg_dir(iteration<T> &iter, SparseMatrix &A, std::vector<T> &x, const std::vector<T> &rhs,
              const std::vector<T> &xd, const std::vector<int> &ld)
    {
    fg_iter_result r;
    fgb200::check(fg_bicg_dir(A.device(), x.data(), rhs.data(), xd.data(), ld.data(), (int)ld.size(),
                              iter.resmax, iter.get_maxiter(), &r), "bicg_dir");
    detail::store(iter, r);
    }
// src/algebra/bicg.h:163-234 ; returns the residual like the reference
template <typename T>
T bicg_dir(iteration<T> &iter, SparseMatrix &A, std::vector<T> &x, const std::vector<T> &rhs,
           const std::vector<int> &ld)
    {
    fg_iter_result r;
    fgb200::check(fg_bicg_dir(A.device(), x.data(), rhs.data(), nullptr, ld.data(), (int)ld.size(),
                              iter.resmax, iter.get_maxiter(), &r), "bicg_dir");
    detail::store(iter, r);
    return r.res / r.rhsnorm;
    }
// src/algebra/cg.h:15-58
template <typename T>
void cg(iteration<T> &iter, SparseMatrix &A, std::vector<T> &x, const std::vector<T> &rhs)
    {
    fg_iter_result r;
    fgb200::check(fg_cg(A.device(), x.data(), rhs.data(), iter.resmax, iter.get_maxiter(), &r), "cg");
    detail::store(iter, r);
    }
// src/algebra/cg.h:68-121
template <typename T>
void cg_dir(iteration<T> &iter, SparseMatrix &A, std::vector<T> &x, const std::vector<T> &rhs,
            const std::vector<T> &xd, const std::vector<int> &ld)
    {
    fg_iter_result r;
    fgb200::check(fg_cg_dir(A.device(), x.data(), rhs.data(), xd.data(), ld.data(), (int)ld.size(),
                            iter.resmax, iter.get_maxiter(), &r), "cg_dir");
    detail::store(iter, r);
    }
}  // namespace algebra

// ---------------------------------------------------------------------------------------------
// LinAlgebra — reference src/linear_algebra.h:33-121 + solver<2> (src/solver.h:26-143)
// ---------------------------------------------------------------------------------------------
class LinAlgebra
    {
public:
    LinAlgebra(fgb200::Settings &s, fgb200::MeshView &my_msh, int device = 0)
        : iter("bicg_dir", s.TOL, s.verbose != 0, s.MAXITER), NOD(my_msh.NOD()), verbose(s.verbose)
        {
        fg_mesh m;
        m.NOD = my_msh.NOD();
        m.node_p = my_msh.node_p.data();
        m.NT = my_msh.NT();
        m.tet_ind = my_msh.tet_ind.data();
        m.tet_reg = my_msh.tet_reg.data();
        m.NF = my_msh.NF();
        m.tri_ind = my_msh.tri_ind.data();
        m.tri_reg = my_msh.tri_reg.data();
        m.tri_dMs = my_msh.tri_dMs.data();
        fg_params p;
        p.nreg_tet = (int)s.paramTetra.size();
        p.prm_tet = s.paramTetra.data();
        p.nreg_tri = (int)s.paramTriangle.size();
        p.prm_tri = s.paramTriangle.data();
        p.npi_tet = s.npi_tet;
        p.npi_tri = s.npi_tri;
        p.tol = s.TOL;
        p.maxiter = s.MAXITER;
        fgb200::check(fg_create(&m, &p, device, &ctx), "LinAlgebra");
        idx_dir = s.recenter ? s.recentering_direction : FG_IDX_UNDEF;  // linear_algebra.h:50-53
        }
    LinAlgebra(const LinAlgebra &) = delete;
    LinAlgebra &operator=(const LinAlgebra &) = delete;
    ~LinAlgebra() { fg_destroy(ctx); }

    /** src/linear_algebra.cpp:3-11: mt19937 seeded with rand(), one U(0,1) draw, times M_2_PI */
    void base_projection() const
        {
        std::mt19937 gen(rand());
        std::uniform_real_distribution<> distrib(0.0, 1.0);
        const double r = distrib(gen);
        last_angle = M_2_PI * r;
        fgb200::check(fg_base_projection(ctx, last_angle), "base_projection");
        }
    /** src/linear_algebra.cpp:26-52 */
    template <class Vector3> void prepareElements(const Vector3 &Hext, const timing &t_prm) const
        {
        fgb200::check(fg_prepare_elements(ctx, Hext.data(), t_prm.get_dt(), t_prm.prefactor, idx_dir,
                                          DW_vz), "prepareElements");
        dt_of_last_prepare = t_prm.get_dt();
        }
    /** src/linear_algebra.cpp:54-81 (mesh.extSpaceField given once with set_ext_space_field) */
    void prepareElements(const double A_Hext, const timing &t_prm) const
        {
        fgb200::check(fg_prepare_elements_space(ctx, A_Hext, t_prm.get_dt(), t_prm.prefactor, idx_dir,
                                                DW_vz), "prepareElements");
        dt_of_last_prepare = t_prm.get_dt();
        }
    /** src/solver.cpp:6-90 ; true = failure */
    bool solve(const timing &t_prm)
        {
        fg_step_result r;
        fgb200::check(fg_solve(ctx, t_prm.get_dt(), &r), "solve");
        iter.assign(r.status, r.iters, r.res, r.rhsnorm);
        v_max = r.v_max;
        if (r.failed && verbose) std::cout << "solver: " << iter.infos() << std::endl;
        return r.failed != 0;
        }
    /** src/linear_algebra.cpp:13-24 (for the system of the last prepareElements) */
    void buildInitGuess(std::vector<double> &G) const
        {
        G.resize(2 * (size_t)NOD);
        fgb200::check(fg_get_system(ctx, dt_of_last_prepare, nullptr, nullptr, G.data()), "buildInitGuess");
        }
    inline void set_DW_vz(const double vz) { DW_vz = vz; }
    inline double get_v_max(void) const { return v_max; }
    void checkBoundaryConditions(void) const {}

    // ---- the mesh-side state the reference keeps in Mesh::mesh ----
    void set_state(const double *u, const double *v, const double *phi, const double *phiv)
        { fgb200::check(fg_set_state(ctx, u, v, phi, phiv), "set_state"); }
    void set_potentials(const double *phi, const double *phiv)
        { fgb200::check(fg_set_potentials(ctx, phi, phiv), "set_potentials"); }
    void get_state(int step, double *u, double *v, double *phi, double *phiv) const
        { fgb200::check(fg_get_state(ctx, step, u, v, phi, phiv), "get_state"); }
    void evolution() { fgb200::check(fg_commit(ctx), "evolution"); }
    void set_ext_space_field(const double *field)
        { fgb200::check(fg_set_ext_space_field(ctx, field), "set_ext_space_field"); }

    algebra::iteration<double> iter;   // solver<DIM>::iter, src/solver.h:66
    fg_ctx *handle() const { return ctx; }
    mutable double last_angle = 0.0;
    mutable double dt_of_last_prepare = 0.0;

private:
    fg_ctx *ctx = nullptr;
    const int NOD;
    int idx_dir = FG_IDX_UNDEF;
    const int verbose;
    double DW_vz = 0.0;   // never initialised in the reference: 0 here
    double v_max = 0.0;
    };
