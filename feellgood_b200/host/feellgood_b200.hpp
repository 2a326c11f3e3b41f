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
//   template <int DIM> class solver      src/solver.h:20-143               (build_shape, buildMat<N>,
//                                        buildVect<N>, K, L_rhs, iter: the base of the side solvers)
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
    // what Fem::time_integration / Fem::saver read (src/default-settings.yml:43,63-71,239)
    double time_step = 5e-13;               // outputs.evol_time_step
    double DUMAX = 0.02;                    // time_integration.max(du)
    std::vector<std::string> evol_columns = {"t", "<Mx>", "<My>", "<Mz>", "E_ex", "E_demag", "E_zeeman", "E_tot"};
    std::vector<std::string> region_names;  // volume region names, index = region (for "name:<Mx>")
    // applied field: uniform (RtoR3, Settings::getField) or amplitude of mesh.extSpaceField (R4toR3,
    // Settings::getFieldTime); set exactly one
    std::function<Vec3(double)> field;
    std::function<double(double)> field_time;
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
    int getNbNodes() const { return NOD(); }
    // Mesh::mesh::edges (src/mesh.h:100-114): every tetrahedron edge once, as a sorted pair, sorted
    using Edge = std::pair<int, int>;
    std::vector<Edge> edges;
    void build_edges()
        {
        edges.clear();
        edges.reserve(tet_reg.size() * 6);
        for (size_t t = 0; t < tet_reg.size(); t++)
            for (int i = 0; i < 3; ++i)
                for (int j = i + 1; j < 4; ++j) edges.push_back(std::minmax(tet_ind[4 * t + i], tet_ind[4 * t + j]));
        std::sort(edges.begin(), edges.end());
        edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
        edges.shrink_to_fit();
        }
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
// solver<DIM_PROBLEM> — reference src/solver.h:20-143, the base of the side solvers (electrostatSolver:
// DIM 1 with cg_dir, spinAcc: DIM 3 with bicg_dir).  Same protected surface: build_shape, buildMat<N>,
// buildVect<N>, members msh, NOD, verbose, iter, K, L_rhs; K is an algebra::SparseMatrix whose mult and
// Krylov solvers run on the GPU (fg_matrix_*).  The element matrix type only needs operator()(i, j)
// (Eigen::Matrix qualifies; fgb200::Dense below is the Eigen-free stand-in).  The material parameter
// vectors of the reference's constructor are not used by this base class and are left to the derived
// solver.
// ---------------------------------------------------------------------------------------------
namespace fgb200
{
template <int R, int C = R> struct Dense
    {
    double a[R * C] = {};
    double &operator()(int i, int j) { return a[i * C + j]; }
    double operator()(int i, int j) const { return a[i * C + j]; }
    };
}  // namespace fgb200

template <int DIM_PROBLEM> class solver
    {
public:
    explicit solver(fgb200::MeshView &_msh, const std::string &name, const double _tol, const bool v,
                    const int max_iter,
                    const std::function<bool(fgb200::MeshView::Edge)> &edge_filter = [](fgb200::MeshView::Edge)
                        { return true; })
        : msh(&_msh), NOD(_msh.getNbNodes()), verbose(v), iter(name, _tol, v, max_iter),
          K(build_shape(edge_filter)), L_rhs((size_t)DIM_PROBLEM * _msh.getNbNodes()) {}
    virtual ~solver() = default;
    /** check boundary conditions (src/solver.h:43) */
    virtual void checkBoundaryConditions(void) const = 0;

protected:
    static const int DIM_PB = DIM_PROBLEM;
    fgb200::MeshView *msh;
    const int NOD;
    const bool verbose;
    algebra::iteration<double> iter;
    algebra::SparseMatrix K;
    std::vector<double> L_rhs;

    /** src/solver.h:75-104: one DIM x DIM block per node and per relevant edge, both directions */
    algebra::MatrixShape build_shape(const std::function<bool(fgb200::MeshView::Edge)> &edge_filter) const
        {
        if (msh->edges.empty() && msh->NT() > 0) msh->build_edges();
        algebra::MatrixShape shape((size_t)DIM_PROBLEM * NOD);
        auto add_block = [&shape](const int i, const int j)
            {
            for (int k = 0; k < DIM_PROBLEM; ++k)
                for (int l = 0; l < DIM_PROBLEM; ++l) shape[(size_t)DIM_PROBLEM * i + k].insert(DIM_PROBLEM * j + l);
            };
        for (int i = 0; i < NOD; ++i) add_block(i, i);
        for (auto edge : msh->edges)
            if (edge_filter(edge))
                {
                add_block(edge.first, edge.second);
                add_block(edge.second, edge.first);
                }
        return shape;
        }
    /** src/solver.h:110-129 */
    template <int N, class Mat> void buildMat(std::array<int, N> &ind, Mat &Ke)
        {
        for (int ie = 0; ie < N; ie++)
            for (int je = 0; je < N; je++)
                for (int di = 0; di < DIM_PROBLEM; di++)
                    for (int dj = 0; dj < DIM_PROBLEM; dj++)
                        K.add(DIM_PROBLEM * ind[ie] + di, DIM_PROBLEM * ind[je] + dj, Ke(di * N + ie, dj * N + je));
        }
    /** src/solver.h:135-143 */
    template <int N> void buildVect(std::array<int, N> &ind, std::vector<double> &Le)
        {
        for (int ie = 0; ie < N; ie++)
            for (int di = 0; di < DIM_PROBLEM; di++) L_rhs[(size_t)DIM_PROBLEM * ind[ie] + di] += Le[(size_t)di * N + ie];
        }
    };

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
    /** Fem::energy terms (src/energy.cpp:5-68) of the NEXT state: exchange, anisotropy, demag, zeeman */
    template <class Vector3> std::array<double, 4> energy(const Vector3 &Hext) const
        {
        std::array<double, 4> E;
        fgb200::check(fg_energy(ctx, Hext.data(), E.data()), "energy");
        return E;
        }
    std::array<double, 4> energy(const double A_Hext) const
        {
        std::array<double, 4> E;
        fgb200::check(fg_energy_space(ctx, A_Hext, E.data()), "energy");
        return E;
        }
    /** mesh::avg (src/mesh.cpp:89-106): what = 0 u | 1 v, all three components */
    std::array<double, 3> avg(int what, int region = -1) const
        {
        std::array<double, 3> a;
        fgb200::check(fg_avg(ctx, what, region, a.data()), "avg");
        return a;
        }
    /** mesh::max_angle (src/mesh.h:295-306) */
    double max_angle() const
        {
        double a = 0.0;
        fgb200::check(fg_max_angle(ctx, &a), "max_angle");
        return a;
        }

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

// ---------------------------------------------------------------------------------------------
// TimeStepper — reference src/time_integration.cpp:11-38
// ---------------------------------------------------------------------------------------------
class TimeStepper
    {
    const double hard_min;
    const double hard_max;
    double soft_max;

public:
    TimeStepper(const double initial, const double min, const double max)
        : hard_min(min), hard_max(max * (1 + std::numeric_limits<float>::epsilon())), soft_max(initial)
        {
        }
    void set_soft_limit(const double max) { soft_max = std::min(soft_max, max); }
    double operator()(const double stride)
        {
        double step = std::min(stride, soft_max);
        if (step > stride - 2 * hard_min && step < stride) step = stride - 2 * hard_min;
        soft_max = std::max(soft_max, std::min(step * 1.1, hard_max));
        return step;
        }
    };

// LogStats — reference src/log-stats.h (Welford on the logarithms)
class LogStats
    {
public:
    void add(double x)
        {
        x = std::log(x);
        n += 1;
        double delta1 = x - m;
        m += delta1 / n;
        double delta2 = x - m;
        s += delta1 * delta2;
        }
    long count() const { return n; }
    double mean() const { return std::exp(m); }
    double stddev() const { return std::sqrt(s / n); }

private:
    long n = 0;
    double m = 0;
    double s = 0;
    };

// Stats — reference src/time_integration.cpp:41-47
struct Stats
    {
    LogStats good_dt, good_dumax, bad_dt;
    double max_angle = 0.0;
    };

const int NB_ENERGY_TERMS = 4;
enum ENERGY_TYPE { EXCHANGE = 0, ANISOTROPY = 1, DEMAG = 2, ZEEMAN = 3 };  // src/fem.h:27-36

// ---------------------------------------------------------------------------------------------
// Fem — the part of the reference's Fem that surrounds the hot path: vmax, E, Etot, Etot0, energy,
// evolution, compute_all, saver (src/fem.h, src/energy.cpp, src/save.cpp:13-140) and
// Fem::time_integration (src/time_integration.cpp:133-245).  The mesh state lives in the
// LinAlgebra's device context; the demag solver (ScalFMM in the reference, outside the path) is the
// callback run where compute_all calls myFMM.calc_demag: it reads the NEXT magnetisation
// (LinAlgebra::get_state) and writes phi / phiv (LinAlgebra::set_potentials).
// ---------------------------------------------------------------------------------------------
class Fem
    {
public:
    Fem(fgb200::Settings &s, LinAlgebra &la, std::function<void(LinAlgebra &)> demag_solver = nullptr)
        : settings(s), linAlg(la), demag(std::move(demag_solver))
        {
        E.fill(0.0);
        }

    double vmax = 0.0;
    std::array<double, NB_ENERGY_TERMS> E;
    double Etot0 = INFINITY;  // avoid "WARNING: energy increased" on first time step
    double Etot = 0.0;
    Stats stats;
    std::vector<std::vector<double>> evol;  // rows written by saver (also streamed to `out`)

    /** src/energy.cpp:5-68 */
    void energy(const double t)
        {
        if (settings.field_time)
            E = linAlg.energy(settings.field_time(t));
        else
            E = linAlg.energy(settings.field ? settings.field(t) : fgb200::Vec3{{0.0, 0.0, 0.0}});
        Etot = 0.0 + ((E[0] + E[1]) + (E[2] + E[3]));  // std::reduce on 4 doubles (libstdc++)
        if (settings.verbose && (Etot > Etot0))
            std::cout << "WARNING: energy increased from " << Etot0 << " to " << Etot << "\n";
        }
    /** src/fem.h:159-163 */
    void evolution()
        {
        linAlg.evolution();
        Etot0 = Etot;
        }
    /** src/fem.h:203-224 */
    void compute_all(const double t)
        {
        if (demag) demag(linAlg);
        energy(t);
        evolution();
        }
    /** the .evol row of Fem::saver, src/save.cpp:13-140 */
    void saver(const timing &t_prm, std::ostream *out, const int nt)
        {
        std::vector<double> row;
        std::array<double, 3> au{}, av{}, H{};
        int au_reg = -2, av_reg = -2;
        bool have_H = false;
        for (const std::string &col_name : settings.evol_columns)
            {
            int region = -1;
            std::string keyVal = col_name;
            const std::string::size_type colon_pos = col_name.rfind(':');
            if (colon_pos != std::string::npos)
                {
                const std::string region_name = col_name.substr(0, colon_pos);
                const auto it = std::find(settings.region_names.begin(), settings.region_names.end(), region_name);
                if (it == settings.region_names.end())
                    throw std::runtime_error("Error: no region named '" + region_name + "'");
                region = (int)(it - settings.region_names.begin());
                keyVal = col_name.substr(colon_pos + 1);
                }
            auto U = [&](int k) { if (au_reg != region) { au = linAlg.avg(0, region); au_reg = region; } return au[k]; };
            auto V = [&](int k) { if (av_reg != region) { av = linAlg.avg(1, region); av_reg = region; } return av[k]; };
            auto Hk = [&](int k)
                {
                if (!have_H)
                    {
                    const fgb200::Vec3 h = settings.field ? settings.field(t_prm.get_t()) : fgb200::Vec3{{NAN, NAN, NAN}};
                    H = {h.v[0], h.v[1], h.v[2]};
                    have_H = true;
                    }
                return H[k];
                };
            if (keyVal == "iter") row.push_back(nt);
            else if (keyVal == "t") row.push_back(t_prm.get_t());
            else if (keyVal == "dt") row.push_back(t_prm.get_dt());
            else if (keyVal == "max_dm") row.push_back(vmax * t_prm.get_dt());
            else if (keyVal == "max_angle") row.push_back(linAlg.max_angle());
            else if (keyVal == "<Mx>") row.push_back(U(0));
            else if (keyVal == "<My>") row.push_back(U(1));
            else if (keyVal == "<Mz>") row.push_back(U(2));
            else if (keyVal == "<dMx/dt>") row.push_back(V(0));
            else if (keyVal == "<dMy/dt>") row.push_back(V(1));
            else if (keyVal == "<dMz/dt>") row.push_back(V(2));
            else if (keyVal == "E_ex") row.push_back(E[EXCHANGE]);
            else if (keyVal == "E_aniso") row.push_back(E[ANISOTROPY]);
            else if (keyVal == "E_demag") row.push_back(E[DEMAG]);
            else if (keyVal == "E_zeeman") row.push_back(E[ZEEMAN]);
            else if (keyVal == "E_tot") row.push_back(Etot);
            else if (keyVal == "Hx") row.push_back(Hk(0));
            else if (keyVal == "Hy") row.push_back(Hk(1));
            else if (keyVal == "Hz") row.push_back(Hk(2));
            else throw std::runtime_error("Error: invalid column name '" + keyVal + "'");
            }
        if (out)
            {
            for (size_t i = 0; i < row.size(); i++) *out << row[i] << (i + 1 == row.size() ? "\n" : "\t");
            *out << std::flush;
            }
        evol.push_back(std::move(row));
        }

    /** src/time_integration.cpp:133-245; `out` receives the .evol rows (precision 16 like the
     * reference), `stop` stands for exit_if_signal_received.  Returns the exit status. */
    int time_integration(timing &t_prm, int &nt, std::ostream *out = nullptr,
                         const std::function<bool()> &stop = nullptr)
        {
        compute_all(t_prm.get_t());
        if (out)
            {
            *out << "## columns: ";
            for (size_t i = 0; i + 1 < settings.evol_columns.size(); i++) *out << settings.evol_columns[i] << '\t';
            *out << settings.evol_columns.back() << "\n";
            out->precision(16);
            }
        int flag(0);
        int nt_output(0);
        int status(0);
        double t_initial = t_prm.get_t();
        double t_step = settings.time_step;
        int step_count = (int)std::round((t_prm.tf - t_initial) / t_step);
        TimeStepper stepper(t_prm.get_dt(), t_prm.DTMIN, t_prm.DTMAX);
        stats = Stats();
        stats.max_angle = linAlg.max_angle();
        nt = 0;
        for (int step_nb = 0; step_nb <= step_count; step_nb++)
            {
            double t_target = t_initial + step_nb * t_step;
            while (t_prm.get_t() < t_target)
                {
                if (stop && stop()) return 1;
                t_prm.set_dt(stepper(t_target - t_prm.get_t()));
                bool last_step = (t_prm.get_dt() == t_target - t_prm.get_t());
                if (settings.verbose)
                    {
                    std::cout << std::string(64, '-') << '\n';
                    if (flag) std::cout << "  TRYING AGAIN with a smaller time step: retry " << flag << '\n';
                    std::cout << "evol step = " << nt_output << ", step = " << nt << ", t = " << t_prm.get_t()
                              << ", dt = " << t_prm.get_dt() << '\n';
                    }
                if (t_prm.is_dt_TooSmall())
                    {
                    std::cout << "\n**ABORTED**: dt < DTMIN\n";
                    return 1;
                    }
                linAlg.base_projection();
                if (settings.field_time)
                    linAlg.prepareElements(settings.field_time(t_prm.get_t()), t_prm);
                else
                    linAlg.prepareElements(settings.field ? settings.field(t_prm.get_t()) : fgb200::Vec3{{0.0, 0.0, 0.0}}, t_prm);
                bool err = linAlg.solve(t_prm);
                vmax = linAlg.get_v_max();
                if (err)
                    {
                    flag++;
                    stepper.set_soft_limit(t_prm.get_dt() / 2);
                    stats.bad_dt.add(t_prm.get_dt());
                    continue;
                    }
                double dumax = t_prm.get_dt() * vmax;
                stats.good_dt.add(t_prm.get_dt());
                stats.good_dumax.add(dumax);
                if (settings.verbose) std::cout << "  -> dumax = " << dumax << ",  vmax = " << vmax << std::endl;
                stepper.set_soft_limit((settings.DUMAX / vmax) * 0.95);
                if (dumax > settings.DUMAX)
                    {
                    flag++;
                    continue;
                    }
                compute_all(t_prm.get_t());
                nt++;
                flag = 0;
                // Prevent rounding errors from making us miss the target.
                if (last_step)
                    t_prm.set_t(t_target);
                else
                    t_prm.inc_t();
                stats.max_angle = std::max(stats.max_angle, linAlg.max_angle());
                }
            saver(t_prm, out, nt_output++);
            }
        return status;
        }

private:
    fgb200::Settings &settings;
    LinAlgebra &linAlg;
    std::function<void(LinAlgebra &)> demag;
    };
