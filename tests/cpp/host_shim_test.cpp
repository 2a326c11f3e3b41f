// host_shim_test.cpp — drives the C++17 drop-in (feellgood_b200/host/feellgood_b200.hpp) the way
// the reference's Fem::time_integration drives LinAlgebra (src/time_integration.cpp:193-206), and
// runs ports of the reference's algebra unit tests (unit-tests/ut_algebra.cpp:160-330) through the
// algebra:: surface.  Needs a GPU; with `--compile-only` semantics the CPU suite just builds it.
#include <cstdio>
#include <cstring>

#include "../../feellgood_b200/host/feellgood_b200.hpp"

static int n_fail = 0;
#define CHECK(c)                                                   \
    do                                                             \
        {                                                          \
        if (!(c))                                                  \
            {                                                      \
            n_fail++;                                              \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); \
            }                                                      \
        } while (0)

// the gmsh-free Cuboid scheme of python-modules/meshMaker.py:440-519
static fgb200::MeshView cuboid(int nx, int ny, int nz, double h)
    {
    fgb200::MeshView m;
    auto nid = [&](int i, int j, int k) { return (nz + 1) * (ny + 1) * i + (nz + 1) * j + k; };
    for (int i = 0; i <= nx; i++)
        for (int j = 0; j <= ny; j++)
            for (int k = 0; k <= nz; k++)
                for (double c : {i * h, j * h, k * h}) m.node_p.push_back(c);
    static const int T[6][4][3] = {
        {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 1, 1}}, {{0, 0, 0}, {1, 0, 0}, {0, 0, 1}, {0, 1, 1}},
        {{0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 0, 0}}, {{1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 1, 1}},
        {{1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 1}}, {{1, 0, 1}, {1, 0, 0}, {1, 1, 1}, {0, 1, 1}}};
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++)
            for (int k = 0; k < nz; k++)
                for (auto &t : T)
                    {
                    for (auto &c : t) m.tet_ind.push_back(nid(i + c[0], j + c[1], k + c[2]));
                    m.tet_reg.push_back(1);
                    }
    return m;
    }

static void test_llg_loop()
    {
    fgb200::MeshView msh = cuboid(8, 4, 3, 2e-9);
    fgb200::Settings s;
    fg_tet_prm def{};
    fg_tet_prm py{};
    py.alpha_LLG = 0.02; py.A = 1.3e-11; py.Ms = 8e5; py.K = 0; py.K3 = 0;
    py.uk[2] = 1; py.ex[0] = 1; py.ey[1] = 1; py.ez[2] = 1;
    def = py; def.Ms = 795774.7;
    s.paramTetra = {def, py};
    s.paramTriangle = {fg_tri_prm{}};
    srand(2);                                       // --seed 2 (ci-tests/full_test.py:54)
    LinAlgebra linAlg(s, msh);
    const int NOD = msh.NOD();
    std::vector<double> u(3 * (size_t)NOD), un(3 * (size_t)NOD), vn(3 * (size_t)NOD), phi(NOD, 0.0);
    for (int a = 0; a < NOD; a++)
        {
        const double x = msh.node_p[3 * a] * 1e8;
        const double c[3] = {std::cos(x), std::sin(x), 0.2};
        const double nn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int k = 0; k < 3; k++) u[3 * a + k] = c[k] / nn;
        }
    linAlg.set_state(u.data(), nullptr, nullptr, nullptr);
    timing t_prm(2e-11, 1e-16, 5e-13);
    t_prm.set_dt(1e-13);
    const fgb200::Vec3 Hext{{0.0, 8000.0, 0.0}};
    for (int step = 0; step < 5; step++)
        {
        linAlg.base_projection();
        linAlg.prepareElements(Hext, t_prm);
        std::vector<double> G;
        if (step == 1)
            {
            linAlg.buildInitGuess(G);
            CHECK(G.size() == 2 * (size_t)NOD);
            }
        const bool err = linAlg.solve(t_prm);
        CHECK(!err);
        CHECK(linAlg.iter.status == algebra::CONVERGED);
        CHECK(linAlg.iter.get_res() <= linAlg.iter.get_rhsnorm() * linAlg.iter.resmax);
        CHECK(linAlg.get_v_max() > 0.0);
        linAlg.get_state(1, un.data(), vn.data(), nullptr, nullptr);
        linAlg.set_potentials(phi.data(), phi.data());  // what the demag solver would write
        linAlg.evolution();
        t_prm.inc_t();
        }
    double worst = 0.0, moved = 0.0;
    for (int a = 0; a < NOD; a++)
        {
        double nn = 0.0, d = 0.0;
        for (int k = 0; k < 3; k++)
            {
            nn += un[3 * a + k] * un[3 * a + k];
            d += std::fabs(un[3 * a + k] - u[3 * a + k]);
            }
        worst = std::max(worst, std::fabs(std::sqrt(nn) - 1.0));
        moved = std::max(moved, d);
        }
    CHECK(worst < 1e-15);
    CHECK(moved > 1e-6);
    std::printf("llg loop: %s v_max=%.6g\n", linAlg.iter.infos().c_str(), linAlg.get_v_max());
    }

// Fem::time_integration through the C++ drop-in: visible steps land on their targets, accepted steps
// respect DUMAX, E_tot is the sum of its terms and decreases under damping in a fixed field.
static void test_fem_time_integration()
    {
    fgb200::MeshView msh = cuboid(8, 4, 3, 2e-9);
    fgb200::Settings s;
    fg_tet_prm py{};
    py.alpha_LLG = 0.5; py.A = 1.3e-11; py.Ms = 8e5; py.K = 1e4; py.K3 = 0;
    py.uk[2] = 1; py.ex[0] = 1; py.ey[1] = 1; py.ez[2] = 1;
    fg_tet_prm def = py;
    s.paramTetra = {def, py};
    s.paramTriangle = {fg_tri_prm{}};
    s.time_step = 1e-12;
    s.DUMAX = 0.05;
    s.evol_columns = {"iter", "t", "dt", "max_dm", "max_angle", "<Mx>", "<My>", "<Mz>", "<dMz/dt>",
                      "E_ex", "E_aniso", "E_demag", "E_zeeman", "E_tot", "Hy"};
    s.field = [](double) { return fgb200::Vec3{{0.0, 8000.0, 0.0}}; };
    srand(2);
    LinAlgebra linAlg(s, msh);
    const int NOD = msh.NOD();
    std::vector<double> u(3 * (size_t)NOD);
    for (int a = 0; a < NOD; a++)
        {
        const double x = msh.node_p[3 * a] * 1e8;
        const double c[3] = {std::cos(x), 0.2, std::sin(x)};
        const double nn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int k = 0; k < 3; k++) u[3 * a + k] = c[k] / nn;
        }
    linAlg.set_state(u.data(), nullptr, nullptr, nullptr);
    std::vector<double> un(3 * (size_t)NOD), phi(NOD), phiv(NOD, 0.0);
    int demag_calls = 0;
    Fem fem(s, linAlg, [&](LinAlgebra &la)
        {  // stand-in for myFMM.calc_demag: a local function of the new magnetisation
        la.get_state(1, un.data(), nullptr, nullptr, nullptr);
        for (int a = 0; a < NOD; a++) phi[a] = 1e-3 * un[3 * a + 2];
        la.set_potentials(phi.data(), phiv.data());
        demag_calls++;
        });
    timing t_prm(3e-12, 1e-16, 5e-13);
    int nt = -1;
    std::ostringstream evol;
    const int status = fem.time_integration(t_prm, nt, &evol);
    CHECK(status == 0);
    CHECK(nt >= 6);
    CHECK(demag_calls == nt + 1);
    CHECK(fem.evol.size() == 4);
    for (size_t k = 0; k < fem.evol.size(); k++)
        {
        const std::vector<double> &r = fem.evol[k];
        CHECK(r.size() == 15);
        CHECK(r[0] == (double)k);
        CHECK(r[1] == (double)k * 1e-12 || std::fabs(r[1] - k * 1e-12) < 1e-27);
        CHECK(r[3] <= s.DUMAX);
        CHECK(r[13] == 0.0 + ((r[9] + r[10]) + (r[11] + r[12])));
        CHECK(r[14] == 8000.0);
        if (k > 0) CHECK(r[13] < fem.evol[k - 1][13]);
        }
    CHECK(t_prm.get_t() == 3e-12);
    CHECK(fem.stats.good_dt.count() >= nt);
    CHECK(fem.stats.max_angle > 0.0 && fem.stats.max_angle < 3.2);
    CHECK(evol.str().rfind("## columns: iter\tt\tdt", 0) == 0);
    std::printf("fem loop: nt=%d good=%ld bad=%ld <dt>=%.3e Etot %.6e -> %.6e\n", nt, fem.stats.good_dt.count(),
                fem.stats.bad_dt.count(), fem.stats.good_dt.mean(), fem.evol.front()[13], fem.evol.back()[13]);
    }

// TimeStepper / LogStats replay for the bit-exact comparison with the reference's own classes
// (tests/test_host_shim.py feeds the call sequence of tests/golden/ref_timestepper.npz)
static int replay_timestepper()
    {
    double init, mn, mx;
    if (std::scanf("%lf %lf %lf", &init, &mn, &mx) != 3) return 2;
    TimeStepper ts(init, mn, mx);
    LogStats ls;
    int op;
    double val;
    while (std::scanf("%d %lf", &op, &val) == 2)
        {
        if (op == 0)
            std::printf("%.17g\n", ts(val));
        else if (op == 1)
            {
            ts.set_soft_limit(val);
            std::printf("nan\n");
            }
        else
            {
            ls.add(val);
            std::printf("%ld %.17g %.17g\n", ls.count(), ls.mean(), ls.stddev());
            }
        }
    return 0;
    }

// unit-tests/ut_algebra.cpp:160-190 (test_cg) and :245-275 (test_bicg): identity solves in 0 iterations
static void test_identity_solves()
    {
    const int N = 1000;
    algebra::MatrixShape shape(N);
    for (int i = 0; i < N; i++) shape[i].insert(i);
    algebra::SparseMatrix A(shape);
    for (int i = 0; i < N; i++) A.set(i, i, 1.0);
    std::vector<double> x(N, 0.0), b(N);
    for (int i = 0; i < N; i++) b[i] = 1.0 + i;
    algebra::iteration<double> it("cg", 1e-6, false, 100);
    x = b;
    algebra::cg(it, A, x, b);
    CHECK(it.status == algebra::CONVERGED && it.get_iteration() == 0);
    algebra::iteration<double> it2("bicg", 1e-6, false, 100);
    algebra::bicg(it2, A, x, b);
    CHECK(it2.status == algebra::CONVERGED && it2.get_iteration() == 0);
    }

// unit-tests/ut_algebra.cpp:192-243 (test_cg_dir) / :277-330 (test_bicg_dir): 1-D Laplacian with
// Dirichlet values 0 and 1 at the ends -> linear ramp
static void test_laplacian_dirichlet()
    {
    const int NOD = 1001;
    algebra::MatrixShape shape(NOD);
    for (int i = 0; i < NOD; i++)
        for (int j : {i - 1, i, i + 1})
            if (j >= 0 && j < NOD) shape[i].insert(j);
    algebra::SparseMatrix A(shape);
    for (int i = 0; i < NOD; i++)
        {
        A.set(i, i, 2.0);
        if (i > 0) A.set(i, i - 1, -1.0);
        if (i < NOD - 1) A.set(i, i + 1, -1.0);
        }
    CHECK(A(3, 4) == -1.0 && A(3, 7) == 0.0);
    std::vector<double> x(NOD, 0.0), rhs(NOD, 0.0), xd(NOD, 0.0);
    std::vector<int> ld = {0, NOD - 1};
    xd[NOD - 1] = 1.0;
    algebra::iteration<double> it("cg_dir", 1e-10, false, 5000);
    algebra::cg_dir(it, A, x, rhs, xd, ld);
    CHECK(it.status == algebra::CONVERGED);
    double worst = 0.0;
    for (int i = 0; i < NOD; i++) worst = std::max(worst, std::fabs(x[i] - (double)i / (NOD - 1)));
    CHECK(worst < 1e-6);
    std::fill(x.begin(), x.end(), 0.0);
    algebra::iteration<double> it2("bicg_dir", 1e-10, false, 5000);
    algebra::bicg_dir(it2, A, x, rhs, xd, ld);
    CHECK(it2.status == algebra::CONVERGED);
    worst = 0.0;
    for (int i = 0; i < NOD; i++) worst = std::max(worst, std::fabs(x[i] - (double)i / (NOD - 1)));
    CHECK(worst < 1e-6);
    // SparseMatrix::mult
    std::vector<double> y(NOD);
    algebra::mult(A, x, y);
    double r2 = 0.0;
    for (int i = 1; i < NOD - 1; i++) r2 += y[i] * y[i];
    CHECK(std::sqrt(r2) < 1e-8);
    }

// A side solver written against the reference's solver<DIM> surface (src/solver.h:20-143), the way
// electrostatSolver does (src/electrostatSolver.h:22-35, electrostatSolver.cpp): DIM 1, P1 Laplacian
// assembled element by element with buildMat<4> / buildVect<4>, Dirichlet values through cg_dir.
class laplaceSolver : public solver<1>
    {
public:
    laplaceSolver(fgb200::MeshView &m, double tol, int max_iter) : solver<1>(m, "cg_dir", tol, false, max_iter) {}
    void checkBoundaryConditions(void) const override {}
    // grad-grad stiffness of every tetrahedron; x = 0 plane held at 0, x = xmax plane at 1
    bool run(std::vector<double> &V, double xmax)
        {
        K.clear();
        std::fill(L_rhs.begin(), L_rhs.end(), 0.0);
        const auto &p = msh->node_p;
        for (int t = 0; t < msh->NT(); t++)
            {
            std::array<int, 4> ind;
            for (int i = 0; i < 4; i++) ind[i] = msh->tet_ind[4 * (size_t)t + i];
            double J[3][3];
            for (int c = 0; c < 3; c++)
                for (int k = 0; k < 3; k++) J[c][k] = p[3 * (size_t)ind[k + 1] + c] - p[3 * (size_t)ind[0] + c];
            const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                               + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
            double inv[3][3];  // inverse of J
            inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
            inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
            inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
            inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
            inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
            double g[4][3];  // gradients of the four hat functions
            for (int c = 0; c < 3; c++)
                {
                g[0][c] = -(inv[0][c] + inv[1][c] + inv[2][c]);
                for (int k = 0; k < 3; k++) g[k + 1][c] = inv[k][c];
                }
            fgb200::Dense<4> Ke;
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++)
                    Ke(i, j) = std::fabs(det) / 6.0 * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
            buildMat<4>(ind, Ke);
            std::vector<double> Le(4, 0.0);
            buildVect<4>(ind, Le);
            }
        std::vector<int> ld;
        std::vector<double> Vd((size_t)NOD, 0.0);
        for (int a = 0; a < NOD; a++)
            {
            const double x = p[3 * (size_t)a];
            if (x < 1e-12 * xmax || x > xmax * (1 - 1e-12))
                {
                ld.push_back(a);
                Vd[a] = x > 0.5 * xmax ? 1.0 : 0.0;
                }
            }
        V.assign((size_t)NOD, 0.0);
        iter.reset();
        algebra::cg_dir(iter, K, V, L_rhs, Vd, ld);
        return iter.status == algebra::CONVERGED;
        }
    int iterations() const { return iter.get_iteration(); }
    int rows() const { return (int)K.size(); }
    };

static void test_solver_template()
    {
    const double h = 2e-9;
    fgb200::MeshView msh = cuboid(10, 4, 3, h);
    laplaceSolver ls(msh, 1e-10, 2000);
    CHECK(ls.rows() == msh.NOD());
    CHECK(!msh.edges.empty());
    std::vector<double> V;
    CHECK(ls.run(V, 10 * h));
    CHECK(ls.iterations() > 0);
    // the discrete harmonic function with these boundary values is exactly V = x / xmax (P1 elements)
    double err = 0.0;
    for (int a = 0; a < msh.NOD(); a++) err = std::max(err, std::fabs(V[a] - msh.node_p[3 * (size_t)a] / (10 * h)));
    CHECK(err < 1e-8);
    }

int main(int argc, char **argv)
    {
    if (argc > 1 && !std::strcmp(argv[1], "--timestepper")) return replay_timestepper();
    if (argc > 1 && !std::strcmp(argv[1], "--link-check"))
        {
        std::printf("fg_version %d\n", fg_version());
        return fg_version() >= 100 ? 0 : 1;
        }
    try
        {
        test_identity_solves();
        test_laplacian_dirichlet();
        test_solver_template();
        test_llg_loop();
        test_fem_time_integration();
        }
    catch (const std::exception &e)
        {
        std::printf("exception: %s\n", e.what());
        return 2;
        }
    std::printf("%s (%d failures)\n", n_fail ? "FAILED" : "ok", n_fail);
    return n_fail ? 1 : 0;
    }
