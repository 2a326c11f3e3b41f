// Host-side run of the CUDA library's element math (the functions are __host__ __device__, the Gauss
// tables on the host come from the same tet_tables() that fills the __constant__ copies): no GPU.
//   * tet_core (general path of Tet::integrales, reference src/tetra.cpp:210-307) on seeded random
//     tetrahedra with uniaxial + cubic anisotropy, and with the recentring drift term;
//   * tet_iso_front / tet_iso_be (fast path, K = K3 = 0) against tet_core on the same inputs: the
//     two paths must agree bit for bit on the host (same formulas, same order of accumulation);
//   * every case is printed as one JSON line {inputs, contrib[4], BE[12]} so that
//     tests/test_device_math.py can check it against the independent dense numpy restatement;
//   * node_set_basis orthonormality, the explicit node-diagonal block against its closed form, and
//     tri_core (Tri::integrales, src/triangle.cpp:6-36), also printed for the numpy check.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "../../feellgood_b200/csrc/fg_llg_kernels.cuh"

using namespace fg;

static void pr(const char *name, const double *v, int n, bool last = false)
    {
    std::printf("\"%s\": [", name);
    for (int k = 0; k < n; k++) std::printf("%s%.17g", k ? ", " : "", v[k]);
    std::printf("]%s", last ? "" : ", ");
    }

template <int NPI> static int run(int ncase, std::mt19937 &gen)
    {
    std::normal_distribution<double> N(0.0, 1.0);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    int fails = 0;
    for (int it = 0; it < ncase; it++)
        {
        const bool aniso = (it % 3) != 0;   // every third case is isotropic: both paths apply
        const bool drift = aniso && (it % 5) == 0;
        // a positively oriented, jittered tetrahedron of ~2 nm
        double p[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int i = 0; i < 4; i++)
            for (int d = 0; d < 3; d++) p[i][d] = 2e-9 * (p[i][d] + 0.2 * (U(gen) - 0.5));
        double J[3][3];
        for (int d = 0; d < 3; d++)
            for (int k = 0; k < 3; k++) J[d][k] = p[k + 1][d] - p[0][d];
        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                           + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        double inv[3][3];
        inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
        inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        const double dadu[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        TetIn T;
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 3; k++)
                T.da[i][k] = dadu[i][0] * inv[0][k] + dadu[i][1] * inv[1][k] + dadu[i][2] * inv[2][k];
        T.detJ = det;
        for (int i = 0; i < 4; i++)
            {
            double n2 = 0;
            for (int d = 0; d < 3; d++) { T.u[i][d] = N(gen); n2 += T.u[i][d] * T.u[i][d]; }
            for (int d = 0; d < 3; d++) { T.u[i][d] /= std::sqrt(n2); T.v[i][d] = 2e9 * N(gen); }
            T.phi[i] = 0.3 * N(gen);
            T.phiv[i] = 1e10 * N(gen);
            }
        const double MU0 = FG_MU0, A = 1.3e-11, Ms = 8e5, alpha = 0.02 + 0.5 * U(gen);
        const double K = aniso ? 3e5 : 0.0, K3 = aniso ? -1.2e4 : 0.0;
        TetRegion R;
        std::memset(&R, 0, sizeof(R));
        R.alpha = alpha; R.A = A; R.K = K; R.K3 = K3; R.Ms = Ms;
        R.Abis = 2.0 * A / (MU0 * Ms); R.Kbis = 2.0 * K / (MU0 * Ms); R.K3bis = 2.0 * K3 / (MU0 * Ms);
        R.has_K = K != 0; R.has_K3 = K3 != 0;
        const double s2 = 1.0 / std::sqrt(2.0);
        const double uk[3] = {0, 1, 0}, ex[3] = {s2, s2, 0}, ey[3] = {-s2, s2, 0}, ez[3] = {0, 0, 1};
        for (int d = 0; d < 3; d++) { R.uk[d] = uk[d]; R.ex[d] = ex[d]; R.ey[d] = ey[d]; R.ez[d] = ez[d]; }
        StepPrm sp;
        sp.dt = 2e-14 * (1 + 4 * U(gen));
        sp.prefactor = 1.0 + 1e-3 * U(gen);
        sp.Hext[0] = -2e4 * N(gen); sp.Hext[1] = 3e3 * N(gen); sp.Hext[2] = 1e3 * N(gen);
        sp.A_Hext = 0.0;
        sp.Vdrift = drift ? 37.0 : 0.0;
        sp.idx_dir = drift ? (it % 3) : FG_IDX_UNDEF;
        double Hext[3][NPI];
        for (int d = 0; d < 3; d++)
            for (int g = 0; g < NPI; g++) Hext[d][g] = sp.Hext[d];
        double contrib[4], BE[3][4];
        tet_core<NPI>(T, R, sp, Hext, contrib, BE);
        if (!drift)
            {  // the lean instantiation (cubic anisotropy and drift compiled out: k_tet_lean) on a material
               // without cubic anisotropy is the general core bit for bit
            TetRegion Ru = R;
            Ru.has_K3 = 0;
            Ru.K3 = Ru.K3bis = 0.0;
            double cg[4], Bg[3][4], cl[4], Bl[3][4];
            tet_core<NPI>(T, Ru, sp, Hext, cg, Bg);
            tet_core<NPI, false, false>(T, Ru, sp, Hext, cl, Bl);
            for (int i = 0; i < 4; i++)
                {
                bool same = cg[i] == cl[i];
                for (int d = 0; d < 3; d++) same = same && Bg[d][i] == Bl[d][i];
                if (!same)
                    {
                    std::fprintf(stderr, "LEAN CORE DIFFERS case %d node %d\n", it, i);
                    fails++;
                    }
                }
            }
        if (!aniso)
            {
            TetIsoIn Ti;
            for (int i = 0; i < 4; i++)
                {
                for (int k = 0; k < 3; k++) { Ti.da[i][k] = T.da[i][k]; Ti.u[i][k] = T.u[i][k]; }
                Ti.phi[i] = T.phi[i];
                Ti.phiv[i] = T.phiv[i];
                }
            Ti.detJ = T.detJ;
            TetIsoMid M;
            double c2[4];
            tet_iso_front<NPI>(Ti, R, sp, M, c2);
            for (int i = 0; i < 4; i++)
                {
                double be[3];
                tet_iso_be<NPI>(Ti.da[i], i, Ti.detJ, R.Abis, M, be);
                bool same = c2[i] == contrib[i];
                for (int d = 0; d < 3; d++) same = same && be[d] == BE[d][i];
                if (!same)
                    {
                    std::fprintf(stderr, "FAST PATH DIFFERS case %d node %d: %.17g %.17g | %.17g %.17g\n", it, i, c2[i],
                                 contrib[i], be[0], BE[0][i]);
                    fails++;
                    }
                }
            }
        std::printf("{\"npi\": %d, \"drift\": %d, \"idx_dir\": %d, \"Vdrift\": %.17g, \"alpha\": %.17g, \"A\": %.17g, \"Ms\": %.17g, "
                    "\"K\": %.17g, \"K3\": %.17g, \"dt\": %.17g, \"prefactor\": %.17g, \"detJ\": %.17g, ",
                    NPI, (int)drift, sp.idx_dir, sp.Vdrift, alpha, A, Ms, K, K3, sp.dt, sp.prefactor, T.detJ);
        pr("uk", uk, 3); pr("ex", ex, 3); pr("ey", ey, 3); pr("ez", ez, 3); pr("Hext", sp.Hext, 3);
        pr("da", &T.da[0][0], 12); pr("u", &T.u[0][0], 12); pr("v", &T.v[0][0], 12); pr("phi", T.phi, 4);
        pr("phiv", T.phiv, 4); pr("contrib", contrib, 4); pr("BE", &BE[0][0], 12, true);
        std::printf("}\n");
        }
    return fails;
    }

// Node::setBasis (device version) and the node-diagonal block: the explicit triple products of
// project_block + gyro_block (what the assembled-block operator and the Jacobi diagonal use) against
// the closed form [[a_w, Ma], [Ma, -a_w]] the matrix-free SpMV applies (fg_common.cuh, OP_NODE3).
static int check_basis_and_diagonal(std::mt19937 &gen)
    {
    std::normal_distribution<double> N(0.0, 1.0);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double worst_on = 0, worst_dg = 0;
    for (int it = 0; it < 100000; it++)
        {
        double u[3] = {N(gen), N(gen), N(gen)};
        if (it % 7 == 0) u[it % 3] = 0.0;
        const double n = std::sqrt(dot3(u, u));
        if (!(n > 0)) continue;
        for (int d = 0; d < 3; d++) u[d] /= n;
        const double r = 0.63661977236758134308 * U(gen);   // M_2_PI * U(0,1), linear_algebra.cpp:3-11
        double ep[3], eq[3];
        node_set_basis(u, std::cos(r), std::sin(r), ep, eq);
        worst_on = std::fmax(worst_on, std::fmax(std::fabs(dot3(ep, u)), std::fabs(dot3(eq, u))));
        worst_on = std::fmax(worst_on, std::fmax(std::fabs(dot3(ep, eq)), std::fabs(dot3(ep, ep) - 1)));
        worst_on = std::fmax(worst_on, std::fabs(dot3(eq, eq) - 1));
        const double Ma = 1e-27 * (0.1 + U(gen)), aw = 1e-27 * (0.1 + U(gen));   // ~ volume / 4 * alpha
        double k00, k01, k10, k11;
        project_block(Ma, ep, eq, ep, eq, k00, k01, k10, k11);
        gyro_block(aw, u, ep, eq, k00, k01, k10, k11);
        const double sc = std::fmax(Ma, aw);
        worst_dg = std::fmax(worst_dg, std::fabs(k00 - aw) / sc);
        worst_dg = std::fmax(worst_dg, std::fabs(k01 - Ma) / sc);
        worst_dg = std::fmax(worst_dg, std::fabs(k10 - Ma) / sc);
        worst_dg = std::fmax(worst_dg, std::fabs(k11 + aw) / sc);
        }
    std::fprintf(stderr, "basis orthonormality %.3e   explicit Dg - closed form %.3e (relative)\n", worst_on, worst_dg);
    int fails = 0;
    if (!(worst_on < 5e-15)) { std::fprintf(stderr, "FAIL basis\n"); fails++; }     // UT_TOL of ut_node.cpp
    if (!(worst_dg < 1e-14)) { std::fprintf(stderr, "FAIL closed-form Dg\n"); fails++; }
    return fails;
    }

// Tri::integrales (tri_core): one JSON line per case for the numpy check
template <int NPI> static void run_tri(int ncase, std::mt19937 &gen)
    {
    std::normal_distribution<double> N(0.0, 1.0);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (int it = 0; it < ncase; it++)
        {
        TriRegion R;
        R.Ks = 2.5e-4 * (0.5 + U(gen));
        double uk[3] = {N(gen), N(gen), N(gen)};
        const double nk = std::sqrt(dot3(uk, uk));
        for (int d = 0; d < 3; d++) R.uk[d] = uk[d] / nk;
        double u[3][3];
        for (int i = 0; i < 3; i++)
            {
            for (int d = 0; d < 3; d++) u[i][d] = N(gen);
            const double n = std::sqrt(dot3(u[i], u[i]));
            for (int d = 0; d < 3; d++) u[i][d] /= n;
            }
        const double surf = 2e-18 * (0.5 + U(gen)), dMs = 8e5;
        double BE[3][3];
        tri_core<NPI>(R, surf, dMs, u, BE);
        std::printf("{\"tri\": 1, \"npi\": %d, \"Ks\": %.17g, \"surf\": %.17g, \"dMs\": %.17g, ", NPI, R.Ks, surf, dMs);
        pr("uk", R.uk, 3); pr("u", &u[0][0], 9); pr("BE", &BE[0][0], 9, true);
        std::printf("}\n");
        }
    }

int main()
    {
    std::mt19937 gen(5489);
    int fails = run<5>(24, gen) + run<1>(12, gen);
    fails += check_basis_and_diagonal(gen);
    run_tri<4>(8, gen);
    run_tri<1>(4, gen);
    std::fprintf(stderr, fails ? "ELEMENT_MATH_FAILED\n" : "ELEMENT_MATH_OK\n");
    return fails ? 1 : 0;
    }
