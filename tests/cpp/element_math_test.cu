// Host-side run of the CUDA library's element math (the functions are __host__ __device__, the Gauss
// tables on the host come from the same tet_tables() that fills the __constant__ copies): no GPU.
//   * tet_core (general path of Tet::integrales, reference src/tetra.cpp:210-307) on seeded random
//     tetrahedra with uniaxial + cubic anisotropy, and with the recentring drift term;
//   * tet_iso_front / tet_iso_be (fast path, K = K3 = 0) against tet_core on the same inputs: the
//     two paths must agree bit for bit on the host (same formulas, same order of accumulation);
//   * every case is printed as one JSON line {inputs, contrib[4], BE[12]} so that
//     tests/test_device_math.py can check it against the independent dense numpy restatement.
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

int main()
    {
    std::mt19937 gen(5489);
    int fails = run<5>(24, gen) + run<1>(12, gen);
    std::fprintf(stderr, fails ? "ELEMENT_MATH_FAILED\n" : "ELEMENT_MATH_OK\n");
    return fails ? 1 : 0;
    }
