// Host-side check of the device math helpers that have no reference counterpart (compiled by nvcc,
// run on the CPU: the functions are __host__ __device__): the unit-quaternion encoding of the
// tangent-plane basis read by the Krylov kernels (feellgood_b200/csrc/fg_common.cuh).
//   * round trip basis -> quaternion -> basis within 2e-15 for bases built like Node::setBasis
//     (reference src/node.h:73-102), every branch of Shepperd's method exercised
//   * the decoded basis is orthonormal to 5e-15 (the tolerance of the reference's own ut_node.cpp)
//   * |q| = 1; degenerate input (u = 0, the pad / non-magnetic rows) gives a finite quaternion
#include <cmath>
#include <cstdio>
#include <random>

#include "../../feellgood_b200/csrc/fg_common.cuh"

static void set_basis(const double u[3], double r, double ep[3], double eq[3])
    {  // Node::setBasis restated on the host
    int k = 0;
    double m = std::fabs(u[0]);
    if (std::fabs(u[1]) < m) { m = std::fabs(u[1]); k = 1; }
    if (std::fabs(u[2]) < m) { k = 2; }
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    const double d = e[0] * u[0] + e[1] * u[1] + e[2] * u[2];
    for (int c = 0; c < 3; c++) e[c] -= d * u[c];
    const double z = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    if (z > 0) for (int c = 0; c < 3; c++) e[c] /= std::sqrt(z);
    const double f[3] = {u[1] * e[2] - u[2] * e[1], u[2] * e[0] - u[0] * e[2], u[0] * e[1] - u[1] * e[0]};
    for (int c = 0; c < 3; c++)
        {
        ep[c] = std::cos(r) * e[c] - std::sin(r) * f[c];
        eq[c] = std::sin(r) * e[c] + std::cos(r) * f[c];
        }
    }

int main()
    {
    std::mt19937 gen(5489);
    std::normal_distribution<double> N(0.0, 1.0);
    std::uniform_real_distribution<double> U(0.0, 6.283185307179586);
    double worst_rt = 0, worst_on = 0, worst_nq = 0;
    int branch[4] = {0, 0, 0, 0}, fails = 0;
    for (int it = 0; it < 200000; it++)
        {
        double u[3] = {N(gen), N(gen), N(gen)};
        if (it % 7 == 0) u[it % 3] = 0.0;              // the strict-minimum branch of setBasis
        if (it % 11 == 0) { u[0] *= 1e-9; u[1] *= 1e-9; }  // nearly axis-aligned magnetisation
        const double n = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (!(n > 0)) continue;
        for (int c = 0; c < 3; c++) u[c] /= n;
        double ep[3], eq[3], e2[3], f2[3];
        set_basis(u, U(gen), ep, eq);
        const double tr = ep[0] + eq[1] + (ep[0] * eq[1] - ep[1] * eq[0]);
        const double r00 = ep[0], r11 = eq[1], r22 = ep[0] * eq[1] - ep[1] * eq[0];
        branch[tr > 0 ? 0 : (r00 > r11 && r00 > r22 ? 1 : (r11 > r22 ? 2 : 3))]++;
        const double4 q = fg::basis_to_quat(ep, eq);
        fg::quat_to_basis(q, e2, f2);
        for (int c = 0; c < 3; c++)
            {
            worst_rt = std::fmax(worst_rt, std::fabs(e2[c] - ep[c]));
            worst_rt = std::fmax(worst_rt, std::fabs(f2[c] - eq[c]));
            }
        const double ee = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2], ff = f2[0] * f2[0] + f2[1] * f2[1] + f2[2] * f2[2],
                     ef = e2[0] * f2[0] + e2[1] * f2[1] + e2[2] * f2[2];
        worst_on = std::fmax(worst_on, std::fmax(std::fabs(ee - 1), std::fmax(std::fabs(ff - 1), std::fabs(ef))));
        worst_nq = std::fmax(worst_nq, std::fabs(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w - 1));
        }
    std::printf("round trip %.3e  orthonormality %.3e  |q|^2-1 %.3e  branches %d %d %d %d\n", worst_rt, worst_on,
                worst_nq, branch[0], branch[1], branch[2], branch[3]);
    if (!(worst_rt < 2e-15)) { std::printf("FAIL round trip\n"); fails++; }
    if (!(worst_on < 5e-15)) { std::printf("FAIL orthonormality\n"); fails++; }
    if (!(worst_nq < 2e-15)) { std::printf("FAIL norm\n"); fails++; }
    for (int b = 0; b < 4; b++)
        if (branch[b] == 0) { std::printf("FAIL branch %d not exercised\n", b); fails++; }
        {  // degenerate rows: u = 0 gives ep = cos r e, eq = sin r e (not a rotation): finite, unit quaternion
        const double u0[3] = {0, 0, 0};
        for (double r : {0.0, 0.3, 1.5707963267948966, 3.141592653589793, 4.0})
            {
            double ep[3], eq[3], e2[3], f2[3];
            set_basis(u0, r, ep, eq);
            const double4 q = fg::basis_to_quat(ep, eq);
            fg::quat_to_basis(q, e2, f2);
            const double nq = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
            bool fin = std::isfinite(nq) && std::fabs(nq - 1) < 2e-15;
            for (int c = 0; c < 3; c++) fin = fin && std::isfinite(e2[c]) && std::isfinite(f2[c]);
            if (!fin) { std::printf("FAIL degenerate r=%g\n", r); fails++; }
            }
        }
    std::printf(fails ? "DEVICE_MATH_FAILED\n" : "DEVICE_MATH_OK\n");
    return fails ? 1 : 0;
    }
