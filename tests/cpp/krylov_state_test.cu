// Host-side emulation of the device BiCGStab (no GPU): the scalar side -- iteration monitor, alpha /
// omega / rho updates, breakdown, overflow and mid-iteration exit rules -- is the library's own code
// (feellgood_b200/csrc/fg_krylov_state.cuh, __host__ __device__), called here in exactly the order and
// under exactly the gates the kernels of fg_krylov.cu use; the vector side is restated with the same
// expressions (fused multiply-adds where the device contracts them, compensated sums like the device's
// TwoSum reduction trees).  tests/test_device_math.py compares status, iteration count and solution
// with the reference's own bicg_dir (src/algebra/bicg.h:163-234, compiled unmodified in oracle/_ref).
//
// A second mode does the same for the Jacobi-preconditioned CG (src/algebra/cg.h:15-58; kernels k_cg_p,
// k_spmv<ST_CG_SETUP|ST_CG_Q>, k_cg_xr).
//
// argv[1]: "bicg" (default) | "cg"
// stdin:  n nnz nmask tol maxiter, then rowptr[n+1], col[nnz], val[nnz], rhs[n], x0[n], mask[nmask]
// stdout: status nit res rhsn failed (LinAlgebra::solve's predicate, src/solver.cpp:62-69), then x[n]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../feellgood_b200/csrc/fg_krylov_state.cuh"

using namespace fg;

struct Sum   // error-free accumulation (the device's reduction trees are compensated, fg_reduce.cuh)
    {
    double s = 0.0, e = 0.0;
    void add(double b)
        {
        const double t = s + b, bb = t - s;
        e += (s - (t - bb)) + (b - bb);
        s = t;
        }
    double get() const { return s + e; }
    };

int main(int argc, char **argv)
    {
    const bool cg = argc > 1 && std::strcmp(argv[1], "cg") == 0;
    int n, nnz, nmask, maxiter;
    double tol;
    if (std::scanf("%d %d %d %lf %d", &n, &nnz, &nmask, &tol, &maxiter) != 5) return 2;
    std::vector<int> rp(n + 1), col(nnz), ld(nmask);
    std::vector<double> val(nnz), b(n), x(n);
    for (auto &v : rp) if (std::scanf("%d", &v) != 1) return 2;
    for (auto &v : col) if (std::scanf("%d", &v) != 1) return 2;
    for (auto &v : val) if (std::scanf("%lf", &v) != 1) return 2;
    for (auto &v : b) if (std::scanf("%lf", &v) != 1) return 2;
    for (auto &v : x) if (std::scanf("%lf", &v) != 1) return 2;
    for (auto &v : ld) if (std::scanf("%d", &v) != 1) return 2;
    std::vector<unsigned char> mask(n, 0);
    for (int i : ld) mask[i] = 1;
    // what fg_bicg_dir prepares: D = 1/diag (0 when masked), rhs masked (bicg.h:176-178)
    std::vector<double> D(n, 0.0);
    for (int i = 0; i < n; i++)
        {
        double d = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (col[j] == i) d = val[j];
        D[i] = mask[i] ? 0.0 : 1.0 / d;
        if (mask[i]) b[i] = 0.0;
        }
    auto spmv = [&](const std::vector<double> &in, std::vector<double> &out)
        {  // k_spmv_csr: y0 += val * x, left to right (one lane group per row)
        for (int i = 0; i < n; i++)
            {
            double y = 0.0;
            for (int j = rp[i]; j < rp[i + 1]; j++) y = std::fma(val[j], in[col[j]], y);
            out[i] = y;
            }
        };
    std::vector<double> r(n), rt(n), p(n), v(n), s(n), t(n), phat(n), shat(n);
    KState st;
    st.hist = nullptr;
    st.hist_cap = 0;
    kstate_reset(&st, tol, maxiter);
    if (cg)
        {
        std::vector<double> &q = t;
            {  // ST_CG_SETUP: r = b - A x (masked); p = D r; ||b||^2, ||r||^2, (Dr, r)
            spmv(x, v);
            Sum bb, rr, zr;
            for (int i = 0; i < n; i++)
                {
                const double ri = mask[i] ? 0.0 : b[i] - v[i];
                const double z = D[i] * ri;
                r[i] = ri;
                p[i] = z;
                bb.add(b[i] * b[i]);
                rr.add(ri * ri);
                zr.add(z * ri);
                }
            const double tot[RED_NV] = {bb.get(), rr.get(), zr.get(), 0.0};
            spmv_finalize<ST_CG_SETUP>(&st, tot);
            }
        for (int guard = 0; guard < maxiter + 3 && !st.done; guard++)
            {
            // k_cg_p: p = (rho/rho_1) p + D r   (nit > 0)
            if (!st.done && st.nit > 0)
                {
                const double f = st.rho1 / st.rho2;
                for (int i = 0; i < n; i++) p[i] = std::fma(p[i], f, D[i] * r[i]);
                }
            // k_spmv<ST_CG_Q>
            if (!st.done)
                {
                spmv(p, q);
                Sum qp;
                for (int i = 0; i < n; i++)
                    {
                    if (mask[i]) q[i] = 0.0;
                    qp.add(q[i] * p[i]);
                    }
                const double tot[RED_NV] = {qp.get(), 0.0, 0.0, 0.0};
                spmv_finalize<ST_CG_Q>(&st, tot);
                }
            // k_cg_xr
            if (!st.done)
                {
                Sum rr, zr;
                for (int i = 0; i < n; i++)
                    {
                    x[i] = std::fma(st.alpha, p[i], x[i]);
                    const double ri = std::fma(-st.alpha, q[i], r[i]);
                    r[i] = ri;
                    rr.add(ri * ri);
                    zr.add((D[i] * ri) * ri);
                    }
                cg_xr_finalize(&st, rr.get(), zr.get());
                }
            }
        std::printf("%d %d %.17g %.17g %d\n", st.status, st.nit, st.res, st.rhsn, (int)solve_failed(&st));
        for (int i = 0; i < n; i++) std::printf("%.17g\n", x[i]);
        return st.done ? 0 : 3;
        }
        {  // ST_BICG_SETUP: r = b - A x (masked); rt = r; ||b||^2, ||r||^2
        spmv(x, v);
        Sum bb, rr;
        for (int i = 0; i < n; i++)
            {
            const double ri = mask[i] ? 0.0 : b[i] - v[i];
            r[i] = rt[i] = ri;
            bb.add(b[i] * b[i]);
            rr.add(ri * ri);
            }
        const double tot[RED_NV] = {bb.get(), rr.get(), 0.0, 0.0};
        spmv_finalize<ST_BICG_SETUP>(&st, tot);
        }
    for (int guard = 0; guard < maxiter + 3; guard++)
        {
        // k_bicg_p
        if (!st.done)
            {
            const bool first = st.nit == 0;
            const double omega = st.omega, beta = first ? 0.0 : bicg_beta(&st);
            for (int i = 0; i < n; i++)
                {
                p[i] = first ? r[i] : bicg_p_value(p[i], v[i], r[i], omega, beta);
                phat[i] = D[i] * p[i];
                }
            }
        // k_spmv<ST_BICG_V>
        if (!st.done)
            {
            spmv(phat, v);
            Sum d;
            for (int i = 0; i < n; i++)
                {
                if (mask[i]) v[i] = 0.0;
                d.add(v[i] * rt[i]);
                }
            const double tot[RED_NV] = {d.get(), 0.0, 0.0, 0.0};
            spmv_finalize<ST_BICG_V>(&st, tot);
            }
        // k_bicg_s
        if (!st.done)
            {
            Sum ss;
            for (int i = 0; i < n; i++)
                {
                s[i] = bicg_s_value(r[i], v[i], st.alpha);
                shat[i] = D[i] * s[i];
                ss.add(s[i] * s[i]);
                }
            bicg_s_finalize(&st, ss.get());
            }
        // k_spmv<ST_BICG_T>
        if (!st.done)
            {
            spmv(shat, t);
            Sum ts, tt;
            for (int i = 0; i < n; i++)
                {
                if (mask[i]) t[i] = 0.0;
                ts.add(t[i] * s[i]);
                tt.add(t[i] * t[i]);
                }
            const double tot[RED_NV] = {ts.get(), tt.get(), 0.0, 0.0};
            spmv_finalize<ST_BICG_T>(&st, tot);
            }
        // k_bicg_xr
        const int fh = st.final_half;
        if (!(st.done && !fh))
            {
            Sum rr, rtr;
            if (fh)
                for (int i = 0; i < n; i++) x[i] = std::fma(st.alpha, phat[i], x[i]);
            else
                for (int i = 0; i < n; i++)
                    {
                    x[i] = std::fma(st.omega, shat[i], std::fma(st.alpha, phat[i], x[i]));
                    const double ri = std::fma(-st.omega, t[i], s[i]);
                    r[i] = ri;
                    rr.add(ri * ri);
                    rtr.add(rt[i] * ri);
                    }
            bicg_xr_finalize(&st, rr.get(), rtr.get(), fh);
            }
        if (st.done && !st.final_half) break;
        }
    std::printf("%d %d %.17g %.17g %d\n", st.status, st.nit, st.res, st.rhsn, (int)solve_failed(&st));
    for (int i = 0; i < n; i++) std::printf("%.17g\n", x[i]);
    return st.done ? 0 : 3;
    }
