"""Oracle pinning of SURVEY §8(f) rank 2 (charges feeding the demag solver) with the reference's own
unit-test formulas: ut_tet_charges.cpp:17-58, ut_tri_charges.cpp:17-52, ut_triangle.cpp:129-241
(Tri_potential_u) and :243-358 (Tri_potential_v); then fmm::calc_charges as a whole against an
independent numpy restatement, and ut_log-stats.cpp for the loop statistics."""
import ctypes as C

import numpy as np

import cases
from feellgood_b200 import meshgen
from feellgood_b200.fem import LogStats
from oracle import fg_oracle_py as fo

UT_TOL = 5e-16


def rand_unit(rng, n):
    th, ph = np.pi * rng.random(n), 2 * np.pi * rng.random(n)
    return np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_tet_charges(oracle):
    """ut_tet_charges.cpp:17-58 on the unit tetrahedron."""
    L = fo.lib()
    rng = np.random.default_rng(5489)
    p = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    ind = np.arange(4, dtype=np.int32)
    da, w = np.zeros(12), np.zeros(5)
    L.fgo_tet_setup(_dp(p), ind.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(5), _dp(da), _dp(w))
    u = rand_unit(rng, 4)
    Ms = rng.random()
    out = np.zeros(5)
    L.fgo_tet_charges(C.c_int(5), C.c_double(Ms), _dp(da), _dp(w), _dp(np.ascontiguousarray(u)), _dp(out))
    D = da.reshape(4, 3)
    vec_nod = u.T                                           # 3 x N
    dudx, dudy, dudz = vec_nod @ D[:, 0], vec_nod @ D[:, 1], vec_nod @ D[:, 2]
    ref = -Ms * (w * (dudx[0] + dudy[1] + dudz[2]))
    assert np.max(np.abs(out - ref)) <= UT_TOL * np.max(np.abs(ref))


def tri_fixture(rng):
    """ut_triangle.cpp:141-154: a small random triangle near (1, 0, 0)."""
    p = np.stack([1 + (rng.random(3) - 0.5) / 10.0, (rng.random(3) - 0.5) / 10.0,
                  (rng.random(3) - 0.5) / 10.0], axis=1)
    vec = rand_unit(rng, 3)
    dMs = rng.random()
    ind = np.arange(3, dtype=np.int32)
    surf, n, w = C.c_double(), np.zeros(3), np.zeros(4)
    fo.lib().fgo_tri_setup(_dp(np.ascontiguousarray(p)), ind.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(4),
                           C.byref(surf), _dp(n), _dp(w))
    return np.ascontiguousarray(p), np.ascontiguousarray(vec), dMs, surf.value, n, w


def test_tri_charges(oracle):
    """ut_tri_charges.cpp:17-52."""
    rng = np.random.default_rng(5489)
    p, vec, dMs, surf, n, w = tri_fixture(rng)
    out = np.zeros(4)
    fo.lib().fgo_tri_charges(C.c_int(4), C.c_double(dMs), _dp(n), _dp(w), _dp(vec), _dp(out))
    a = np.ctypeslib.as_array(fo.lib().fgo_tri_a(4), shape=(12,)).reshape(3, 4)
    ref = dMs * (w * ((vec.T @ a).T @ n))
    assert np.max(np.abs(out - ref)) <= UT_TOL * np.max(np.abs(ref))
    fo.lib().fgo_tri_charges(C.c_int(4), C.c_double(0.0), _dp(n), _dp(w), _dp(vec), _dp(out))
    assert np.all(out == 0.0)                               # triangle.cpp:50: nothing when dMs == 0


def potential_ref_code(p, s, surf, dMs, i):
    """The 'ref code' of ut_triangle.cpp:156-232 (MuMag_potential.cc), verbatim arithmetic."""
    ii, iii = (i + 1) % 3, (i + 2) % 3
    (x1, y1, z1), (x2, y2, z2), (x3, y3, z3) = p[i], p[ii], p[iii]
    b = np.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2 + (z2 - z1) ** 2)
    t = (x2 - x1) * (x3 - x1) + (y2 - y1) * (y3 - y1) + (z2 - z1) * (z3 - z1)
    h = 2. * surf
    t /= b
    h /= b
    a = t / h
    c = (t - b) / h
    s1, s2, s3 = s[i], s[ii], s[iii]
    l_ = s1
    j = (s2 - s1) / b
    k = t / b / h * (s1 - s2) + (s3 - s1) / h
    cc1 = c * c + 1
    r = np.sqrt(h * h + (c * h + b) * (c * h + b))
    ll = np.log((cc1 * h + c * b + np.sqrt(cc1) * r) / (b * (c + np.sqrt(cc1))))
    pot1 = b * b / cc1 ** 1.5 * ll + c * b * r / cc1 + h * r - c * b * b / cc1 - np.sqrt(a * a + 1) * h * h
    pot1 *= j / 2.
    pot2 = -c * b * b / cc1 ** 1.5 * ll + b * r / cc1 - h * h / 2. + h * h * np.log(c * h + b + r) - b * b / cc1
    pot2 *= k / 2.
    pot3 = h * np.log(c * h + b + r) - h + b / np.sqrt(cc1) * ll
    pot3 *= l_
    pot = pot1 + pot2 + pot3 + h * (k * h / 2. + l_) * (1 - np.log(h * (a + np.sqrt(a * a + 1)))) - k * h * h / 4.
    return dMs * pot


def test_tri_potential(oracle):
    """ut_triangle.cpp:129-241 / :243-358 (tolerance 10*UT_TOL there; the two closed forms differ
    by cancellation, so relative 1e-11 on a few dozen seeded triangles, all three corners)."""
    L = fo.lib()
    rng = np.random.default_rng(5489)
    for _ in range(40):
        p, vec, dMs, surf, n, w = tri_fixture(rng)
        s = vec @ n
        for i in range(3):
            got = L.fgo_tri_potential(_dp(p), _dp(vec), C.c_double(surf), _dp(n), C.c_double(dMs), C.c_int(i))
            ref = potential_ref_code(p, s, surf, dMs, i)
            assert abs(got - ref) <= 1e-11 * max(abs(ref), dMs * np.sqrt(surf))


def test_calc_charges_vs_numpy(oracle):
    """fmm::calc_charges (src/fmm_demag.h:155-185): source order, signs, corrections."""
    case = cases.small_cuboid()
    oc = cases.oracle_ctx(case)
    oc.set_state(case.u, case.v, case.phi, case.phiv)
    m = case.mesh
    a = np.ctypeslib.as_array(fo.lib().fgo_tet_a(5), shape=(20,)).reshape(4, 5)
    at = np.ctypeslib.as_array(fo.lib().fgo_tri_a(4), shape=(12,)).reshape(3, 4)
    pt = np.ctypeslib.as_array(fo.lib().fgo_tri_pds(4), shape=(4,))
    ind, (da, w), mag = oc.tet_ind(), oc.tet_geom(), oc.masks()[0]
    for which, field in ((0, case.u), (1, case.v)):
        src, corr = oc.calc_charges(which)
        pos = oc.source_positions()
        ref_src, ref_pos, ref_corr = [], [], np.zeros(m.NOD)
        for t in range(m.NT):
            Ms = case.tet_regions[m.tet_reg[t]].get("Ms", 795774.7)
            if not Ms > 0:
                continue
            div = np.trace(field[ind[t]].T @ da[t].reshape(4, 3))
            ref_src += list(-Ms * w[t] * div)
            ref_pos += list((m.node_p[ind[t]].T @ a).T)
        for f in range(m.NF):
            i3 = m.tri_ind[f]
            if not mag[i3].all() or case.tri_regions[m.tri_reg[f]].get("suppress_charges", False):
                continue
            p = m.node_p[i3]
            nv = np.cross(p[1] - p[0], p[2] - p[0])
            surf, n = 0.5 * np.linalg.norm(nv), nv / np.linalg.norm(nv)
            q = m.tri_dMs[f] * (2 * surf * pt) * ((field[i3].T @ at).T @ n)
            g = (p.T @ at).T
            ref_src += list(q)
            ref_pos += list(g)
            for i in range(3):
                ref_corr[i3[i]] -= np.sum(q / np.linalg.norm(p[i] - g, axis=1))
                ref_corr[i3[i]] += potential_ref_code(p, field[i3] @ n, surf, m.tri_dMs[f], i)
        assert src.size == len(ref_src) == oc.n_sources()
        assert cases.rel_max(src, np.array(ref_src)) < 1e-13
        assert cases.rel_max(pos, np.array(ref_pos)) < 1e-15
        assert cases.rel_max(corr, ref_corr) < 1e-9        # cancellation in the corrections
    # the all-pairs potential: phi = (sum q / r + corr) / 4 pi on the magnetic nodes only
    src, corr = oc.calc_charges(0)
    pos = oc.source_positions()
    oc.demag_direct(True)
    phi = oc.get_state(1)[2]
    r = np.linalg.norm(m.node_p[:, None, :] - pos[None, :, :], axis=2)
    ref = (np.sum(src[None, :] / r, axis=1) + corr) / (4 * np.pi)
    assert cases.rel_max(phi[mag], ref[mag]) < 1e-13
    assert np.array_equal(phi[~mag], case.phi[~mag])       # non-magnetic nodes are not targets
    oc.close()


def test_log_stats_vs_naive():
    """ut_log-stats.cpp: Welford vs the naive two-pass formulas, tolerance 1e-12."""
    rng = np.random.default_rng(5489)
    xs = np.exp(rng.normal(np.log(0.1), 3, size=10000))
    st = LogStats()
    for x in xs:
        st.add(x)
    lg = np.log(xs)
    assert st.count() == xs.size
    assert abs(st.mean() - np.exp(lg.mean())) <= 1e-12 * np.exp(lg.mean())
    assert abs(st.stddev() - np.sqrt(np.mean((lg - lg.mean()) ** 2))) <= 1e-12 * lg.std()
