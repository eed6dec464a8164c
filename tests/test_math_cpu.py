"""epimethex_b200/csrc/epx_math.h (the header the CUDA kernels include) compiled for the CPU with gcc and
checked against mpmath / the oracle. This is a unit test of the device math source, not a fallback path."""
import ctypes as C
import os

import mpmath as mp
import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mt():
    L = C.CDLL(os.path.join(HERE, "_mathtest", "libepx_mathtest.so"))
    d = C.c_double
    for name, args in (("mt_t_two_sided_p", 2), ("mt_pkstwo", 1), ("mt_ks_p_asym", 3), ("mt_log_fc", 2),
                       ("mt_linear_fc", 2), ("mt_pearson_p", 2), ("mt_lbeta_half", 1)):
        f = getattr(L, name); f.restype = d; f.argtypes = [d] * args
    L.mt_ks_p_exact.restype = d
    L.mt_ks_p_exact.argtypes = [d, C.c_int, C.c_int, C.POINTER(d)]
    L.mt_welch.restype = C.c_int
    L.mt_welch.argtypes = [d] * 6 + [C.POINTER(d)] * 2
    return L


def test_student_tail_vs_mpmath(mt):
    mp.mp.dps = 40
    worst = 0.0
    for df in (1, 1.5, 2, 3.7, 10, 39.9, 40.1, 157.3, 314, 471, 998, 2999.5, 5998, 8998, 17998):
        for t in (1e-8, 1e-3, 0.1, 0.5, 1, 1.5, 2, 3, 5, 8, 12, 20, 40):
            x = mp.mpf(df) / (mp.mpf(df) + mp.mpf(t) ** 2)
            want = mp.betainc(mp.mpf(df) / 2, mp.mpf(1) / 2, 0, x, regularized=True)
            got = mt.mt_t_two_sided_p(t, df)
            err = abs(got - want)
            assert err <= 1e-10 or err <= 1e-8 * want, (t, df, got, want)   # BASELINE tolerance
            assert err <= 2e-12 * want + 1e-300, (t, df, got, want)          # what it actually achieves
            worst = max(worst, float(err / want))
    assert mt.mt_t_two_sided_p(0.0, 10.0) == 1.0
    assert mt.mt_t_two_sided_p(float("inf"), 10.0) == 0.0
    assert np.isnan(mt.mt_t_two_sided_p(float("nan"), 10.0))


def test_lbeta_half(mt):
    mp.mp.dps = 40
    for a in (0.5, 1, 7.25, 19.99, 20, 20.01, 78.5, 235.5, 4499):
        want = mp.log(mp.beta(a, mp.mpf(1) / 2))
        assert abs(mt.mt_lbeta_half(a) - want) < 5e-15 * max(1, abs(want))


def test_ks_pvalues_match_oracle(mt):
    for x in np.linspace(0.05, 3.0, 60):
        assert mt.mt_pkstwo(x) == pytest.approx(oracle.pkstwo(x), rel=1e-15, abs=1e-300)
    u = (C.c_double * 256)()
    for m, n in ((2, 2), (10, 10), (33, 34), (99, 99), (7, 50)):
        for dnum in range(0, m * n + 1, max(1, m * n // 17)):
            want = min(1.0, max(0.0, 1 - oracle.psmirnov2x(dnum / (m * n), m, n)))
            assert mt.mt_ks_p_exact(float(dnum), m, n, u) == pytest.approx(want, rel=1e-15, abs=1e-300)
    for m in (100, 157, 158, 333, 3000):
        for d in (0, 1, m // 10, m // 3, m):
            D = d * m / (m * m)
            want = min(1.0, max(0.0, 1 - oracle.pkstwo(np.sqrt(m * m / (2.0 * m)) * D)))
            assert mt.mt_ks_p_asym(float(d * m), float(m), float(m)) == pytest.approx(want, rel=1e-14, abs=1e-300)


def test_welch_and_fc_match_oracle(mt):
    rng = np.random.default_rng(1)
    for _ in range(20):
        a, b = rng.normal(0.5, 0.1, 157), rng.normal(0.5, 0.2, 158)
        rc, t, df, p = oracle.welch(a, b)
        tt, dd = C.c_double(), C.c_double()
        assert mt.mt_welch(a.mean(), a.var(ddof=1), 157.0, b.mean(), b.var(ddof=1), 158.0, C.byref(tt), C.byref(dd)) == 0
        assert tt.value == pytest.approx(t, rel=1e-11) and dd.value == pytest.approx(df, rel=1e-11)
        assert mt.mt_t_two_sided_p(tt.value, dd.value) == pytest.approx(p, rel=1e-9, abs=1e-12)
    for a, b in ((4.0, 2.0), (2.0, 4.0), (1.5, -0.5), (0.0, 2.0), (2.0, 2.0), (-3.0, -1.0), (0.0, 0.0)):
        got, want = mt.mt_linear_fc(a, b), oracle.linear_fc(a, b)
        assert (np.isnan(got) and np.isnan(want)) or got == want
        assert mt.mt_log_fc(a, b) == oracle.log_fc(a, b)


def test_pearson_p_table_vs_mpmath(mt):
    """the pair kernel interpolates a per-n table of h(r) = log p - (df/2) log(1-r^2); check the whole scheme"""
    mt.mt_pearson_h.restype = C.c_double; mt.mt_pearson_h.argtypes = [C.c_double] * 2
    mt.mt_pearson_grid.restype = C.c_int; mt.mt_pearson_grid.argtypes = [C.c_double]
    mt.mt_pearson_p_tab.restype = C.c_double
    mt.mt_pearson_p_tab.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_int, C.c_double]
    mp.mp.dps = 30
    rng = np.random.default_rng(5)
    for n in (6, 30, 99, 473, 1000, 9000):
        df = n - 2.0
        N = mt.mt_pearson_grid(df)
        tab = np.array([mt.mt_pearson_h(i / N, df) for i in range(N + 1)])
        ptr = tab.ctypes.data_as(C.POINTER(C.c_double))
        rs = np.concatenate([rng.random(120), rng.random(120) * 8 / np.sqrt(df) * rng.random(120), 1 - rng.random(20) * 2 / N,
                             [0.0, 1e-12, 0.5]])
        for r in rs[rs < 1]:
            want = mp.betainc(mp.mpf(df) / 2, mp.mpf(1) / 2, 0, 1 - mp.mpf(float(r)) ** 2, regularized=True)
            got = mt.mt_pearson_p_tab(-float(r), ptr, N, df)
            err = abs(got - want)
            assert err <= 1e-11 or err <= 1e-9 * want, (n, r, got, want)
            if want > 1e-300:
                assert abs(got - oracle.pearson_p_from_r(float(r), n)) <= 1e-10 + 1e-8 * want
        assert mt.mt_pearson_p_tab(1.0, ptr, N, df) == 0.0 and mt.mt_pearson_p_tab(-1.0, ptr, N, df) == 0.0


def test_welch_p_table_vs_mpmath(mt):
    """t mode of the pair kernel: p(t, df) by 4 x 4 Lagrange interpolation of the (df, r) table of h(r; df) over the
    Welch-Satterthwaite range m - 1 <= df <= 2(m - 1); must stay far inside 1e-10 abs / 1e-8 rel everywhere"""
    mt.mt_welch_p_tab.restype = C.c_double
    mt.mt_welch_p_tab.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    dp = C.POINTER(C.c_double)
    mp.mp.dps = 30
    rng = np.random.default_rng(3)
    for m in (30, 157, 333, 3000):
        N, K, df0, ddf = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        mt.mt_welch_grid(m, C.byref(N), C.byref(K), C.byref(df0), C.byref(ddf))
        ld = N.value + 1
        tab = np.empty((K.value, ld))
        mt.mt_welch_table(m, tab.ctypes.data_as(dp), ld)
        assert df0.value < m - 1 and df0.value + (K.value - 2) * ddf.value >= 2 * (m - 1)
        cases = [(0.0, m - 1.0), (1e-9, 2.0 * (m - 1)), (40.0, 1.5 * m), (-3.0, m - 1.0), (2.5, 2.0 * (m - 1))]
        for _ in range(150):
            cases.append((float(rng.choice([rng.normal(0, 1.5), rng.normal(0, 6), rng.random() * 0.01])), (m - 1) * (1 + rng.random())))
        for t, df in cases:
            want = mp.betainc(mp.mpf(df) / 2, mp.mpf(1) / 2, 0, mp.mpf(df) / (mp.mpf(df) + mp.mpf(t) ** 2), regularized=True)
            got = mt.mt_welch_p_tab(t, df, tab.ctypes.data_as(dp), N.value, ld, K.value, df0.value, 1.0 / ddf.value)
            assert abs(got - want) <= 0.02 * (1e-10 + 1e-8 * want), (m, t, df, got, want)
            assert abs(got - mt.mt_t_two_sided_p(t, df)) <= 1e-10 + 1e-8 * want     # the continued fraction it replaces
        assert mt.mt_welch_p_tab(float("inf"), 200.0, tab.ctypes.data_as(dp), N.value, ld, K.value, df0.value, 1.0 / ddf.value) == 0.0
        assert np.isnan(mt.mt_welch_p_tab(float("nan"), 200.0, tab.ctypes.data_as(dp), N.value, ld, K.value, df0.value, 1.0 / ddf.value))
