"""ctypes binding of the CPU oracle (oracle/epx_oracle.c).

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED (see oracle/epx_oracle.h).  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module; the product package ``epimethex_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libepx_oracle.so")

PAIR_COLS = 11
GENE_COLS = 15

PAIR_COLUMN_NAMES = [
    "medianDown", "medianMedium", "medianUP", "bd_UPvsMID", "bd_UPvsDOWN", "bd_MIDvsDOWN",
    "pvalue_UPvsMID", "pvalue_UPvsDOWN", "pvalue_MIDvsDOWN", "pearson_correlation",
    "pvalue_pearson_correlation",
]


def build(force: bool = False) -> str:
    """Compile the oracle with the Makefile beside it (gcc only)."""
    src = os.path.join(_HERE, "epx_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


class _Options(C.Structure):
    _fields_ = [("data_linear", C.c_int32), ("test_expr_t", C.c_int32), ("test_meth_t", C.c_int32),
                ("fc_from_means", C.c_int32), ("threads", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.epo_order_desc.argtypes = [dp, C.c_int, ip]
        L.epo_quantile7.argtypes = [dp, C.c_int, C.c_double]; L.epo_quantile7.restype = C.c_double
        L.epo_binvar_counts.argtypes = [dp, C.c_int, ip]; L.epo_binvar_counts.restype = C.c_int
        for f in (L.epo_log_fc, L.epo_linear_fc):
            f.argtypes = [C.c_double, C.c_double]; f.restype = C.c_double
        L.epo_guard.argtypes = [dp, C.c_int, dp, C.c_int]; L.epo_guard.restype = C.c_int
        L.epo_welch.argtypes = [dp, C.c_int, dp, C.c_int, dp, dp, dp]; L.epo_welch.restype = C.c_int
        L.epo_pt.argtypes = [C.c_double, C.c_double]; L.epo_pt.restype = C.c_double
        L.epo_pbeta.argtypes = [C.c_double] * 3; L.epo_pbeta.restype = C.c_double
        L.epo_ks.argtypes = [dp, C.c_int, dp, C.c_int, C.POINTER(C.c_int64), dp, C.POINTER(C.c_int)]
        L.epo_ks.restype = C.c_int
        L.epo_psmirnov2x.argtypes = [C.c_double, C.c_int, C.c_int]; L.epo_psmirnov2x.restype = C.c_double
        L.epo_pkstwo.argtypes = [C.c_double]; L.epo_pkstwo.restype = C.c_double
        L.epo_median.argtypes = [dp, C.c_int]; L.epo_median.restype = C.c_double
        L.epo_pearson.argtypes = [dp, dp, C.c_int, dp, dp]; L.epo_pearson.restype = C.c_int
        L.epo_analyze.restype = C.c_int
        L.epo_analyze.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.POINTER(_Options)] + [C.c_void_p] * 10
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def order_desc(x):
    a, p = _d(x)
    out = np.empty(a.size, np.int32)
    lib().epo_order_desc(p, a.size, out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def quantile7(asc, p):
    a, ptr = _d(asc)
    return lib().epo_quantile7(ptr, a.size, float(p))


def binvar_counts(x_desc):
    a, ptr = _d(x_desc)
    out = np.zeros(3, np.int32)
    rc = lib().epo_binvar_counts(ptr, a.size, out.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise ValueError("'breaks' are not unique")
    return tuple(int(v) for v in out)  # (nUP, nM, nD)


def log_fc(a, b):
    return lib().epo_log_fc(float(a), float(b))


def linear_fc(a, b):
    return lib().epo_linear_fc(float(a), float(b))


def guard(a, b):
    a, pa = _d(a); b, pb = _d(b)
    return lib().epo_guard(pa, a.size, pb, b.size)


def welch(a, b):
    a, pa = _d(a); b, pb = _d(b)
    t, df, p = C.c_double(), C.c_double(), C.c_double()
    rc = lib().epo_welch(pa, a.size, pb, b.size, C.byref(t), C.byref(df), C.byref(p))
    return rc, t.value, df.value, p.value


def pt(q, df):
    return lib().epo_pt(float(q), float(df))


def pbeta(x, a, b):
    return lib().epo_pbeta(float(x), float(a), float(b))


def ks(a, b):
    a, pa = _d(a); b, pb = _d(b)
    dn, p, ex = C.c_int64(), C.c_double(), C.c_int()
    rc = lib().epo_ks(pa, a.size, pb, b.size, C.byref(dn), C.byref(p), C.byref(ex))
    return rc, dn.value, p.value, bool(ex.value)


def psmirnov2x(stat, m, n):
    return lib().epo_psmirnov2x(float(stat), int(m), int(n))


def pkstwo(x):
    return lib().epo_pkstwo(float(x))


def median(x):
    a, p = _d(x)
    return lib().epo_median(p, a.size)


def pearson_p_from_r(r, n):
    """two-sided p of a correlation r over n samples, as psych::corr.test computes it"""
    t = r * np.sqrt(n - 2.0) / np.sqrt(1.0 - r * r)
    return 2.0 * pt(-abs(t), n - 2.0)


def pearson(x, y):
    x, px = _d(x); y, py = _d(y)
    r, p = C.c_double(), C.c_double()
    rc = lib().epo_pearson(px, py, x.size, C.byref(r), C.byref(p))
    return rc, r.value, p.value


@dataclass
class OracleResult:
    pair: np.ndarray        # [N_p, 11]
    pair_na: np.ndarray     # [N_p] uint16 bit mask
    pair_stat: np.ndarray   # [N_p, 3] Welch t or KS D
    ks_dnum: np.ndarray     # [N_p, 3] int32, -1 = no KS
    gene: np.ndarray        # [G, 15]
    gene_na: np.ndarray     # [G] uint16
    gene_ks_dnum: np.ndarray  # [G, 3]
    perm: np.ndarray        # [G, n] int32
    counts: tuple           # (nUP, nM, nD)
    ref_gene: int


def analyze(expr, meth, row_ptr, probe_idx, *, data_linear=True, test_expr_t=True, test_meth_t=False,
            fc_from_means=False, threads=1) -> OracleResult:
    """expr [G, n] gene-major, meth [P, n] probe-major (C-contiguous rows, any leading dimension)."""
    expr = np.ascontiguousarray(expr, np.float64)
    G, n = expr.shape
    meth = np.asarray(meth, np.float64)
    assert meth.ndim == 2 and meth.strides[1] == 8 and meth.shape[1] >= n
    ld = meth.strides[0] // 8 if meth.shape[0] > 1 else meth.shape[1]
    row_ptr = np.ascontiguousarray(row_ptr, np.int64)
    probe_idx = np.ascontiguousarray(probe_idx, np.int32)
    npairs = int(row_ptr[-1])
    assert row_ptr.size == G + 1 and probe_idx.size == npairs
    res = OracleResult(
        pair=np.empty((npairs, PAIR_COLS)), pair_na=np.zeros(npairs, np.uint16),
        pair_stat=np.empty((npairs, 3)), ks_dnum=np.empty((npairs, 3), np.int32),
        gene=np.empty((G, GENE_COLS)), gene_na=np.zeros(G, np.uint16),
        gene_ks_dnum=np.empty((G, 3), np.int32), perm=np.empty((G, n), np.int32), counts=(0, 0, 0), ref_gene=-1)
    counts = np.zeros(3, np.int32)
    ref = C.c_int32(-1)
    opt = _Options(int(data_linear), int(test_expr_t), int(test_meth_t), int(fc_from_means), int(threads))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().epo_analyze(vp(expr), n, G, vp(meth), ld, vp(row_ptr), vp(probe_idx), C.byref(opt),
                           vp(res.pair), vp(res.pair_na), vp(res.pair_stat), vp(res.ks_dnum),
                           vp(res.gene), vp(res.gene_na), vp(res.gene_ks_dnum), vp(res.perm),
                           vp(counts), C.cast(C.byref(ref), C.c_void_p))
    if rc == -2:
        raise ValueError("bin.var: 'breaks' are not unique")
    if rc != 0:
        raise ValueError(f"epo_analyze failed with code {rc}")
    res.counts = tuple(int(v) for v in counts)
    res.ref_gene = int(ref.value)
    return res


@dataclass
class GroupResult:
    group: np.ndarray       # [n_groups, 11] same columns as the pair table
    group_na: np.ndarray    # [n_groups] uint16 bit mask
    group_stat: np.ndarray  # [n_groups, 3] Welch t or KS D
    group_dnum: np.ndarray  # [n_groups, 3] int64 KS numerator over (k*m)^2, -1 = no KS


def analyze_groups(expr, meth, row_ptr, probe_idx, group_ptr, group_pairs, *, test_meth_t=False, threads=1) -> GroupResult:
    """Pooled analyses (CG_of_a_genes / by position / by island): see epo_analyze_groups in epx_oracle.h."""
    expr = np.ascontiguousarray(expr, np.float64)
    G, n = expr.shape
    meth = np.asarray(meth, np.float64)
    assert meth.ndim == 2 and meth.strides[1] == 8 and meth.shape[1] >= n
    ld = meth.strides[0] // 8 if meth.shape[0] > 1 else meth.shape[1]
    row_ptr = np.ascontiguousarray(row_ptr, np.int64)
    probe_idx = np.ascontiguousarray(probe_idx, np.int32)
    group_ptr = np.ascontiguousarray(group_ptr, np.int64)
    group_pairs = np.ascontiguousarray(group_pairs, np.int32)
    ng = group_ptr.size - 1
    assert group_pairs.size == int(group_ptr[-1])
    res = GroupResult(group=np.empty((ng, PAIR_COLS)), group_na=np.zeros(ng, np.uint16),
                      group_stat=np.empty((ng, 3)), group_dnum=np.empty((ng, 3), np.int64))
    opt = _Options(1, 1, int(test_meth_t), 0, int(threads))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    L = lib()
    L.epo_analyze_groups.restype = C.c_int
    L.epo_analyze_groups.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    rc = L.epo_analyze_groups(vp(expr), n, G, vp(meth), ld, vp(row_ptr), vp(probe_idx), C.cast(C.byref(opt), C.c_void_p),
                              ng, vp(group_ptr), vp(group_pairs), vp(res.group), vp(res.group_na), vp(res.group_stat),
                              vp(res.group_dnum))
    if rc != 0:
        raise ValueError(f"epo_analyze_groups failed with code {rc}")
    return res
