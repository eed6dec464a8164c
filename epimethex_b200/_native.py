"""ctypes binding of libepx.so (C ABI in include/epx.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible, every entry
point raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C epimethex_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libepx.so")

PAIR_COLS = 11
GENE_COLS = 15

# symbols include/epx.h declares; tests check that the library exports every one of them
EXPORTS = [
    "epx_abi_version", "epx_device_count", "epx_device_info", "epx_options_default", "epx_host_alloc",
    "epx_host_free", "epx_session_create", "epx_session_destroy", "epx_session_set_methylation_rows",
    "epx_session_set_methylation_cols", "epx_session_set_methylation_device", "epx_session_set_expression",
    "epx_session_set_index", "epx_session_run", "epx_session_run_prepare", "epx_session_run_pairs",
    "epx_session_pairs_chunk", "epx_session_wait", "epx_session_fetch", "epx_session_timing",
    "epx_session_result_device_ptr", "epx_analyze", "epx_analyze_host", "epx_multi_create", "epx_multi_destroy",
    "epx_multi_set_methylation_rows", "epx_multi_set_methylation_cols", "epx_multi_set_methylation_sliced",
    "epx_multi_analyze", "epx_multi_analyze_host", "epx_multi_groups", "epx_session_run_groups", "epx_session_fetch_groups",
    "epx_write_csv", "epx_format_real15",
]

ERRORS = {-1: "EPX_EINVAL", -2: "EPX_ECUDA", -3: "EPX_ENOMEM", -4: "EPX_ESTATE", -5: "EPX_EBREAKS",
          -6: "EPX_ESTRICT"}


class EpxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("data_linear", C.c_int32), ("test_expr_t", C.c_int32),
                ("test_meth_t", C.c_int32), ("fc_from_means", C.c_int32), ("strict_n3", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class _Result(C.Structure):
    _fields_ = [("pair", C.c_void_p), ("pair_na", C.c_void_p), ("pair_stat", C.c_void_p),
                ("pair_ks_dnum", C.c_void_p), ("gene", C.c_void_p), ("gene_na", C.c_void_p),
                ("gene_ks_dnum", C.c_void_p), ("perm", C.c_void_p), ("counts", C.c_int32 * 3),
                ("ref_gene", C.c_int32)]


class _GroupResult(C.Structure):
    _fields_ = [("group", C.c_void_p), ("group_na", C.c_void_p), ("group_stat", C.c_void_p),
                ("group_ks_dnum", C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("gene_sort_ms", C.c_float), ("gene_tests_ms", C.c_float),
                ("probe_rank_ms", C.c_float), ("pair_stats_ms", C.c_float),
                ("n_pairs", C.c_int64), ("n_unique_probes", C.c_int64), ("n_genes", C.c_int64),
                ("n_samples", C.c_int64), ("algorithmic_bytes", C.c_int64), ("kernel_launches", C.c_int32),
                ("sort_full_rows", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_lib = None


def lib():
    """Load libepx.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C epimethex_b200/csrc` "
                          "(epimethex_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    e = [C.c_char_p, C.c_size_t]
    vp, i64 = C.c_void_p, C.c_int64
    L.epx_abi_version.restype = C.c_int
    L.epx_device_count.argtypes = [C.POINTER(C.c_int)] + e
    L.epx_device_info.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(i64)] + e
    L.epx_options_default.argtypes = [C.POINTER(Options)]
    L.epx_options_default.restype = None
    L.epx_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t] + e
    L.epx_host_free.argtypes = [vp]
    L.epx_host_free.restype = None
    L.epx_session_create.argtypes = [C.c_int, C.POINTER(vp)] + e
    L.epx_session_destroy.argtypes = [vp]
    L.epx_session_destroy.restype = None
    L.epx_session_set_methylation_rows.argtypes = [vp, vp, i64, i64, i64] + e
    L.epx_session_set_methylation_cols.argtypes = [vp, C.POINTER(vp), i64, i64] + e
    L.epx_session_set_methylation_device.argtypes = [vp, vp, i64, i64, i64] + e
    L.epx_session_set_expression.argtypes = [vp, vp, i64, i64, C.c_int] + e
    L.epx_session_set_index.argtypes = [vp, vp, vp, i64] + e
    L.epx_session_run.argtypes = [vp, C.POINTER(Options), vp] + e
    L.epx_session_run_prepare.argtypes = [vp, C.POINTER(Options), vp] + e
    L.epx_session_run_pairs.argtypes = [vp, C.c_int, C.c_int, vp] + e
    L.epx_session_pairs_chunk.argtypes = [vp, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)] + e
    L.epx_session_wait.argtypes = [vp, vp] + e
    L.epx_session_fetch.argtypes = [vp, C.POINTER(_Result), vp] + e
    L.epx_session_timing.argtypes = [vp, C.POINTER(Timing)] + e
    L.epx_session_result_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)] + e
    L.epx_analyze.argtypes = [vp, vp, i64, i64, vp, vp, C.POINTER(Options), C.POINTER(_Result)] + e
    L.epx_analyze_host.argtypes = [vp, vp, i64, i64, vp, i64, i64, vp, vp, C.POINTER(Options), C.POINTER(_Result)] + e
    L.epx_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)] + e
    L.epx_multi_destroy.argtypes = [vp]
    L.epx_multi_destroy.restype = None
    L.epx_multi_set_methylation_rows.argtypes = [vp, vp, i64, i64, i64] + e
    L.epx_multi_set_methylation_cols.argtypes = [vp, C.POINTER(vp), i64, i64] + e
    L.epx_multi_groups.argtypes = [vp, vp, vp, i64, C.POINTER(_GroupResult)] + e
    L.epx_multi_set_methylation_sliced.argtypes = [vp, vp, i64, i64, i64, vp, vp, i64] + e
    L.epx_multi_analyze.argtypes = [vp, vp, i64, i64, vp, vp, C.POINTER(Options), C.POINTER(_Result)] + e
    L.epx_multi_analyze_host.argtypes = [vp, vp, i64, i64, vp, i64, i64, vp, vp, C.POINTER(Options), C.POINTER(_Result)] + e
    L.epx_session_run_groups.argtypes = [vp, vp, vp, i64, vp] + e
    L.epx_session_fetch_groups.argtypes = [vp, C.POINTER(_GroupResult), vp] + e
    L.epx_write_csv.argtypes = [C.c_char_p, i64, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), vp, vp,
                                C.POINTER(C.c_char_p), C.c_char_p, C.c_int, C.c_int] + e
    L.epx_format_real15.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
    for name in EXPORTS:
        f = getattr(L, name)
        if f.restype is C.c_int and name not in ("epx_abi_version",):
            f.restype = C.c_int
    _lib = L
    return L


def _check(rc: int, buf) -> None:
    if rc != 0:
        raise EpxError(rc, buf.value.decode("utf-8", "replace"))


def _errbuf():
    return C.create_string_buffer(512)


def format_real15(x: float) -> str:
    """One double as R prints it with 15 significant digits (as.character / write.csv)."""
    buf = C.create_string_buffer(64)
    if lib().epx_format_real15(float(x), buf, len(buf)) != 0:
        raise ValueError("epx_format_real15 failed")
    return buf.value.decode()


def write_csv(path: str, row_names, col_names, values: np.ndarray, *, text_col=None, text_col_name: str = "",
              quote_numbers: bool = False, threads: int = 0) -> None:
    """R's write.csv of a table (epx_write_csv): NaN cells are written as NA."""
    values = np.ascontiguousarray(values, np.float64)
    n_rows, n_cols = values.shape if values.ndim == 2 else (len(row_names), 0)
    enc = lambda seq: (C.c_char_p * len(seq))(*[None if v is None else str(v).encode() for v in seq])
    rn, cn = enc(list(row_names)), enc(list(col_names))
    tc = enc(list(text_col)) if text_col is not None else None
    b = _errbuf()
    _check(lib().epx_write_csv(os.fsencode(path), n_rows, n_cols, rn, cn, values.ctypes.data, None, tc,
                               text_col_name.encode(), int(quote_numbers), threads or (os.cpu_count() or 1), b, len(b)), b)


def device_count() -> int:
    n, b = C.c_int(0), _errbuf()
    _check(lib().epx_device_count(C.byref(n), b, len(b)), b)
    return n.value


def device_info(device: int = 0) -> dict:
    name, b = C.create_string_buffer(256), _errbuf()
    sm, mem = C.c_int(0), C.c_int64(0)
    _check(lib().epx_device_info(device, name, len(name), C.byref(sm), C.byref(mem), b, len(b)), b)
    return {"name": name.value.decode(), "sm_count": sm.value, "mem_bytes": mem.value}


def make_options(data_linear=True, test_expr_t=True, test_meth_t=False, fc_from_means=False,
                 strict_n3=False, serialize=False) -> Options:
    o = Options()
    lib().epx_options_default(C.byref(o))
    o.data_linear, o.test_expr_t, o.test_meth_t = int(data_linear), int(test_expr_t), int(test_meth_t)
    o.fc_from_means, o.strict_n3 = int(fc_from_means), int(strict_n3)
    o.reserved[0] = 1 if serialize else 0   # profiling aid: no overlap of the gene kernels with the probe ranking
    return o


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory from epx_host_alloc (freed with the array)."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dtype.itemsize, 1)
    p, b = C.c_void_p(), _errbuf()
    _check(lib().epx_host_alloc(C.byref(p), nbytes, b, len(b)), b)
    raw = (C.c_ubyte * nbytes).from_address(p.value)
    weakref.finalize(raw, lib().epx_host_free, p.value)
    return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)


@dataclass
class Result:
    pair: np.ndarray          # [N_p, 11]
    pair_na: np.ndarray       # [N_p] uint16 NA bit mask
    pair_stat: np.ndarray     # [N_p, 3]
    ks_dnum: np.ndarray       # [N_p, 3]
    gene: np.ndarray          # [G, 15]
    gene_na: np.ndarray       # [G]
    gene_ks_dnum: np.ndarray  # [G, 3]
    perm: np.ndarray          # [G, n] uint16
    counts: tuple = (0, 0, 0)
    ref_gene: int = -1
    timing: dict = field(default_factory=dict)


@dataclass
class GroupResult:
    group: np.ndarray        # [n_groups, 11] same columns as Result.pair
    group_na: np.ndarray     # [n_groups] uint16 NA bit mask
    group_stat: np.ndarray   # [n_groups, 3] Welch t or KS D
    group_dnum: np.ndarray   # [n_groups, 3] int64 KS numerator over (k*m)^2


class Session:
    """Device-resident state of one GPU: the methylation matrix survives across gene-range calls."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        b = _errbuf()
        _check(lib().epx_session_create(device, C.byref(self._h), b, len(b)), b)
        self.device = device
        self.n_pairs = 0
        self.n_genes = 0
        self.n_samples = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().epx_session_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- inputs
    def set_methylation(self, rows: np.ndarray):
        """rows: [P, n] probe-major float64 (C-contiguous rows, any row stride)."""
        rows = np.asarray(rows)
        if rows.dtype != np.float64 or rows.ndim != 2 or rows.strides[1] != 8:
            rows = np.ascontiguousarray(rows, np.float64)
        ld = rows.strides[0] // 8 if rows.shape[0] > 1 else rows.shape[1]
        b = _errbuf()
        _check(lib().epx_session_set_methylation_rows(self._h, rows.ctypes.data, rows.shape[0], rows.shape[1], ld,
                                                      b, len(b)), b)

    def set_methylation_columns(self, cols):
        """cols: sequence of n float64 arrays of length P (one per sample), e.g. data.frame columns."""
        cols = [np.ascontiguousarray(c, np.float64) for c in cols]
        P = cols[0].shape[0]
        assert all(c.shape == (P,) for c in cols)
        ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        b = _errbuf()
        _check(lib().epx_session_set_methylation_cols(self._h, ptrs, len(cols), P, b, len(b)), b)

    def set_methylation_device(self, ptr: int, n_probes: int, n_samples: int, ld: int):
        b = _errbuf()
        _check(lib().epx_session_set_methylation_device(self._h, ptr, n_probes, n_samples, ld, b, len(b)), b)

    def set_expression(self, expr: np.ndarray):
        expr = np.ascontiguousarray(expr, np.float64)
        b = _errbuf()
        _check(lib().epx_session_set_expression(self._h, expr.ctypes.data, expr.shape[0], expr.shape[1], 0, b, len(b)), b)
        self.n_genes, self.n_samples = expr.shape

    def set_expression_device(self, ptr: int, n_genes: int, n_samples: int):
        b = _errbuf()
        _check(lib().epx_session_set_expression(self._h, ptr, n_genes, n_samples, 1, b, len(b)), b)
        self.n_genes, self.n_samples = n_genes, n_samples

    def set_index(self, row_ptr: np.ndarray, probe_idx: np.ndarray):
        row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        probe_idx = np.ascontiguousarray(probe_idx, np.int32)
        assert probe_idx.size == int(row_ptr[-1])
        b = _errbuf()
        _check(lib().epx_session_set_index(self._h, row_ptr.ctypes.data, probe_idx.ctypes.data, row_ptr.size - 1,
                                           b, len(b)), b)
        self.n_pairs = int(row_ptr[-1])

    # -- compute
    def run(self, options: Options | None = None, stream: int | None = None):
        b = _errbuf()
        opt = options if options is not None else make_options()
        _check(lib().epx_session_run(self._h, C.byref(opt), C.c_void_p(stream or 0), b, len(b)), b)

    def run_prepare(self, options: Options | None = None, stream: int | None = None):
        b = _errbuf()
        opt = options if options is not None else make_options()
        _check(lib().epx_session_run_prepare(self._h, C.byref(opt), C.c_void_p(stream or 0), b, len(b)), b)

    def run_pairs(self, chunk: int, n_chunks: int, stream: int | None = None):
        b = _errbuf()
        _check(lib().epx_session_run_pairs(self._h, chunk, n_chunks, C.c_void_p(stream or 0), b, len(b)), b)

    def pairs_chunk(self, chunk: int, n_chunks: int):
        first, count, b = C.c_int64(), C.c_int64(), _errbuf()
        _check(lib().epx_session_pairs_chunk(self._h, chunk, n_chunks, C.byref(first), C.byref(count), b, len(b)), b)
        return first.value, count.value

    def wait(self, stream: int | None = None):
        b = _errbuf()
        _check(lib().epx_session_wait(self._h, C.c_void_p(stream or 0), b, len(b)), b)

    def timing(self) -> dict:
        t, b = Timing(), _errbuf()
        _check(lib().epx_session_timing(self._h, C.byref(t), b, len(b)), b)
        return t.as_dict()

    def result_device_ptr(self, which: int):
        p, n, b = C.c_void_p(), C.c_int64(), _errbuf()
        _check(lib().epx_session_result_device_ptr(self._h, which, C.byref(p), C.byref(n), b, len(b)), b)
        return p.value, n.value

    def _alloc_result(self, full: bool, pinned: bool = False):
        Np, G, n = self.n_pairs, self.n_genes, self.n_samples
        if pinned and not full:
            cap = getattr(self, "_pinned_cap", (0, 0))
            if Np > cap[0] or G > cap[1]:          # grow only: page-locking 40 MB costs tens of milliseconds
                cap = (max(Np, cap[0]), max(G, cap[1]))
                self._pinned = (pinned_empty((cap[0], PAIR_COLS)), pinned_empty((cap[0],), np.uint16),
                                pinned_empty((cap[1], GENE_COLS)), pinned_empty((cap[1],), np.uint16))
                self._pinned_cap = cap
            pr, pn, gn, gm = (self._pinned[0][:Np], self._pinned[1][:Np], self._pinned[2][:G], self._pinned[3][:G])
            res = Result(pair=pr, pair_na=pn, pair_stat=np.empty((0, 3)), ks_dnum=np.empty((0, 3), np.int32), gene=gn,
                         gene_na=gm, gene_ks_dnum=np.empty((0, 3), np.int32), perm=np.empty((0, n), np.uint16))
            r = _Result()
            r.pair, r.pair_na, r.gene, r.gene_na = pr.ctypes.data, pn.ctypes.data, gn.ctypes.data, gm.ctypes.data
            return res, r
        res = Result(pair=np.empty((Np, PAIR_COLS)), pair_na=np.zeros(Np, np.uint16),
                     pair_stat=np.empty((Np, 3)) if full else np.empty((0, 3)),
                     ks_dnum=np.empty((Np, 3), np.int32) if full else np.empty((0, 3), np.int32),
                     gene=np.empty((G, GENE_COLS)), gene_na=np.zeros(G, np.uint16),
                     gene_ks_dnum=np.empty((G, 3), np.int32) if full else np.empty((0, 3), np.int32),
                     perm=np.empty((G, n), np.uint16) if full else np.empty((0, n), np.uint16))
        r = _Result()
        r.pair, r.pair_na = res.pair.ctypes.data, res.pair_na.ctypes.data
        r.gene, r.gene_na = res.gene.ctypes.data, res.gene_na.ctypes.data
        if full:
            r.pair_stat, r.pair_ks_dnum = res.pair_stat.ctypes.data, res.ks_dnum.ctypes.data
            r.gene_ks_dnum, r.perm = res.gene_ks_dnum.ctypes.data, res.perm.ctypes.data
        return res, r

    def fetch(self, full: bool = True, stream: int | None = None) -> Result:
        res, r = self._alloc_result(full)
        b = _errbuf()
        _check(lib().epx_session_fetch(self._h, C.byref(r), C.c_void_p(stream or 0), b, len(b)), b)
        res.counts, res.ref_gene = tuple(r.counts), r.ref_gene
        res.timing = self.timing()
        return res

    def run_groups(self, group_ptr, group_pairs, stream: int | None = None):
        """Pooled analyses of probe groups (CG_of_a_genes / CG_by_position / CG_by_islands rows); asynchronous.
        Call after run()/run_prepare() of the same expression + index."""
        group_ptr = np.ascontiguousarray(group_ptr, np.int64)
        group_pairs = np.ascontiguousarray(group_pairs, np.int32)
        assert group_ptr.ndim == 1 and group_ptr.size >= 1 and group_pairs.size == int(group_ptr[-1])
        self.n_groups = group_ptr.size - 1
        b = _errbuf()
        _check(lib().epx_session_run_groups(self._h, group_ptr.ctypes.data, group_pairs.ctypes.data, self.n_groups,
                                            C.c_void_p(stream or 0), b, len(b)), b)

    def fetch_groups(self, stream: int | None = None) -> GroupResult:
        q = self.n_groups
        res = GroupResult(group=np.empty((q, 11)), group_na=np.zeros(q, np.uint16), group_stat=np.empty((q, 3)),
                          group_dnum=np.empty((q, 3), np.int64))
        r = _GroupResult(res.group.ctypes.data, res.group_na.ctypes.data, res.group_stat.ctypes.data,
                         res.group_dnum.ctypes.data)
        b = _errbuf()
        _check(lib().epx_session_fetch_groups(self._h, C.byref(r), C.c_void_p(stream or 0), b, len(b)), b)
        return res

    def analyze_groups(self, group_ptr, group_pairs) -> GroupResult:
        self.run_groups(group_ptr, group_pairs)
        return self.fetch_groups()

    def analyze(self, expr, row_ptr, probe_idx, options: Options | None = None, full: bool = True,
                pinned_results: bool = False) -> Result:
        """One blocking epx_analyze call with host buffers (what the R `.Call` shim does).
        pinned_results: write the result tables into page-locked buffers owned by the session (reused by the next
        call: copy what you keep)."""
        expr = np.ascontiguousarray(expr, np.float64)
        row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        probe_idx = np.ascontiguousarray(probe_idx, np.int32)
        self.n_genes, self.n_samples = expr.shape
        self.n_pairs = int(row_ptr[-1])
        res, r = self._alloc_result(full, pinned_results)
        opt = options if options is not None else make_options()
        b = _errbuf()
        _check(lib().epx_analyze(self._h, expr.ctypes.data, expr.shape[0], expr.shape[1], row_ptr.ctypes.data,
                                 probe_idx.ctypes.data, C.byref(opt), C.byref(r), b, len(b)), b)
        res.counts, res.ref_gene = tuple(r.counts), r.ref_gene
        res.timing = self.timing()
        return res

    def analyze_host(self, meth_rows, expr, row_ptr, probe_idx, options: Options | None = None, full: bool = True,
                     pinned_results: bool = False) -> Result:
        """One blocking epx_analyze_host call: the methylation matrix stays in host memory ([P, n] probe-major float64,
        any row stride) and only the rows the index names are uploaded -- fetched by the device itself when the matrix
        is page-locked (pinned_empty), else the whole matrix through the staged upload. The session then holds a
        partial matrix: the next call is analyze_host again, or set_methylation."""
        rows = np.asarray(meth_rows)
        if rows.dtype != np.float64 or rows.ndim != 2 or rows.strides[1] != 8:
            rows = np.ascontiguousarray(rows, np.float64)
        ld = rows.strides[0] // 8 if rows.shape[0] > 1 else rows.shape[1]
        expr = np.ascontiguousarray(expr, np.float64)
        row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        probe_idx = np.ascontiguousarray(probe_idx, np.int32)
        if rows.shape[1] != expr.shape[1]:
            raise ValueError(f"sample counts differ: methylation {rows.shape[1]}, expression {expr.shape[1]}")
        self.n_genes, self.n_samples = expr.shape
        self.n_pairs = int(row_ptr[-1])
        res, r = self._alloc_result(full, pinned_results)
        opt = options if options is not None else make_options()
        b = _errbuf()
        _check(lib().epx_analyze_host(self._h, rows.ctypes.data, rows.shape[0], ld, expr.ctypes.data, expr.shape[0],
                                      expr.shape[1], row_ptr.ctypes.data, probe_idx.ctypes.data, C.byref(opt),
                                      C.byref(r), b, len(b)), b)
        res.counts, res.ref_gene = tuple(r.counts), r.ref_gene
        res.timing = self.timing()
        return res


class MultiSession:
    """Several GPUs driven from this one process (what the R shim would use): genes are split by pair count,
    every device computes its pairs and writes them straight into the host result tables."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        b = _errbuf()
        _check(lib().epx_multi_create(devs, len(devices), C.byref(self._h), b, len(b)), b)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().epx_multi_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_methylation(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, np.float64)
        b = _errbuf()
        _check(lib().epx_multi_set_methylation_rows(self._h, rows.ctypes.data, rows.shape[0], rows.shape[1],
                                                    rows.shape[1], b, len(b)), b)

    def set_methylation_columns(self, cols):
        """cols: n float64 arrays of length P (one per sample: data.frame columns); replicated on every device"""
        cols = [np.ascontiguousarray(c, np.float64) for c in cols]
        P = cols[0].shape[0]
        assert all(c.shape == (P,) for c in cols)
        ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        b = _errbuf()
        _check(lib().epx_multi_set_methylation_cols(self._h, ptrs, len(cols), P, b, len(b)), b)

    def analyze_groups(self, group_ptr, group_pairs) -> GroupResult:
        """pooled tables of the index of the last analyze(): every group runs on the device that owns its gene"""
        group_ptr = np.ascontiguousarray(group_ptr, np.int64)
        group_pairs = np.ascontiguousarray(group_pairs, np.int32)
        q = group_ptr.size - 1
        res = GroupResult(group=np.empty((q, 11)), group_na=np.zeros(q, np.uint16), group_stat=np.empty((q, 3)),
                          group_dnum=np.empty((q, 3), np.int64))
        r = _GroupResult(res.group.ctypes.data, res.group_na.ctypes.data, res.group_stat.ctypes.data,
                         res.group_dnum.ctypes.data)
        b = _errbuf()
        _check(lib().epx_multi_groups(self._h, group_ptr.ctypes.data, group_pairs.ctypes.data, q, C.byref(r), b, len(b)), b)
        return res

    def set_methylation_sliced(self, rows: np.ndarray, row_ptr, probe_idx):
        """probe-sliced copies: each device gets only the rows its share of the genes references"""
        rows = np.ascontiguousarray(rows, np.float64)
        row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        probe_idx = np.ascontiguousarray(probe_idx, np.int32)
        b = _errbuf()
        _check(lib().epx_multi_set_methylation_sliced(self._h, rows.ctypes.data, rows.shape[0], rows.shape[1],
                                                      rows.shape[1], row_ptr.ctypes.data, probe_idx.ctypes.data,
                                                      row_ptr.size - 1, b, len(b)), b)

    def analyze_host(self, meth_rows, expr, row_ptr, probe_idx, options: Options | None = None, full: bool = True,
                     pinned_results: bool = False) -> Result:
        """One blocking epx_multi_analyze_host call: the matrix stays in (page-locked) host memory, every device fetches
        the rows its own gene range names."""
        return self.analyze(expr, row_ptr, probe_idx, options, full, pinned_results, meth_rows=meth_rows)

    def analyze(self, expr, row_ptr, probe_idx, options: Options | None = None, full: bool = True,
                pinned_results: bool = False, meth_rows=None) -> Result:
        """One blocking epx_multi_analyze call. full / pinned_results as in Session.analyze: the two result tables only,
        written by every device straight into page-locked buffers this object keeps (copy what you keep)."""
        expr = np.ascontiguousarray(expr, np.float64)
        row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        probe_idx = np.ascontiguousarray(probe_idx, np.int32)
        G, n = expr.shape
        Np = int(row_ptr[-1])
        r = _Result()
        if pinned_results and not full:
            cap = getattr(self, "_pinned_cap", (0, 0))
            if Np > cap[0] or G > cap[1]:
                cap = (max(Np, cap[0]), max(G, cap[1]))
                self._pinned = (pinned_empty((cap[0], PAIR_COLS)), pinned_empty((cap[0],), np.uint16),
                                pinned_empty((cap[1], GENE_COLS)), pinned_empty((cap[1],), np.uint16))
                self._pinned_cap = cap
            pr, pn, gn, gm = (self._pinned[0][:Np], self._pinned[1][:Np], self._pinned[2][:G], self._pinned[3][:G])
            res = Result(pair=pr, pair_na=pn, pair_stat=np.empty((0, 3)), ks_dnum=np.empty((0, 3), np.int32), gene=gn,
                         gene_na=gm, gene_ks_dnum=np.empty((0, 3), np.int32), perm=np.empty((0, n), np.uint16))
            r.pair, r.pair_na, r.gene, r.gene_na = pr.ctypes.data, pn.ctypes.data, gn.ctypes.data, gm.ctypes.data
        else:
            res = Result(pair=np.empty((Np, PAIR_COLS)), pair_na=np.zeros(Np, np.uint16),
                         pair_stat=np.empty((Np, 3)) if full else np.empty((0, 3)),
                         ks_dnum=np.empty((Np, 3), np.int32) if full else np.empty((0, 3), np.int32),
                         gene=np.empty((G, GENE_COLS)), gene_na=np.zeros(G, np.uint16),
                         gene_ks_dnum=np.empty((G, 3), np.int32) if full else np.empty((0, 3), np.int32),
                         perm=np.empty((G, n), np.uint16) if full else np.empty((0, n), np.uint16))
            r.pair, r.pair_na = res.pair.ctypes.data, res.pair_na.ctypes.data
            r.gene, r.gene_na = res.gene.ctypes.data, res.gene_na.ctypes.data
            if full:
                r.pair_stat, r.pair_ks_dnum = res.pair_stat.ctypes.data, res.ks_dnum.ctypes.data
                r.gene_ks_dnum, r.perm = res.gene_ks_dnum.ctypes.data, res.perm.ctypes.data
        opt = options if options is not None else make_options()
        b = _errbuf()
        if meth_rows is None:
            _check(lib().epx_multi_analyze(self._h, expr.ctypes.data, G, n, row_ptr.ctypes.data, probe_idx.ctypes.data,
                                           C.byref(opt), C.byref(r), b, len(b)), b)
        else:
            rows = np.asarray(meth_rows)
            if rows.dtype != np.float64 or rows.ndim != 2 or rows.strides[1] != 8:
                rows = np.ascontiguousarray(rows, np.float64)
            if rows.shape[1] != n:
                raise ValueError(f"sample counts differ: methylation {rows.shape[1]}, expression {n}")
            ld = rows.strides[0] // 8 if rows.shape[0] > 1 else rows.shape[1]
            _check(lib().epx_multi_analyze_host(self._h, rows.ctypes.data, rows.shape[0], ld, expr.ctypes.data, G, n,
                                                row_ptr.ctypes.data, probe_idx.ctypes.data, C.byref(opt), C.byref(r),
                                                b, len(b)), b)
        res.counts, res.ref_gene = tuple(r.counts), r.ref_gene
        return res
