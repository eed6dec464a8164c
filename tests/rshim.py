"""ctypes binding of tests/_rstub/libepx_rshim_test.so: the .Call shim of r_package/src/epx_rshim.c compiled against
the fake R C API of tests/_rstub/ (see Rinternals.h there) plus a plain-C driver.  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_rstub", "libepx_rshim_test.so")
NA_INT = -2147483648


class Report(C.Structure):
    _fields_ = [("rc", C.c_int), ("protect_depth", C.c_int), ("live_finalizers", C.c_int), ("collections", C.c_long),
                ("ralloc_live", C.c_long), ("msg", C.c_char * 512)]

    @property
    def message(self):
        return self.msg.decode("utf-8", "replace")


_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", os.path.join(HERE, "_rstub")], check=True, capture_output=True)
        L = C.CDLL(LIB)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.rshim_upload.restype = vp
        L.rshim_upload.argtypes = [C.POINTER(vp), C.c_int, C.c_long, ip, ip, C.c_int, C.c_int, C.POINTER(Report)]
        L.rshim_analyze.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_long, vp, C.c_long, vp, C.c_int, C.c_int,
                                    vp, vp, vp, vp, vp, vp, C.POINTER(Report)]
        L.rshim_groups.argtypes = [vp, vp, C.c_long, vp, C.c_long, vp, C.POINTER(Report)]
        L.rshim_release.argtypes = [vp, C.POINTER(Report)]
        L.rshim_device_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Report)]
        L.rshim_finish.argtypes = [C.POINTER(Report)]
        L.rshim_finish.restype = None
        L.rshim_registered.argtypes = [C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


def registered():
    buf = C.create_string_buffer(1024)
    dynamic = lib().rshim_registered(buf, len(buf))
    return dict((k, int(v)) for k, v in (item.split("/") for item in buf.value.decode().strip(";").split(";"))), dynamic


def upload(cols, devices=(0,), as_int=None, bad_kind=0):
    """cols: list of n float64 arrays of length P. Returns (session handle or None, Report)."""
    cols = [np.ascontiguousarray(c, np.float64) for c in cols]
    ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    dev = (C.c_int * len(devices))(*devices)
    flags = (C.c_int * len(cols))(*(as_int if as_int is not None else [0] * len(cols)))
    rep = Report()
    h = lib().rshim_upload(ptrs, len(cols), cols[0].shape[0], flags, dev, len(devices), bad_kind, C.byref(rep))
    return h, rep


def analyze(session, expr_gene_major, row_ptr, probe_idx, opts=(1, 1, 0, 0, 0), interrupt_at=0, n_opts=None):
    """expr_gene_major [G, n] C-contiguous == R's n x G column-major matrix. probe_idx: -1 -> NA_integer_."""
    x = np.ascontiguousarray(expr_gene_major, np.float64)
    G, n = x.shape
    rp = np.ascontiguousarray(row_ptr, np.float64)
    pi = np.ascontiguousarray(probe_idx, np.int32).copy()
    pi[pi < 0] = NA_INT
    Np = int(rp[-1]) if rp.size else 0
    o = np.ascontiguousarray(opts, np.int32)
    out = dict(pair=np.empty((Np, 11)), gene=np.empty((G, 15)), perm=np.empty((G, n), np.int32),
               dnum=np.empty((Np, 3), np.int32), counts=np.zeros(4, np.int32), is_na=np.zeros((Np, 11), np.int32))
    rep = Report()
    p = lambda a: a.ctypes.data
    lib().rshim_analyze(session, p(x), n, G, p(rp), rp.size, p(pi), pi.size, p(o), o.size if n_opts is None else n_opts,
                        interrupt_at, p(out["pair"]), p(out["gene"]), p(out["perm"]), p(out["dnum"]), p(out["counts"]),
                        p(out["is_na"]), C.byref(rep))
    return out, rep


def groups(session, group_ptr, group_pairs):
    gp = np.ascontiguousarray(group_ptr, np.float64)
    gl = np.ascontiguousarray(group_pairs, np.int32)
    out = np.empty((gp.size - 1, 11))
    rep = Report()
    lib().rshim_groups(session, gp.ctypes.data, gp.size, gl.ctypes.data, gl.size, out.ctypes.data, C.byref(rep))
    return out, rep


def release(session):
    rep = Report()
    lib().rshim_release(session, C.byref(rep))
    return rep


def finish():
    rep = Report()
    lib().rshim_finish(C.byref(rep))
    return rep
