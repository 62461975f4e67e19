"""ORACLE (test infrastructure): ctypes binding of oracle/libfsd_oracle.so.

CPU restatement (plain C, fp64) of the reference's hot path; see oracle/fsd_oracle.h.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfsd_oracle.so")

MAX_SORTED, MAX_WV, HORIZON = 12, 32, 40

STATUS_BITS = {
    "NO_LEFT": 1 << 0, "NO_RIGHT": 1 << 1, "FEW_CONES": 1 << 2, "FEW_MATCHES": 1 << 3,
    "FIT1_FAILED": 1 << 4, "PATH_TOO_FAR": 1 << 5, "MPC_FAILED": 1 << 6, "TIE_P": 1 << 7,
    "OVERFLOW": 1 << 8, "REF_RAISES": 1 << 9, "UNSUPPORTED": 1 << 10,
}


class Result(C.Structure):
    _fields_ = [
        ("n_left", C.c_int), ("n_right", C.c_int),
        ("left_idx", C.c_int * MAX_SORTED), ("right_idx", C.c_int * MAX_SORTED),
        ("n_left_wv", C.c_int), ("n_right_wv", C.c_int),
        ("left_wv", (C.c_double * 2) * MAX_WV), ("right_wv", (C.c_double * 2) * MAX_WV),
        ("l2r", C.c_int * MAX_WV), ("r2l", C.c_int * MAX_WV),
        ("path", (C.c_double * 4) * HORIZON),
        ("P", C.c_int), ("n_trim", C.c_int), ("status", C.c_uint),
        ("first_k", (C.c_int * 2) * 2), ("n_configs", C.c_int * 2), ("n_pops", C.c_int * 2),
    ]


RESULT_DTYPE = np.dtype(Result)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) if the shared object is missing or stale."""
    srcs = [os.path.join(_HERE, f) for f in ("fitpack.c", "sort.c", "match.c", "path.c", "api.c",
                                             "fsd_oracle.h", "oracle_internal.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libfsd_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
        L.fsd_oracle_plan_frame.argtypes = [dp, up, C.c_int, dp, dp, C.c_int, dp, C.POINTER(Result)]
        L.fsd_oracle_plan_batch.argtypes = [dp, up, ip, C.c_int, dp, dp, C.POINTER(C.c_short), C.c_int, C.c_void_p]
        L.fsd_oracle_initial_path.argtypes = [dp]
        L.fsd_oracle_parcur.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double, dp, ip, dp, dp]
        L.fsd_oracle_parcur.restype = C.c_int
        L.fsd_oracle_splev.argtypes = [dp, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.fsd_oracle_sort.argtypes = [dp, up, C.c_int, dp, dp, C.POINTER(Result)]
        L.fsd_oracle_match.argtypes = [dp, C.c_int, dp, C.c_int, dp, dp, C.POINTER(Result)]
        L.fsd_oracle_path.argtypes = [dp, C.c_int, dp, C.c_int, ip, ip, dp, dp, C.c_int, dp, C.POINTER(Result)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def plan_batch(batch, force_P=None, threads: int = 1) -> np.ndarray:
    """Plan every frame of a FrameBatch; returns a structured array (RESULT_DTYPE) of length B."""
    xy = np.ascontiguousarray(batch.cones_xy, dtype=np.float64)
    ty = np.ascontiguousarray(batch.cones_type, dtype=np.uint8)
    off = np.ascontiguousarray(batch.offsets, dtype=np.int32)
    pos = np.ascontiguousarray(batch.pos, dtype=np.float64)
    dr = np.ascontiguousarray(batch.dir, dtype=np.float64)
    B = len(off) - 1
    res = np.zeros(B, dtype=RESULT_DTYPE)
    fp = None
    if force_P is not None:
        fp_arr = np.ascontiguousarray(force_P, dtype=np.int16)
        fp = fp_arr.ctypes.data_as(C.POINTER(C.c_short))
    lib().fsd_oracle_plan_batch(_dp(xy), ty.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                off.ctypes.data_as(C.POINTER(C.c_int)), B, _dp(pos), _dp(dr), fp,
                                int(threads), res.ctypes.data)
    return res


def initial_path() -> np.ndarray:
    out = np.zeros((HORIZON, 4))
    lib().fsd_oracle_initial_path(_dp(out))
    return out


def splprep(points: np.ndarray, s: float, k: int | None = None):
    """Oracle counterpart of scipy.interpolate.splprep(points.T, s=s, k=k, u=chord length).
    Returns (t, cx, cy, k, fp, ier, u)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    m = len(pts)
    if k is None:
        k = int(np.clip(m - 1, 1, 3))
    u = np.concatenate(([0.0], np.cumsum(np.linalg.norm(np.diff(pts, axis=0), axis=1))))
    u = np.ascontiguousarray(u)
    nest = m + 2 * k
    t = np.zeros(nest)
    c = np.zeros(2 * nest)
    n = C.c_int(0)
    fp = C.c_double(0)
    ier = lib().fsd_oracle_parcur(_dp(pts), _dp(u), m, k, float(s), _dp(t), C.byref(n), _dp(c), C.byref(fp))
    nn = n.value
    return t[:nn].copy(), c[: nn - k - 1].copy(), c[nest : nest + nn - k - 1].copy(), k, fp.value, ier, u


def splev(x: np.ndarray, t: np.ndarray, c: np.ndarray, k: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    cc = np.zeros(len(t))
    cc[: len(c)] = c
    y = np.zeros(len(x))
    lib().fsd_oracle_splev(_dp(t), len(t), _dp(cc), k, _dp(x), len(x), _dp(y))
    return y
