"""ctypes binding of the HOST-CHECK build of the kernels' per-frame code (tests only).

`libfsdplan_hostcheck.so` is ft_fsd_path_planning_b200/csrc/hostcheck.cpp: the same sort / match /
spline / path sources the CUDA kernels are compiled from, built by g++ with a warp of one lane.
It exists so that the CPU-only test tier can check the kernels' logic; the product never loads it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CSRC = os.path.join(_ROOT, "ft_fsd_path_planning_b200", "csrc")
_LIB = os.path.join(_CSRC, "libfsdplan_hostcheck.so")
MAX_SORTED, MAX_WV, HORIZON = 12, 32, 40


class Params(C.Structure):
    _fields_ = [
        ("max_n_neighbors", C.c_int32), ("max_length", C.c_int32), ("max_dist", C.c_double),
        ("max_dist_to_first", C.c_double), ("threshold_directional_angle", C.c_double),
        ("threshold_absolute_angle", C.c_double), ("car_size", C.c_double), ("max_dfs_pops", C.c_int32),
        ("reserved0", C.c_int32), ("min_track_width", C.c_double), ("max_search_range", C.c_double),
        ("max_search_angle", C.c_double), ("smoothing", C.c_double), ("predict_every", C.c_double),
        ("maximal_distance_for_valid_path", C.c_double), ("mpc_path_length", C.c_double),
        ("refit_smoothing", C.c_double),
    ]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".cpp"))]
    srcs.append(os.path.join(_ROOT, "include", "fsdplan.h"))
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if stale:
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=off",
             "-o", _LIB, os.path.join(_CSRC, "hostcheck.cpp")])
    return _LIB


_LIB_BIG = os.path.join(_CSRC, "libfsdplan_hostcheck_big.so")
_lib_big = None


def lib_big():
    """The same sources with the large static bounds of csrc/kernels_big.cu (2 048 path points, 64 knots per fit)."""
    global _lib_big
    if _lib_big is None:
        srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".cpp"))]
        if not os.path.exists(_LIB_BIG) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_BIG) for s in srcs):
            subprocess.check_call(
                ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=off",
                 "-DFSD_PCAP=2048", "-DFSD_NCAP=64", "-o", _LIB_BIG, os.path.join(_CSRC, "hostcheck.cpp")])
        _lib_big = C.CDLL(_LIB_BIG)
    return _lib_big


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def set_caps(cap: int = 0, pcap: int = 0, resume: bool = False, start_cap: int = 0) -> None:
    """Knot records / path points of the working memory of later calls (0, 0: the compiled NCAP / PCAP).  With cap > NCAP
    the records continue behind the PathSmem image -- the memory layout of path_kernel's in-kernel second chance.
    resume=True holds the extra records in reserve: a fit starts with NCAP records, is SUSPENDED when it outgrows them and
    resumed with all `cap` (what path_kernel does); resume=False gives every fit all `cap` records from the start.
    start_cap (8 .. NCAP): the records a fit starts with when resume=True -- small values make ordinary frames suspend."""
    lib().fsd_hostcheck_set_caps(int(cap), int(pcap), int(bool(resume)), int(start_cap))


def default_params() -> Params:
    p = Params()
    lib().fsd_hostcheck_params_default(C.byref(p))
    return p


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def initial_path() -> np.ndarray:
    out = np.zeros((HORIZON, 4))
    p = default_params()
    lib().fsd_hostcheck_initial_path(C.byref(p), _p(out, C.c_double))
    return out


def plan_batch(batch, force_P=None, prev=None):
    xy = np.ascontiguousarray(batch.cones_xy, dtype=np.float64)
    ty = np.ascontiguousarray(batch.cones_type, dtype=np.uint8)
    off = np.ascontiguousarray(batch.offsets, dtype=np.int32)
    pos = np.ascontiguousarray(batch.pos, dtype=np.float64)
    dr = np.ascontiguousarray(batch.dir, dtype=np.float64)
    B = len(off) - 1
    out = {
        "path": np.zeros((B, HORIZON, 4)),
        "left_idx": np.zeros((B, MAX_SORTED), np.int16), "right_idx": np.zeros((B, MAX_SORTED), np.int16),
        "n_wv": np.zeros((B, 2), np.int16),
        "left_wv": np.zeros((B, MAX_WV, 2)), "right_wv": np.zeros((B, MAX_WV, 2)),
        "l2r": np.zeros((B, MAX_WV), np.int16), "r2l": np.zeros((B, MAX_WV), np.int16),
        "grid": np.zeros((B, 2), np.int16), "sort_dbg": np.zeros((B, 8), np.int16),
        "status": np.zeros(B, np.uint32),
    }
    fp = None
    if force_P is not None:
        fpa = np.ascontiguousarray(force_P, dtype=np.int16)
        fp = _p(fpa, C.c_int16)
    pv = None
    if prev is not None:
        pva = np.ascontiguousarray(prev, dtype=np.float64)
        pv = _p(pva, C.c_double)
    p = default_params()
    i16 = C.c_int16
    lib().fsd_hostcheck_plan(
        C.byref(p), B, _p(xy, C.c_double), _p(ty, C.c_uint8), _p(off, C.c_int32), _p(pos, C.c_double),
        _p(dr, C.c_double), fp, pv, _p(out["path"], C.c_double), _p(out["left_idx"], i16),
        _p(out["right_idx"], i16), _p(out["n_wv"], i16), _p(out["left_wv"], C.c_double),
        _p(out["right_wv"], C.c_double), _p(out["l2r"], i16), _p(out["r2l"], i16), _p(out["grid"], i16),
        _p(out["sort_dbg"], i16), _p(out["status"], C.c_uint32))
    return out


def global_path(gpath, pos, direction, force_P=None, prev=None):
    """Poses [n, 2] along a global path [M, 2] through the kernels' path code; returns dict(path, grid, status)."""
    gp = np.ascontiguousarray(gpath, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
    dr = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 2)
    n = len(pos)
    out = np.zeros((n, HORIZON, 4))
    grid = np.zeros((n, 2), np.int16)
    status = np.zeros(n, np.uint32)
    fp = None
    if force_P is not None:
        fpa = np.ascontiguousarray(force_P, dtype=np.int16)
        fp = _p(fpa, C.c_int16)
    pv, stride = None, 0
    if prev is not None:
        pva = np.ascontiguousarray(prev, dtype=np.float64)
        pv, stride = _p(pva, C.c_double), (0 if pva.size == HORIZON * 4 else HORIZON * 4)
    p = default_params()
    lib_big().fsd_hostcheck_global_path(C.byref(p), n, _p(pos, C.c_double), _p(dr, C.c_double), _p(gp, C.c_double), len(gp),
                                    fp, pv, stride, _p(out, C.c_double), _p(grid, C.c_int16), _p(status, C.c_uint32))
    return {"path": out, "grid": grid, "status": status}


def knn(batch):
    """(nbr [total, 2, 5] uint8, deg [total, 2] uint8) of the kernels' k-NN graph code (layout of fsd_knn_batch)."""
    xy = np.ascontiguousarray(batch.cones_xy, dtype=np.float64)
    ty = np.ascontiguousarray(batch.cones_type, dtype=np.uint8)
    off = np.ascontiguousarray(batch.offsets, dtype=np.int32)
    nbr = np.zeros((len(xy), 2, 5), np.uint8)
    deg = np.zeros((len(xy), 2), np.uint8)
    p = default_params()
    lib().fsd_hostcheck_knn(C.byref(p), len(off) - 1, _p(xy, C.c_double), _p(ty, C.c_uint8), _p(off, C.c_int32),
                            _p(nbr, C.c_uint8), _p(deg, C.c_uint8))
    return nbr, deg


def fit(points: np.ndarray, s: float):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    m = len(pts)
    t = np.zeros(64)
    c = np.zeros(128)
    n = C.c_int(0)
    k = C.c_int(0)
    ier = lib().fsd_hostcheck_fit(_p(pts, C.c_double), m, C.c_double(float(s)), _p(t, C.c_double), C.byref(n),
                                  _p(c, C.c_double), C.byref(k))
    nn, kk = n.value, k.value
    cc = c[: 2 * (nn - kk - 1)].reshape(-1, 2)
    return t[:nn].copy(), cc[:, 0].copy(), cc[:, 1].copy(), kk, ier


# ---- skidpad ------------------------------------------------------------------------------------------------------

def skidpad_relocalize(cones_xy, pos, orig_pos, orig_dir, jitter, ref):
    xy = np.ascontiguousarray(cones_xy, dtype=np.float64)
    out = np.zeros(8)
    nacc = C.c_int(0)
    d = lambda a: _p(np.ascontiguousarray(a, dtype=np.float64), C.c_double)
    lib().fsd_hostcheck_skidpad_relocalize(_p(xy, C.c_double), len(xy), d(pos), d(orig_pos), d(orig_dir), d(jitter),
                                           d(ref), _p(out, C.c_double), C.byref(nacc))
    return out, nacc.value


def skidpad_steps(reloc8, table, pos, direction, state, force_P=None, prev=None):
    """Steps of ONE trajectory; prev: [40,4] shared or [n,40,4]; returns dict."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    direction = np.ascontiguousarray(direction, dtype=np.float64)
    table = np.ascontiguousarray(table, dtype=np.float64)
    n = len(pos)
    out = np.zeros((n, HORIZON, 4))
    internal = np.zeros((n, HORIZON, 4))
    index = np.zeros(n, np.int32)
    status = np.zeros(n, np.uint32)
    grid = np.zeros((n, 2), np.int16)
    st = C.c_int(int(state))
    prev = np.ascontiguousarray(prev if prev is not None else initial_path(), dtype=np.float64)
    stride = 0 if prev.size == HORIZON * 4 else HORIZON * 4
    fp = None
    if force_P is not None:
        fpa = np.ascontiguousarray(force_P, dtype=np.int16)
        fp = _p(fpa, C.c_int16)
    p = default_params()
    r8 = np.ascontiguousarray(reloc8, dtype=np.float64)
    lib().fsd_hostcheck_skidpad_steps(C.byref(p), _p(r8, C.c_double), _p(table, C.c_double), len(table),
                                      _p(pos, C.c_double), _p(direction, C.c_double), n, C.byref(st), fp,
                                      _p(prev, C.c_double), stride, _p(out, C.c_double), _p(internal, C.c_double),
                                      _p(index, C.c_int32), _p(status, C.c_uint32), _p(grid, C.c_int16))
    return {"path": out, "internal": internal, "index": index, "status": status, "grid": grid, "state": st.value}
