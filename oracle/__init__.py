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


class Reloc(C.Structure):
    _fields_ = [("relocalized", C.c_int), ("n_accepted", C.c_int), ("translation", C.c_double * 2),
                ("rotation", C.c_double), ("right_ref", C.c_double * 2), ("right_calc", C.c_double * 2)]


RESULT_DTYPE = np.dtype(Result)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) if the shared object is missing or stale."""
    srcs = [os.path.join(_HERE, f) for f in ("fitpack.c", "sort.c", "match.c", "path.c", "api.c", "skidpad.c",
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
        L.fsd_oracle_path_global.argtypes = [dp, C.c_int, dp, dp, C.c_int, dp, C.POINTER(Result)]
        L.fsd_oracle_adjacency.argtypes = [dp, up, C.c_int, C.c_int, ip, ip]
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


def adjacency(batch):
    """Both sides' k-NN graphs of every frame (create_adjacency_matrix): (nbr [total, 2, 5] int, deg [total, 2] int) in
    the layout of fsd_knn_batch (side 0 = left graph, 1 = right graph; frames with fewer than 3 cones: degree 0)."""
    xy = np.ascontiguousarray(batch.cones_xy, dtype=np.float64)
    ty = np.ascontiguousarray(batch.cones_type, dtype=np.uint8)
    off = np.asarray(batch.offsets)
    total = len(xy)
    nbr = np.zeros((total, 2, 5), dtype=np.int32)
    deg = np.zeros((total, 2), dtype=np.int32)
    L = lib()
    for b in range(len(off) - 1):
        lo, n = int(off[b]), int(off[b + 1] - off[b])
        if n < 3:
            continue
        for s, side in enumerate((2, 1)):
            nb = np.zeros((n, 5), dtype=np.int32)
            dg = np.zeros(n, dtype=np.int32)
            L.fsd_oracle_adjacency(_dp(xy[lo:lo + n]), ty[lo:lo + n].ctypes.data_as(C.POINTER(C.c_ubyte)), n, side,
                                   nb.ctypes.data_as(C.POINTER(C.c_int)), dg.ctypes.data_as(C.POINTER(C.c_int)))
            nbr[lo:lo + n, s] = nb
            deg[lo:lo + n, s] = dg
    return nbr, deg


def path_global(global_path, pos, direction, force_P=0, prev_path=None) -> Result:
    """run_path_calculation with a global path for one pose; returns the Result struct (path, P, n_trim, status)."""
    gp = np.ascontiguousarray(global_path, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    dr = np.ascontiguousarray(direction, dtype=np.float64)
    pv = None if prev_path is None else np.ascontiguousarray(prev_path, dtype=np.float64)
    res = Result()
    lib().fsd_oracle_path_global(_dp(gp), len(gp), _dp(pos), _dp(dr), int(force_P), None if pv is None else _dp(pv),
                                 C.byref(res))
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


# ---- skidpad mission ---------------------------------------------------------------------------------------------

def skidpad_constants(path_table: np.ndarray):
    """(global path = table[::2], reference centres [right, left], jitter) -- host-side constants.
    Reference centres: hyper circle fit of the table's points with y < -2 / y > 2
    (skidpad_relocalizer.py:172-183); jitter: numpy RandomState(42).randn (skidpad_relocalizer.py:38, 53)."""
    def fit(pts):
        x, y = pts[:, 0], pts[:, 1]
        xi, yi = x - x.mean(), y - y.mean()
        zi = xi * xi + yi * yi
        n = len(x)
        mxy, mxx, myy = (xi * yi).sum() / n, (xi * xi).sum() / n, (yi * yi).sum() / n
        mxz, myz, mzz = (xi * zi).sum() / n, (yi * zi).sum() / n, (zi * zi).sum() / n
        mz = mxx + myy
        cov = mxx * myy - mxy * mxy
        var = mzz - mz * mz
        a2 = 4 * cov - 3 * mz * mz - mzz
        a1 = var * mz + 4.0 * cov * mz - mxz * mxz - myz * myz
        a0 = mxz * (mxz * myy - myz * mxy) + myz * (myz * mxx - mxz * mxy) - var * cov
        a22 = a2 + a2
        yv, xv = a0, 0.0
        for _ in range(99):
            dy = a1 + xv * (a22 + 16.0 * xv * xv)
            xn = xv - yv / dy
            if xn == xv or not np.isfinite(xn):
                break
            yn = a0 + xn * (a1 + xn * (a2 + 4.0 * xn * xn))
            if abs(yn) >= abs(yv):
                break
            xv, yv = xn, yn
        det = xv * xv - xv * mz + cov
        return np.array([(mxz * (myy - xv) - myz * mxy) / det / 2.0 + x.mean(),
                         (myz * (mxx - xv) - mxz * mxy) / det / 2.0 + y.mean()])

    ref = np.stack([fit(path_table[path_table[:, 1] < -2]), fit(path_table[path_table[:, 1] > 2])])
    jitter = np.random.RandomState(42).randn(1140 * 6)
    return np.ascontiguousarray(path_table[::2]), ref, jitter


class SkidpadOracle:
    """Sequential skidpad planner (one trajectory), mirroring PathPlanner(MissionTypes.skidpad)."""

    def __init__(self, path_table: np.ndarray):
        self.path, self.ref, self.jitter = skidpad_constants(np.asarray(path_table, dtype=np.float64))
        self.reloc = Reloc()
        self.orig = None
        self.index = C.c_int(0)
        self.prev = initial_path()
        L = lib()
        dp = C.POINTER(C.c_double)
        L.fsd_oracle_skidpad_relocalize.argtypes = [dp, C.c_int, dp, dp, dp, dp, dp, C.POINTER(Reloc)]
        L.fsd_oracle_skidpad_step.argtypes = [dp, C.c_int, C.POINTER(C.c_int), C.POINTER(Reloc), dp, dp, C.c_int, dp,
                                              dp, C.POINTER(Result)]

    def step(self, cones_by_type, pos, direction, force_P=0):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        direction = np.ascontiguousarray(direction, dtype=np.float64)
        if not self.reloc.relocalized:
            if self.orig is None:
                self.orig = (pos.copy(), direction.copy())
            xy = np.ascontiguousarray(np.concatenate([np.asarray(c, float).reshape(-1, 2) for c in cones_by_type]))
            lib().fsd_oracle_skidpad_relocalize(_dp(xy), len(xy), _dp(pos), _dp(self.orig[0]), _dp(self.orig[1]),
                                                _dp(self.jitter), _dp(np.ascontiguousarray(self.ref)),
                                                C.byref(self.reloc))
        res = Result()
        internal = np.zeros((HORIZON, 4))
        lib().fsd_oracle_skidpad_step(_dp(self.path), len(self.path), C.byref(self.index), C.byref(self.reloc),
                                      _dp(pos), _dp(direction), int(force_P), _dp(self.prev), _dp(internal),
                                      C.byref(res))
        self.prev = internal
        return np.ctypeslib.as_array(res.path).copy(), res
