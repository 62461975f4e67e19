"""ctypes binding of libfsdplan.so (include/fsdplan.h).  Loading fails loudly when the library has not been built, and
every CUDA entry point returns FSD_ERR_NO_DEVICE without a GPU -- nothing falls back to the host.  The one host entry
point, fsd_plan_batch_cpu (the kernels' own per-frame sources compiled for the host), is used only when the caller asks
for device="cpu" explicitly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("FSD_LIBFSDPLAN", os.path.join(CSRC, "libfsdplan.so"))  # override: kernel A/B experiments
INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "fsdplan.h")

MAX_CONES, MAX_SORTED, MAX_WV, HORIZON = 256, 12, 32, 40
MISSION_AUTOCROSS, MISSION_TRACKDRIVE = 3, 4

STATUS_BITS = {
    "NO_LEFT": 1 << 0, "NO_RIGHT": 1 << 1, "FEW_CONES": 1 << 2, "FEW_MATCHES": 1 << 3,
    "FIT1_FAILED": 1 << 4, "PATH_TOO_FAR": 1 << 5, "MPC_FAILED": 1 << 6, "TIE_P": 1 << 7,
    "OVERFLOW": 1 << 8, "REF_RAISES": 1 << 9, "UNSUPPORTED": 1 << 10,
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class Params(C.Structure):
    """struct fsd_params"""
    _fields_ = [
        ("max_n_neighbors", C.c_int32), ("max_length", C.c_int32), ("max_dist", C.c_double),
        ("max_dist_to_first", C.c_double), ("threshold_directional_angle", C.c_double),
        ("threshold_absolute_angle", C.c_double), ("car_size", C.c_double), ("max_dfs_pops", C.c_int32),
        ("reserved0", C.c_int32), ("min_track_width", C.c_double), ("max_search_range", C.c_double),
        ("max_search_angle", C.c_double), ("smoothing", C.c_double), ("predict_every", C.c_double),
        ("maximal_distance_for_valid_path", C.c_double), ("mpc_path_length", C.c_double),
        ("refit_smoothing", C.c_double),
    ]


class Intermediate(C.Structure):
    """struct fsd_intermediate (device pointers)"""
    _fields_ = [(n, C.c_void_p) for n in
                ("path_f64", "n_wv", "left_wv", "right_wv", "l2r", "r2l", "grid", "sort_dbg")]


MAX_PEERS = 16


class Gather(C.Structure):
    """struct fsd_gather: the all-gather of the output paths fused into the path kernel (peer-mapped device pointers)"""
    _fields_ = [("n_peers", C.c_int32), ("reserved0", C.c_int32), ("first_row", C.c_int64),
                ("peer_out_path", C.c_void_p * MAX_PEERS), ("multicast_out_path", C.c_void_p)]


def sources():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h")) or f == "cpu_backend.cpp"]
    return files + [INCLUDE]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/kernels.cu for sm_100a into csrc/libfsdplan.so (in-tree, so it travels with the repo)."""
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in sources())
    if stale:
        # kernels.cu: the CUDA kernels and the C-ABI; cpu_backend.cpp: the same per-frame sources compiled for the host
        # (fsd_plan_batch_cpu, an explicit entry point -- the CUDA entry points never fall back to it)
        cmd = ["nvcc", *NVCC_FLAGS, "-Xcompiler", "-Wno-unknown-pragmas", "-o", LIB_PATH, os.path.join(CSRC, "kernels.cu"),
               os.path.join(CSRC, "kernels_big.cu"), os.path.join(CSRC, "cpu_backend.cpp")]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()').  ft_fsd_path_planning_b200 has no CPU implementation.")
        L = C.CDLL(LIB_PATH)
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        L.fsd_abi_version.restype = C.c_int
        L.fsd_strerror.restype = C.c_char_p
        L.fsd_strerror.argtypes = [C.c_int]
        L.fsd_params_default.argtypes = [C.POINTER(Params)]
        L.fsd_workspace_bytes.restype = sz
        L.fsd_workspace_bytes.argtypes = [i32, i32]
        L.fsd_plan_launches.restype = i32
        L.fsd_plan_launches.argtypes = [i32]
        L.fsd_initial_path.argtypes = [C.POINTER(Params), vp, vp]
        plan_args = [C.POINTER(Params), i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(Intermediate), vp, vp, i32,
                     vp, vp, sz, vp]
        L.fsd_plan_batch.argtypes = plan_args
        L.fsd_plan_batch_f64.argtypes = plan_args
        L.fsd_plan_first_chunk.restype = i32
        L.fsd_plan_first_chunk.argtypes = [i32]
        L.fsd_plan_batch_ex.argtypes = [C.POINTER(Params), i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                        C.POINTER(Intermediate), vp, vp, i32, vp, vp, sz, vp, vp]
        L.fsd_plan_batch_gather.argtypes = [C.POINTER(Params), i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                            C.POINTER(Intermediate), vp, vp, i32, vp, vp, sz, vp, vp, C.POINTER(Gather)]
        L.fsd_plan_batch_cpu.argtypes = [C.POINTER(Params), i32, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                         C.POINTER(Intermediate), vp, vp, i32, vp, i32]
        L.fsd_initial_path_cpu.argtypes = [C.POINTER(Params), vp]
        L.fsd_knn_batch.argtypes = [C.POINTER(Params), i32, i32, vp, vp, vp, vp, vp, vp]
        L.fsd_sort_batch.argtypes = [C.POINTER(Params), i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.fsd_match_batch.argtypes = [C.POINTER(Params), i32, vp, vp, vp, vp, vp, vp, C.POINTER(Intermediate), vp, vp]
        L.fsd_sort_match_batch.argtypes = [C.POINTER(Params), i32, i32, vp, vp, vp, vp, vp, vp, vp,
                                           C.POINTER(Intermediate), vp, vp]
        L.fsd_path_batch.argtypes = [C.POINTER(Params), i32, i32, vp, vp, C.POINTER(Intermediate), vp, vp, i32, vp, vp,
                                     vp, sz, vp]
        L.fsd_path_batch_gather.argtypes = [C.POINTER(Params), i32, i32, vp, vp, C.POINTER(Intermediate), vp, vp, i32, vp,
                                            vp, vp, sz, vp, C.POINTER(Gather)]
        L.fsd_global_path_workspace_bytes.restype = sz
        L.fsd_global_path_workspace_bytes.argtypes = [i32]
        L.fsd_global_path_batch.argtypes = [C.POINTER(Params), i32, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, sz, vp]
        L.fsd_skidpad_workspace_bytes.restype = sz
        L.fsd_skidpad_workspace_bytes.argtypes = [i32]
        L.fsd_skidpad_relocalize_batch.argtypes = [C.POINTER(Params), i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.fsd_skidpad_plan_batch.argtypes = [C.POINTER(Params), i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, vp,
                                             vp, vp, vp, vp, vp, vp, sz, vp]
        if L.fsd_abi_version() != 1:
            raise RuntimeError("libfsdplan.so ABI version mismatch")
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise RuntimeError(f"libfsdplan: {lib().fsd_strerror(code).decode()} ({code})")


def default_params() -> Params:
    p = Params()
    check(lib().fsd_params_default(C.byref(p)))
    return p
