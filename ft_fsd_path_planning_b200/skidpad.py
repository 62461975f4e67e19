"""Host side of the skidpad mission: constants (canonical path, reference circle centres, jitter sequence) and the
batched relocalization / tracking calls into libfsdplan.so.

Reference: fsd_path_planning/relocalization/skidpad/skidpad_relocalizer.py (constants :172-196, :38, :242),
fsd_path_planning/calculate_path/skidpad_calculate_path.py, fsd_path_planning/full_pipeline/full_pipeline.py:122-194.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _lib

HORIZON = _lib.HORIZON
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "skidpad_path.npy")
N_TRIPLES = 1140  # C(20, 3)


def _hyper_circle_centre(pts: np.ndarray) -> np.ndarray:
    """Centre of the algebraic (hyper) circle fit, the reference's circle_fit (utils/math_utils.py:579-646)."""
    x, y = pts[:, 0], pts[:, 1]
    n = len(x)
    xi, yi = x - x.mean(), y - y.mean()
    zi = xi * xi + yi * yi
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


def skidpad_constants():
    """(tracked path = table[::2], reference centres [right, left], jitter)"""
    table = np.load(_DATA)
    ref = np.stack([_hyper_circle_centre(table[table[:, 1] < -2]), _hyper_circle_centre(table[table[:, 1] > 2])])
    jitter = np.random.RandomState(42).randn(N_TRIPLES * 6)  # the sequence skidpad_relocalizer.py:38, 53 consumes
    return np.ascontiguousarray(table[::2]), ref, jitter


def to_known_frame(reloc8: np.ndarray, p: np.ndarray) -> np.ndarray:
    """transform_pose of skidpad_relocalizer.py:133-147 (host, for relocalization_info)."""
    tx, ty, rot, rrx, rry = reloc8[:5]
    c, s = np.cos(rot), np.sin(rot)
    x, y = p[0] + tx - rrx, p[1] + ty - rry
    return np.array([x * c - y * s + rrx, x * s + y * c + rry])


class SkidpadBatchPlanner:
    """Batched skidpad planner: `relocalize` (K1) once per trajectory, `plan` (K2 + MPC tail) for many steps."""

    def __init__(self, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("ft_fsd_path_planning_b200 needs a CUDA device: the planner has no CPU implementation")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.lib()
        self.params = _lib.default_params()
        path, ref, jitter = skidpad_constants()
        dev = self.device
        self.path = torch.from_numpy(path).to(dev)
        self.ref = torch.from_numpy(np.ascontiguousarray(ref.reshape(-1))).to(dev)
        self.jitter = torch.from_numpy(jitter).to(dev)
        self.ref_host = ref

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def relocalize(self, cones_xy: torch.Tensor, offsets: torch.Tensor, pos: torch.Tensor, orig_pos: torch.Tensor,
                   orig_dir: torch.Tensor):
        """fp64 device tensors; returns (reloc [T, 8] float64, n_accepted [T] int32)."""
        T = offsets.numel() - 1
        reloc = torch.zeros((T, 8), dtype=torch.float64, device=self.device)
        nacc = torch.zeros((T,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fsd_skidpad_relocalize_batch(
                C.byref(self.params), T, cones_xy.data_ptr(), offsets.data_ptr(), pos.data_ptr(), orig_pos.data_ptr(),
                orig_dir.data_ptr(), self.jitter.data_ptr(), self.ref.data_ptr(), reloc.data_ptr(), nacc.data_ptr(),
                self._stream()))
        return reloc, nacc

    def plan(self, step_offsets: torch.Tensor, pos: torch.Tensor, direction: torch.Tensor, reloc: torch.Tensor,
             index_state: torch.Tensor, *, force_P: Optional[torch.Tensor] = None,
             prev_path: Optional[torch.Tensor] = None):
        """step_offsets [T+1] int32, pos/direction [S, 2] float64, reloc [T, 8], index_state [T] int32 (updated in
        place).  Returns a dict of device tensors."""
        T = step_offsets.numel() - 1
        S = pos.shape[0]
        dev = self.device
        out = {
            "path": torch.empty((S, HORIZON, 4), dtype=torch.float32, device=dev),
            "path_f64": torch.empty((S, HORIZON, 4), dtype=torch.float64, device=dev),
            "internal": torch.empty((S, HORIZON, 4), dtype=torch.float64, device=dev),
            "index": torch.empty((S,), dtype=torch.int32, device=dev),
            "grid": torch.empty((S, 2), dtype=torch.int16, device=dev),
            "status": torch.empty((S,), dtype=torch.int32, device=dev),
        }
        ws = torch.empty((int(self.lib.fsd_skidpad_workspace_bytes(S)),), dtype=torch.uint8, device=dev)
        stride = 0
        if prev_path is not None:
            stride = 0 if prev_path.numel() == HORIZON * 4 else HORIZON * 4
        with torch.cuda.device(dev):
            _lib.check(self.lib.fsd_skidpad_plan_batch(
                C.byref(self.params), T, S, step_offsets.data_ptr(), pos.data_ptr(), direction.data_ptr(),
                reloc.data_ptr(), index_state.data_ptr(), self.path.data_ptr(), self.path.shape[0],
                None if force_P is None else force_P.data_ptr(), None if prev_path is None else prev_path.data_ptr(),
                stride, out["path"].data_ptr(), out["path_f64"].data_ptr(), out["internal"].data_ptr(),
                out["index"].data_ptr(), out["grid"].data_ptr(), out["status"].data_ptr(), ws.data_ptr(), ws.numel(),
                self._stream()))
        out["_workspace"] = ws  # keep alive until the stream has consumed it
        return out
