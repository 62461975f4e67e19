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


def skidpad_constants():
    """(tracked path = table[::2], reference centres [right, left], jitter)"""
    table = np.load(_DATA)
    # circle centres of the two loops of the canonical path (skidpad_relocalizer.py:172-183): constants of the track
    # definition, extracted from the reference by tools/extract_skidpad_path.py like the path table itself
    ref = np.load(os.path.join(os.path.dirname(_DATA), "skidpad_ref_centers.npy"))
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
            raise RuntimeError("SkidpadBatchPlanner needs a CUDA device and never falls back to the host")
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


def plan_skidpad_sharded(planner: SkidpadBatchPlanner, step_offsets: np.ndarray, pos: torch.Tensor,
                         direction: torch.Tensor, reloc: torch.Tensor, index_state: torch.Tensor, *, group=None,
                         plan_fn=None):
    """Skidpad on several GPUs: the steps of one trajectory are sequential (`index_along_path`, the
    previous-path fallback), trajectories are independent -> every rank plans WHOLE trajectories
    (`distributed.shard_trajectories`, balanced by step count) and one all-gather per output tensor collects the
    results in trajectory order.  step_offsets: host int array [T + 1]; pos / direction [S, 2], reloc [T, 8],
    index_state [T] are the FULL tensors on every rank (each rank reads its slice; index_state is updated for the
    rank's own trajectories and then gathered).  `plan_fn(step_offsets, pos, dir, reloc, index_state) -> dict` defaults to
    `planner.plan`; tests pass a CPU stand-in to exercise the partition and the gathers on gloo.  Returns the gathered
    dict (path, path_f64, internal, index, grid, status) with S rows on every rank, and the gathered index_state."""
    import torch.distributed as dist

    from .distributed import all_gather_ragged, shard_trajectories

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    off = np.asarray(step_offsets, dtype=np.int64)
    cuts = [shard_trajectories(off, r, world) for r in range(world)]
    t_lo, t_hi = cuts[rank]
    s_lo, s_hi = int(off[t_lo]), int(off[t_hi])
    dev = pos.device
    local_off = torch.from_numpy((off[t_lo : t_hi + 1] - off[t_lo]).astype(np.int32)).to(dev)
    local_state = index_state[t_lo:t_hi].clone()
    fn = plan_fn if plan_fn is not None else planner.plan
    if t_hi > t_lo:
        out = fn(local_off, pos[s_lo:s_hi].contiguous(), direction[s_lo:s_hi].contiguous(),
                 reloc[t_lo:t_hi].contiguous(), local_state)
    else:
        out = None
    step_counts = [int(off[c[1]] - off[c[0]]) for c in cuts]
    traj_counts = [c[1] - c[0] for c in cuts]
    specs = {"path": ((HORIZON, 4), torch.float32), "path_f64": ((HORIZON, 4), torch.float64),
             "internal": ((HORIZON, 4), torch.float64), "index": ((), torch.int32), "grid": ((2,), torch.int16),
             "status": ((), torch.int32)}
    gathered = {}
    for k, (tail, dt) in specs.items():
        local = out[k] if out is not None else torch.zeros((0, *tail), dtype=dt, device=dev)
        gathered[k] = all_gather_ragged(local, step_counts, group)
    state = all_gather_ragged(local_state, traj_counts, group)
    index_state.copy_(state)
    return gathered, state
