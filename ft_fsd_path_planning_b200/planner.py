"""Host side of the batched planner: torch tensors as device buffers, libfsdplan.so through ctypes.

`BatchPlanner` is the batched entry point (thousands of independent frames per call, "fresh planner per
frame" semantics).  `PathPlanner` mirrors the reference's facade
(fsd_path_planning/full_pipeline/full_pipeline.py:53-217): same constructor, same
`calculate_path_in_global_frame` signature / return values / ValueError, same statefulness
(the previous path feeds the fallbacks of the next call, core_calculate_path.py:103-110, 561-573).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .enums import ConeTypes, MissionTypes
from .synth import FrameBatch, pack_frames

HORIZON, MAX_SORTED, MAX_WV = _lib.HORIZON, _lib.MAX_SORTED, _lib.MAX_WV


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_offsets(offsets: np.ndarray, total: int) -> None:
    """CSR offsets seen on the host: start at >= 0, non-decreasing, end within the cone arrays."""
    if len(offsets) and (offsets[0] < 0 or offsets[-1] > total or (np.diff(offsets) < 0).any()):
        raise ValueError("offsets must be non-decreasing with 0 <= offsets[0] and offsets[-1] <= number of cones")


@dataclass
class PlanResult:
    """Device tensors produced by one batched call."""

    path: torch.Tensor  # [B, 40, 4] float32: u, x, y, curvature
    left_idx: torch.Tensor  # [B, 12] int16, -1 padded (indices into the frame's cone list)
    right_idx: torch.Tensor  # [B, 12] int16
    status: torch.Tensor  # [B] int32 bit field (see _lib.STATUS_BITS)
    path_f64: Optional[torch.Tensor] = None  # [B, 40, 4] float64
    n_wv: Optional[torch.Tensor] = None  # [B, 2] int16
    left_wv: Optional[torch.Tensor] = None  # [B, 32, 2] float64
    right_wv: Optional[torch.Tensor] = None
    l2r: Optional[torch.Tensor] = None  # [B, 32] int16
    r2l: Optional[torch.Tensor] = None
    grid: Optional[torch.Tensor] = None  # [B, 2] int16: P, points entering the last re-fit
    sort_dbg: Optional[torch.Tensor] = None  # [B, 8] int16


class BatchPlanner:
    """Plans batches of independent frames on one GPU.

    Every buffer is a torch tensor; the C library only sees raw pointers and the stream.  Outputs are fresh
    tensors per call (or the caller's `out=` result, re-used); only the library's scratch workspace is cached
    (one batch size resident) and calls on different streams are ordered on it with an event.
    """

    def __init__(self, device: Union[str, torch.device, int] = "cuda", mission: int = _lib.MISSION_TRACKDRIVE):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchPlanner needs a CUDA device and never falls back to the host (the host build of the "
                               "kernel sources is a separate, explicit choice: CpuBatchPlanner / PathPlanner(device='cpu'))")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError("BatchPlanner runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.lib()
        self.params = _lib.default_params()
        self.mission = int(mission)
        self._ws, self._ws_key = None, None
        self._last_stream, self._last_event = None, None
        self._prev0: Optional[torch.Tensor] = None
        self._pinned_key = None
        self._pinned: dict = {}
        self.kernel_events: list = []

    def _default_prev(self) -> torch.Tensor:
        if self._prev0 is None:
            self._prev0 = self.initial_path()
        return self._prev0

    def first_chunk(self, B: int) -> int:
        """Leading frames of a B-frame call whose outputs are final when `chunk_ready` fires (B: the call is not split)."""
        with torch.cuda.device(self.device):
            return int(self.lib.fsd_plan_first_chunk(B))

    def kernel_times_ms(self, clear: bool = True):
        """[(sort_match_ms, path_ms), ...] for the calls made with kernel_events=True (synchronize first)."""
        out = [(e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])) for e in self.kernel_events]
        if clear:
            self.kernel_events = []
        return out

    # -- buffers -------------------------------------------------------------------------------------------
    def _workspace(self, B: int) -> torch.Tensor:
        """Scratch of the C library for a batch of B frames: cached (one batch size resident), never handed out."""
        if self._ws_key != B:
            with torch.cuda.device(self.device):  # fsd_workspace_bytes sizes the scratch for the CURRENT device
                nbytes = int(self.lib.fsd_workspace_bytes(B, 0))
            self._ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
            self._ws_key = B
        return self._ws

    def _outputs(self, B: int, intermediates: bool) -> dict:
        """Fresh output tensors for one call: results of earlier calls are never overwritten."""
        dev = self.device
        bufs = {
            "path": torch.empty((B, HORIZON, 4), dtype=torch.float32, device=dev),
            "left_idx": torch.empty((B, MAX_SORTED), dtype=torch.int16, device=dev),
            "right_idx": torch.empty((B, MAX_SORTED), dtype=torch.int16, device=dev),
            "status": torch.empty((B,), dtype=torch.int32, device=dev),
        }
        if intermediates:
            bufs.update({
                "path_f64": torch.empty((B, HORIZON, 4), dtype=torch.float64, device=dev),
                "n_wv": torch.empty((B, 2), dtype=torch.int16, device=dev),
                "left_wv": torch.empty((B, MAX_WV, 2), dtype=torch.float64, device=dev),
                "right_wv": torch.empty((B, MAX_WV, 2), dtype=torch.float64, device=dev),
                "l2r": torch.empty((B, MAX_WV), dtype=torch.int16, device=dev),
                "r2l": torch.empty((B, MAX_WV), dtype=torch.int16, device=dev),
                "grid": torch.empty((B, 2), dtype=torch.int16, device=dev),
                "sort_dbg": torch.empty((B, 8), dtype=torch.int16, device=dev),
            })
        return bufs

    def _order_after_previous_call(self, stream: torch.cuda.Stream) -> None:
        """The workspace is shared by all calls of this planner: a call on another stream than the previous one waits
        for the previous call (calls on one stream are ordered anyway)."""
        if self._last_stream is not None and self._last_stream != stream and self._last_event is not None:
            stream.wait_event(self._last_event)

    def _mark_call(self, stream: torch.cuda.Stream) -> None:
        if self._last_stream is not None and self._last_stream != stream or self._last_event is None:
            self._last_event = torch.cuda.Event()
        self._last_stream = stream
        self._last_event.record(stream)

    # -- the batched call ------------------------------------------------------------------------------------
    def plan(self, cones_xy: torch.Tensor, cones_type: torch.Tensor, offsets: torch.Tensor, pos: torch.Tensor,
             direction: torch.Tensor, *, force_P: Optional[torch.Tensor] = None,
             prev_path: Optional[torch.Tensor] = None, intermediates: bool = False,
             kernel_events: bool = False, out: Optional[PlanResult] = None,
             chunk_ready: Optional[torch.cuda.Event] = None, gather: Optional["_lib.Gather"] = None) -> PlanResult:
        """cones_xy [total, 2] float32|float64, cones_type [total] uint8, offsets [B+1] int32, pos/direction [B, 2]
        (same dtype as cones_xy); all on this planner's device.  Asynchronous on the current stream.  The CSR
        `offsets` must be non-decreasing with offsets[-1] <= total (checked for free by plan_host / plan_pinned, which
        see them on the host; checking a device tensor here would cost a synchronisation per call).

        Returns fresh tensors, or writes into `out` (a PlanResult of an earlier call with the same B and
        `intermediates`) to avoid allocations in a loop.

        chunk_ready: a CUDA event recorded as soon as the outputs of the first `first_chunk(B)` frames are final
        (fsd_plan_batch_ex) -- the multi-GPU pipeline starts their all-gather on a side stream at that point.

        gather: a `_lib.Gather` descriptor (distributed.PeerGather.descriptor()) -- the path kernel then stores every
        frame's path into all peers' gathered buffers as well (fsd_plan_batch_gather: the all-gather fused into the
        kernel over NVLink peer memory / NVSwitch multicast); call PeerGather.finish() before reading the gathered buffer.

        kernel_events=True issues the two launches through the stage entry points (fsd_sort_match_batch,
        fsd_path_batch) with CUDA events around each; the events are kept in `self.kernel_events`
        (read them with `kernel_times_ms()` after a synchronize)."""
        B = offsets.numel() - 1
        f64 = cones_xy.dtype == torch.float64
        if B == 0:
            e = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=self.device)
            return PlanResult(e(0, HORIZON, 4), e(0, MAX_SORTED, dt=torch.int16), e(0, MAX_SORTED, dt=torch.int16),
                              e(0, dt=torch.int32))
        for t, dt in ((cones_xy, None), (cones_type, torch.uint8), (offsets, torch.int32), (pos, cones_xy.dtype),
                      (direction, cones_xy.dtype)):
            if t.device != self.device or not t.is_contiguous() or (dt is not None and t.dtype != dt):
                raise ValueError("inputs must be contiguous tensors of the documented dtype on the planner's device")
        if cones_xy.dtype not in (torch.float32, torch.float64):
            raise ValueError("cones_xy must be float32 or float64")
        if cones_xy.dim() != 2 or cones_xy.shape[1] != 2 or cones_type.numel() != cones_xy.shape[0]:
            raise ValueError("cones_xy must be [total, 2] and cones_type [total]")
        if pos.numel() != 2 * B or direction.numel() != 2 * B:
            raise ValueError("pos and direction must be [B, 2] with B = len(offsets) - 1")
        intermediates = intermediates or kernel_events
        if out is not None:
            if out.path.shape[0] != B or out.path.device != self.device or (intermediates and out.path_f64 is None):
                raise ValueError("out= must be the PlanResult of a call with the same batch size and intermediates")
            bufs = {k: getattr(out, k) for k in ("path", "left_idx", "right_idx", "status", "path_f64", "n_wv", "left_wv",
                                                 "right_wv", "l2r", "r2l", "grid", "sort_dbg") if getattr(out, k) is not None}
            intermediates = intermediates or out.path_f64 is not None
        else:
            bufs = self._outputs(B, intermediates)
        ws = self._workspace(B)
        inter = None
        if intermediates:
            inter = _lib.Intermediate(*[bufs[n].data_ptr() for n in
                                        ("path_f64", "n_wv", "left_wv", "right_wv", "l2r", "r2l", "grid", "sort_dbg")])
        stride = 0
        if prev_path is not None:
            if prev_path.dtype != torch.float64 or not prev_path.is_contiguous() or prev_path.device != self.device:
                raise ValueError("prev_path must be a contiguous float64 tensor on the planner's device")
            stride = 0 if prev_path.numel() == HORIZON * 4 else HORIZON * 4
            if stride and prev_path.numel() != B * HORIZON * 4:
                raise ValueError("prev_path must be [40, 4] or [B, 40, 4]")
        if force_P is not None and (force_P.dtype != torch.int16 or force_P.numel() != B or
                                    force_P.device != self.device or not force_P.is_contiguous()):
            raise ValueError("force_P must be a contiguous int16 [B] tensor on the planner's device")
        fn = self.lib.fsd_plan_batch_f64 if f64 else self.lib.fsd_plan_batch
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            self._order_after_previous_call(cur)
            stream = cur.cuda_stream
            if kernel_events:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
                rc = self.lib.fsd_sort_match_batch(
                    C.byref(self.params), B, int(f64), cones_xy.data_ptr(), cones_type.data_ptr(), offsets.data_ptr(),
                    pos.data_ptr(), direction.data_ptr(), bufs["left_idx"].data_ptr(), bufs["right_idx"].data_ptr(),
                    C.byref(inter), bufs["status"].data_ptr(), stream)
                _lib.check(rc)
                ev[1].record()
                if prev_path is None:
                    prev_path = self._default_prev()
                rc = self.lib.fsd_path_batch(
                    C.byref(self.params), B, int(f64), pos.data_ptr(), direction.data_ptr(), C.byref(inter),
                    _ptr(force_P), prev_path.data_ptr(), stride, bufs["path"].data_ptr(), bufs["status"].data_ptr(),
                    ws.data_ptr(), ws.numel(), stream)
                ev[2].record()
                self.kernel_events.append(ev)
            elif gather is not None:
                if chunk_ready is not None:
                    chunk_ready.record(cur)
                rc = self.lib.fsd_plan_batch_gather(
                    C.byref(self.params), self.mission, B, int(f64), cones_xy.data_ptr(), cones_type.data_ptr(),
                    offsets.data_ptr(), pos.data_ptr(), direction.data_ptr(), bufs["path"].data_ptr(),
                    bufs["left_idx"].data_ptr(), bufs["right_idx"].data_ptr(),
                    C.byref(inter) if inter is not None else None, _ptr(force_P), _ptr(prev_path), stride,
                    bufs["status"].data_ptr(), ws.data_ptr(), ws.numel(), stream,
                    chunk_ready.cuda_event if chunk_ready is not None else None, C.byref(gather))
            elif chunk_ready is not None:
                chunk_ready.record(cur)  # creates the lazily-initialised CUDA event; re-recorded by the library
                rc = self.lib.fsd_plan_batch_ex(
                    C.byref(self.params), self.mission, B, int(f64), cones_xy.data_ptr(), cones_type.data_ptr(),
                    offsets.data_ptr(), pos.data_ptr(), direction.data_ptr(), bufs["path"].data_ptr(),
                    bufs["left_idx"].data_ptr(), bufs["right_idx"].data_ptr(),
                    C.byref(inter) if inter is not None else None, _ptr(force_P), _ptr(prev_path), stride,
                    bufs["status"].data_ptr(), ws.data_ptr(), ws.numel(), stream, chunk_ready.cuda_event)
            else:
                rc = fn(C.byref(self.params), self.mission, B, cones_xy.data_ptr(), cones_type.data_ptr(),
                        offsets.data_ptr(), pos.data_ptr(), direction.data_ptr(), bufs["path"].data_ptr(),
                        bufs["left_idx"].data_ptr(), bufs["right_idx"].data_ptr(),
                        C.byref(inter) if inter is not None else None, _ptr(force_P), _ptr(prev_path), stride,
                        bufs["status"].data_ptr(), ws.data_ptr(), ws.numel(), stream)
            self._mark_call(cur)
        _lib.check(rc)
        if out is not None:
            return out
        res = PlanResult(bufs["path"], bufs["left_idx"], bufs["right_idx"], bufs["status"])
        if intermediates:
            for n in ("path_f64", "n_wv", "left_wv", "right_wv", "l2r", "r2l", "grid", "sort_dbg"):
                setattr(res, n, bufs[n])
        return res

    def knn(self, cones_xy: torch.Tensor, cones_type: torch.Tensor, offsets: torch.Tensor,
            out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """The cost-matrix step in isolation (fsd_knn_batch): both sides' k-NN graphs of every frame as adjacency lists.
        Returns (nbr [total, 2, 5] uint8, deg [total, 2] uint8), frame-local indices, ascending; asynchronous."""
        B = offsets.numel() - 1
        total = int(cones_xy.shape[0])
        if out is None:
            out = (torch.empty((total, 2, 5), dtype=torch.uint8, device=self.device),
                   torch.empty((total, 2), dtype=torch.uint8, device=self.device))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fsd_knn_batch(
                C.byref(self.params), B, int(cones_xy.dtype == torch.float64), cones_xy.data_ptr(), cones_type.data_ptr(),
                offsets.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def plan_global_path(self, global_path: torch.Tensor, pos: torch.Tensor, direction: torch.Tensor, *,
                         force_P: Optional[torch.Tensor] = None, prev_path: Optional[torch.Tensor] = None) -> dict:
        """Path calculation along a GLOBAL PATH (fsd_global_path_batch; core_calculate_path.py:516-528): `global_path`
        [M, 2] float64 shared by all poses, pos / direction [n, 2] float64, all on this planner's device.  Returns a dict of
        device tensors: path [n, 40, 4] float32, path_f64, grid [n, 2] int16, status [n] int32.  Asynchronous."""
        n = pos.shape[0]
        dev = self.device
        for t in (global_path, pos, direction):
            if t.device != dev or t.dtype != torch.float64 or not t.is_contiguous():
                raise ValueError("global_path, pos and direction must be contiguous float64 tensors on the planner's device")
        out = {"path": torch.empty((n, HORIZON, 4), dtype=torch.float32, device=dev),
               "path_f64": torch.empty((n, HORIZON, 4), dtype=torch.float64, device=dev),
               "grid": torch.empty((n, 2), dtype=torch.int16, device=dev),
               "status": torch.empty((n,), dtype=torch.int32, device=dev)}
        stride = 0
        if prev_path is not None:
            stride = 0 if prev_path.numel() == HORIZON * 4 else HORIZON * 4
        with torch.cuda.device(dev):
            ws = torch.empty((int(self.lib.fsd_global_path_workspace_bytes(n)),), dtype=torch.uint8, device=dev)
            _lib.check(self.lib.fsd_global_path_batch(
                C.byref(self.params), n, pos.data_ptr(), direction.data_ptr(), global_path.data_ptr(), global_path.shape[0],
                _ptr(force_P), _ptr(prev_path), stride, out["path"].data_ptr(), out["path_f64"].data_ptr(),
                out["grid"].data_ptr(), out["status"].data_ptr(), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(dev).cuda_stream))
        out["_workspace"] = ws  # keep alive until the stream has consumed it
        return out

    def plan_host(self, batch: FrameBatch, *, force_P: Optional[np.ndarray] = None,
                  prev_path: Optional[np.ndarray] = None, intermediates: bool = False) -> PlanResult:
        """Convenience: host FrameBatch in, device PlanResult out (copies on the current stream)."""
        dev = self.device
        _check_offsets(np.asarray(batch.offsets), len(batch.cones_xy))
        dt = torch.float64 if batch.cones_xy.dtype == np.float64 else torch.float32
        xy = torch.from_numpy(np.ascontiguousarray(batch.cones_xy)).to(dev, dt)
        ty = torch.from_numpy(np.ascontiguousarray(batch.cones_type, dtype=np.uint8)).to(dev)
        off = torch.from_numpy(np.ascontiguousarray(batch.offsets, dtype=np.int32)).to(dev)
        pos = torch.from_numpy(np.ascontiguousarray(batch.pos)).to(dev, dt)
        dr = torch.from_numpy(np.ascontiguousarray(batch.dir)).to(dev, dt)
        fp = None if force_P is None else torch.from_numpy(np.ascontiguousarray(force_P, dtype=np.int16)).to(dev)
        pv = None if prev_path is None else torch.from_numpy(np.ascontiguousarray(prev_path, dtype=np.float64)).to(dev)
        if xy.numel() == 0:
            xy = torch.zeros((1, 2), dtype=dt, device=dev)
            ty = torch.zeros((1,), dtype=torch.uint8, device=dev)
        return self.plan(xy, ty, off, pos, dr, force_P=fp, prev_path=pv, intermediates=intermediates)

    def plan_pinned(self, cones_xy: torch.Tensor, cones_type: torch.Tensor, offsets: torch.Tensor, pos: torch.Tensor,
                    direction: torch.Tensor, out_path: torch.Tensor, out_left_idx: torch.Tensor,
                    out_right_idx: torch.Tensor, out_status: torch.Tensor, *, chunks=None,
                    zero_copy: bool = True, gather: Optional["_lib.Gather"] = None) -> None:
        """Host-to-host entry point: all arguments are PINNED host tensors (inputs as for `plan`, outputs
        [B, 40, 4] float32 / [B, 12] int16 / [B, 12] int16 / [B] int32).  The batch is cut into `chunks` contiguous
        chunks (default: two equal chunks from 5 000 frames on).  Each chunk's host->device copies, its sort + match launches
        (free-running kernels: small chunks cost nothing) and the device->host copy of its sort indices are queued on a
        stream of its own, so the copies of one chunk overlap the kernels of the others; the path stage then runs ONCE
        over the whole batch (its CTA-synchronous rounds want a full grid).  With `zero_copy` (default) the path kernel
        stores every frame's 40 x 4 path STRAIGHT INTO `out_path` -- pinned host memory is mapped into the device's
        address space (unified addressing), the stores are posted writes over PCIe that overlap the kernel's own work --
        so no device->host copy of the paths follows the kernel; only the status words (4 B per frame) are copied.
        `gather`: as for `plan` -- the path kernel also stores the paths into every peer GPU's gathered buffer.
        Asynchronous: everything is ordered on the caller's current stream (synchronize it before reading the
        outputs)."""
        B = offsets.numel() - 1
        if B <= 0:
            return
        for t in (cones_xy, cones_type, offsets, pos, direction, out_path, out_left_idx, out_right_idx, out_status):
            if t.device.type != "cpu" or not t.is_pinned() or not t.is_contiguous():
                raise ValueError("plan_pinned takes contiguous pinned host tensors")
        if cones_xy.dtype not in (torch.float32, torch.float64) or pos.dtype != cones_xy.dtype or \
                direction.dtype != cones_xy.dtype or cones_type.dtype != torch.uint8 or offsets.dtype != torch.int32:
            raise ValueError("inputs must have the dtypes documented for plan()")
        if out_path.dtype != torch.float32 or out_left_idx.dtype != torch.int16 or out_right_idx.dtype != torch.int16 \
                or out_status.dtype != torch.int32 or out_path.numel() != B * HORIZON * 4 \
                or out_left_idx.numel() != B * MAX_SORTED or out_right_idx.numel() != B * MAX_SORTED \
                or out_status.numel() != B:
            raise ValueError("outputs must be [B,40,4] float32, [B,12] int16, [B,12] int16, [B] int32")
        _check_offsets(offsets.numpy(), int(cones_xy.shape[0]))
        f64 = cones_xy.dtype == torch.float64
        # chunks: None = two equal chunks (one below 5 000 frames; measured best, tools/pinned_probe.py: a short first chunk
        # whose copy is exposed for less loses more in kernel launches than it gains); an int = that many equal chunks; a
        # sequence of fractions = the upper bounds of the chunks (ending in 1.0)
        if chunks is None:
            fr = [0.5, 1.0] if B >= 5000 else [1.0]
        elif isinstance(chunks, int):
            fr = [(k + 1) / max(1, min(chunks, B)) for k in range(max(1, min(chunks, B)))]
        else:
            fr = [float(f) for f in chunks]
        bounds = sorted({0, B, *(min(B, max(0, int(round(f * B)))) for f in fr)})
        K = len(bounds) - 1
        dev = self.device
        key = ("pinned", B, int(cones_xy.shape[0]), f64, tuple(bounds))
        if self._pinned_key != key:
            e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
            self._pinned = {
                "xy": e((max(int(cones_xy.shape[0]), 1), 2), cones_xy.dtype), "ty": e((max(int(cones_xy.shape[0]), 1),), torch.uint8),
                "off": e((B + 1,), torch.int32), "pos": e((B, 2), cones_xy.dtype), "dir": e((B, 2), cones_xy.dtype),
                "li": e((B, MAX_SORTED), torch.int16),
                "ri": e((B, MAX_SORTED), torch.int16), "st": e((B,), torch.int32), "bounds": bounds,
                "n_wv": e((B, 2), torch.int16), "left_wv": e((B, MAX_WV, 2), torch.float64),
                "right_wv": e((B, MAX_WV, 2), torch.float64), "l2r": e((B, MAX_WV), torch.int16),
                "r2l": e((B, MAX_WV), torch.int16),
                "streams": [torch.cuda.Stream(dev) for _ in range(K)],
                "sorted": [torch.cuda.Event() for _ in range(K)],
            }
            self._pinned_key = key
        P = self._pinned
        if not zero_copy and "path" not in P:
            P["path"] = torch.empty((B, HORIZON, 4), dtype=torch.float32, device=dev)
        ws = self._workspace(B)
        esz = cones_xy.element_size()
        off_host = offsets.numpy()
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            self._order_after_previous_call(cur)
            for k in range(K):
                lo, hi = P["bounds"][k], P["bounds"][k + 1]
                c0, c1 = int(off_host[lo]), int(off_host[hi])
                st = P["streams"][k]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    if c1 > c0:
                        P["xy"][c0:c1].copy_(cones_xy[c0:c1], non_blocking=True)
                        P["ty"][c0:c1].copy_(cones_type[c0:c1], non_blocking=True)
                    P["off"][lo:hi + 1].copy_(offsets[lo:hi + 1], non_blocking=True)
                    P["pos"][lo:hi].copy_(pos[lo:hi], non_blocking=True)
                    P["dir"][lo:hi].copy_(direction[lo:hi], non_blocking=True)
                    # sort + match of the chunk (the CSR offsets are absolute: a chunk is the same arrays entered at frame lo)
                    inter = _lib.Intermediate(None, P["n_wv"].data_ptr() + 4 * lo, P["left_wv"].data_ptr() + 16 * MAX_WV * lo,
                                              P["right_wv"].data_ptr() + 16 * MAX_WV * lo, P["l2r"].data_ptr() + 2 * MAX_WV * lo,
                                              P["r2l"].data_ptr() + 2 * MAX_WV * lo, None, None)
                    _lib.check(self.lib.fsd_sort_match_batch(
                        C.byref(self.params), hi - lo, int(f64), P["xy"].data_ptr(), P["ty"].data_ptr(),
                        P["off"].data_ptr() + 4 * lo, P["pos"].data_ptr() + 2 * esz * lo, P["dir"].data_ptr() + 2 * esz * lo,
                        P["li"].data_ptr() + 2 * MAX_SORTED * lo, P["ri"].data_ptr() + 2 * MAX_SORTED * lo, C.byref(inter),
                        P["st"].data_ptr() + 4 * lo, st.cuda_stream))
                    P["sorted"][k].record(st)  # the path stage waits for the chunk's kernels, not for its copies
                    out_left_idx[lo:hi].copy_(P["li"][lo:hi], non_blocking=True)
                    out_right_idx[lo:hi].copy_(P["ri"][lo:hi], non_blocking=True)
            for k in range(K):
                cur.wait_event(P["sorted"][k])
            # the path stage over the whole batch, then the paths and the status words back to the host
            inter = _lib.Intermediate(None, P["n_wv"].data_ptr(), P["left_wv"].data_ptr(), P["right_wv"].data_ptr(),
                                      P["l2r"].data_ptr(), P["r2l"].data_ptr(), None, None)
            _lib.check(self.lib.fsd_path_batch_gather(
                C.byref(self.params), B, int(f64), P["pos"].data_ptr(), P["dir"].data_ptr(), C.byref(inter), None,
                self._default_prev().data_ptr(), 0, (out_path if zero_copy else P["path"]).data_ptr(), P["st"].data_ptr(),
                ws.data_ptr(), ws.numel(), cur.cuda_stream, C.byref(gather) if gather is not None else None))
            if not zero_copy:
                out_path.copy_(P["path"], non_blocking=True)
            out_status.copy_(P["st"], non_blocking=True)
            for k in range(K):  # the sort-index copies still running on the chunk streams
                cur.wait_stream(P["streams"][k])
            self._mark_call(cur)

    def initial_path(self) -> torch.Tensor:
        """The constant path of a fresh planner (core_calculate_path.py:103-107), computed on the device."""
        out = torch.empty((HORIZON, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fsd_initial_path(C.byref(self.params), out.data_ptr(),
                                                 torch.cuda.current_stream(self.device).cuda_stream))
        return out


class CpuBatchPlanner:
    """The planner on the host: fsd_plan_batch_cpu, i.e. the kernels' own per-frame sources compiled for the CPU
    (csrc/cpu_backend.cpp).  BASELINE config 1 ("CPU plumbing, no GPU") and machines without a GPU; chosen EXPLICITLY with
    device="cpu" -- the CUDA planner never falls back to it.  Same results as the CUDA path up to fp64 rounding."""

    def __init__(self, mission: int = _lib.MISSION_TRACKDRIVE, threads: Optional[int] = None):
        import os

        self.lib = _lib.lib()
        self.params = _lib.default_params()
        self.mission = int(mission)
        self.threads = int(threads or os.cpu_count() or 1)
        self.device = torch.device("cpu")

    def initial_path(self) -> torch.Tensor:
        out = np.zeros((HORIZON, 4))
        _lib.check(self.lib.fsd_initial_path_cpu(C.byref(self.params), out.ctypes.data))
        return torch.from_numpy(out)

    def plan_host(self, batch: FrameBatch, *, force_P: Optional[np.ndarray] = None,
                  prev_path: Optional[np.ndarray] = None, intermediates: bool = False) -> PlanResult:
        """Host FrameBatch in, PlanResult of CPU tensors out (synchronous)."""
        _check_offsets(np.asarray(batch.offsets), len(batch.cones_xy))
        B = batch.n_frames
        xy = np.ascontiguousarray(batch.cones_xy, dtype=np.float64).reshape(-1, 2)
        ty = np.ascontiguousarray(batch.cones_type, dtype=np.uint8)
        off = np.ascontiguousarray(batch.offsets, dtype=np.int32)
        pos = np.ascontiguousarray(batch.pos, dtype=np.float64)
        dr = np.ascontiguousarray(batch.dir, dtype=np.float64)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt)
        res = PlanResult(z((B, HORIZON, 4), torch.float32), z((B, MAX_SORTED), torch.int16), z((B, MAX_SORTED), torch.int16),
                         z((B,), torch.int32))
        inter = None
        if intermediates:
            res.path_f64, res.n_wv = z((B, HORIZON, 4), torch.float64), z((B, 2), torch.int16)
            res.left_wv, res.right_wv = z((B, MAX_WV, 2), torch.float64), z((B, MAX_WV, 2), torch.float64)
            res.l2r, res.r2l = z((B, MAX_WV), torch.int16), z((B, MAX_WV), torch.int16)
            res.grid, res.sort_dbg = z((B, 2), torch.int16), z((B, 8), torch.int16)
            inter = _lib.Intermediate(*[getattr(res, n).data_ptr() for n in
                                        ("path_f64", "n_wv", "left_wv", "right_wv", "l2r", "r2l", "grid", "sort_dbg")])
        fp = None if force_P is None else np.ascontiguousarray(force_P, dtype=np.int16)
        if fp is not None and fp.size != B:
            raise ValueError("force_P must be int16 [B]")
        pv = None if prev_path is None else np.ascontiguousarray(prev_path, dtype=np.float64)
        stride = 0
        if pv is not None:
            stride = 0 if pv.size == HORIZON * 4 else HORIZON * 4
            if stride and pv.size != B * HORIZON * 4:
                raise ValueError("prev_path must be [40, 4] or [B, 40, 4]")
        if B:
            _lib.check(self.lib.fsd_plan_batch_cpu(
                C.byref(self.params), self.mission, B, xy.ctypes.data if xy.size else None, ty.ctypes.data if ty.size else None,
                off.ctypes.data, pos.ctypes.data, dr.ctypes.data, res.path.data_ptr(), res.left_idx.data_ptr(),
                res.right_idx.data_ptr(), C.byref(inter) if inter is not None else None,
                None if fp is None else fp.ctypes.data, None if pv is None else pv.ctypes.data, stride,
                res.status.data_ptr(), self.threads))
        return res


@dataclass
class RelocalizationInformation:
    """fsd_path_planning/relocalization/relocalization_information.py:13-35"""

    translation: np.ndarray
    rotation: float


class _AccelerationRelocalization:
    """Relocalization of the acceleration / EBS missions (AccelerationRelocalizer,
    fsd_path_planning/relocalization/acceleration/acceleration_relocalization.py:120-172): the cones 0 .. 2 m to the left of
    the car are taken for the left boundary; of 100 random 3-cone subsets the straight line with the smallest squared
    error gives the track direction, and the map frame is the first pose's position turned by that direction.  A
    one-off, ~100 tiny line fits: host code (numpy), like the constants of the skidpad mission.  The reference draws the
    subsets from numpy's GLOBAL unseeded RNG; `seed` (None = the same global RNG) makes the draw reproducible --
    np.random.seed(seed) before the reference's first call gives the same subsets."""

    def __init__(self, seed: Optional[int] = None):
        self._rng = np.random if seed is None else np.random.RandomState(seed)
        self.origin: Optional[np.ndarray] = None
        self.angle: Optional[float] = None

    @property
    def done(self) -> bool:
        return self.angle is not None

    @staticmethod
    def _turn(points: np.ndarray, theta: float) -> np.ndarray:
        c, s = np.cos(theta), np.sin(theta)
        return np.dot(points, np.array(((c, -s), (s, c))).T)

    def attempt(self, cones: Sequence[np.ndarray], position: np.ndarray, direction: np.ndarray) -> None:
        if self.done:
            return
        if self.origin is None:
            self.origin = np.array(position, dtype=np.float64)
        pts = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in cones])
        if len(pts) < 3:
            return
        yaw = np.arctan2(direction[1], direction[0])
        local = self._turn(pts - position, -yaw)
        left = local[(local[:, 1] > 0) & (local[:, 1] < 2)]
        left = left[left[:, 0].argsort()]
        if len(left) < 4:
            return
        best, best_err = None, np.inf
        for _ in range(100):  # the draws must be consumed one by one, in this order
            sub = left[self._rng.choice(left.shape[0], 3, replace=False)]
            coeff = np.polyfit(sub[:, 0], sub[:, 1], 1)
            err = float(np.sum((sub[:, 1] - np.polyval(coeff, sub[:, 0])) ** 2))
            if err < best_err:
                best, best_err = coeff, err
        self.angle = float(np.arctan(best[0]) + yaw)

    def to_map(self, position: np.ndarray, yaw: float):
        return self._turn(np.asarray(position, dtype=np.float64) - self.origin, -self.angle), yaw - self.angle

    def to_world(self, points: np.ndarray) -> np.ndarray:
        return self._turn(points, self.angle) + self.origin


class ReferenceRaisesError(RuntimeError):
    """The reference raises an exception on this input (e.g. the IndexError of functional_cone_matching.py:130 when a
    sorted side holds exactly one cone) or takes its latent-bug path (core_calculate_path.py:482-483).  The batched
    planner flags such frames (FSD_ST_REF_RAISES / FSD_ST_UNSUPPORTED); the drop-in facade raises like the reference
    and, like it, leaves the planner's state (the previous path) untouched."""


class PathPlanner:
    """Drop-in for fsd_path_planning.PathPlanner on the trackdrive / autocross path.

    on_reference_error: "raise" (default, the reference's behaviour: an exception, state untouched) or "previous"
    (return the previous path instead of raising; state untouched as well)."""

    def __init__(self, mission: MissionTypes, experimental_performance_improvements: bool = False,
                 device: Union[str, torch.device, int] = "cuda", on_reference_error: str = "raise",
                 relocalization_seed: Optional[int] = None) -> None:
        if on_reference_error not in ("raise", "previous"):
            raise ValueError('on_reference_error must be "raise" or "previous"')
        self.on_reference_error = on_reference_error
        self.mission = MissionTypes(mission)
        self._accel: Optional[_AccelerationRelocalization] = None
        if self.mission in (MissionTypes.acceleration, MissionTypes.ebs_test):
            if str(device) == "cpu":
                raise NotImplementedError("the host planner (device='cpu') covers trackdrive / autocross only")
            self._accel = _AccelerationRelocalization(relocalization_seed)
        # the experimental sorting cache of the reference changes results and is not reproduced
        self.experimental_performance_improvements = experimental_performance_improvements
        self._cpu = str(device) == "cpu"
        if self._cpu and self.mission == MissionTypes.skidpad:
            raise NotImplementedError("the host planner (device='cpu') covers trackdrive / autocross only")
        # device="cpu" selects the host build of the kernels' sources (fsd_plan_batch_cpu) EXPLICITLY; the default
        # (CUDA) never falls back to it and raises without a GPU
        self._planner = CpuBatchPlanner(threads=1) if self._cpu else BatchPlanner(device, mission=_lib.MISSION_TRACKDRIVE)
        self.global_path = None
        self._prev_path: Optional[torch.Tensor] = None  # previous_paths[-1] of the reference
        self._skid = None
        if self.mission == MissionTypes.skidpad:
            from .skidpad import SkidpadBatchPlanner

            dev = self._planner.device
            self._skid = SkidpadBatchPlanner(dev)
            self._reloc = torch.zeros((1, 8), dtype=torch.float64, device=dev)
            self._reloc_host = np.zeros(8)
            self._orig_pose = None
            self._index_state = torch.zeros((1,), dtype=torch.int32, device=dev)

    @staticmethod
    def _convert_direction_to_array(direction: Any) -> np.ndarray:
        direction = np.squeeze(np.array(direction))
        if direction.shape == (2,):
            return direction.astype(np.float64)
        if direction.shape in [(1,), ()]:
            yaw = float(np.asarray(direction).reshape(-1)[0])
            return np.array([np.cos(yaw), np.sin(yaw)])
        raise ValueError("direction must be a float or a 2 element array")

    def set_global_path(self, global_path):
        """full_pipeline.py:81: from now on the path is calculated along this (M, 2) line (core_calculate_path.py:516-528)."""
        self.global_path = None if global_path is None else np.ascontiguousarray(global_path, dtype=np.float64).reshape(-1, 2)
        self._global_path_dev = None

    def _global_path_step(self, position: np.ndarray, direction: np.ndarray) -> np.ndarray:
        """One call with a global path set: fsd_global_path_batch on one pose, stateful like the reference."""
        if self._cpu:
            raise NotImplementedError("global-path planning is not part of the host planner (device='cpu')")
        dev = self._planner.device
        if getattr(self, "_global_path_dev", None) is None:
            self._global_path_dev = torch.from_numpy(self.global_path).to(dev)
        if self._prev_path is None:
            self._prev_path = self._planner.initial_path()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(1, 2)).to(dev)
        res = self._planner.plan_global_path(self._global_path_dev, t(position), t(direction), prev_path=self._prev_path)
        status = int(res["status"][0].item())
        if status & (_lib.STATUS_BITS["REF_RAISES"] | _lib.STATUS_BITS["UNSUPPORTED"]):
            if self.on_reference_error == "raise":
                raise ReferenceRaisesError(f"the reference planner raises on this input (status 0x{status:x})")
            return self._prev_path.cpu().numpy()
        self._prev_path = res["path_f64"][0].clone()
        return self._prev_path.cpu().numpy()

    def _acceleration_step(self, cones, position, direction, return_intermediate_results):
        """full_pipeline.py:122-194 for the acceleration / EBS missions: relocalize once, then follow the known map."""
        self._accel.attempt(cones, position, direction)
        e2, ei = np.zeros((0, 2)), np.zeros(0, dtype=int)
        if self._accel.done:
            yaw = float(np.arctan2(direction[1], direction[0]))
            pos_map, yaw_map = self._accel.to_map(position, yaw)
            if self.global_path is None:
                import os

                self.set_global_path(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                                                         "acceleration_path.npy")))
            path = self._global_path_step(pos_map, np.array([np.cos(yaw_map), np.sin(yaw_map)])).copy()
            path[:, 1:3] = self._accel.to_world(path[:, 1:3])
        else:
            # not relocalized yet: no sorting / matching in these missions, so both sides are empty and the planner
            # returns the path of the previous call (core_calculate_path.py:531-536)
            batch = pack_frames([([e2] * 5, position, direction)], dtype=np.float64)
            if self._prev_path is None:
                self._prev_path = self._planner.initial_path()
            res = self._plan_with_prev(batch)
            self._prev_path = res.path_f64[0].clone()
            path = self._prev_path.cpu().numpy()
        if not return_intermediate_results:
            return path
        return path, e2, e2, e2, e2, ei, ei

    @property
    def relocalization_info(self) -> Optional[RelocalizationInformation]:
        """relocalization_information.py:13-35: where the SLAM origin and the x axis land in the map frame."""
        if self._accel is not None:
            if not self._accel.done:
                return None
            o, _ = self._accel.to_map(np.zeros(2), 0.0)
            e, _ = self._accel.to_map(np.array([1.0, 0.0]), 0.0)
            return RelocalizationInformation(o, float(np.arctan2(e[1] - o[1], e[0] - o[0])))
        if self._skid is None or self._reloc_host[7] == 0.0:
            return None
        from .skidpad import to_known_frame

        o = to_known_frame(self._reloc_host, np.zeros(2))
        e = to_known_frame(self._reloc_host, np.array([1.0, 0.0]))
        return RelocalizationInformation(o, float(np.arctan2(e[1] - o[1], e[0] - o[0])))

    def _skidpad_step(self, cones, position, direction, return_intermediate_results):
        """full_pipeline.py:122-140, 165-194 for MissionTypes.skidpad (sorting and matching are skipped)."""
        dev = self._planner.device
        t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
        pos_d, dir_d = t(position.reshape(1, 2)), t(direction.reshape(1, 2))
        if self._reloc_host[7] == 0.0:
            if self._orig_pose is None:
                self._orig_pose = (pos_d.clone(), dir_d.clone())
            xy = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in cones])
            if len(xy) > 0:
                off = t(np.array([0, len(xy)]), torch.int32)
                self._reloc, _ = self._skid.relocalize(t(xy), off, pos_d, self._orig_pose[0], self._orig_pose[1])
                self._reloc_host = self._reloc[0].cpu().numpy()
        if self._prev_path is None:
            self._prev_path = self._planner.initial_path()
        res = self._skid.plan(t(np.array([0, 1]), torch.int32), pos_d, dir_d, self._reloc, self._index_state,
                              prev_path=self._prev_path)
        self._prev_path = res["internal"][0].clone()
        path = res["path_f64"][0].cpu().numpy()
        if not return_intermediate_results:
            return path
        e2, ei = np.zeros((0, 2)), np.zeros(0, dtype=int)
        return path, e2, e2, e2, e2, ei, ei

    def calculate_path_in_global_frame(
        self,
        cones: List[np.ndarray],
        vehicle_position: np.ndarray,
        vehicle_direction: Union[np.ndarray, float],
        return_intermediate_results: bool = False,
    ):
        """Same contract as the reference: returns a (40, 4) float64 array [u, x, y, curvature], or the
        7-tuple (path, sorted_left, sorted_right, left_with_virtual, right_with_virtual, l2r, r2l)."""
        direction = self._convert_direction_to_array(vehicle_direction)
        if self._skid is not None:
            return self._skidpad_step(cones, np.asarray(vehicle_position, dtype=np.float64).reshape(2), direction,
                                      return_intermediate_results)
        position = np.asarray(vehicle_position, dtype=np.float64).reshape(2)
        if self._accel is not None:
            return self._acceleration_step(cones, position, direction, return_intermediate_results)
        batch = pack_frames([(cones, position, direction)], dtype=np.float64)
        if self._prev_path is None:
            self._prev_path = self._planner.initial_path()
        if self.global_path is not None:
            # set_global_path: sorting and matching still run (their results are returned as intermediates), the path
            # follows the global line (core_calculate_path.py:516-528)
            out_path = self._global_path_step(position, direction)
            if not return_intermediate_results:
                return out_path
            res = self._plan_with_prev(batch)
            return (out_path, *self._intermediates(batch, res))
        res = self._plan_with_prev(batch)
        status = int(res.status[0].item())
        if status & (_lib.STATUS_BITS["REF_RAISES"] | _lib.STATUS_BITS["UNSUPPORTED"]):
            # the reference raises here and never reaches `self.previous_paths = ...` (core_calculate_path.py:572)
            if self.on_reference_error == "raise":
                raise ReferenceRaisesError(f"the reference planner raises on this input (status 0x{status:x})")
            res.path_f64[0].copy_(self._prev_path)
        else:
            self._prev_path = res.path_f64[0].clone()
        out_path = res.path_f64[0].cpu().numpy()
        if not return_intermediate_results:
            return out_path
        return (out_path, *self._intermediates(batch, res))

    @staticmethod
    def _intermediates(batch: FrameBatch, res: PlanResult):
        """(sorted_left, sorted_right, left_with_virtual, right_with_virtual, l2r, r2l) of a one-frame result."""
        xy = batch.cones_xy
        li = res.left_idx[0].cpu().numpy()
        ri = res.right_idx[0].cpu().numpy()
        n_wv = res.n_wv[0].cpu().numpy()
        nl, nr = int(n_wv[0]), int(n_wv[1])
        return (
            xy[li[li >= 0]],
            xy[ri[ri >= 0]],
            res.left_wv[0, :nl].cpu().numpy(),
            res.right_wv[0, :nr].cpu().numpy(),
            res.l2r[0, :nl].cpu().numpy().astype(np.int64),
            res.r2l[0, :nr].cpu().numpy().astype(np.int64),
        )

    def _plan_with_prev(self, batch: FrameBatch) -> PlanResult:
        if self._cpu:
            return self._planner.plan_host(batch, prev_path=self._prev_path.numpy(), intermediates=True)
        dev = self._planner.device
        xy = torch.from_numpy(batch.cones_xy).to(dev)
        ty = torch.from_numpy(batch.cones_type).to(dev)
        if xy.numel() == 0:
            xy = torch.zeros((1, 2), dtype=torch.float64, device=dev)
            ty = torch.zeros((1,), dtype=torch.uint8, device=dev)
        return self._planner.plan(xy, ty, torch.from_numpy(batch.offsets).to(dev), torch.from_numpy(batch.pos).to(dev),
                                  torch.from_numpy(batch.dir).to(dev), prev_path=self._prev_path, intermediates=True)

    def calculate_paths_batched(self, cones_xy: torch.Tensor, cones_type: torch.Tensor, offsets: torch.Tensor,
                                positions: torch.Tensor, directions: torch.Tensor, *,
                                return_intermediate: bool = False) -> PlanResult:
        """Batched entry point next to the single-frame one: torch CUDA tensors in, PlanResult out; every frame
        is planned by a fresh planner (no state is read or written)."""
        return self._planner.plan(cones_xy, cones_type, offsets, positions, directions,
                                  intermediates=return_intermediate)
