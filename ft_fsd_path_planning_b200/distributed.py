"""Multi-GPU plumbing: frames are independent, so a batch is partitioned over the ranks with no data-path
collective; all-gathers of the fixed-size outputs collect the paths.

One process per GPU (`torchrun`), `torch.distributed` with the nccl backend on GPUs; the same code runs on
the gloo backend with CPU tensors, which is how tests/test_dist_gloo.py covers it without a GPU.

Two partitions:
  * `shard_bounds`      one contiguous block of ceil(B / G) frames per rank, one all-gather (`all_gather_frames`);
  * `shard_blocks`      `parts` blocks per rank (block p of rank r = frames [p * B/parts + r * h, ... + h)), so that the
                        all-gather of part p fills the contiguous slice p of the result and can run on a side stream
                        while the rank plans part p + 1 (`GatherPipeline`).
`PeerGather` removes the collective altogether: the path kernel stores every finished frame's path straight into ALL
GPUs' gathered buffers (peer-mapped symmetric memory over NVLink, or one multimem store through the NVSwitch multicast
address), so the "all-gather" overlaps the planning frame by frame and only a cross-GPU barrier remains.

Skidpad: steps of one trajectory are sequential, trajectories are independent ->
`shard_trajectories` gives every rank whole trajectories.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of ceil(B / G) frames per rank; trailing ranks may get fewer (or none)."""
    per = (n_frames + world_size - 1) // world_size
    lo = min(rank * per, n_frames)
    hi = min(lo + per, n_frames)
    return lo, hi


def shard_blocks(n_frames: int, rank: int, world_size: int, parts: int = 2) -> List[Tuple[int, int]]:
    """`parts` blocks per rank.  The batch is cut into `parts` equal slices (n_frames must be a multiple of
    parts * world_size); slice p is block-partitioned over the ranks.  Returns [(lo, hi)] * parts."""
    if n_frames % (parts * world_size) != 0:
        raise ValueError("n_frames must be a multiple of parts * world_size")
    h = n_frames // (parts * world_size)
    return [(p * (n_frames // parts) + rank * h, p * (n_frames // parts) + (rank + 1) * h) for p in range(parts)]


def shard_parts(part_sizes: List[int], rank: int, world_size: int) -> List[Tuple[int, int]]:
    """Uneven version of shard_blocks: every rank plans sum(part_sizes) frames as len(part_sizes) blocks; the global
    batch is [part 0 of rank 0 .. part 0 of rank G-1 | part 1 of rank 0 .. | ...].  Returns this rank's [(lo, hi)]."""
    out, base = [], 0
    for sz in part_sizes:
        out.append((base + rank * sz, base + (rank + 1) * sz))
        base += world_size * sz
    return out


def shard_trajectories(step_offsets, rank: int, world_size: int) -> Tuple[int, int]:
    """Skidpad: trajectories [t_lo, t_hi) of this rank.  Whole trajectories only (their steps are sequential); the
    cut points balance the number of STEPS per rank (greedy prefix split of the step counts).  `step_offsets` is the
    [T + 1] CSR array of fsd_skidpad_plan_batch."""
    off = [int(v) for v in step_offsets]
    T, total = len(off) - 1, off[-1] - off[0]
    cuts = [0]
    for r in range(1, world_size):
        target = off[0] + total * r / world_size
        t = cuts[-1]
        while t < T and off[t + 1] <= target:
            t += 1
        # the trajectory straddling the target goes to the side with the smaller imbalance
        if t < T and (target - off[t]) > (off[t + 1] - target):
            t += 1
        cuts.append(max(t, cuts[-1]))
    cuts.append(T)
    return cuts[rank], cuts[rank + 1]


def all_gather_frames(local: torch.Tensor, n_frames: int, group: Optional[dist.ProcessGroup] = None,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gather per-frame outputs ([n_local, ...], rank-contiguous blocks as in shard_bounds) into [n_frames, ...].

    Every rank contributes a block padded to ceil(B / G) rows so that one equal-count all-gather suffices.
    `out` ([G * ceil(B / G), ...]) is re-used when given (no allocation per step)."""
    world = dist.get_world_size(group)
    per = (n_frames + world - 1) // world
    tail = local.shape[1:]
    if local.shape[0] != per:
        padded = local.new_zeros((per, *tail))
        padded[: local.shape[0]] = local
    else:
        padded = local.contiguous()
    if out is None:
        out = local.new_empty((world * per, *tail))
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_frames]


def all_gather_ragged(local: torch.Tensor, counts: List[int], group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Gather blocks of different length (rank r contributes counts[r] rows): padded to max(counts), one equal-count
    all-gather, then compacted.  Used for the skidpad outputs (whole trajectories per rank)."""
    world = dist.get_world_size(group)
    per = max(counts) if counts else 0
    tail = local.shape[1:]
    padded = local.new_zeros((per, *tail))
    padded[: local.shape[0]] = local
    out = local.new_empty((world * per, *tail))
    if per > 0:
        # gathered as raw bytes: every backend moves uint8 (gloo has no int16)
        dist.all_gather_into_tensor(out.view(-1).view(torch.uint8), padded.view(-1).view(torch.uint8), group=group)
    return torch.cat([out[r * per : r * per + counts[r]] for r in range(world)], 0)


class GatherPipeline:
    """All-gather of part p on a side stream while the caller's stream plans part p + 1.

        pipe = GatherPipeline(n_frames, (40, 4), torch.float32, device, parts=2)
        for p, (lo, hi) in enumerate(shard_blocks(n_frames, rank, world, 2)):
            res = planner.plan(...frames lo:hi...)
            pipe.gather(p, res.path)       # returns at once; the gather waits for the planner on the side stream
        full = pipe.finish()               # [n_frames, 40, 4]; the caller's stream waits for the gathers

    The result buffer is allocated once.  On CPU tensors (gloo, tests) the gathers run inline."""

    def __init__(self, n_frames: int, tail: Tuple[int, ...], dtype: torch.dtype, device: torch.device, parts: int = 2,
                 group: Optional[dist.ProcessGroup] = None, part_sizes: Optional[List[int]] = None):
        """part_sizes: frames per rank in each part (shard_parts layout); default: `parts` equal parts (shard_blocks)."""
        self.world = dist.get_world_size(group)
        if part_sizes is None:
            if n_frames % (parts * self.world) != 0:
                raise ValueError("n_frames must be a multiple of parts * world_size")
            part_sizes = [n_frames // (parts * self.world)] * parts
        if sum(part_sizes) * self.world != n_frames:
            raise ValueError("part_sizes must add up to n_frames / world_size")
        parts = len(part_sizes)
        self.bounds = [0]
        for sz in part_sizes:
            self.bounds.append(self.bounds[-1] + self.world * sz)
        self.group, self.parts, self.n = group, parts, n_frames
        self.out = torch.empty((n_frames, *tail), dtype=dtype, device=device)
        self.cuda = device.type == "cuda"
        self.stream = torch.cuda.Stream(device) if self.cuda else None
        self.events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if self.cuda else None
                       for _ in range(parts)]

    def gather(self, part: int, local: torch.Tensor, after: Optional["torch.cuda.Event"] = None) -> None:
        """Queue the all-gather of this rank's block of part `part`.  The gather starts after everything queued so far
        on the caller's stream, or -- when `after` is given -- as soon as that event has fired (the planner's
        `chunk_ready`: the block is final although the caller's stream still has work queued)."""
        dst = self.out[self.bounds[part] : self.bounds[part + 1]]
        if not self.cuda:
            dist.all_gather_into_tensor(dst, local.contiguous(), group=self.group)
            return
        cur = torch.cuda.current_stream(local.device)
        if after is not None:
            self.stream.wait_event(after)
        else:
            self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.events[part][0].record()
            dist.all_gather_into_tensor(dst, local, group=self.group)
            self.events[part][1].record()
        local.record_stream(self.stream)

    def finish(self) -> torch.Tensor:
        if self.cuda:
            torch.cuda.current_stream(self.out.device).wait_stream(self.stream)
        return self.out

    def gather_ms(self) -> float:
        """Device time of the gathers of the last step (sum over the parts); synchronize first."""
        if not self.cuda:
            return 0.0
        return float(sum(a.elapsed_time(b) for a, b in self.events))


class PeerGather:
    """The all-gather of the output paths FUSED into the path kernel (fsdplan.h: fsd_gather, fsd_plan_batch_gather).

        pg = PeerGather(n_global, device)                       # collective: every rank of the group calls it
        res = planner.plan(..., gather=pg.descriptor(lo))        # lo = global row of this rank's first frame
        full = pg.finish()                                       # [n_global, 40, 4]; a cross-GPU barrier, no copy

    The gathered buffers live in torch symmetric memory (CUDA VMM allocations exchanged between the ranks' processes and
    mapped into every GPU's address space): the kernel writes row `lo + b` of every peer's buffer through the peer
    pointers -- or, where the NVSwitch fabric offers a multicast address for the allocation, with ONE multimem store per
    value that the switch replicates to all GPUs.  `buffers` gathered buffers are used in turn, so a rank that is already
    planning step k + 1 never overwrites the step-k result a slower peer is still reading.  CUDA + NCCL group only; the
    gloo / CPU path keeps `GatherPipeline`."""

    def __init__(self, n_frames: int, device: torch.device, group: Optional[dist.ProcessGroup] = None, buffers: int = 2,
                 multicast: bool = True, tail: Tuple[int, ...] = (40, 4)):
        import torch.distributed._symmetric_memory as symm_mem

        g = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(g), dist.get_rank(g)
        if self.world > 16:
            raise ValueError("fsd_gather holds at most 16 peers")
        self.n, self.device = n_frames, torch.device(device)
        self.bufs, self.hdls = [], []
        with torch.cuda.device(self.device):
            for _ in range(buffers):
                t = symm_mem.empty((n_frames, *tail), dtype=torch.float32, device=self.device)
                try:
                    h = symm_mem.rendezvous(t, g)
                except TypeError:
                    h = symm_mem.rendezvous(t, g.group_name)
                self.bufs.append(t)
                self.hdls.append(h)
        self.multicast = bool(multicast) and all(int(getattr(h, "multicast_ptr", 0) or 0) != 0 for h in self.hdls)
        self.cur = 0

    @staticmethod
    def make_descriptor(peer_ptrs: List[int], first_row: int, multicast_ptr: int = 0):
        """fsd_gather from raw device pointers (one per rank, every GPU's [n_global, 40, 4] fp32 buffer)."""
        from . import _lib

        if len(peer_ptrs) > _lib.MAX_PEERS:
            raise ValueError("too many peers")
        d = _lib.Gather()
        d.n_peers, d.first_row = (0 if multicast_ptr else len(peer_ptrs)), int(first_row)
        for r, ptr in enumerate(peer_ptrs):
            d.peer_out_path[r] = int(ptr)
        d.multicast_out_path = int(multicast_ptr) or None
        return d

    def descriptor(self, first_row: int):
        """Descriptor of the buffer of the current step; `first_row`: global row of the caller's frame 0."""
        h = self.hdls[self.cur]
        return self.make_descriptor([int(p) for p in h.buffer_ptrs], first_row,
                                    int(h.multicast_ptr) if self.multicast else 0)

    def finish(self) -> torch.Tensor:
        """Cross-GPU barrier on the current stream (every rank's kernels, hence its peer stores, have completed when it
        passes), returns the gathered buffer of this step and moves on to the next buffer."""
        with torch.cuda.device(self.device):
            self.hdls[self.cur].barrier()
        out = self.bufs[self.cur]
        self.cur = (self.cur + 1) % len(self.bufs)
        return out
