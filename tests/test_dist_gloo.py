"""world_size-2 runs of the multi-GPU plumbing on the gloo backend (CPU tensors): block sharding + one all-gather,
the two-part pipeline (gather of part p next to the planning of part p + 1) and the skidpad split by trajectory."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ft_fsd_path_planning_b200.distributed import (GatherPipeline, all_gather_frames, all_gather_ragged, shard_blocks,
                                                   shard_bounds, shard_trajectories)


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _frame_values(lo, hi):
    # stand-in for the per-frame planner output: a deterministic function of the global frame index
    idx = torch.arange(lo, hi, dtype=torch.float32)
    return idx[:, None, None] * torch.ones((hi - lo, 40, 4)) + torch.arange(4.0)


def _worker(rank, world, port, n_frames, ret):
    _init(rank, world, port)
    lo, hi = shard_bounds(n_frames, rank, world)
    full = all_gather_frames(_frame_values(lo, hi), n_frames)
    ret[rank] = bool(torch.equal(full, _frame_values(0, n_frames)))
    status = all_gather_frames(torch.full((hi - lo,), rank, dtype=torch.int32), n_frames)
    ret[rank] = ret[rank] and status.shape[0] == n_frames and int(status[0]) == 0
    # pre-allocated result buffer, re-used across steps
    per = (n_frames + world - 1) // world
    buf = torch.empty((world * per, 40, 4))
    again = all_gather_frames(_frame_values(lo, hi), n_frames, out=buf)
    ret[rank] = ret[rank] and again.data_ptr() == buf.data_ptr() and bool(torch.equal(again, full))
    dist.destroy_process_group()


def _pipeline_worker(rank, world, port, n_frames, ret):
    _init(rank, world, port)
    pipe = GatherPipeline(n_frames, (40, 4), torch.float32, torch.device("cpu"), parts=2)
    blocks = shard_blocks(n_frames, rank, world, 2)
    ok = True
    for step in range(2):  # the result buffer is re-used from step to step
        for p, (lo, hi) in enumerate(blocks):
            pipe.gather(p, _frame_values(lo, hi) + step)
        ok = ok and bool(torch.equal(pipe.finish(), _frame_values(0, n_frames) + step))
    # the blocks of all ranks tile the batch exactly once
    cover = torch.zeros(n_frames, dtype=torch.int32)
    for r in range(world):
        for lo, hi in shard_blocks(n_frames, r, world, 2):
            cover[lo:hi] += 1
    ret[rank] = ok and bool((cover == 1).all())
    dist.destroy_process_group()


def _sequential_stand_in(step_offsets, pos, direction, reloc, index_state):
    """CPU stand-in for SkidpadBatchPlanner.plan with the property that matters for the split: a step depends on every
    earlier step of ITS trajectory (running sum) and on nothing else."""
    off = step_offsets.tolist()
    S = pos.shape[0]
    run = torch.zeros(S, dtype=torch.float64)
    for t in range(len(off) - 1):
        run[off[t]:off[t + 1]] = torch.cumsum(pos[off[t]:off[t + 1], 0] * reloc[t, 0], 0) + index_state[t]
        index_state[t] += off[t + 1] - off[t]
    base = run[:, None, None] * torch.ones((S, 40, 4), dtype=torch.float64)
    return {"path": base.float(), "path_f64": base, "internal": base + 1.0, "index": run.to(torch.int32),
            "grid": torch.stack([run.to(torch.int16), run.to(torch.int16)], 1), "status": torch.zeros(S, dtype=torch.int32)}


def _skidpad_worker(rank, world, port, ret):
    from ft_fsd_path_planning_b200.skidpad import plan_skidpad_sharded

    _init(rank, world, port)
    rng = np.random.default_rng(0)
    counts = np.array([5, 40, 3, 17, 9, 1, 22])  # ragged trajectories
    off = np.concatenate([[0], np.cumsum(counts)])
    S, T = int(off[-1]), len(counts)
    pos = torch.from_numpy(rng.integers(1, 5, (S, 2)).astype(np.float64))
    direction = torch.ones((S, 2), dtype=torch.float64)
    reloc = torch.from_numpy(rng.integers(1, 3, (T, 8)).astype(np.float64))
    state = torch.arange(T, dtype=torch.int32)
    expect = _sequential_stand_in(torch.from_numpy(off.astype(np.int32)), pos, direction, reloc, state.clone())
    got, new_state = plan_skidpad_sharded(None, off, pos, direction, reloc, state, plan_fn=_sequential_stand_in)
    ok = all(torch.equal(got[k], expect[k]) for k in expect)
    ok = ok and bool(torch.equal(new_state, torch.arange(T, dtype=torch.int32) + torch.from_numpy(counts).int()))
    # whole trajectories only, every trajectory exactly once, roughly balanced by steps
    cuts = [shard_trajectories(off, r, world) for r in range(world)]
    ok = ok and cuts[0][0] == 0 and cuts[-1][1] == T and all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
    steps = [int(off[c[1]] - off[c[0]]) for c in cuts]
    ok = ok and max(steps) <= S // world + int(counts.max())
    ragged = all_gather_ragged(torch.full((rank + 1, 2), float(rank)), [r + 1 for r in range(world)])
    ok = ok and ragged.shape[0] == world * (world + 1) // 2
    ret[rank] = ok
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_shard_and_gather():
    for n_frames in (7, 16):
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(2, _free_port(), n_frames, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]


def test_two_rank_gather_pipeline():
    ret = mp.Manager().dict()
    mp.spawn(_pipeline_worker, args=(2, _free_port(), 16, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_two_rank_skidpad_split_by_trajectory():
    ret = mp.Manager().dict()
    mp.spawn(_skidpad_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_shard_trajectories_single_trajectory_is_replicas_only():
    # one trajectory cannot be split: rank 0 gets it, the others get nothing
    assert shard_trajectories([0, 256], 0, 4) in ((0, 1), (0, 0))
    got = [shard_trajectories([0, 256], r, 4) for r in range(4)]
    assert sum(hi - lo for lo, hi in got) == 1
