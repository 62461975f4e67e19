"""world_size-2 run of the multi-GPU plumbing on the gloo backend (CPU tensors): block sharding + one all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ft_fsd_path_planning_b200.distributed import all_gather_frames, shard_bounds


def _worker(rank, world, port, n_frames, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_frames, rank, world)
    # stand-in for the per-frame planner output: a deterministic function of the global frame index
    idx = torch.arange(lo, hi, dtype=torch.float32)
    local = idx[:, None, None] * torch.ones((hi - lo, 40, 4)) + torch.arange(4.0)
    full = all_gather_frames(local, n_frames)
    expect = torch.arange(n_frames, dtype=torch.float32)[:, None, None] * torch.ones((n_frames, 40, 4)) + torch.arange(4.0)
    ret[rank] = bool(torch.equal(full, expect))
    status = all_gather_frames(torch.full((hi - lo,), rank, dtype=torch.int32), n_frames)
    ret[rank] = ret[rank] and status.shape[0] == n_frames and int(status[0]) == 0
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
