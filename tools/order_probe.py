"""Developer probe (GPU box): does grouping frames of similar size (cone count) into the same CTA round help?"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402
from ft_fsd_path_planning_b200.synth import FrameBatch  # noqa: E402

n = 10240
batch = synth.gen_autocross(2, n)


def permute(b, perm):
    cnt = np.diff(b.offsets)[perm]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    idx = np.concatenate([np.arange(b.offsets[p], b.offsets[p + 1]) for p in perm])
    return FrameBatch(b.cones_xy[idx], b.cones_type[idx], off, b.pos[perm], b.dir[perm])


dev = torch.device("cuda:0")
bp = BatchPlanner(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, perm in (("original", np.arange(n)), ("sorted by N", np.argsort(np.diff(batch.offsets), kind="stable")),
                   ("sorted by N desc", np.argsort(-np.diff(batch.offsets), kind="stable"))):
    b = permute(batch, perm)
    xy, ty, off = (torch.from_numpy(a).to(dev) for a in (b.cones_xy, b.cones_type, b.offsets))
    pos, dr = torch.from_numpy(b.pos).to(dev), torch.from_numpy(b.dir).to(dev)
    for _ in range(3):
        bp.plan(xy, ty, off, pos, dr, kernel_events=True)
    torch.cuda.synchronize()
    bp.kernel_times_ms()
    for _ in range(10):
        flush.fill_(1)
        bp.plan(xy, ty, off, pos, dr, kernel_events=True)
    torch.cuda.synchronize()
    kt = np.array(bp.kernel_times_ms())
    print(f"{name:18s}: sort_match {kt[:, 0].mean():.3f} ms, path {kt[:, 1].mean():.3f} ms")
