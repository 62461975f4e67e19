"""Small target for ncu: a few planner calls on the bench workload (device-resident inputs)."""
import sys

import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
import os

batch = synth.gen_autocross(2, n, workers=min(os.cpu_count() or 1, 16))
bp = BatchPlanner("cuda:0")
dev = bp.device
xy, ty, off = (torch.from_numpy(a).to(dev) for a in (batch.cones_xy, batch.cones_type, batch.offsets))
pos, dr = torch.from_numpy(batch.pos).to(dev), torch.from_numpy(batch.dir).to(dev)
stage = len(sys.argv) > 3 and sys.argv[3] == "stage"  # the two stage entry points over the whole batch (no chunking)
for _ in range(iters):
    bp.plan(xy, ty, off, pos, dr, kernel_events=stage)
if stage:
    bp.knn(xy, ty, off)  # the cost-matrix step in isolation (fsd_knn_batch)
torch.cuda.synchronize()
print("done", n, iters)
