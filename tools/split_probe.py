"""Developer probe (GPU box): the batch planned as K chunks on K streams (tails of one kernel overlap the next)."""
import sys

import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
batch = synth.gen_autocross(2, n)
dev = torch.device("cuda:0")
xy, ty, off = (torch.from_numpy(a).to(dev) for a in (batch.cones_xy, batch.cones_type, batch.offsets))
pos, dr = torch.from_numpy(batch.pos).to(dev), torch.from_numpy(batch.dir).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for K in (1, 2, 3, 4, 8):
    planners = [BatchPlanner(dev) for _ in range(K)]
    streams = [torch.cuda.Stream(dev) for _ in range(K)]
    bounds = [n * i // K for i in range(K + 1)]

    def step():
        cur = torch.cuda.current_stream()
        for k in range(K):
            streams[k].wait_stream(cur)
            with torch.cuda.stream(streams[k]):
                lo, hi = bounds[k], bounds[k + 1]
                planners[k].plan(xy, ty, off[lo:hi + 1], pos[lo:hi], dr[lo:hi])
        for k in range(K):
            cur.wait_stream(streams[k])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"K={K}: median {ts[len(ts) // 2]:.3f} ms, min {ts[0]:.3f} ms -> {n / ts[len(ts) // 2] * 1e3:.0f} frames/s")
