"""Developer probe (GPU box): stage times of one rank's shard of the multi-GPU bench (frames [rank * n, (rank + 1) * n) of
the bench stream) -- rank 5's shard holds a frame that overflows the knot bound of the path kernel."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

n = 10240
for rank in [int(a) for a in sys.argv[1:]] or [0, 5]:
    batch = synth.gen_autocross(2, n, start=rank * n, workers=16)
    bp = BatchPlanner("cuda:0")
    args = tuple(torch.from_numpy(a).to(bp.device) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=bp.device)
    out = None
    for _ in range(3):
        out = bp.plan(*args, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = bp.plan(*args, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    for _ in range(10):
        flush.zero_()
        bp.plan(*args, kernel_events=True)
    torch.cuda.synchronize()
    kt = bp.kernel_times_ms()
    st = out.status.cpu().numpy().astype(np.uint32)
    print(f"rank {rank}: step {np.median(ts):.3f} ms, sort+match {np.median([t[0] for t in kt]):.3f} ms, path {np.median([t[1] for t in kt]):.3f} ms, "
          f"overflow-flagged frames {int(((st & 0x100) != 0).sum())}")
