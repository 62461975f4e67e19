"""Developer probe (GPU box): plan_pinned end-to-end time for different chunk counts."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

n = 10240
batch = synth.gen_autocross(2, n)
bp = BatchPlanner("cuda:0")
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = [pin(a) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir)]
out = [torch.zeros((n, 40, 4), dtype=torch.float32).pin_memory(), torch.zeros((n, 12), dtype=torch.int16).pin_memory(),
       torch.zeros((n, 12), dtype=torch.int16).pin_memory(), torch.zeros((n,), dtype=torch.int32).pin_memory()]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
ref = None
for zc in (False, True):
    for K in (2, None, (0.0625, 0.25, 0.5, 1.0), (0.25, 1.0), (0.125, 0.375, 0.6875, 1.0)):
        for o in out:
            o.zero_()
        for _ in range(3):
            bp.plan_pinned(*h, *out, chunks=K, zero_copy=zc)
        torch.cuda.synchronize()
        got = [o.clone() for o in out]
        if ref is None:
            ref = got
        same = all(torch.equal(a, b) for a, b in zip(ref, got))
        ts = []
        for _ in range(30):  # the host does not touch the output buffers between the timed calls
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bp.plan_pinned(*h, *out, chunks=K, zero_copy=zc)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"zero_copy={zc} chunks={K}: mean {np.mean(ts):.3f} ms, median {np.median(ts):.3f} ms, min {np.min(ts):.3f} ms, outputs identical: {same}")
