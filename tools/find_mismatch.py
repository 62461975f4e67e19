"""Developer probe (GPU box): frames of the 8 x 10 240-frame bench stream whose CUDA path differs from the oracle's."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import oracle  # noqa: E402
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

bp = BatchPlanner("cuda:0")
n = 10240
for rank in range(8):
    batch = synth.gen_autocross(2, n, start=rank * n, workers=16)
    ref = oracle.plan_batch(batch.astype(np.float64), threads=16)
    r = bp.plan_host(batch, intermediates=True)
    torch.cuda.synchronize()
    p64 = r.path_f64.cpu().numpy()
    grid = r.grid.cpu().numpy()
    same = grid[:, 0] == ref["P"]
    err = np.abs(p64 - ref["path"]).reshape(n, -1).max(1)
    bad = np.nonzero(same & (err > 1e-6))[0]
    for b in bad:
        print(f"rank {rank} frame {rank * n + b}: err {err[b]:.3e} status {int(r.status[b]):#x} ref status {int(ref['status'][b]):#x} grid {grid[b]} "
              f"ref P {ref['P'][b]} n_trim ref {ref.get('n_trim', [None] * n)[b] if isinstance(ref, dict) else None}")
        np.save(f"gpurun_out/bad_{rank * n + b}_cuda.npy", p64[b])
    print(f"rank {rank}: {len(bad)} frames above 1e-6, max err {err[same].max():.3e}", flush=True)
