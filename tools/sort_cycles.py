"""Developer probe (GPU box, probe build -DFSD_FRAME_CYCLES): per-frame time of the sort kernel (cycles / 64, in sort_dbg[7])
for one rank's shard of the bench stream -> which frames make the kernel's tail."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

n = 10240
for rank in [int(a) for a in sys.argv[1:]] or [0, 7]:
    batch = synth.gen_autocross(2, n, start=rank * n, workers=16)
    bp = BatchPlanner("cuda:0")
    r = bp.plan_host(batch, intermediates=True)
    torch.cuda.synchronize()
    dbg = r.sort_dbg.cpu().numpy()
    cyc = dbg[:, 7].astype(np.int64) * 64
    top = np.argsort(-cyc)[:8]
    print(f"rank {rank}: sort cycles mean {cyc.mean():.0f} p50 {np.percentile(cyc, 50):.0f} p99 {np.percentile(cyc, 99):.0f} max {cyc.max()}")
    for b in top:
        print(f"   frame {b}: {cyc[b]} cycles, cones {int(batch.offsets[b + 1] - batch.offsets[b])}, first_k {dbg[b, :4].tolist()} n_configs {dbg[b, 4:6].tolist()} pops_left {dbg[b, 6]}, "
              f"sorted {int((r.left_idx[b] >= 0).sum())}/{int((r.right_idx[b] >= 0).sum())} status {int(r.status[b]):#x}")
    np.save(f"gpurun_out/sort_cycles_rank{rank}.npy", cyc)
