"""Developer probe (GPU box): fits that outgrow their frame slot by MANY knots.

With small smoothing parameters the spline fits want far more knots than a frame slot's 34 records: many frames are
suspended and resumed with the arena extended over the CTA's shared memory, and fits that end up with >= 39 knots write band
rows on top of the NEIGHBOURING slots' headers (point-buffer pointers, arena capacity).  The CUDA path must still give what
the host build of the same sources gives (fsd_plan_batch_cpu, large static bounds), round after round.

    python tools/resume_probe.py [frames] [smoothing] [refit_smoothing]        (FSD_LIBFSDPLAN selects the library)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, CpuBatchPlanner, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s1 = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
s2 = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0001
batch = synth.gen_autocross(2, n, workers=16)
gp, cp = BatchPlanner("cuda:0"), CpuBatchPlanner(threads=16)
for p in (gp, cp):
    p.params.smoothing = s1
    p.params.refit_smoothing = s2
ref = cp.plan_host(batch.astype(np.float64), intermediates=True)
res = gp.plan_host(batch, intermediates=True)
torch.cuda.synchronize()
g = lambda t: t.cpu().numpy()
st_g, st_c = g(res.status).astype(np.uint32), g(ref.status).astype(np.uint32)
over_g, over_c = (st_g & 0x100) != 0, (st_c & 0x100) != 0
same = (g(res.grid)[:, 0] == g(ref.grid)[:, 0]) & ~over_g & ~over_c
err = np.abs(g(res.path_f64) - g(ref.path_f64)).reshape(n, -1).max(1)
if len(sys.argv) > 4 and sys.argv[4] == "json":  # tests/test_gpu_parity.py
    import json

    print(json.dumps({"frames": n, "compared": int(same.sum()), "path_max_err": float(err[same].max()),
                      "frames_above_1e-6": int((err[same] > 1e-6).sum()),
                      "status_differs": int(((st_g & 0xFFFFFE7F) != (st_c & 0xFFFFFE7F)).sum()),
                      "sort_idx_differ": int((g(res.left_idx) != g(ref.left_idx)).any(1).sum()),
                      "overflow_gpu": int(over_g.sum()), "overflow_cpu": int(over_c.sum())}))
    sys.exit(0)
print(f"frames {n} smoothing {s1} refit {s2}: overflow-flagged gpu {int(over_g.sum())} cpu {int(over_c.sum())}, "
      f"status differs on {int(((st_g & 0xFFFFFE7F) != (st_c & 0xFFFFFE7F)).sum())}, compared {int(same.sum())}, "
      f"path max err {float(err[same].max()):.3e}, frames above 1e-6: {int((err[same] > 1e-6).sum())}, "
      f"sort idx differ {int((g(res.left_idx) != g(ref.left_idx)).any(1).sum())}")
