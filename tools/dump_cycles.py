"""Developer probe (GPU box, probe build -DFSD_FRAME_CYCLES -DFSD_NO_LOCKSTEP): per-frame path-machine time (cycles / 256)
and the matching intermediates of the bench batch -> gpurun_out/frame_cycles.npz (input of the work-predictor study)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

out = {}
for kind, seed, n in (("color", 2, 10240), ("colorless", 3, 10000), ("mixed", 5, 8192)):
    batch = synth.gen_mixed(seed, n, workers=16) if kind == "mixed" else synth.gen_autocross(seed, n, workers=16)
    if kind == "colorless":
        batch = synth.remove_color_info(batch)
    bp = BatchPlanner("cuda:0")
    r = bp.plan_host(batch, intermediates=True)
    torch.cuda.synchronize()
    out[kind + "_cycles"] = r.grid.cpu().numpy()[:, 1].astype(np.int32)
    out[kind + "_n_wv"] = r.n_wv.cpu().numpy()
    out[kind + "_left_wv"] = r.left_wv.cpu().numpy().astype(np.float32)
    out[kind + "_right_wv"] = r.right_wv.cpu().numpy().astype(np.float32)
    out[kind + "_l2r"] = r.l2r.cpu().numpy()
    out[kind + "_r2l"] = r.r2l.cpu().numpy()
    out[kind + "_pos"] = batch.pos
    out[kind + "_dir"] = batch.dir
np.savez_compressed("gpurun_out/frame_cycles.npz", **out)
print({k: v.shape for k, v in out.items() if k.endswith("cycles")})
