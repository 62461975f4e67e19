"""Experiment: how much faster do the kernels run when the 16 warps of a CTA execute identical instruction streams?
(groups of 16 consecutive frames are copies of one frame -> perfect instruction-cache sharing inside a CTA)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth

def timeit(bp, batch, iters=10):
    dev = bp.device
    xy, ty, off = (torch.from_numpy(a).to(dev) for a in (batch.cones_xy, batch.cones_type, batch.offsets))
    pos, dr = torch.from_numpy(batch.pos).to(dev), torch.from_numpy(batch.dir).to(dev)
    for _ in range(3):
        bp.plan(xy, ty, off, pos, dr, kernel_events=True)
    torch.cuda.synchronize(); bp.kernel_times_ms()
    for _ in range(iters):
        bp.plan(xy, ty, off, pos, dr, kernel_events=True)
    torch.cuda.synchronize()
    t = np.array(bp.kernel_times_ms())
    return t.mean(0)

bp = BatchPlanner("cuda:0")
base = synth.gen_autocross(2, 10240)
print("distinct frames      sort/path ms:", timeit(bp, base))
frames = [base.frame(b) for b in range(0, 640)]
rep = synth.pack_frames([f for f in frames for _ in range(16)], dtype=np.float32)
print("16 copies per CTA    sort/path ms:", timeit(bp, rep))
rng = np.random.default_rng(0)
perm = rng.permutation(10240)
shuf = synth.pack_frames([frames[p // 16] for p in perm], dtype=np.float32)
print("same frames shuffled sort/path ms:", timeit(bp, shuf))
