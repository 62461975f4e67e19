"""Developer script (GPU box): device-resident throughput of the other BASELINE configs (the bench line is config 1/2's
frame shape at the metric's 10k batch): config 2 (1 024 coloured), config 3 (10 000 colourless), config 4 (skidpad, 16
trajectories x 256 steps), config 5's per-GPU shard (8 192 mixed).  L2 flushed between iterations."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, SkidpadBatchPlanner, synth  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


bp = BatchPlanner(dev)
for name, batch in (("config 2: 1 024 coloured frames", synth.gen_autocross(2, 1024)),
                    ("config 3: 10 000 colourless frames", synth.remove_color_info(synth.gen_autocross(3, 10000))),
                    ("config 5 shard: 8 192 mixed frames", synth.gen_mixed(5, 8192))):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (t(batch.cones_xy), t(batch.cones_type), t(batch.offsets), t(batch.pos), t(batch.dir))
    ms = timed(lambda: bp.plan(*args))
    print(f"{name}: {ms:.3f} ms -> {batch.n_frames / ms * 1e3:,.0f} frames/s  ({batch.total_cones / batch.n_frames:.1f} cones/frame)")

sp = SkidpadBatchPlanner(dev)
T, S = 16, 256
xy, ty, off, pos, dirs = synth.gen_skidpad(4, T, S)
t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
a_xy, a_off, a_p0, a_d0 = t64(xy), t32(off), t64(pos[:, 0]), t64(dirs[:, 0])
a_so, a_pos, a_dir = t32(np.arange(T + 1) * S), t64(pos.reshape(-1, 2)), t64(dirs.reshape(-1, 2))
ms_r = timed(lambda: sp.relocalize(a_xy, a_off, a_p0, a_p0, a_d0))
reloc, _ = sp.relocalize(a_xy, a_off, a_p0, a_p0, a_d0)


def steps():
    state = torch.zeros((T,), dtype=torch.int32, device=dev)
    sp.plan(a_so, a_pos, a_dir, reloc, state)


ms_p = timed(steps)
print(f"config 4: skidpad, {T} trajectories x {S} steps: relocalization {ms_r:.3f} ms ({T} trajectories), "
      f"{T * S} pose steps {ms_p:.3f} ms -> {T * S / ms_p * 1e3:,.0f} steps/s")
