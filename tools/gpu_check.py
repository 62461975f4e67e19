"""Developer script (GPU box): plan a synthetic batch with the CUDA path, compare with the oracle, time it."""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import oracle  # noqa: E402  (checker only)
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--seed", type=int, default=2)
ap.add_argument("--f64", action="store_true")
ap.add_argument("--colorless", action="store_true")
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()

batch = synth.gen_autocross(args.seed, args.frames)
if args.colorless:
    batch = synth.remove_color_info(batch)
if args.f64:
    batch = batch.astype(np.float64)
B = batch.n_frames
t0 = time.time()
ref = oracle.plan_batch(batch.astype(np.float64), threads=16)
print(f"oracle: {B / (time.time() - t0):.0f} frames/s (16 threads)")
bp = BatchPlanner("cuda:0")
res = bp.plan_host(batch, force_P=ref["P"].astype(np.int16), intermediates=True)
torch.cuda.synchronize()
li, ri = res.left_idx.cpu().numpy(), res.right_idx.cpu().numpy()
sort_bad = np.where((li != ref["left_idx"]).any(1) | (ri != ref["right_idx"]).any(1))[0]
nwv = res.n_wv.cpu().numpy()
nwv_bad = np.where((nwv[:, 0] != ref["n_left_wv"]) | (nwv[:, 1] != ref["n_right_wv"]))[0]
m_bad = np.where((res.l2r.cpu().numpy() != ref["l2r"]).any(1) | (res.r2l.cpu().numpy() != ref["r2l"]).any(1))[0]
p64 = res.path_f64.cpu().numpy()
perr = np.abs(p64 - ref["path"]).reshape(B, -1).max(1)
perr32 = np.abs(res.path.cpu().numpy() - ref["path"]).reshape(B, -1).max(1)
st = res.status.cpu().numpy().astype(np.uint32)
grid = res.grid.cpu().numpy()
print(f"B={B} sort_bad={len(sort_bad)} nwv_bad={len(nwv_bad)} match_bad={len(m_bad)} "
      f"path_f64_bad(>1e-6)={(~(perr <= 1e-6)).sum()} max={np.nanmax(perr):.3e} path_f32_bad(>1e-4)={(~(perr32 <= 1e-4)).sum()} "
      f"max32={np.nanmax(perr32):.3e} status_or={np.bitwise_or.reduce(st):#x} "
      f"status_mismatch={((st & 0xffffff7f) != (ref["status"] & 0xffffff7f)).sum()} grid_bad={(grid[:, 0] != ref['P']).sum()},{(grid[:, 1] != ref['n_trim']).sum()}")
for b in sort_bad[:5]:
    print(" sort", b, li[b], ref["left_idx"][b], ri[b], ref["right_idx"][b], res.sort_dbg[b].cpu().numpy(), ref["first_k"][b].ravel(), ref["n_configs"][b], ref["n_pops"][b])
for b in np.where(~(perr <= 1e-6))[0][:5]:
    print(" path", b, perr[b], hex(st[b]), hex(ref["status"][b]), grid[b], ref["P"][b], ref["n_trim"][b])
# timing: device-resident inputs
dev = bp.device
dt = torch.float64 if args.f64 else torch.float32
xy = torch.from_numpy(batch.cones_xy).to(dev, dt)
ty = torch.from_numpy(batch.cones_type).to(dev)
off = torch.from_numpy(batch.offsets).to(dev)
pos = torch.from_numpy(batch.pos).to(dev, dt)
dr = torch.from_numpy(batch.dir).to(dev, dt)
for _ in range(3):
    bp.plan(xy, ty, off, pos, dr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    bp.plan(xy, ty, off, pos, dr)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
print(f"GPU: {ms:.3f} ms per batch of {B} -> {B / ms * 1e3:.0f} frames/s")
