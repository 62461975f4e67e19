"""Small target for compute-sanitizer: the planner on a few hundred frames through every entry point touched in round 2
(plan, plan with fused gather into local 'peer' buffers, plan_pinned with zero-copy output)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402
from ft_fsd_path_planning_b200.distributed import PeerGather  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
batch = synth.concat_batches([synth.gen_mixed(5, n), synth.gen_autocross(2, 16, start=59376)])
B = batch.n_frames
bp = BatchPlanner("cuda:0")
dev = bp.device
args = tuple(torch.from_numpy(a).to(dev) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir))
ref = bp.plan(*args, intermediates=True)
peers = [torch.zeros((B + 8, 40, 4), dtype=torch.float32, device=dev) for _ in range(2)]
res = bp.plan(*args, gather=PeerGather.make_descriptor([t.data_ptr() for t in peers], 8))
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = [pin(a) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir)]
out = [torch.zeros((B, 40, 4), dtype=torch.float32).pin_memory(), torch.zeros((B, 12), dtype=torch.int16).pin_memory(),
       torch.zeros((B, 12), dtype=torch.int16).pin_memory(), torch.zeros((B,), dtype=torch.int32).pin_memory()]
bp.plan_pinned(*h, *out, chunks=2)
torch.cuda.synchronize()
ok = torch.equal(res.path, ref.path) and torch.equal(peers[0][8:8 + B], ref.path) and torch.equal(out[0], ref.path.cpu())
print("outputs identical:", ok, "frames", B, "flagged", int(((ref.status.cpu().numpy().astype(np.uint32) & 0x80000100) != 0).sum()))
