"""Developer probe (GPU box): per-kernel times (stage entry points, whole batch) for a coloured and a colourless batch."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ft_fsd_path_planning_b200 import BatchPlanner, synth  # noqa: E402

dev = torch.device("cuda:0")
bp = BatchPlanner(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, b in (("coloured 10240", synth.gen_autocross(2, 10240)), ("colourless 4096", synth.remove_color_info(synth.gen_autocross(3, 4096))),
                ("colourless 10000", synth.remove_color_info(synth.gen_autocross(3, 10000)))):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (t(b.cones_xy), t(b.cones_type), t(b.offsets), t(b.pos), t(b.dir))
    for _ in range(3):
        bp.plan(*args, kernel_events=True)
    torch.cuda.synchronize()
    bp.kernel_times_ms()
    for _ in range(10):
        flush.zero_()
        bp.plan(*args, kernel_events=True)
    torch.cuda.synchronize()
    kt = np.array(bp.kernel_times_ms())
    print(f"{name:18s}: sort_match {np.median(kt[:, 0]):.3f} ms, path {np.median(kt[:, 1]):.3f} ms")
