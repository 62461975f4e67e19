"""Developer script (THIS container only: needs /root/reference): the oracle against the unmodified reference on fresh
synthetic frames that are not among the committed goldens - an additional pin of the oracle.

    python tools/oracle_vs_reference.py [frames per kind, default 600]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
sys.path.insert(0, "tests/golden")
import oracle  # noqa: E402
import ref_harness as rh  # noqa: E402
from ft_fsd_path_planning_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
assert rh.reference_available(), "needs /root/reference"
rh.load_reference()


def augmented(name, seed):
    """recorded frames of a golden set under a random rigid motion + 2 cm jitter, stored as fp32"""
    from conftest import load_golden

    base, _ = load_golden(name)
    rng = np.random.default_rng(seed)
    frame_of = np.repeat(np.arange(base.n_frames), np.diff(base.offsets))
    th = rng.uniform(-np.pi, np.pi, base.n_frames)
    tr = rng.uniform(-200, 200, (base.n_frames, 2))
    c, s = np.cos(th), np.sin(th)
    mv = lambda p, f: np.stack([c[f] * p[:, 0] - s[f] * p[:, 1], s[f] * p[:, 0] + c[f] * p[:, 1]], 1) + tr[f]
    xy = mv(base.cones_xy, frame_of) + rng.normal(0, 0.02, base.cones_xy.shape)
    return synth.FrameBatch(xy.astype(np.float32), base.cones_type, base.offsets,
                            mv(base.pos, np.arange(base.n_frames)).astype(np.float32),
                            np.stack([c * base.dir[:, 0] - s * base.dir[:, 1], s * base.dir[:, 0] + c * base.dir[:, 1]], 1).astype(np.float32))


for kind, seed in (("colour", 201), ("colourless", 202), ("mixed", 203), ("fsg augmented", 204), ("fss augmented", 205)):
    if kind.endswith("augmented"):
        b = augmented("fsg_color" if kind.startswith("fsg") else "fss_color", seed)
        n = b.n_frames
    else:
        b = synth.gen_mixed(seed, n) if kind == "mixed" else synth.gen_autocross(seed, n)
    if kind == "colourless":
        b = synth.remove_color_info(b)
    b64 = b.astype(np.float64)
    t0 = time.time()
    ref = rh.run_batch(b64)
    dt = time.time() - t0
    ora = oracle.plan_batch(b64, force_P=ref["P"].astype(np.int16), threads=8)  # P-conditioned
    ok = ref["error"] == 0
    sort_bad = ((ora["left_idx"] != ref["left_idx"]).any(1) | (ora["right_idx"] != ref["right_idx"]).any(1)) & ok
    same_P = ok & (ora["P"] == ref["P"])
    err = np.abs(ora["path"] - ref["path"]).reshape(n, -1).max(1)
    print(f"{kind:14s}: {n} frames, reference {n / dt:.0f} frames/s (1 core); reference raised on {int((~ok).sum())}; "
          f"sort indices differ on {int(sort_bad.sum())}; same grid size P on {int(same_P.sum())}, "
          f"max |path - reference| over those {np.nanmax(np.where(same_P, err, 0)):.2e}", flush=True)
