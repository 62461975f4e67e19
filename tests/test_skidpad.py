"""Skidpad mission (BASELINE config 4): oracle and kernel sources (host-check) against the
sequential reference runs stored in tests/golden/skidpad.npz.  CPU only."""
import os

import numpy as np
import pytest

import hostcheck
import oracle
from conftest import GOLDEN_DIR, ROOT

TABLE = np.load(os.path.join(ROOT, "ft_fsd_path_planning_b200", "data", "skidpad_path.npy"))


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN_DIR, "skidpad.npz")))


def _log_frames(g):
    off = g["log_offsets"]
    for b in range(len(off) - 1):
        xy = g["log_cones_xy"][off[b]:off[b + 1]]
        ty = g["log_cones_type"][off[b]:off[b + 1]]
        yield [xy[ty == t] for t in range(5)], g["log_pos"][b], g["log_dir"][b]


def test_oracle_replays_recorded_skidpad_log(gold):
    so = oracle.SkidpadOracle(TABLE)
    worst = 0.0
    for b, (cones, pos, direction) in enumerate(_log_frames(gold)):
        path, res = so.step(cones, pos, direction, force_P=int(gold["log_P"][b]))
        worst = max(worst, np.abs(path - gold["log_path"][b]).max())
        assert bool(so.reloc.relocalized) == bool(gold["log_relocalized"][b]), b
        assert so.index.value == gold["log_index"][b], b
    assert worst < 1e-8
    # RelocalizationInformation (relocalization_information.py:13-35)
    from ft_fsd_path_planning_b200.skidpad import to_known_frame

    r8 = np.array([*so.reloc.translation, so.reloc.rotation, *so.reloc.right_ref, *so.reloc.right_calc, 1.0])
    o, e = to_known_frame(r8, np.zeros(2)), to_known_frame(r8, np.array([1.0, 0.0]))
    assert np.allclose(o, gold["log_info"][:2], atol=1e-8)
    assert abs(np.arctan2(e[1] - o[1], e[0] - o[0]) - gold["log_info"][2]) < 1e-10


def test_oracle_synthetic_trajectories(gold):
    from ft_fsd_path_planning_b200 import synth

    T, S = int(gold["syn_T"]), int(gold["syn_S"])
    xy, ty, off, pos, dirs = synth.gen_skidpad(4, T, S)
    for t in range(T):
        so = oracle.SkidpadOracle(TABLE)
        c = xy[off[t]:off[t + 1]]
        cones = [c[ty[off[t]:off[t + 1]] == k] for k in range(5)]
        for s in range(S):
            path, res = so.step(cones, pos[t, s], dirs[t, s], force_P=int(gold["syn_P"][t, s]))
            assert np.abs(path - gold["syn_path"][t, s]).max() < 1e-8, (t, s)
            assert so.index.value == gold["syn_index"][t, s]


def test_kernel_sources_replay_skidpad_log(gold):
    """hostcheck = the kernels' own sources with a one-lane warp; stateful replay like PathPlanner(skidpad)."""
    path_tab, ref, jitter = oracle.skidpad_constants(TABLE)
    reloc8, orig, state, prev = np.zeros(8), None, 0, hostcheck.initial_path()
    worst = 0.0
    for b, (cones, pos, direction) in enumerate(_log_frames(gold)):
        if reloc8[7] == 0:
            if orig is None:
                orig = (pos.copy(), direction.copy())
            reloc8, _ = hostcheck.skidpad_relocalize(np.concatenate([c.reshape(-1, 2) for c in cones]), pos, orig[0],
                                                     orig[1], jitter, ref)
        h = hostcheck.skidpad_steps(reloc8, path_tab, pos[None], direction[None], state, force_P=[gold["log_P"][b]],
                                    prev=prev)
        state, prev = h["state"], h["internal"][0]
        worst = max(worst, np.abs(h["path"][0] - gold["log_path"][b]).max())
        assert bool(reloc8[7]) == bool(gold["log_relocalized"][b]), b
        assert state == gold["log_index"][b] or not gold["log_relocalized"][b]
    assert worst < 1e-7


def test_kernel_sources_batched_steps_equal_sequential(gold):
    """All steps of a trajectory in ONE call (prev = initial path for every step) equal the sequential reference as
    long as no previous-path fallback fires (it never does on these trajectories)."""
    from ft_fsd_path_planning_b200 import synth

    T, S = int(gold["syn_T"]), int(gold["syn_S"])
    xy, ty, off, pos, dirs = synth.gen_skidpad(4, T, S)
    path_tab, ref, jitter = oracle.skidpad_constants(TABLE)
    for t in range(2):
        reloc8, nacc = hostcheck.skidpad_relocalize(xy[off[t]:off[t + 1]], pos[t, 0], pos[t, 0], dirs[t, 0], jitter, ref)
        assert reloc8[7] == 1.0 and nacc >= 3
        h = hostcheck.skidpad_steps(reloc8, path_tab, pos[t], dirs[t], 0, force_P=gold["syn_P"][t])
        assert (h["index"] == gold["syn_index"][t]).all()
        assert np.abs(h["path"] - gold["syn_path"][t]).max() < 1e-7
        assert not (h["status"] & 0x7FC).any()
