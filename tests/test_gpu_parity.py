"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes -> libfsdplan.so), against the golden
vectors of the unmodified reference, against the oracle on seeded synthetic batches, and -- at BASELINE.json's
full sizes -- through size-independent properties (determinism, batch-order and shard independence, rigid-motion
equivariance).  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import oracle  # the checker, never the thing under test
from conftest import GOLDEN_SETS, TIE_DIST_TOL, TIE_KAPPA_TOL, compare_with_golden, load_golden, tie_frame_deviation
from ft_fsd_path_planning_b200 import synth

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from ft_fsd_path_planning_b200 import BatchPlanner, MissionTypes, PathPlanner, _lib  # noqa: E402


@pytest.fixture(scope="module")
def planner():
    return BatchPlanner("cuda:0")


def _np(res):
    return {k: getattr(res, k).cpu().numpy() for k in
            ("path", "left_idx", "right_idx", "status", "path_f64", "n_wv", "left_wv", "right_wv", "l2r", "r2l", "grid")}


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_cuda_matches_reference_P_conditioned(planner, name):
    """Gate (i) of SURVEY 8(d): sort indices / matches exact, path within 1e-4 with the reference's P."""
    batch, g = load_golden(name)
    r = _np(planner.plan_host(batch, force_P=g["P"], intermediates=True))
    compare_with_golden(name, g, r["left_idx"], r["right_idx"], r["n_wv"], r["left_wv"], r["right_wv"], r["l2r"],
                        r["r2l"], r["path_f64"], path_tol=1e-7)
    # the fp32 output tensor of the ABI: tolerance of the north star, 1e-4
    compare_with_golden(name, g, r["left_idx"], r["right_idx"], r["n_wv"], r["left_wv"], r["right_wv"], r["l2r"],
                        r["r2l"], r["path"], path_tol=1e-4)
    ok = g["error"] == 0
    assert (r["grid"][ok, 1] == g["n_trim"][ok]).all()
    assert not (r["status"].astype(np.uint32) & ((1 << 8) | (1 << 10))).any()


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_cuda_tie_normalised(planner, name):
    """Gate (ii): default evaluation-grid rule vs the reference patched to the same rule; (iii) strict: reported."""
    batch, g = load_golden(name)
    r = _np(planner.plan_host(batch, intermediates=True))
    ok = g["error"] == 0
    assert (r["grid"][ok, 0] == g["tie_P"][ok]).all()
    err = np.abs(r["path_f64"] - g["tie_path"]).reshape(len(ok), -1).max(1)
    assert (err[ok] <= 1e-7).all()
    strict = np.abs(r["path_f64"] - g["path"]).reshape(len(ok), -1).max(1) <= 1e-4
    print(f"{name}: strict parity {int(strict[ok].sum())}/{int(ok.sum())} (reference's own P coin flip)")
    # strict gate, made a gate: where the reference's coin flip agrees with the tie rule the default output is within
    # 1e-4; elsewhere it is the same curve sampled one grid step apart -- bounded geometrically
    other = ok & (g["P"] != g["tie_P"])
    assert strict[ok & ~other].all()
    for b in np.where(other)[0]:
        dist, kappa = tie_frame_deviation(r["path_f64"][b], g["path"][b])
        assert dist <= TIE_DIST_TOL and kappa <= TIE_KAPPA_TOL, (name, b, dist, kappa)


@pytest.mark.parametrize("gen,seed,n", [("color", 11, 2048), ("colorless", 12, 2048), ("mixed", 13, 2048)])
def test_cuda_f32_batches_match_oracle(planner, gen, seed, n):
    """fp32 coordinates in HBM (the BASELINE layout); the oracle is fed the same values widened to fp64."""
    batch = synth.gen_mixed(seed, n) if gen == "mixed" else synth.gen_autocross(seed, n)
    if gen == "colorless":
        batch = synth.remove_color_info(batch)
    ref = oracle.plan_batch(batch.astype(np.float64), threads=8)
    r = _np(planner.plan_host(batch, force_P=ref["P"].astype(np.int16), intermediates=True))
    assert (r["left_idx"] == ref["left_idx"]).all() and (r["right_idx"] == ref["right_idx"]).all()
    assert (r["n_wv"][:, 0] == ref["n_left_wv"]).all() and (r["n_wv"][:, 1] == ref["n_right_wv"]).all()
    assert (r["l2r"] == ref["l2r"]).all() and (r["r2l"] == ref["r2l"]).all()
    assert np.abs(r["left_wv"] - ref["left_wv"]).max() <= 1e-9 and np.abs(r["right_wv"] - ref["right_wv"]).max() <= 1e-9
    assert np.abs(r["path_f64"] - ref["path"]).max() <= 1e-7
    assert np.abs(r["path"] - ref["path"]).max() <= 1e-4
    assert ((r["status"].astype(np.uint32) & 0xFFFFFF7F) == (ref["status"] & 0xFFFFFF7F)).all()


def test_augmented_real_frames_match_oracle(planner):
    """SURVEY 8d "augmented-real" variant: the recorded FSG / FS-Spain frames (the golden inputs) under random rigid
    motions plus 2 cm jitter (seed 6), stored as fp32 - FSG-shaped statistics at a few thousand frames, CUDA against the
    oracle (sort indices, matches bit-identical; path 1e-4)."""
    rng = np.random.default_rng(6)
    parts = []
    for name in ("fsg_color", "fss_color", "fsg_colorless"):
        base, _ = load_golden(name)
        frame_of = np.repeat(np.arange(base.n_frames), np.diff(base.offsets))
        for _ in range(3):
            th = rng.uniform(-np.pi, np.pi, base.n_frames)
            tr = rng.uniform(-200, 200, (base.n_frames, 2))
            c, s = np.cos(th), np.sin(th)
            mv = lambda p, f: np.stack([c[f] * p[:, 0] - s[f] * p[:, 1], s[f] * p[:, 0] + c[f] * p[:, 1]], 1) + tr[f]
            xy = mv(base.cones_xy, frame_of) + rng.normal(0, 0.02, base.cones_xy.shape)
            parts.append(synth.FrameBatch(
                xy.astype(np.float32), base.cones_type, base.offsets, mv(base.pos, np.arange(base.n_frames)).astype(np.float32),
                np.stack([c * base.dir[:, 0] - s * base.dir[:, 1], s * base.dir[:, 0] + c * base.dir[:, 1]], 1).astype(np.float32)))
    offs, tot = [np.zeros(1, np.int32)], 0
    for p in parts:
        offs.append((p.offsets[1:] + tot).astype(np.int32))
        tot += p.total_cones
    batch = synth.FrameBatch(np.concatenate([p.cones_xy for p in parts]), np.concatenate([p.cones_type for p in parts]),
                             np.concatenate(offs), np.concatenate([p.pos for p in parts]), np.concatenate([p.dir for p in parts]))
    assert batch.n_frames > 3000
    ref = oracle.plan_batch(batch.astype(np.float64), threads=8)
    r = _np(planner.plan_host(batch, force_P=ref["P"].astype(np.int16), intermediates=True))
    assert (r["left_idx"] == ref["left_idx"]).all() and (r["right_idx"] == ref["right_idx"]).all()
    assert (r["l2r"] == ref["l2r"]).all() and (r["r2l"] == ref["r2l"]).all()
    assert np.abs(r["path_f64"] - ref["path"]).max() <= 1e-7 and np.abs(r["path"] - ref["path"]).max() <= 1e-4
    assert ((r["status"].astype(np.uint32) & 0xFFFFFF7F) == (ref["status"] & 0xFFFFFF7F)).all()


def _oracle_parity(planner, batch, tag, shard_of=None):
    """CUDA (through the C-ABI) against the oracle on the same batch: sort indices, seeds / #configurations / DFS pops
    (sort_dbg), matches and status bit-identical; with-virtual cones 1e-9; path 1e-7 (fp64) / 1e-4 (fp32 output),
    P-conditioned (SURVEY 8d gate i).  Returns the statistics the round report quotes."""
    ref = oracle.plan_batch(batch.astype(np.float64), threads=os.cpu_count() or 8)
    res = planner.plan_host(batch, force_P=ref["P"].astype(np.int16), intermediates=True)
    r = _np(res)
    dbg = res.sort_dbg.cpu().numpy()
    B = batch.n_frames
    where = lambda m: f"{tag}: frames {np.where(m)[0][:8]}" + (f" of shard {shard_of}" if shard_of is not None else "")
    bad = (r["left_idx"] != ref["left_idx"]).any(1) | (r["right_idx"] != ref["right_idx"]).any(1)
    assert not bad.any(), "sort indices differ, " + where(bad)
    bad = (dbg[:, :4] != ref["first_k"].reshape(B, 4)).any(1)
    assert not bad.any(), "seeds differ, " + where(bad)
    bad = (dbg[:, 4:6] != ref["n_configs"]).any(1) | (dbg[:, 6:8] != np.minimum(ref["n_pops"], 32767)).any(1)
    assert not bad.any(), "number of configurations / DFS pops differ, " + where(bad)
    bad = (r["n_wv"][:, 0] != ref["n_left_wv"]) | (r["n_wv"][:, 1] != ref["n_right_wv"]) | \
          (r["l2r"] != ref["l2r"]).any(1) | (r["r2l"] != ref["r2l"]).any(1)
    assert not bad.any(), "matches differ, " + where(bad)
    assert np.abs(r["left_wv"] - ref["left_wv"]).max() <= 1e-9 and np.abs(r["right_wv"] - ref["right_wv"]).max() <= 1e-9
    # A frame on which a static bound of the batched path kernel overflows (more than 32 knots in one spline fit: 1 frame
    # of the 65 536 of config 5) is re-planned by the large-bounds kernel (csrc/kernels_big.cu) and must match as well;
    # only a frame that exceeds even those bounds stays flagged FSD_ST_OVERFLOW (none is expected).
    over = (r["status"].astype(np.uint32) & 0x100) != 0
    assert over.sum() == 0, f"{int(over.sum())} frames overflowed a static bound, " + where(over)
    assert not (r["status"].astype(np.uint32) >> 16).any(), "internal marker bits leaked into out_status"
    bad = ~over & ((r["status"].astype(np.uint32) & 0xFFFFFF7F) != (ref["status"] & 0xFFFFFF7F))
    assert not bad.any(), "status differs, " + where(bad)
    assert (r["grid"][~over, 1] == ref["n_trim"][~over]).all()
    e64 = np.where(over, 0.0, np.abs(r["path_f64"] - ref["path"]).reshape(B, -1).max(1))
    e32 = np.where(over, 0.0, np.abs(r["path"] - ref["path"]).reshape(B, -1).max(1))
    assert e64.max() <= 1e-7, f"fp64 path differs by {e64.max()}, " + where(e64 > 1e-7)
    assert e32.max() <= 1e-4, f"fp32 path differs by {e32.max()}, " + where(e32 > 1e-4)
    return {"frames": B, "path_max_err_f64": float(e64.max()), "path_max_err_f32": float(e32.max()),
            "frames_with_2plus_configs": int((ref["n_configs"] >= 2).any(1).sum()),
            "flagged": int(((ref["status"] & 0x700) != 0).sum()), "overflowed": int(over.sum()), "result": r}


def test_config2_1024_coloured_frames_match_oracle(planner):
    """BASELINE config 2, the exact batch: gen_autocross(seed=2, B=1024), colours known."""
    batch = synth.gen_autocross(2, 1024, workers=min(os.cpu_count() or 1, 16))
    st = _oracle_parity(planner, batch, "config 2")
    print(f"config 2: {st['frames']} frames, fp64 path err {st['path_max_err_f64']:.2e}, fp32 {st['path_max_err_f32']:.2e}, "
          f"{st['frames_with_2plus_configs']} frames decided by the cost function, {st['flagged']} flagged")
    assert st["frames_with_2plus_configs"] > 0


def test_config3_10000_colourless_frames_match_oracle(planner):
    """BASELINE config 3 at full size: remove_color_info(gen_autocross(seed=3, B=10000)) against the oracle, plus the
    properties that need no oracle (determinism, shard independence, well-formed outputs)."""
    B = 10000
    batch = synth.remove_color_info(synth.gen_autocross(3, B, workers=min(os.cpu_count() or 1, 16)))
    st = _oracle_parity(planner, batch, "config 3")
    print(f"config 3: {st['frames']} frames, fp64 path err {st['path_max_err_f64']:.2e}, fp32 {st['path_max_err_f32']:.2e}, "
          f"{st['frames_with_2plus_configs']} frames decided by the cost function, {st['flagged']} flagged")
    a = _np(planner.plan_host(batch, intermediates=True))
    b = _np(planner.plan_host(batch, intermediates=True))
    for k in ("path", "left_idx", "right_idx", "status", "path_f64"):
        assert np.array_equal(a[k], b[k]), f"{k} not deterministic"
    # shard independence: two halves planned separately equal the full batch (multi-GPU partition)
    lo = _np(planner.plan_host(batch.slice(0, B // 2), intermediates=True))
    hi = _np(planner.plan_host(batch.slice(B // 2, B), intermediates=True))
    assert np.array_equal(np.concatenate([lo["path"], hi["path"]]), a["path"])
    assert np.array_equal(np.concatenate([lo["left_idx"], hi["left_idx"]]), a["left_idx"])
    # every frame produced a usable path: u strictly increasing, finite values, |curvature| <= 1
    assert np.isfinite(a["path"]).all()
    assert (np.diff(a["path"][:, :, 0], axis=1) > 0).all()
    assert (np.abs(a["path"][:, :, 3]) <= 1.0 + 1e-6).all()
    # sort indices: valid, unique within a side, -1 padded at the end only
    n = np.diff(batch.offsets)
    for idx in (a["left_idx"], a["right_idx"]):
        valid = idx >= 0
        assert (idx[valid] < np.repeat(n, valid.sum(1))).all()
        assert (np.diff(valid.astype(int), axis=1) <= 0).all()
        s = np.sort(idx, axis=1)
        assert ((np.diff(s, axis=1) != 0) | (s[:, 1:] < 0)).all()


def test_config5_65536_mixed_frames_in_eight_shards_match_oracle(planner):
    """BASELINE config 5, the exact batch: gen_mixed(seed=5, B=65536), block-sharded 8 192 frames per GPU.  On one GPU:
    every shard against the oracle, and the eight shards planned separately are byte-identical to the slices of the
    whole batch (no state crosses frames)."""
    batch = synth.gen_mixed(5, 65536, workers=min(os.cpu_count() or 1, 32))
    B = batch.n_frames
    assert B == 65536
    whole = planner.plan_host(batch)
    a = {k: getattr(whole, k).clone() for k in ("path", "left_idx", "right_idx", "status")}
    tot = {"frames": 0, "frames_with_2plus_configs": 0, "flagged": 0, "overflowed": 0, "e64": 0.0, "e32": 0.0}
    for g in range(8):
        lo, hi = g * 8192, (g + 1) * 8192
        shard = batch.slice(lo, hi)
        st = _oracle_parity(planner, shard, "config 5", shard_of=g)
        for k in ("frames", "frames_with_2plus_configs", "flagged", "overflowed"):
            tot[k] += st[k]
        tot["e64"], tot["e32"] = max(tot["e64"], st["path_max_err_f64"]), max(tot["e32"], st["path_max_err_f32"])
        # default grid rule (no force_P): the shard alone equals its slice of the whole batch
        r = planner.plan_host(shard)
        for k in a:
            assert torch.equal(getattr(r, k), a[k][lo:hi]), f"shard {g}: {k} differs from the whole batch"
    print(f"config 5: {tot['frames']} frames in 8 shards, fp64 path err {tot['e64']:.2e}, fp32 {tot['e32']:.2e}, "
          f"{tot['frames_with_2plus_configs']} frames decided by the cost function, {tot['flagged']} flagged "
          f"(inputs on which the reference raises / takes its latent-bug path), {tot['overflowed']} overflowed a static bound")
    st = a["status"].cpu().numpy().astype(np.uint32)
    assert ((st & 0x100) != 0).sum() == 0, "static bounds overflowed"
    assert ((st & 0x600) != 0).mean() < 0.01


@pytest.mark.parametrize("kind", ["color", "colorless", "mixed"])
def test_knn_stage_adjacency_is_bit_exact(planner, kind):
    """fsd_knn_batch (the cost-matrix step in isolation) against the oracle's create_adjacency_matrix on 4 096 synthetic
    frames per colour mode and on the recorded FSG / FS-Spain frames: degrees and neighbour lists identical."""
    from test_hostcheck import adjacency_equal

    batches = [synth.gen_mixed(33, 4096, workers=16) if kind == "mixed" else synth.gen_autocross(32, 4096, workers=16)]
    if kind == "colorless":
        batches[0] = synth.remove_color_info(batches[0])
    batches.append(load_golden("fsg_" + ("colorless" if kind == "colorless" else "color"))[0])
    batches.append(load_golden("fss_" + ("colorless" if kind == "colorless" else "color"))[0])
    dev = planner.device
    for batch in batches:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        nbr, deg = planner.knn(t(batch.cones_xy), t(batch.cones_type), t(batch.offsets))
        torch.cuda.synchronize()
        ref_nbr, ref_deg = oracle.adjacency(batch)
        assert adjacency_equal(nbr.cpu().numpy(), deg.cpu().numpy(), ref_nbr, ref_deg)


def test_rigid_motion_equivariance(planner):
    """Planning a rotated + translated copy of a frame gives the same sort indices and the transformed path."""
    B = 1024
    batch = synth.gen_autocross(21, B).astype(np.float64)
    rng = np.random.default_rng(0)
    th = rng.uniform(-np.pi, np.pi, B)
    tr = rng.uniform(-50, 50, (B, 2))
    c, s = np.cos(th), np.sin(th)
    frame_of = np.repeat(np.arange(B), np.diff(batch.offsets))

    def move(p, f):
        return np.stack([c[f] * p[:, 0] - s[f] * p[:, 1], s[f] * p[:, 0] + c[f] * p[:, 1]], 1) + tr[f]

    moved = synth.FrameBatch(move(batch.cones_xy, frame_of), batch.cones_type, batch.offsets,
                             move(batch.pos, np.arange(B)),
                             np.stack([c * batch.dir[:, 0] - s * batch.dir[:, 1], s * batch.dir[:, 0] + c * batch.dir[:, 1]], 1))
    a = _np(planner.plan_host(batch, intermediates=True))
    force = torch.from_numpy(a["grid"][:, 0].copy())
    m = _np(planner.plan_host(moved, force_P=a["grid"][:, 0], intermediates=True))
    same = (a["left_idx"] == m["left_idx"]).all(1) & (a["right_idx"] == m["right_idx"]).all(1)
    assert same.mean() > 0.999, f"sort indices changed under a rigid motion in {int((~same).sum())} frames"
    exp_xy = np.stack([c[:, None] * a["path_f64"][:, :, 1] - s[:, None] * a["path_f64"][:, :, 2] + tr[:, None, 0],
                       s[:, None] * a["path_f64"][:, :, 1] + c[:, None] * a["path_f64"][:, :, 2] + tr[:, None, 1]], -1)
    # frames whose previous-path fallback fired are tied to the global origin and are not equivariant
    clean = same & ((a["status"].astype(np.uint32) & 0x7C) == 0) & ((m["status"].astype(np.uint32) & 0x7C) == 0)
    err = np.abs(m["path_f64"][:, :, 1:3] - exp_xy).reshape(B, -1).max(1)
    kerr = np.abs(m["path_f64"][:, :, 3] - a["path_f64"][:, :, 3]).max(1)
    assert clean.mean() > 0.8
    assert np.percentile(err[clean], 99) < 1e-6 and np.percentile(kerr[clean], 99) < 1e-6


def test_stage_entry_points_compose(planner):
    """fsd_sort_batch -> fsd_match_batch -> fsd_path_batch equals fsd_plan_batch."""
    lib, dev = planner.lib, planner.device
    batch = synth.gen_autocross(31, 512)
    B = batch.n_frames
    full = _np(planner.plan_host(batch, intermediates=True))
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev) if dt is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    xy, ty, off, pos, dr = t(batch.cones_xy), t(batch.cones_type), t(batch.offsets), t(batch.pos), t(batch.dir)
    li = torch.empty((B, 12), dtype=torch.int16, device=dev)
    ri = torch.empty_like(li)
    dbg = torch.empty((B, 8), dtype=torch.int16, device=dev)
    st = torch.empty((B,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    p = planner.params
    _lib.check(lib.fsd_sort_batch(C.byref(p), B, xy.data_ptr(), ty.data_ptr(), off.data_ptr(), pos.data_ptr(),
                                  dr.data_ptr(), li.data_ptr(), ri.data_ptr(), dbg.data_ptr(), st.data_ptr(), stream))
    assert np.array_equal(li.cpu().numpy(), full["left_idx"]) and np.array_equal(ri.cpu().numpy(), full["right_idx"])
    bufs = {k: torch.empty_like(getattr(planner.plan_host(batch, intermediates=True), k)) for k in
            ("path_f64", "n_wv", "left_wv", "right_wv", "l2r", "r2l", "grid", "sort_dbg")}
    inter = _lib.Intermediate(*[bufs[k].data_ptr() for k in
                                ("path_f64", "n_wv", "left_wv", "right_wv", "l2r", "r2l", "grid", "sort_dbg")])
    st2 = torch.empty_like(st)
    _lib.check(lib.fsd_match_batch(C.byref(p), B, xy.data_ptr(), off.data_ptr(), pos.data_ptr(), dr.data_ptr(),
                                   li.data_ptr(), ri.data_ptr(), C.byref(inter), st2.data_ptr(), stream))
    assert np.array_equal(bufs["n_wv"].cpu().numpy(), full["n_wv"])
    assert np.array_equal(bufs["l2r"].cpu().numpy(), full["l2r"])
    assert np.array_equal(bufs["left_wv"].cpu().numpy(), full["left_wv"])
    out = torch.empty((B, 40, 4), dtype=torch.float32, device=dev)
    st3 = torch.zeros_like(st)
    prev = planner.initial_path()
    ws = torch.empty((int(lib.fsd_workspace_bytes(B, 0)),), dtype=torch.uint8, device=dev)
    _lib.check(lib.fsd_path_batch(C.byref(p), B, 0, pos.data_ptr(), dr.data_ptr(), C.byref(inter), None,
                                  prev.data_ptr(), 0, out.data_ptr(), st3.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), full["path"])
    assert np.array_equal((st.cpu().numpy() | st2.cpu().numpy() | st3.cpu().numpy()), full["status"])


def test_kernel_event_mode_equals_single_call(planner):
    batch = synth.gen_autocross(41, 256)
    a = _np(planner.plan_host(batch, intermediates=True))
    dev = planner.device
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    res = planner.plan(t(batch.cones_xy), t(batch.cones_type), t(batch.offsets), t(batch.pos), t(batch.dir),
                       kernel_events=True)
    torch.cuda.synchronize()
    times = planner.kernel_times_ms()
    assert len(times) == 1 and times[0][0] > 0 and times[0][1] > 0
    assert np.array_equal(res.path.cpu().numpy(), a["path"]) and np.array_equal(res.status.cpu().numpy(), a["status"])


def test_two_chunk_plan_equals_stage_entry_points(planner):
    """fsd_plan_batch may plan a large batch as two chunks on two streams (plan mode, fsd_plan_launches); the result must
    be bit-identical to the unsplit stage entry points, and the call must stay ordered on the caller's stream."""
    B = 6000
    batch = synth.gen_mixed(42, B)
    dev = planner.device
    assert planner.lib.fsd_plan_launches(256) == 4  # sort, match, path, second chance
    assert planner.lib.fsd_plan_launches(B) in (4, 5, 8, 10)  # + the path keys of a work-ordered path stage; one or two chunks
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    args = (t(batch.cones_xy), t(batch.cones_type), t(batch.offsets), t(batch.pos), t(batch.dir))
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):  # a non-default caller stream; outputs are read on it right after the call
        res = planner.plan(*args, intermediates=True)
        a = {k: getattr(res, k).clone() for k in ("path", "left_idx", "right_idx", "status", "path_f64", "l2r", "r2l")}
    side.synchronize()
    res = planner.plan(*args, kernel_events=True)
    torch.cuda.synchronize()
    planner.kernel_times_ms()
    for k, v in a.items():
        assert torch.equal(v, getattr(res, k)), f"{k} differs between the chunked call and the stage entry points"


def test_two_chunk_plan_offsets_every_per_frame_argument(planner):
    """The second chunk of a large batch must see its own slice of force_P and of a per-frame prev_path: the chunked
    call equals two unsplit calls on the halves."""
    B = 6000
    batch = synth.gen_autocross(44, B)
    rng = np.random.default_rng(44)
    force = np.where(rng.random(B) < 0.5, 120, 121).astype(np.int16)
    prev0 = planner.initial_path().cpu().numpy()
    prev = np.repeat(prev0[None], B, 0)
    prev[:, :, 1] += rng.normal(0, 0.5, (B, 1))  # a different previous path per frame
    whole = _np(planner.plan_host(batch, force_P=force, prev_path=prev, intermediates=True))
    assert (whole["grid"][:, 0] == force).all()
    for lo, hi in ((0, 3000), (3000, 6000)):
        part = _np(planner.plan_host(batch.slice(lo, hi), force_P=force[lo:hi], prev_path=prev[lo:hi], intermediates=True))
        for k in ("path", "path_f64", "left_idx", "right_idx", "status", "grid"):
            assert np.array_equal(part[k], whole[k][lo:hi]), f"{k} differs in frames [{lo}, {hi})"


def test_plan_pinned_equals_plan(planner):
    """The pipelined host-to-host entry point (chunks on streams of their own) gives the same bytes as plan()."""
    B = 7000
    batch = synth.gen_autocross(43, B)
    ref = _np(planner.plan_host(batch, intermediates=True))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h = [pin(a) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir)]
    for chunks, zero_copy in ((None, True), (1, True), (3, True), (None, False), (3, False)):
        # zero_copy: the path kernel stores straight into the pinned host buffer (no device->host copy of the paths)
        out = [torch.zeros((B, 40, 4), dtype=torch.float32).pin_memory(), torch.zeros((B, 12), dtype=torch.int16).pin_memory(),
               torch.zeros((B, 12), dtype=torch.int16).pin_memory(), torch.zeros((B,), dtype=torch.int32).pin_memory()]
        planner.plan_pinned(*h, *out, chunks=chunks, zero_copy=zero_copy)
        torch.cuda.synchronize()
        assert np.array_equal(out[0].numpy(), ref["path"]) and np.array_equal(out[1].numpy(), ref["left_idx"])
        assert np.array_equal(out[2].numpy(), ref["right_idx"]) and np.array_equal(out[3].numpy(), ref["status"])
    with pytest.raises(ValueError):
        planner.plan_pinned(torch.zeros((4, 2)), *h[1:], *out)  # not pinned


def test_fused_gather_stores_every_row_into_every_peer_buffer(planner):
    """fsd_plan_batch_gather / fsd_path_batch_gather: the path kernel stores each frame's path into row first_row + b of
    every peer's gathered buffer (here: three buffers on the one GPU stand in for the peers' -- the stores are the same
    instructions whether the pointer maps local or NVLink peer memory; the multi-process run is bench.py --gpus N)."""
    from ft_fsd_path_planning_b200.distributed import PeerGather

    B, first_row, n_global = 6000, 777, 8000  # large enough for the two-chunk plan: chunk B's rows must follow chunk A's
    batch = synth.gen_autocross(47, B)
    dev = planner.device
    args = tuple(torch.from_numpy(a).to(dev) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir))
    ref = planner.plan(*args)
    ref_path = ref.path.clone()
    peers = [torch.full((n_global, 40, 4), -7.0, dtype=torch.float32, device=dev) for _ in range(3)]
    res = planner.plan(*args, gather=PeerGather.make_descriptor([t.data_ptr() for t in peers], first_row))
    torch.cuda.synchronize()
    assert torch.equal(res.path, ref_path)
    for t in peers:
        assert torch.equal(t[first_row:first_row + B], ref_path)
        assert bool((t[:first_row] == -7.0).all()) and bool((t[first_row + B:] == -7.0).all())
    # the host-to-host entry point with the fused gather: paths to pinned host memory AND to the peers
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h = [pin(a) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir)]
    out = [torch.zeros((B, 40, 4), dtype=torch.float32).pin_memory(), torch.zeros((B, 12), dtype=torch.int16).pin_memory(),
           torch.zeros((B, 12), dtype=torch.int16).pin_memory(), torch.zeros((B,), dtype=torch.int32).pin_memory()]
    for t in peers:
        t.fill_(-7.0)
    planner.plan_pinned(*h, *out, gather=PeerGather.make_descriptor([t.data_ptr() for t in peers], 0))
    torch.cuda.synchronize()
    assert np.array_equal(out[0].numpy(), ref_path.cpu().numpy())
    for t in peers:
        assert torch.equal(t[:B], ref_path)


def test_frame_that_outgrows_its_knot_records_is_resumed_inside_the_path_kernel(planner):
    """Frame 59 383 of the bench stream needs 33 knots in one fit (a frame slot holds 34 records).  A fit that outgrows
    its slot is suspended and resumed at the end of its round with the arena extended over the CTA's shared memory: no
    OVERFLOW flag, result equal to the oracle's, the neighbours in its round untouched."""
    batch = synth.gen_autocross(2, 64, start=59352)
    ref = oracle.plan_batch(batch.astype(np.float64), threads=4)
    r = _np(planner.plan_host(batch, intermediates=True))
    assert not (r["status"].astype(np.uint32) & 0x80000100).any()
    assert np.array_equal(r["left_idx"], ref["left_idx"]) and np.array_equal(r["right_idx"], ref["right_idx"])
    assert np.array_equal(r["status"].astype(np.uint32) & 0xFFFFFF7F, ref["status"].astype(np.uint32) & 0xFFFFFF7F)
    same_P = r["grid"][:, 0] == ref["P"]
    assert same_P[31]
    assert np.abs(r["path_f64"] - ref["path"])[same_P].max() <= 1e-7
    assert np.abs(r["path"].astype(np.float64) - ref["path"])[same_P].max() <= 1e-4


def test_suspend_and_resume_is_invisible_in_the_results():
    """FSD_TEST_CAP=14 makes the fits of ordinary frames outgrow their (artificially small) arena: about every second
    frame is suspended and resumed with the CTA's shared memory, or -- a second one in the same round -- handed to the
    large-bounds kernel.  A child process (the library reads the variable once) must produce the very bytes of the
    default run."""
    import subprocess
    import sys
    import tempfile

    B = 3000
    batch = synth.gen_mixed(31, B)
    ref = _np(BatchPlanner("cuda:0").plan_host(batch, intermediates=True))
    with tempfile.TemporaryDirectory() as tmp:
        code = (f"import sys, numpy as np, torch; sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r});"
                "from ft_fsd_path_planning_b200 import BatchPlanner, synth;"
                f"r = BatchPlanner('cuda:0').plan_host(synth.gen_mixed(31, {B}), intermediates=True); torch.cuda.synchronize();"
                f"np.savez({os.path.join(tmp, 'o.npz')!r}, path=r.path_f64.cpu().numpy(), status=r.status.cpu().numpy(), grid=r.grid.cpu().numpy())")
        env = dict(os.environ, FSD_TEST_CAP="14")
        subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=600)
        z = np.load(os.path.join(tmp, "o.npz"))
        assert np.array_equal(z["status"], ref["status"]) and np.array_equal(z["grid"], ref["grid"])
        assert np.array_equal(z["path"], ref["path_f64"])


def test_filing_frames_under_their_path_keys_in_the_match_kernel_changes_nothing():
    """fsd_plan_batch's matching kernel also computes the path keys and files the frames in their bins' lists, where the
    path kernel looks its rounds up.  FSD_PLAN_MODE bit 8 brings back the separate key / counting-sort kernels and the order
    array: a child process with it must produce the very bytes of the default run, on a batch large enough for the
    work-ordered path stage and on a small one (no ordering)."""
    import subprocess
    import sys
    import tempfile

    for B in (6000, 300):
        batch = synth.gen_mixed(37, B)
        ref = _np(BatchPlanner("cuda:0").plan_host(batch, intermediates=True))
        with tempfile.TemporaryDirectory() as tmp:
            code = (f"import sys, numpy as np, torch; sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r});"
                    "from ft_fsd_path_planning_b200 import BatchPlanner, synth;"
                    f"r = BatchPlanner('cuda:0').plan_host(synth.gen_mixed(37, {B}), intermediates=True); torch.cuda.synchronize();"
                    f"np.savez({os.path.join(tmp, 'o.npz')!r}, path=r.path_f64.cpu().numpy(), status=r.status.cpu().numpy(), "
                    "grid=r.grid.cpu().numpy(), li=r.left_idx.cpu().numpy(), ri=r.right_idx.cpu().numpy(), l2r=r.l2r.cpu().numpy(), "
                    "r2l=r.r2l.cpu().numpy())")
            env = dict(os.environ, FSD_PLAN_MODE=str(29 + 256))
            subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=600)
            z = np.load(os.path.join(tmp, "o.npz"))
            for k, kr in (("status", "status"), ("grid", "grid"), ("li", "left_idx"), ("ri", "right_idx"), ("l2r", "l2r"),
                          ("r2l", "r2l"), ("path", "path_f64")):
                assert np.array_equal(z[k], ref[kr]), (B, k)


def test_fits_that_outgrow_their_slot_by_many_knots():
    """With tiny smoothing parameters the spline fits want far more knots than a frame slot's 34 records: most rounds hold
    suspended frames, the resumed fits run with the arena extended over the CTA's shared memory, and fits that end up with
    >= 39 knots write band rows on top of the NEIGHBOURING slots' headers (point-buffer pointers, arena capacity) -- which
    must be bound afresh before the next round (they were not before r2_zc: illegal memory access on this very input,
    tools/resume_probe.py).  The CUDA path, round after round, must give what the host build of the same sources gives
    (large static bounds, one lane): identical decisions; paths equal up to the summation order of a 32-lane warp, which a
    near-interpolating fit amplifies on a handful of frames.  A child process: an illegal address would poison this one."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "resume_probe.py"), "4096", "1e-5", "1e-6", "json"],
                         check=True, cwd=root, capture_output=True, text=True, timeout=600).stdout
    r = __import__("json").loads(out.strip().splitlines()[-1])
    assert r["status_differs"] == 0 and r["sort_idx_differ"] == 0
    assert r["compared"] >= 0.99 * r["frames"]
    assert r["frames_above_1e-6"] <= 0.005 * r["frames"] and r["path_max_err"] < 5e-3, r


def test_edge_cases(planner):
    """Empty batch, frames without cones, ragged frames, more than FSD_MAX_CONES cones."""
    z = np.zeros((0, 2))
    frames = [([z, z, z, z, z], np.zeros(2), np.array([1.0, 0.0]))]
    cones, pos, direction = synth.gen_autocross_frame(2, 0)
    frames.append((cones, pos, direction))
    frames.append(([z, z, z, z, z], np.ones(2), np.array([0.0, 1.0])))
    rng = np.random.default_rng(1)
    big = [rng.uniform(-60, 60, (300, 2)), z, z, z, z]  # 300 cones > FSD_MAX_CONES
    frames.append((big, np.zeros(2), np.array([1.0, 0.0])))
    batch = synth.pack_frames(frames)
    r = _np(planner.plan_host(batch, intermediates=True))
    ref = oracle.plan_batch(batch.slice(0, 3), force_P=r["grid"][:3, 0])
    assert np.abs(r["path_f64"][:3] - ref["path"]).max() < 1e-7
    assert (r["left_idx"][:3] == ref["left_idx"]).all()
    assert r["status"][0] & _lib.STATUS_BITS["FEW_CONES"] and r["status"][3] & _lib.STATUS_BITS["OVERFLOW"]
    assert np.isfinite(r["path"]).all()
    # B = 0
    empty = synth.pack_frames([])
    res = planner.plan_host(empty)
    assert res.path.shape[0] == 0


def test_previous_path_input(planner):
    """Stateful use: a per-frame previous path feeds the fallbacks (core_calculate_path.py:531-536)."""
    z = np.zeros((0, 2))
    frames = [([z, z, z, z, z], np.array([1.0, 0.5]), np.array([1.0, 0.1])) for _ in range(3)]
    batch = synth.pack_frames(frames)
    init = oracle.initial_path()
    other = init.copy()
    other[:, 2] += 0.05 * other[:, 1]  # a slightly different previous path
    prev = np.stack([init, other, init])
    r = _np(planner.plan_host(batch, prev_path=prev, intermediates=True))
    assert np.array_equal(r["path_f64"][0], r["path_f64"][2])
    assert not np.array_equal(r["path_f64"][0], r["path_f64"][1])
    lib = oracle.lib()
    res = oracle.Result()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    xy, ty = np.zeros((1, 2)), np.zeros(1, np.uint8)
    pos, dr = np.array([1.0, 0.5]), np.array([1.0, 0.1])
    lib.fsd_oracle_plan_frame(dp(xy), ty.ctypes.data_as(C.POINTER(C.c_ubyte)), 0, dp(pos), dp(dr), int(r["grid"][1, 0]),
                              dp(np.ascontiguousarray(other)), C.byref(res))
    assert np.abs(np.ctypeslib.as_array(res.path) - r["path_f64"][1]).max() < 1e-7


def test_drop_in_path_planner_sequential():
    """The facade, called like the reference (one frame after the other, stateful), on the recorded FSG log."""
    batch, g = load_golden("fsg_color")
    pp = PathPlanner(MissionTypes.trackdrive)
    worst = 0.0
    for b in range(0, 40):
        cones, pos, direction = batch.frame(b)
        out = pp.calculate_path_in_global_frame(cones, pos, direction, return_intermediate_results=True)
        path, sl, sr, lwv, rwv, l2r, r2l = out
        assert path.shape == (40, 4) and path.dtype == np.float64
        worst = max(worst, np.abs(path - g["tie_path"][b]).max())
        nl, nr = int(g["n_left_wv"][b]), int(g["n_right_wv"][b])
        assert lwv.shape == (nl, 2) and rwv.shape == (nr, 2)
        assert (l2r == g["l2r"][b][:nl]).all() and (r2l == g["r2l"][b][:nr]).all()
        li = g["left_idx"][b]
        assert np.array_equal(sl, batch.cones_xy[batch.offsets[b]:batch.offsets[b + 1]][li[li >= 0]])
    assert worst < 1e-7
    yaw = float(np.arctan2(direction[1], direction[0]))
    p2 = PathPlanner(MissionTypes.trackdrive).calculate_path_in_global_frame(cones, pos, yaw)
    assert np.abs(p2 - PathPlanner(MissionTypes.trackdrive).calculate_path_in_global_frame(cones, pos, direction / np.linalg.norm(direction))).max() < 1e-6
    with pytest.raises(ValueError, match="direction must be a float or a 2 element array"):
        pp.calculate_path_in_global_frame(cones, pos, [1.0, 0.0, 0.0])


# ---- skidpad mission (BASELINE config 4) --------------------------------------------------------

def _skid_gold():
    import os
    from conftest import GOLDEN_DIR

    return dict(np.load(os.path.join(GOLDEN_DIR, "skidpad.npz")))


def test_skidpad_drop_in_replays_recorded_log():
    """PathPlanner(MissionTypes.skidpad), called frame after frame like the reference, on the recorded skidpad log."""
    g = _skid_gold()
    pp = PathPlanner(MissionTypes.skidpad)
    off = g["log_offsets"]
    worst, compared = 0.0, 0
    for b in range(len(off) - 1):
        xy, ty = g["log_cones_xy"][off[b]:off[b + 1]], g["log_cones_type"][off[b]:off[b + 1]]
        cones = [xy[ty == t] for t in range(5)]
        path = pp.calculate_path_in_global_frame(cones, g["log_pos"][b], g["log_dir"][b])
        assert path.shape == (40, 4)
        assert (pp.relocalization_info is not None) == bool(g["log_relocalized"][b]), b
        if g["log_P"][b] == 120:  # the reference's own grid-size coin flip; 120 is the tie-rule value
            worst = max(worst, np.abs(path - g["log_path"][b]).max())
            compared += 1
    assert compared > 100 and worst < 1e-7
    info = pp.relocalization_info
    assert np.allclose(info.translation, g["log_info"][:2], atol=1e-8) and abs(info.rotation - g["log_info"][2]) < 1e-10
    assert int(pp._index_state.item()) == g["log_index"][-1]


def test_skidpad_batched_config4_matches_oracle_and_reference():
    """16 trajectories x 256 steps = 4096 batched pose steps: K1 for all trajectories in one launch, K2 + MPC tail for
    all steps in two launches; checked against the sequential oracle, and against the reference's golden runs."""
    import os
    from conftest import ROOT
    from ft_fsd_path_planning_b200 import SkidpadBatchPlanner

    sp = SkidpadBatchPlanner("cuda:0")
    dev = sp.device
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    table = np.load(os.path.join(ROOT, "ft_fsd_path_planning_b200", "data", "skidpad_path.npy"))

    def run(T, S, force=None):
        xy, ty, off, pos, dirs = synth.gen_skidpad(4, T, S)
        reloc, nacc = sp.relocalize(t64(xy), t32(off), t64(pos[:, 0]), t64(pos[:, 0]), t64(dirs[:, 0]))
        state = torch.zeros((T,), dtype=torch.int32, device=dev)
        fp = None if force is None else torch.from_numpy(np.ascontiguousarray(force, dtype=np.int16).reshape(-1)).to(dev)
        out = sp.plan(t32(np.arange(T + 1) * S), t64(pos.reshape(-1, 2)), t64(dirs.reshape(-1, 2)), reloc, state,
                      force_P=fp)
        torch.cuda.synchronize()
        return (xy, ty, off, pos, dirs), reloc.cpu().numpy(), {k: v.cpu().numpy() for k, v in out.items() if k[0] != "_"}

    # (a) against the reference's golden trajectories
    g = _skid_gold()
    T, S = int(g["syn_T"]), int(g["syn_S"])
    _, reloc, out = run(T, S, force=g["syn_P"])
    assert (reloc[:, 7] == 1.0).all()
    assert (out["index"].reshape(T, S) == g["syn_index"]).all()
    assert np.abs(out["path_f64"].reshape(T, S, 40, 4) - g["syn_path"]).max() < 1e-7
    # trajectory 3 relocalises onto the mirrored solution: its steps take the previous-path fallback, which only the
    # sequential fix-up pass (skid_fixup_kernel) can reproduce
    assert (out["status"].astype(np.uint32) & 0x20).any() and not (out["status"].astype(np.uint32) & 0x700).any()
    # (b) full config 4 against the oracle
    T, S = 16, 256
    (xy, ty, off, pos, dirs), reloc, out = run(T, S)
    # a trajectory whose first attempt fails would retry with the next frame in a sequential run; those are left to
    # the stateful facade (test_skidpad_drop_in_replays_recorded_log) and skipped here
    assert (reloc[:, 7] == 1.0).sum() >= 12
    paths = out["path_f64"].reshape(T, S, 40, 4)
    grid = out["grid"].reshape(T, S, 2)
    worst = 0.0
    for t in range(0, T, 3):
        if reloc[t, 7] != 1.0:
            continue
        so = oracle.SkidpadOracle(table)
        c = xy[off[t]:off[t + 1]]
        cones = [c[ty[off[t]:off[t + 1]] == k] for k in range(5)]
        for s in range(S):
            p, res = so.step(cones, pos[t, s], dirs[t, s], force_P=int(grid[t, s, 0]))
            worst = max(worst, np.abs(p - paths[t, s]).max())
            assert so.index.value == out["index"].reshape(T, S)[t, s]
    assert worst < 1e-7
    assert np.abs(out["path"] - out["path_f64"]).max() < 1e-3  # fp32 copy; SLAM coordinates reach ~1e3 m
