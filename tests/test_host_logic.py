"""Host-side logic: packed frame format, synthetic generator, sharding arithmetic, facade argument handling."""
import numpy as np
import pytest

from ft_fsd_path_planning_b200 import synth
from ft_fsd_path_planning_b200.distributed import shard_bounds
from ft_fsd_path_planning_b200.enums import ConeTypes, MissionTypes


def test_enums_match_reference_values():
    assert [int(c) for c in ConeTypes] == [0, 1, 2, 3, 4]
    assert ConeTypes.YELLOW == ConeTypes.RIGHT == 1 and ConeTypes.BLUE == ConeTypes.LEFT == 2
    assert MissionTypes.trackdrive == 4 and MissionTypes.autocross == 3 and MissionTypes.skidpad == 2


def test_generator_is_a_pure_function_of_seed_and_index():
    a = synth.gen_autocross(2, 8)
    b = synth.gen_autocross(2, 4, start=4)
    c = a.slice(4, 8)
    assert np.array_equal(b.cones_xy, c.cones_xy) and np.array_equal(b.offsets, c.offsets)
    assert np.array_equal(b.pos, c.pos) and np.array_equal(b.cones_type, c.cones_type)
    n = np.diff(a.offsets)
    assert n.min() >= 40 and n.max() <= synth.MAX_CONES_PER_FRAME
    assert a.cones_xy.dtype == np.float32


def test_pack_and_unpack_round_trip():
    frames = [synth.gen_autocross_frame(3, i) for i in range(5)]
    batch = synth.pack_frames(frames)
    for i, (cones, pos, direction) in enumerate(frames):
        c2, p2, d2 = batch.frame(i)
        for t in range(5):
            assert np.array_equal(np.asarray(cones[t]).reshape(-1, 2), c2[t])
        assert np.array_equal(pos, p2) and np.array_equal(direction, d2)
    # cones are stored in ConeTypes order inside a frame
    lo, hi = batch.offsets[0], batch.offsets[1]
    assert (np.diff(batch.cones_type[lo:hi].astype(int)) >= 0).all()


def test_remove_color_and_mixed():
    b = synth.gen_autocross(5, 6)
    u = synth.remove_color_info(b)
    assert (u.cones_type == 0).all() and np.array_equal(u.cones_xy, b.cones_xy)
    m = synth.gen_mixed(5, 6)
    assert np.array_equal(np.diff(m.offsets), np.diff(b.offsets))
    assert (m.slice(1, 2).cones_type == b.slice(1, 2).cones_type).all()  # odd frames keep their colours
    assert (m.slice(0, 1).cones_type == 0).any()


def test_algorithmic_bytes_formula():
    b = synth.gen_autocross(2, 3)
    assert b.algorithmic_bytes() == 9 * b.total_cones + 708 * 3


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (65536, 8), (5, 8), (0, 2)])
def test_shard_bounds_partition(n, world):
    covered = []
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        assert 0 <= lo <= hi <= n
        covered.extend(range(lo, hi))
    assert covered == list(range(n))


def test_direction_conversion_like_reference():
    from ft_fsd_path_planning_b200.planner import PathPlanner

    conv = PathPlanner._convert_direction_to_array
    assert np.allclose(conv(np.pi / 2), [0.0, 1.0])
    assert np.allclose(conv([0.0, 2.0]), [0.0, 2.0])
    assert np.allclose(conv(np.array([[0.3]])), [np.cos(0.3), np.sin(0.3)])
    with pytest.raises(ValueError, match="direction must be a float or a 2 element array"):
        conv([1.0, 2.0, 3.0])


def test_json_log_round_trip(tmp_path):
    from ft_fsd_path_planning_b200.io import load_data_json, save_data_json

    batch = synth.gen_autocross(9, 4).astype(np.float64)
    save_data_json(batch, tmp_path / "log.json")
    again = load_data_json(tmp_path / "log.json")
    assert np.array_equal(again.cones_xy, batch.cones_xy) and np.array_equal(again.offsets, batch.offsets)
    assert np.array_equal(again.cones_type, batch.cones_type) and np.array_equal(again.pos, batch.pos)
    unknown = load_data_json(tmp_path / "log.json", remove_color_info=True)
    assert (unknown.cones_type == 0).all() and np.array_equal(unknown.cones_xy, batch.cones_xy)


def test_json_loader_reads_reference_log_format():
    """The goldens were packed from the reference's own logs; loading the same file through io.py gives the same batch
    (only where the reference tree is present)."""
    import os

    from conftest import load_golden
    from ft_fsd_path_planning_b200.io import load_data_json

    path = "/root/reference/fsd_path_planning/demo/fsg_19_2_laps.json"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    batch, _ = load_golden("fsg_color")
    loaded = load_data_json(path)
    assert np.array_equal(loaded.cones_xy, batch.cones_xy) and np.array_equal(loaded.offsets, batch.offsets)


def test_host_planner_is_explicit_and_matches_goldens():
    """fsd_plan_batch_cpu (the kernels' per-frame sources compiled for the host, BASELINE config 1 "CPU plumbing"): all 440
    recorded FSG frames against the reference's goldens, through the batched entry point and through the drop-in facade
    with device="cpu".  The default device stays CUDA-only: nothing falls back to the host planner."""
    import pytest
    import torch

    from conftest import compare_with_golden, load_golden
    from ft_fsd_path_planning_b200 import BatchPlanner, CpuBatchPlanner, MissionTypes, PathPlanner

    batch, g = load_golden("fsg_color")
    r = CpuBatchPlanner(threads=4).plan_host(batch, force_P=g["P"], intermediates=True)
    n = lambda t: t.numpy()
    compare_with_golden("fsg_color", g, n(r.left_idx), n(r.right_idx), n(r.n_wv), n(r.left_wv), n(r.right_wv), n(r.l2r),
                        n(r.r2l), n(r.path_f64), path_tol=1e-7)
    compare_with_golden("fsg_color", g, n(r.left_idx), n(r.right_idx), n(r.n_wv), n(r.left_wv), n(r.right_wv), n(r.l2r),
                        n(r.r2l), n(r.path), path_tol=1e-4)
    pp = PathPlanner(MissionTypes.trackdrive, device="cpu")
    worst = 0.0
    for b in range(0, 440, 11):
        cones, pos, direction = batch.frame(b)
        path, sl, sr, lwv, rwv, l2r, r2l = pp.calculate_path_in_global_frame(cones, pos, direction, return_intermediate_results=True)
        worst = max(worst, float(np.abs(path - g["tie_path"][b]).max()))
        assert (l2r == g["l2r"][b][: len(l2r)]).all()
    assert worst < 1e-7
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="needs a CUDA device"):
            BatchPlanner("cuda")
        with pytest.raises(RuntimeError):
            PathPlanner(MissionTypes.trackdrive)  # the default device is CUDA: no silent fallback to the host planner


def test_peer_gather_descriptor_layout():
    """struct fsd_gather as the kernels see it: peer pointers, first row, multicast address (no GPU needed)."""
    import ctypes as C

    from ft_fsd_path_planning_b200 import _lib
    from ft_fsd_path_planning_b200.distributed import PeerGather

    d = PeerGather.make_descriptor([0x1000, 0x2000, 0x3000], 4096)
    assert (d.n_peers, d.first_row) == (3, 4096) and [d.peer_out_path[i] for i in range(3)] == [0x1000, 0x2000, 0x3000]
    assert d.multicast_out_path is None and d.peer_out_path[3] is None
    m = PeerGather.make_descriptor([0x1000, 0x2000], 0, multicast_ptr=0x9000)
    assert m.n_peers == 0 and m.multicast_out_path == 0x9000  # one multimem store instead of a store per peer
    assert C.sizeof(_lib.Gather) == 4 + 4 + 8 + 8 * _lib.MAX_PEERS + 8
    with pytest.raises(ValueError):
        PeerGather.make_descriptor(list(range(1, 18)), 0)
