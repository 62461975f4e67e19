"""Generate the committed golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_goldens.py            # writes tests/golden/*.npz

Each .npz holds the packed fp64 inputs of a frame set and, per frame, the reference's outputs
(ref_harness.run_batch): sort indices, with-virtual cone lists, matches, the (40, 4) path, the size P of the
last evaluation grid and the number of points entering the last re-fit -- once for the unmodified reference
("strict") and once with the tie-normalised evaluation-grid rule of SURVEY.md 8(d)(ii) ("tie_path", "tie_P").
fitpack.npz holds scipy.interpolate.splprep(full_output=1) results for spline fits the reference issues.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_harness as rh  # noqa: E402
from ft_fsd_path_planning_b200 import synth  # noqa: E402


def fixtures():
    """Hand-made tracks of the reference's demo (fsd_path_planning/demo/streamlit_demo/common.py:72-130),
    restated: "Straight", "Simple Corner", "Corner Missing Blue", "Corner Missing Blue Alt"."""
    out = []
    pos, direction = np.zeros(2), np.array([1.0, 0.0])
    rng = np.random.default_rng(0)
    xs = np.arange(-2, 20, 4)
    left = np.column_stack((xs, np.ones(len(xs)) * 1.5 + rng.uniform(-0.3, 0.3, len(xs))))
    right = np.column_stack((xs, np.ones(len(xs)) * -1.5 + rng.uniform(-0.3, 0.3, len(xs))))
    out.append(("straight", left, right))

    def unit(a):
        return np.stack([np.cos(a), np.sin(a)], -1)

    def rot(p, th):
        c, s = np.cos(th), np.sin(th)
        return p @ np.array(((c, -s), (s, c))).T

    inner = unit(np.arange(0, np.pi / 2, np.pi / 15)) * 9
    outer = unit(np.arange(0, np.pi / 2, np.pi / 20)) * 12
    centre = np.mean((inner[:2] + outer[:2]) / 2, axis=0)
    left, right = rot(inner - centre, -np.pi / 2), rot(outer - centre, -np.pi / 2)
    out.append(("simple_corner", left, right))
    keep = np.ones(len(left), bool)
    keep[3:7] = False
    out.append(("corner_missing_blue", left[keep], right))
    keep = np.ones(len(left), bool)
    keep[1:4] = False
    out.append(("corner_missing_blue_alt", left[keep], right))
    frames = []
    for _, l, r in out:
        for colour in (True, False):
            if colour:
                cones = [np.zeros((0, 2)), r, l, np.zeros((0, 2)), np.zeros((0, 2))]
            else:
                cones = [np.concatenate([r, l]), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 2))]
            frames.append((cones, pos, direction))
    return frames


def edge_frames():
    """Degenerate inputs: empty frame, very few cones, one colour only, car far from every cone."""
    frames = []
    z = np.zeros((0, 2))
    frames.append(([z, z, z, z, z], np.zeros(2), np.array([1.0, 0.0])))
    frames.append(([z, z, z, z, z], np.array([3.0, -2.0]), np.array([0.0, 2.0])))
    for i in range(24):
        cones, pos, direction = synth.gen_autocross_frame(7, i)
        xy = np.concatenate(cones, 0)
        ty = np.concatenate([np.full(len(c), t) for t, c in enumerate(cones)])
        order = np.argsort(np.linalg.norm(xy - pos, axis=1), kind="stable")
        if i < 14:
            sel = np.sort(order[:i])  # the i nearest cones
            xy2, ty2 = xy[sel], ty[sel]
        elif i < 17:
            xy2, ty2 = xy[ty != synth.YELLOW], ty[ty != synth.YELLOW]
        elif i < 20:
            xy2, ty2 = xy[ty != synth.BLUE], ty[ty != synth.BLUE]
        elif i < 22:
            xy2, ty2 = xy, ty
            pos = pos + np.array([9.0, 7.0])  # off the track
        else:
            xy2, ty2 = xy, ty
            direction = -direction  # driving the wrong way
        frames.append(([xy2[ty2 == t] for t in range(5)], pos, direction))
    return frames


def save(name, batch, extra=None):
    print(f"[{name}] {batch.n_frames} frames", flush=True)
    g = rh.run_batch(batch)
    gt = rh.run_batch(batch, tie_rule=True)
    out = {
        "cones_xy": batch.cones_xy.astype(np.float64), "cones_type": batch.cones_type, "offsets": batch.offsets,
        "pos": batch.pos.astype(np.float64), "dir": batch.dir.astype(np.float64),
        "tie_path": gt["path"], "tie_P": gt["P"],
    }
    for k, v in g.items():
        if k != "error_text":
            out[k] = v
    if extra:
        out.update(extra)
    print(f"   reference errors: {int(g['error'].sum())} {list(g['error_text'][:3])}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def fitpack_goldens():
    """splprep(full_output=1) on the fits the reference issues for every 8th FSG frame + 60 synthetic frames."""
    from scipy.interpolate import splev, splprep

    rh.load_reference()
    from fsd_path_planning.utils import spline_fit as sf

    calls = []
    orig = sf.SplineFitterFactory.fit

    def hook(self, trace, periodic=False):
        calls.append((np.array(trace, float).copy(), float(self.smoothing)))
        return orig(self, trace, periodic)

    sf.SplineFitterFactory.fit = hook
    for f in rh.load_demo_log("fsg_19_2_laps.json")[::8]:
        rh.run_frame(*f)
    b = synth.gen_autocross(2, 60).astype(np.float64)
    for i in range(b.n_frames):
        rh.run_frame(*b.frame(i))
    sf.SplineFitterFactory.fit = orig
    calls = [c for c in calls if not (len(c[0]) == 40 and c[1] == 0.2)][:400]
    pts, meta, knots, coefs, evals = [], [], [], [], []
    for p, s in calls:
        m = len(p)
        k = int(np.clip(m - 1, 1, 3))
        u = np.concatenate(([0.0], np.cumsum(np.linalg.norm(np.diff(p, axis=0), axis=1))))
        (tck, _), fp, ier, _ = splprep(p.T, s=s, k=k, u=u, full_output=1)
        t, c = tck[0], np.array(tck[1])
        ue = np.linspace(-0.5, u[-1] + 0.5, 16)
        ev = np.array(splev(ue, tck)).T
        meta.append([m, k, s, fp, ier, len(t), len(pts)])
        pts.extend(p.tolist())
        knots.append(np.pad(t, (0, 64 - len(t))))
        coefs.append(np.pad(c, ((0, 0), (0, 64 - c.shape[1]))))
        evals.append(np.concatenate([ue[:, None], ev], 1))
    np.savez_compressed(os.path.join(HERE, "fitpack.npz"), points=np.array(pts), meta=np.array(meta),
                        knots=np.array(knots), coefs=np.array(coefs), evals=np.array(evals))
    print(f"[fitpack] {len(calls)} fits")


def skidpad_goldens():
    """Sequential (stateful) reference runs of MissionTypes.skidpad: the recorded log and 6 synthetic trajectories."""
    mod = rh.load_reference()
    rec = rh._state["rec"]

    def run(frames):
        pp = mod.PathPlanner(mod.MissionTypes.skidpad)
        paths, Ps, reloc_flag, index = [], [], [], []
        for cones, pos, direction in frames:
            rec.clear()
            paths.append(pp.calculate_path_in_global_frame(cones, pos, direction))
            Ps.append(int(rec.get("P", 0)))
            reloc_flag.append(bool(pp.relocalizer.is_relocalized))
            index.append(int(pp.pathing.index_along_path))
        info = pp.relocalization_info
        return (np.array(paths), np.array(Ps, np.int16), np.array(reloc_flag), np.array(index, np.int32),
                np.zeros(3) if info is None else np.array([info.translation[0], info.translation[1], info.rotation]))

    log = rh.load_demo_log("skidpad.json")
    batch = synth.pack_frames(log)
    paths, Ps, flag, index, info = run(log)
    out = {"log_cones_xy": batch.cones_xy, "log_cones_type": batch.cones_type, "log_offsets": batch.offsets,
           "log_pos": batch.pos, "log_dir": batch.dir, "log_path": paths, "log_P": Ps, "log_relocalized": flag,
           "log_index": index, "log_info": info}
    T, S = 6, 96
    xy, ty, off, pos, dirs = synth.gen_skidpad(4, T, S)
    syn_paths, syn_P, syn_flag, syn_index, syn_info = [], [], [], [], []
    for t in range(T):
        c = xy[off[t]:off[t + 1]]
        cones = [c[ty[off[t]:off[t + 1]] == k] for k in range(5)]
        p, P, f, ix, info_t = run([(cones, pos[t, s], dirs[t, s]) for s in range(S)])
        syn_paths.append(p), syn_P.append(P), syn_flag.append(f), syn_index.append(ix), syn_info.append(info_t)
        print(f"  skidpad synthetic trajectory {t}: relocalized at step {int(np.argmax(f)) if f.any() else -1}", flush=True)
    out.update({"syn_T": T, "syn_S": S, "syn_path": np.array(syn_paths), "syn_P": np.array(syn_P),
                "syn_relocalized": np.array(syn_flag), "syn_index": np.array(syn_index), "syn_info": np.array(syn_info)})
    np.savez_compressed(os.path.join(HERE, "skidpad.npz"), **out)
    print("[skidpad] log", len(log), "frames; synthetic", T, "x", S)


def global_path_goldens():
    """Reference runs of the GLOBAL-PATH branch of run_path_calculation (core_calculate_path.py:516-528):
    (a) PathPlanner(trackdrive).set_global_path(line) on poses along a synthetic closed racing line, fresh planner per pose
        and one sequential (stateful) run;
    (b) MissionTypes.acceleration, sequential, on a synthetic acceleration track (two straight cone rows under a rigid
        motion); the relocalizer's unseeded np.random.choice (acceleration_relocalization.py:32) is pinned by seeding
        numpy's global RNG at run time -- no source edit."""
    mod = rh.load_reference()
    rec = rh._state["rec"]
    rng = np.random.default_rng(7)
    centre, tangent = synth._track(np.random.default_rng([7, 0]))
    line = centre[::2].copy()  # ~0.3 m spacing, closed
    idx = rng.integers(0, len(centre), 48)
    nrm = np.stack([-tangent[idx, 1], tangent[idx, 0]], 1)
    pos = centre[idx] + nrm * rng.normal(0, 0.4, (48, 1))
    yaw = np.arctan2(tangent[idx, 1], tangent[idx, 0]) + rng.normal(0, np.deg2rad(6), 48)
    dirs = np.stack([np.cos(yaw), np.sin(yaw)], 1)
    empty = [np.zeros((0, 2)) for _ in range(5)]
    fresh, fresh_P = [], []
    for i in range(48):
        pp = mod.PathPlanner(mod.MissionTypes.trackdrive)
        pp.set_global_path(line)
        rec.clear()
        fresh.append(pp.calculate_path_in_global_frame(empty, pos[i], dirs[i]))
        fresh_P.append(int(rec.get("P", 0)))
    order = np.argsort(idx)  # one lap in driving order, one planner
    pp = mod.PathPlanner(mod.MissionTypes.trackdrive)
    pp.set_global_path(line)
    seq, seq_P = [], []
    for i in order:
        rec.clear()
        seq.append(pp.calculate_path_in_global_frame(empty, pos[i], dirs[i]))
        seq_P.append(int(rec.get("P", 0)))
    out = {"line": line, "pos": pos, "dir": dirs, "fresh_path": np.array(fresh), "fresh_P": np.array(fresh_P, np.int16),
           "seq_order": order, "seq_path": np.array(seq), "seq_P": np.array(seq_P, np.int16)}
    # (b) acceleration: cones every 5 m on both sides of a 75 m straight, 3 m wide, under a rigid motion + 3 cm noise
    from fsd_path_planning.relocalization.acceleration.acceleration_relocalization import BASE_ACCELERATION_PATH

    th, tr = 0.7, np.array([12.0, -30.0])
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    xs = np.arange(0.0, 80.0, 5.0)
    blue = np.stack([xs, np.full_like(xs, 1.5)], 1) + rng.normal(0, 0.03, (len(xs), 2))
    yellow = np.stack([xs, np.full_like(xs, -1.5)], 1) + rng.normal(0, 0.03, (len(xs), 2))
    cones = [np.zeros((0, 2)), yellow @ rot.T + tr, blue @ rot.T + tr, np.zeros((0, 2)), np.zeros((0, 2))]
    S = 40
    px = np.linspace(-2.0, 70.0, S)
    ppos = np.stack([px, rng.normal(0, 0.1, S)], 1) @ rot.T + tr
    pyaw = th + rng.normal(0, np.deg2rad(2), S)
    pdir = np.stack([np.cos(pyaw), np.sin(pyaw)], 1)
    seed = 1234
    np.random.seed(seed)
    pa = mod.PathPlanner(mod.MissionTypes.acceleration)
    acc, acc_P, acc_flag = [], [], []
    for s in range(S):
        rec.clear()
        acc.append(pa.calculate_path_in_global_frame(cones, ppos[s], pdir[s]))
        acc_P.append(int(rec.get("P", 0)))
        acc_flag.append(bool(pa.relocalizer.is_relocalized))
    info = pa.relocalization_info
    out.update({"acc_map": np.asarray(BASE_ACCELERATION_PATH, dtype=np.float64), "acc_seed": seed,
                "acc_cones_xy": np.concatenate(cones), "acc_cones_type": np.concatenate([np.full(len(c), t, np.uint8) for t, c in enumerate(cones)]),
                "acc_pos": ppos, "acc_dir": pdir, "acc_path": np.array(acc), "acc_P": np.array(acc_P, np.int16),
                "acc_relocalized": np.array(acc_flag),
                "acc_info": np.array([info.translation[0], info.translation[1], info.rotation])})
    np.savez_compressed(os.path.join(HERE, "global_path.npz"), **out)
    print("[global_path] trackdrive", len(fresh), "poses; acceleration", S, "steps, relocalized at",
          int(np.argmax(acc_flag)) if any(acc_flag) else -1)


if __name__ == "__main__":
    only = sys.argv[1:] or None
    fsg = synth.pack_frames(rh.load_demo_log("fsg_19_2_laps.json"))
    fss = synth.pack_frames(rh.load_demo_log("fss_19_4_laps.json")[::4])
    jobs = {
        "fsg_color": lambda: save("fsg_color", fsg),
        "fsg_colorless": lambda: save("fsg_colorless", synth.remove_color_info(fsg)),
        "fss_color": lambda: save("fss_color", fss),
        "fss_colorless": lambda: save("fss_colorless", synth.remove_color_info(fss)),
        "fixtures": lambda: save("fixtures", synth.pack_frames(fixtures() + edge_frames())),
        "synth_color": lambda: save("synth_color", synth.gen_autocross(2, 256).astype(np.float64)),
        "synth_colorless": lambda: save("synth_colorless", synth.remove_color_info(synth.gen_autocross(3, 256)).astype(np.float64)),
        "synth_mixed": lambda: save("synth_mixed", synth.gen_mixed(5, 256).astype(np.float64)),
        "fitpack": fitpack_goldens,
        "skidpad": skidpad_goldens,
        "global_path": global_path_goldens,
    }
    for name, fn in jobs.items():
        if only is None or name in only:
            fn()
