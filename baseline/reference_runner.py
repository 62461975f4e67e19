#!/usr/bin/env python
"""Time the UNMODIFIED reference (papalotis/ft-fsd-path-planning, numba + scipy) on the host cores.

Bench infrastructure only (bench.py's `cpu_baseline.reference_numba` leg and `--impl reference`); the product never
imports this.  BASELINE.md section 4: a fresh `PathPlanner(MissionTypes.trackdrive)` per frame (the batched GPU
semantics), frames = a subset of exactly the synthetic batch the GPU plans (fp32 coordinates widened to fp64), BLAS / OMP
threads = 1, import and JIT warm-up excluded, 1 process and `multiprocessing.Pool(n)` (wall time of the slowest
worker).  The reference is looked for in baseline/_ref (baseline/install_reference.sh) and then in /root/reference.

    python baseline/reference_runner.py --seed 2 --frames 2048 [--single-frames 512] [--workers N] [--check]

prints ONE JSON line: {"available": true, "source": ..., "cores": n, "frames": F, "single": {...}, "pool": {...},
"jit_warmup_s": ..., "parity": {...}} or {"available": false, "reason": "..."}.
"""
import argparse
import importlib.util
import json
import os
import sys
import time

for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/fsd_numba_cache")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def locate():
    for cand in (os.path.join(HERE, "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "fsd_path_planning", "sorting_cones")):
            return cand
    return None


def _load_synth():
    spec = importlib.util.spec_from_file_location("fsd_synth", os.path.join(ROOT, "ft_fsd_path_planning_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fsd_synth"] = mod  # dataclasses look the module up by name
    spec.loader.exec_module(mod)
    return mod


_REF = {}


def _import_reference(src):
    if "pp" in _REF:
        return _REF["pp"], _REF["mt"]
    import types
    import warnings

    warnings.filterwarnings("ignore")
    if src not in sys.path:
        sys.path.insert(0, src)
    if "icecream" not in sys.modules and not os.path.exists(os.path.join(src, "icecream.py")):
        stub = types.ModuleType("icecream")
        stub.ic = lambda *a, **k: (a[0] if a else None)
        sys.modules["icecream"] = stub
    from fsd_path_planning import MissionTypes, PathPlanner

    _REF["pp"], _REF["mt"] = PathPlanner, MissionTypes
    return PathPlanner, MissionTypes


def _plan_frames(src, frames, keep):
    """Fresh planner per frame.  Returns (seconds, number of frames that raised, [paths] if keep)."""
    import numpy as np

    PathPlanner, MissionTypes = _import_reference(src)
    out, raised = [], 0
    t0 = time.perf_counter()
    for cones, pos, direction in frames:
        try:
            path = PathPlanner(MissionTypes.trackdrive).calculate_path_in_global_frame(cones, pos, direction)
        except Exception:  # the reference raises on a few degenerate inputs (e.g. one cone on a side)
            raised += 1
            path = np.full((40, 4), np.nan)
        if keep:
            out.append(path)
    return time.perf_counter() - t0, raised, out


def _worker(args):
    src, seed, lo, hi, warm = args
    synth = _load_synth()
    frames = [synth.gen_autocross_frame(seed, i) for i in range(lo, hi)]
    _plan_frames(src, frames[:warm], False)  # import + JIT (loaded from the numba cache) + first-call overheads
    dt, raised, _ = _plan_frames(src, frames, False)
    return dt, raised, hi - lo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--frames", type=int, default=2048, help="frames of the pool leg (frames 0 .. F-1 of the batch)")
    ap.add_argument("--single-frames", type=int, default=512, help="frames of the 1-process leg")
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--check", action="store_true", help="compare the paths of the 1-process leg with the oracle port")
    a = ap.parse_args()
    src = locate()
    if src is None:
        print(json.dumps({"available": False, "reason": "no reference under baseline/_ref or /root/reference "
                                                        "(run baseline/install_reference.sh where /root/reference exists)"}))
        return
    try:
        import numba  # noqa: F401
        import scipy  # noqa: F401
        import sklearn  # noqa: F401
    except Exception as e:  # pragma: no cover
        print(json.dumps({"available": False, "reason": f"reference dependencies not importable: {e!r}"}))
        return
    import numpy as np

    synth = _load_synth()
    cores = a.workers or os.cpu_count() or 1
    t0 = time.perf_counter()
    warm = [synth.gen_autocross_frame(a.seed, i) for i in range(3)]
    try:
        _plan_frames(src, warm, False)
    except Exception as e:
        print(json.dumps({"available": False, "reason": f"reference failed to run: {e!r}"}))
        return
    jit_s = time.perf_counter() - t0
    ns = min(a.single_frames, a.frames)
    frames = [synth.gen_autocross_frame(a.seed, i) for i in range(ns)]
    dt1, raised1, paths = _plan_frames(src, frames, a.check)
    out = {"available": True, "source": src, "cores": cores, "jit_warmup_s": round(jit_s, 1),
           "single": {"frames": ns, "seconds": dt1, "frames_per_s": ns / dt1, "raised": raised1}}
    if a.check:
        sys.path.insert(0, ROOT)
        import oracle

        batch = synth.pack_frames(frames, dtype=np.float64)
        ref = np.stack(paths)
        ok = np.isfinite(ref).all(axis=(1, 2))
        port = oracle.plan_batch(batch, threads=cores)
        # the reference's own grid size P is not returned by its API: compare under the port's P only where the last
        # sample index agrees (u[39] = (P - 1) * step)
        same = ok & (np.abs(ref[:, -1, 0] - port["path"][:, -1, 0]) < 1e-6)
        err = np.abs(ref[same] - port["path"][same]).max() if same.any() else None
        out["parity"] = {"frames": int(ns), "same_grid": int(same.sum()), "path_max_err_vs_port": err,
                         "note": "frames whose grid size P differs (the reference's 120/121 coin flip, SURVEY Q13) are not compared"}
    # pool leg: contiguous blocks of the first `frames` frames, one per worker
    import multiprocessing as mp

    F = a.frames
    bounds = [F * w // cores for w in range(cores + 1)]
    jobs = [(src, a.seed, bounds[w], bounds[w + 1], 3) for w in range(cores) if bounds[w + 1] > bounds[w]]
    ctx = mp.get_context("spawn")
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_worker, jobs)
    slowest = max(r[0] for r in res)
    out["pool"] = {"frames": F, "workers": len(jobs), "seconds_slowest_worker": slowest, "frames_per_s": F / slowest,
                   "raised": int(sum(r[1] for r in res))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
