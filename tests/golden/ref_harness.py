"""Run the UNMODIFIED reference (papalotis/ft-fsd-path-planning) and record golden vectors.

Only usable where the reference source tree exists (the build container: /root/reference).
Nothing in `pytest -m gpu`, `bench.py` or `smoke()` imports this module; it is the
generator of the committed fixtures under tests/golden/*.npz and the live cross-check used
by tests that are skipped when the reference is absent.

The reference is imported as-is; the only additions are
  * a one-line stub for the unused `icecream` import
    (fsd_path_planning/cone_matching/functional_cone_matching.py:15),
  * run-time hooks (monkeypatches, no source edits) that RECORD values the public API does
    not return: the sort indices chosen by
    `calc_final_configs_for_left_and_right` (sorting_cones/trace_sorter/combine_traces.py:21)
    and the size P of the last evaluation grid (`PathParameterizer._refit_spline`,
    calculate_path/path_parameterization.py:125) -- see SURVEY.md Q13,
  * optionally the "tie-normalised" evaluation-grid rule of SURVEY.md section 8(d)(ii).
"""
from __future__ import annotations

import math
import os
import sys
import types
from typing import Dict, List, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("FSD_REFERENCE_ROOT", "/root/reference")
MAX_SORTED = 12
MAX_WV = 32

_state: Dict[str, object] = {}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fsd_path_planning"))


def load_reference():
    """Import the reference package with the hooks installed (idempotent)."""
    if "mod" in _state:
        return _state["mod"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    if "icecream" not in sys.modules:
        stub = types.ModuleType("icecream")
        stub.ic = lambda *a, **k: (a[0] if a else None)
        sys.modules["icecream"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    warnings.filterwarnings("ignore")
    import fsd_path_planning as mod
    from fsd_path_planning.calculate_path import path_parameterization as pp
    from fsd_path_planning.sorting_cones.trace_sorter import core_trace_sorter as cts
    from fsd_path_planning.utils import spline_fit as sf

    rec: Dict[str, object] = {}
    _state["rec"] = rec

    orig_combine = cts.calc_final_configs_for_left_and_right

    def combine_hook(*args, **kwargs):
        left, right = orig_combine(*args, **kwargs)
        rec["left_idx"] = np.asarray(left)
        rec["right_idx"] = np.asarray(right)
        return left, right

    cts.calc_final_configs_for_left_and_right = combine_hook

    orig_refit = pp.PathParameterizer._refit_spline

    def refit_hook(self, path):
        spline = orig_refit(self, path)
        rec["n_trim"] = len(path)
        rec["P"] = len(spline.calculate_u_eval())
        return spline

    pp.PathParameterizer._refit_spline = refit_hook

    orig_u_eval = sf.SplineEvaluator.calculate_u_eval

    def u_eval_hook(self, max_u=None):
        if not _state.get("tie_rule", False):
            return orig_u_eval(self, max_u)
        if max_u is None:
            max_u = self.max_u
        q = max_u / self.predict_every
        r = round(q)
        n = r if abs(q - r) < 1e-9 else math.ceil(q)
        return np.arange(n) * self.predict_every

    sf.SplineEvaluator.calculate_u_eval = u_eval_hook
    _state["mod"] = mod
    return mod


def set_tie_rule(enabled: bool) -> None:
    """P = round(q) when |q - round(q)| < 1e-9 else ceil(q)."""
    _state["tie_rule"] = bool(enabled)


def run_frame(cones_by_type, pos, direction, mission: str = "trackdrive") -> Dict[str, object]:
    """Fresh PathPlanner per frame (the batch semantic)."""
    mod = load_reference()
    rec = _state["rec"]
    rec.clear()
    planner = mod.PathPlanner(getattr(mod.MissionTypes, mission))
    rec.clear()  # the constructor runs one parameterisation of the initial path
    out: Dict[str, object] = {"error": ""}
    try:
        res = planner.calculate_path_in_global_frame(
            [np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in cones_by_type],
            np.asarray(pos, dtype=np.float64),
            np.asarray(direction, dtype=np.float64),
            return_intermediate_results=True,
        )
    except Exception as exc:  # the reference raises on a few degenerate inputs
        out["error"] = f"{type(exc).__name__}: {exc}"
        res = None
    out["left_idx"] = rec.get("left_idx", np.zeros(0, int))
    out["right_idx"] = rec.get("right_idx", np.zeros(0, int))
    out["P"] = int(rec.get("P", -1))
    out["n_trim"] = int(rec.get("n_trim", -1))
    if res is not None:
        (out["path"], out["sorted_left"], out["sorted_right"], out["left_wv"], out["right_wv"],
         out["l2r"], out["r2l"]) = res
    return out


def _pad_idx(a, n, fill=-1):
    o = np.full(n, fill, dtype=np.int16)
    a = np.asarray(a).astype(np.int64)
    o[: len(a)] = a
    return o


def _pad_xy(a, n):
    o = np.full((n, 2), np.nan)
    a = np.asarray(a, dtype=np.float64).reshape(-1, 2)
    o[: len(a)] = a
    return o


def run_batch(batch, tie_rule: bool = False, progress: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Golden record for every frame of a FrameBatch (see synth.FrameBatch)."""
    set_tie_rule(tie_rule)
    B = batch.n_frames
    g = {
        "left_idx": np.full((B, MAX_SORTED), -1, np.int16),
        "right_idx": np.full((B, MAX_SORTED), -1, np.int16),
        "n_left_wv": np.zeros(B, np.int16),
        "n_right_wv": np.zeros(B, np.int16),
        "left_wv": np.full((B, MAX_WV, 2), np.nan),
        "right_wv": np.full((B, MAX_WV, 2), np.nan),
        "l2r": np.full((B, MAX_WV), -2, np.int16),
        "r2l": np.full((B, MAX_WV), -2, np.int16),
        "path": np.full((B, 40, 4), np.nan),
        "P": np.zeros(B, np.int16),
        "n_trim": np.zeros(B, np.int16),
        "error": np.zeros(B, np.uint8),
    }
    errors: List[str] = []
    for b in range(B):
        cones, pos, direction = batch.frame(b)
        r = run_frame(cones, pos, direction)
        g["left_idx"][b] = _pad_idx(r["left_idx"], MAX_SORTED)
        g["right_idx"][b] = _pad_idx(r["right_idx"], MAX_SORTED)
        g["P"][b] = r["P"]
        g["n_trim"][b] = r["n_trim"]
        if r["error"]:
            g["error"][b] = 1
            errors.append(f"{b}: {r['error']}")
            continue
        g["n_left_wv"][b] = len(r["left_wv"])
        g["n_right_wv"][b] = len(r["right_wv"])
        g["left_wv"][b] = _pad_xy(r["left_wv"], MAX_WV)
        g["right_wv"][b] = _pad_xy(r["right_wv"], MAX_WV)
        g["l2r"][b] = _pad_idx(r["l2r"], MAX_WV, fill=-2)
        g["r2l"][b] = _pad_idx(r["r2l"], MAX_WV, fill=-2)
        g["path"][b] = r["path"]
        if progress and (b + 1) % progress == 0:
            print(f"  reference: {b + 1}/{B}", flush=True)
    set_tie_rule(False)
    g["error_text"] = np.array(errors)
    return g


def load_demo_log(name: str):
    """Frames of one of the reference's recorded logs as [(cones_by_type, pos, dir), ...]
    (format: fsd_path_planning/demo/json_demo.py:255-275)."""
    import json

    path = os.path.join(REFERENCE_ROOT, "fsd_path_planning", "demo", name)
    data = json.load(open(path))
    frames = []
    for d in data:
        cones = [np.array(c, dtype=np.float64).reshape(-1, 2) for c in d["slam_cones"]]
        frames.append((cones, np.array(d["car_position"], float), np.array(d["car_direction"], float)))
    return frames
