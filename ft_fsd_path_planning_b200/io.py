"""Log / wire-format loader: the reference's recorded SLAM logs (a JSON list of
{"car_position": [x, y], "car_direction": [dx, dy], "slam_cones": [5 lists of [x, y]]}) -> packed frame batches.

Reference: load_data_json, fsd_path_planning/demo/json_demo.py:255-275 (including remove_color_info :266-273).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Union

import numpy as np

from .synth import FrameBatch, pack_frames, remove_color_info as _remove_color


def load_data_json(data_path: Union[str, Path], remove_color_info: bool = False, dtype=np.float64) -> FrameBatch:
    """All frames of a log as one FrameBatch (cones in ConeTypes order inside each frame)."""
    data = json.loads(Path(data_path).read_text())
    frames = []
    for d in data:
        cones = [np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in d["slam_cones"]]
        if len(cones) != 5:
            raise ValueError("slam_cones must hold 5 lists (one per ConeTypes value)")
        frames.append((cones, np.asarray(d["car_position"], dtype=np.float64).reshape(2),
                       np.asarray(d["car_direction"], dtype=np.float64).reshape(2)))
    batch = pack_frames(frames, dtype=dtype)
    return _remove_color(batch) if remove_color_info else batch


def save_data_json(batch: FrameBatch, data_path: Union[str, Path]) -> None:
    """Inverse of load_data_json (round-trip helper for tests and for exporting synthetic batches)."""
    out = []
    for b in range(batch.n_frames):
        cones, pos, direction = batch.frame(b)
        out.append({"car_position": pos.tolist(), "car_direction": direction.tolist(),
                    "slam_cones": [c.tolist() for c in cones]})
    Path(data_path).write_text(json.dumps(out))
