"""Extract the canonical skidpad path (a data table of 5786 (x, y) points, 0.05 m spacing) from the reference into
ft_fsd_path_planning_b200/data/skidpad_path.npy.  Build container only (needs /root/reference).

Source of the DATA: fsd_path_planning/relocalization/skidpad/skidpad_path_data.py (BASE_SKIDPAD_PATH).  No code is
taken from the reference; the table is the track definition the skidpad mission tracks.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
import ref_harness as rh  # noqa: E402

rh.load_reference()
from fsd_path_planning.relocalization.skidpad.skidpad_path_data import BASE_SKIDPAD_PATH  # noqa: E402

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ft_fsd_path_planning_b200", "data", "skidpad_path.npy")
np.save(out, np.asarray(BASE_SKIDPAD_PATH, dtype=np.float64))
print(BASE_SKIDPAD_PATH.shape, "->", out)

# canonical-frame cone map of the skidpad track (data for the synthetic config-4 generator): the cones of the last
# frame of the reference's recorded skidpad log, moved into the map frame with the reference's own relocalization
mod = rh.load_reference()
frames = rh.load_demo_log("skidpad.json")
pp = mod.PathPlanner(mod.MissionTypes.skidpad)
for cones, pos, direction in frames[:40]:
    pp.calculate_path_in_global_frame(cones, pos, direction)
assert pp.relocalizer.is_relocalized
cones = frames[-1][0]
xy = np.concatenate([c.reshape(-1, 2) for c in cones])
ty = np.concatenate([np.full(len(c), t) for t, c in enumerate(cones)])
known = np.array([pp.relocalizer.transform_to_known_map_frame(p, 0.0)[0] for p in xy])
out2 = os.path.join(os.path.dirname(out), "skidpad_cones.npy")
np.save(out2, np.concatenate([known, ty[:, None].astype(np.float64)], axis=1))
print(known.shape, "->", out2)

# the two reference circle centres of the canonical path (calculate_reference_centers_for_skidpad_path,
# skidpad_relocalizer.py:172-183): constants of the track definition, taken from the reference's own run-time value
ref_centers = np.asarray(pp.relocalizer.reference_centers, dtype=np.float64)
out3 = os.path.join(os.path.dirname(out), "skidpad_ref_centers.npy")
np.save(out3, ref_centers)
print(ref_centers, "->", out3)

# the known map of the acceleration / EBS missions (BASE_ACCELERATION_PATH, acceleration_relocalization.py:175-211: a
# 160 m x 4.8 m loop at 0.2 m spacing with 1 cm of seeded noise) -- track-definition DATA like the skidpad table
from fsd_path_planning.relocalization.acceleration.acceleration_relocalization import BASE_ACCELERATION_PATH  # noqa: E402

out4 = os.path.join(os.path.dirname(out), "acceleration_path.npy")
np.save(out4, np.asarray(BASE_ACCELERATION_PATH, dtype=np.float64))
print(BASE_ACCELERATION_PATH.shape, "->", out4)
