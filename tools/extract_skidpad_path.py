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
