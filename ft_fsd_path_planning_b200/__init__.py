"""ft_fsd_path_planning_b200 -- B200-native batched cone-track path planner.

Drop-in for the sorting_cones -> cone_matching -> calculate_path pipeline of
papalotis/ft-fsd-path-planning: `PathPlanner(mission).calculate_path_in_global_frame(...)` keeps the
reference's signature; `BatchPlanner.plan(...)` plans thousands of independent frames per call with
hand-written sm_100a CUDA kernels (csrc/) behind the C-ABI of include/fsdplan.h.
"""
from .enums import ConeTypes, MissionTypes  # noqa: F401
from .synth import FrameBatch, gen_autocross, gen_mixed, pack_frames, remove_color_info  # noqa: F401

__all__ = ["ConeTypes", "MissionTypes", "FrameBatch", "gen_autocross", "gen_mixed", "pack_frames",
           "remove_color_info", "PathPlanner", "BatchPlanner", "SkidpadBatchPlanner", "PlanResult",
           "RelocalizationInformation", "ReferenceRaisesError", "CpuBatchPlanner", "build"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree (nvcc, sm_100a)."""
    from . import _lib

    return _lib.build(force=force, verbose=verbose)


def __getattr__(name):
    # torch is imported lazily so that the host-side helpers (synth, enums) stay importable without it
    if name in ("PathPlanner", "BatchPlanner", "PlanResult", "RelocalizationInformation", "ReferenceRaisesError",
                "CpuBatchPlanner"):
        from . import planner

        return getattr(planner, name)
    if name == "SkidpadBatchPlanner":
        from . import skidpad

        return skidpad.SkidpadBatchPlanner
    raise AttributeError(name)
