"""The C-ABI library loads and exports every symbol include/fsdplan.h declares; host-side behaviour that needs no
GPU (defaults, workspace sizing, error strings, argument errors).  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from ft_fsd_path_planning_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "fsdplan.h")).read()
    declared = sorted(set(re.findall(r"\b(fsd_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 11
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fsdplan.h but not exported by libfsdplan.so"


def test_header_cites_reference_interfaces():
    header = open(os.path.join(ROOT, "include", "fsdplan.h")).read()
    for cite in ("full_pipeline.py:84-207", "core_cone_sorting.py:117-136", "core_cone_matching.py:87-124",
                 "core_calculate_path.py:514-575"):
        assert cite in header


def test_defaults_are_the_reference_values(lib):
    p = _lib.default_params()
    assert (p.max_n_neighbors, p.max_length) == (5, 12)
    assert (p.max_dist, p.max_dist_to_first, p.car_size) == (6.5, 6.0, 2.1)
    assert abs(p.threshold_directional_angle - 0.6981317007977318) < 1e-15
    assert abs(p.threshold_absolute_angle - 1.1344640137963142) < 1e-15
    assert (p.min_track_width, p.max_search_range) == (3.0, 5.0)
    assert (p.smoothing, p.predict_every, p.refit_smoothing) == (0.2, 0.1, 0.01)
    assert (p.maximal_distance_for_valid_path, p.mpc_path_length) == (5.0, 20.0)


def test_abi_version_strerror_workspace(lib):
    assert lib.fsd_abi_version() == 1
    assert lib.fsd_strerror(0) == b"ok"
    assert b"CPU" in lib.fsd_strerror(-4)
    small, big = lib.fsd_workspace_bytes(1, 0), lib.fsd_workspace_bytes(1024, 0)
    assert 0 < small < big and big >= 1024 * (1280 + 1024 + 128)


def test_argument_errors_do_not_need_a_device(lib):
    p = _lib.default_params()
    assert lib.fsd_params_default(None) == -1
    # null offsets / outputs are rejected before any CUDA call
    rc = lib.fsd_plan_batch(C.byref(p), 4, 8, None, None, None, None, None, None, None, None, None, None, None, 0,
                            None, None, 0, None)
    assert rc == -1
    # wrong mission
    one = C.c_int(0)
    addr = C.addressof(one)
    rc = lib.fsd_plan_batch(C.byref(p), 2, 8, addr, addr, addr, addr, addr, addr, addr, addr, None, None, None, 0,
                            addr, None, 0, None)
    assert rc == -5


def test_cuda_planner_never_falls_back_to_the_host():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ft_fsd_path_planning_b200 import BatchPlanner, MissionTypes, PathPlanner

    with pytest.raises(RuntimeError, match="never falls back to the host"):
        BatchPlanner()
    with pytest.raises(RuntimeError, match="never falls back to the host"):
        PathPlanner(MissionTypes.trackdrive)  # default device: CUDA.  The host planner needs device="cpu"


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "ft_fsd_path_planning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "fsd_oracle" not in text and "libfsd_oracle" not in text, f
