"""The oracle (oracle/*.c) against the golden vectors generated from the unmodified reference, and its
FITPACK restatement against scipy's results stored in tests/golden/fitpack.npz.  CPU only."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN_DIR, compare_with_golden


def test_oracle_matches_reference_P_conditioned(golden):
    name, batch, g = golden
    r = oracle.plan_batch(batch, force_P=g["P"], threads=4)
    compare_with_golden(name, g, r["left_idx"], r["right_idx"], np.stack([r["n_left_wv"], r["n_right_wv"]], 1),
                        r["left_wv"], r["right_wv"], r["l2r"], r["r2l"], r["path"], path_tol=1e-8)
    ok = g["error"] == 0
    assert (r["n_trim"][ok] == g["n_trim"][ok]).all()
    assert not (r["status"] & (oracle.STATUS_BITS["UNSUPPORTED"] | oracle.STATUS_BITS["OVERFLOW"])).any()


def test_oracle_tie_normalised_and_strict(golden):
    """default rule vs the reference run with the same rule patched in; (iii) strict: informational."""
    name, batch, g = golden
    r = oracle.plan_batch(batch, threads=4)
    ok = g["error"] == 0
    err = np.abs(r["path"] - g["tie_path"]).reshape(len(ok), -1).max(1)
    assert (r["P"][ok] == g["tie_P"][ok]).all(), name
    assert (err[ok] <= 1e-8).all(), name
    strict = np.abs(r["path"] - g["path"]).reshape(len(ok), -1).max(1) <= 1e-4
    print(f"{name}: strict parity on {int(strict[ok].sum())}/{int(ok.sum())} frames (reference P coin flip, SURVEY Q13)")


def test_oracle_fitpack_matches_scipy():
    f = np.load(os.path.join(GOLDEN_DIR, "fitpack.npz"))
    pts, meta = f["points"], f["meta"]
    worst = 0.0
    for i, (m, k, s, fp, ier, n, start) in enumerate(meta):
        m, k, n, start, ier = int(m), int(k), int(n), int(start), int(ier)
        p = pts[start : start + m]
        t, cx, cy, kk, fp2, ier2, u = oracle.splprep(p, s, k)
        assert kk == k and ier2 == ier and len(t) == n
        assert np.array_equal(t, f["knots"][i][:n])
        nk1 = n - k - 1
        worst = max(worst, np.abs(cx - f["coefs"][i][0][:nk1]).max(), np.abs(cy - f["coefs"][i][1][:nk1]).max())
        ev = f["evals"][i]
        assert np.abs(oracle.splev(ev[:, 0], t, cx, k) - ev[:, 1]).max() < 1e-9
        assert np.abs(oracle.splev(ev[:, 0], t, cy, k) - ev[:, 2]).max() < 1e-9
    assert worst < 1e-9


def test_oracle_empty_frame_known_answer():
    """an empty frame returns the constant initial path pushed through the MPC tail."""
    from conftest import load_golden

    batch, g = load_golden("fixtures")
    r = oracle.plan_batch(batch.slice(8, 9), force_P=g["P"][8:9])  # first edge frame: no cones, pose (0,0)/(1,0)
    p = r["path"][0]
    assert np.allclose(p[1, :3], [0.495, 0.495, 1.29e-4], atol=1e-3)
    assert abs(p[0, 3] - 1.00039e-3) < 1e-5
    assert r["status"][0] & oracle.STATUS_BITS["FEW_CONES"]
    assert oracle.initial_path().shape == (40, 4)


def test_oracle_simple_corner_known_answer():
    """SURVEY section 4 known-answer vector ("Simple Corner", colours known)."""
    from conftest import load_golden

    batch, g = load_golden("fixtures")
    r = oracle.plan_batch(batch.slice(2, 3), force_P=g["P"][2:3])
    assert r["n_left"][0] == 8 and r["n_right"][0] == 10
    assert list(r["l2r"][0][:8]) == [0, 1, 3, 4, 5, 7, 8, 9]
    assert list(r["r2l"][0][:10]) == [0, 1, 1, 2, 3, 4, 5, 5, 6, 7]
    assert np.allclose(r["path"][0][20], [10.07905, 8.14601, 5.16626, 0.09695], atol=2e-5)


def test_numpy_pairwise_sum_restatement():
    rng = np.random.default_rng(0)
    import ctypes as C

    # the oracle's evaluation-grid size depends on numpy's pairwise summation order: check P on goldens instead
    # of re-deriving it; here only make sure that the default rule never yields a grid below the horizon.
    from conftest import load_golden

    batch, g = load_golden("synth_color")
    r = oracle.plan_batch(batch.slice(0, 32))
    assert (r["P"] >= 40).all()
