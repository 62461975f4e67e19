"""The kernels' own per-frame sources (sort.cuh / match.cuh / spline.cuh / path.cuh) compiled for the host with a
one-lane warp (csrc/hostcheck.cpp) against the golden vectors and scipy's spline fits.  CPU only; this is a check
of the kernel logic, not a product path."""
import os

import numpy as np
import pytest

import hostcheck
from conftest import GOLDEN_DIR, TIE_DIST_TOL, TIE_KAPPA_TOL, compare_with_golden, tie_frame_deviation


def test_kernel_sources_match_reference(golden):
    name, batch, g = golden
    h = hostcheck.plan_batch(batch, force_P=g["P"])
    compare_with_golden(name, g, h["left_idx"], h["right_idx"], h["n_wv"], h["left_wv"], h["right_wv"], h["l2r"],
                        h["r2l"], h["path"], path_tol=1e-7)
    ok = g["error"] == 0
    assert (h["grid"][ok, 1] == g["n_trim"][ok]).all()
    assert not (h["status"] & ((1 << 8) | (1 << 10))).any(), "overflow / unsupported flagged"


def test_kernel_sources_tie_rule(golden):
    name, batch, g = golden
    h = hostcheck.plan_batch(batch)
    ok = g["error"] == 0
    assert (h["grid"][ok, 0] == g["tie_P"][ok]).all()
    err = np.abs(h["path"] - g["tie_path"]).reshape(len(ok), -1).max(1)
    assert (err[ok] <= 1e-7).all(), name
    assert (h["status"][ok] & (1 << 7)).mean() > 0.9, "run-time grids are ties by construction (SURVEY Q13)"
    # Default mode against the UNMODIFIED reference (gate iii, strict): on frames where the reference's coin flip landed
    # on the other grid size the samples differ by up to one spacing, but they lie on the same curve -- gated
    # geometrically so that a regression of the tie rule (or of the curve) is caught, not just printed.
    strict = np.abs(h["path"] - g["path"]).reshape(len(ok), -1).max(1) <= 1e-4
    other = ok & (g["P"] != g["tie_P"])
    assert strict[ok & ~other].all(), f"{name}: default mode differs from the reference although P agrees"
    for b in np.where(other)[0]:
        dist, kappa = tie_frame_deviation(h["path"][b], g["path"][b])
        assert dist <= TIE_DIST_TOL and kappa <= TIE_KAPPA_TOL, (name, b, dist, kappa)


def test_normal_equation_spline_fit_matches_scipy():
    """Knot vectors identical to scipy.interpolate.splprep, coefficients to 1e-9 (normal equations + Cholesky vs
    FITPACK's Givens sweep)."""
    f = np.load(os.path.join(GOLDEN_DIR, "fitpack.npz"))
    pts, meta = f["points"], f["meta"]
    worst = 0.0
    for i, (m, k, s, fp, ier, n, start) in enumerate(meta):
        m, k, n, start, ier = int(m), int(k), int(n), int(start), int(ier)
        p = pts[start : start + m]
        t, cx, cy, kk, ier2 = hostcheck.fit(p, s)
        assert kk == k and len(t) == n and ier2 == ier, (i, m, s, len(t), n, ier2, ier)
        assert np.array_equal(t, f["knots"][i][:n]), i
        nk1 = n - k - 1
        worst = max(worst, np.abs(cx - f["coefs"][i][0][:nk1]).max(), np.abs(cy - f["coefs"][i][1][:nk1]).max())
    assert worst < 1e-9, worst


def test_initial_path_matches_oracle():
    import oracle

    assert np.abs(hostcheck.initial_path() - oracle.initial_path()).max() < 1e-9


def adjacency_equal(nbr, deg, ref_nbr, ref_deg):
    """Degrees equal and the first `degree` entries of every neighbour list equal (entries past the degree are unspecified)."""
    if not np.array_equal(np.asarray(deg, dtype=np.int32), ref_deg):
        return False
    live = np.arange(5)[None, None, :] < ref_deg[:, :, None]
    return bool((np.asarray(nbr, dtype=np.int32)[live] == ref_nbr[live]).all())


def test_knn_graph_sources_match_oracle_adjacency(golden):
    """the cost-matrix step (create_adjacency_matrix, adjacency_matrix.py:60-128) bit-exact on every golden
    frame -- kernel sources (host-check build) against the oracle's restatement of the reference's dense formulation."""
    import oracle

    name, batch, g = golden
    if name == "fixtures":
        # hand-made frames on an exact lattice: equal distances everywhere, and the k-th neighbour among equals is
        # unspecified in the reference itself (np.argsort's default sort is unstable)
        pytest.skip("lattice-exact coordinates: k-NN ties are unspecified in the reference")
    nbr, deg = hostcheck.knn(batch)
    ref_nbr, ref_deg = oracle.adjacency(batch)
    assert adjacency_equal(nbr, deg, ref_nbr, ref_deg), name


def test_suspended_fit_resumed_with_an_extended_arena():
    """A fit that wants more knots than its arena holds is SUSPENDED (not truncated) and resumed -- nothing recomputed --
    once the arena is larger: path_kernel does this for the rare frame that outgrows its NCAP knot records, with the
    CTA's whole shared memory as the larger arena.  Same sources, one lane: frames that start with 12 / 16 / 24 records
    and are resumed with 192 give the same bytes as fits that had the large arena all along; frame 59 383 of the bench
    stream (33 knots in one fit, in rank 5's shard of the 8-GPU bench) equals the oracle, which has no static bounds."""
    import oracle
    from ft_fsd_path_planning_b200 import synth

    batch = synth.concat_batches([synth.gen_autocross(2, 4, start=59382), synth.gen_autocross(7, 60),
                                  synth.remove_color_info(synth.gen_autocross(8, 32))]).astype(np.float64)
    try:
        hostcheck.set_caps(192, 8 * 704)
        ext = hostcheck.plan_batch(batch)
        outs = {}
        for start in (12, 16, 24, 0):
            hostcheck.set_caps(192, 0, resume=True, start_cap=start)
            outs[start] = hostcheck.plan_batch(batch)
    finally:
        hostcheck.set_caps(0, 0)
    assert not (ext["status"] & 0x100).any()
    for start, res in outs.items():
        for k in ("path", "grid", "status", "left_idx", "right_idx"):
            assert np.array_equal(res[k], ext[k]), (start, k)
    ref = oracle.plan_batch(batch, threads=2)
    assert np.array_equal(ext["left_idx"], ref["left_idx"]) and np.array_equal(ext["right_idx"], ref["right_idx"])
    same_P = ext["grid"][:, 0] == ref["P"]
    assert same_P[1] and same_P.mean() > 0.5
    assert np.abs(ext["path"] - ref["path"])[same_P].max() <= 1e-8
