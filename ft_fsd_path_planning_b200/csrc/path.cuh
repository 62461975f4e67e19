// Path calculation for ONE frame by ONE warp (P1-P4 of SURVEY.md section 8a).
//
// Behaviour follows the reference's
//   fsd_path_planning/calculate_path/core_calculate_path.py:151-575      (centre line, fallbacks, MPC tail)
//   fsd_path_planning/calculate_path/path_parameterization.py:49-328     (re-fit, curvature, 40 samples)
//   fsd_path_planning/calculate_path/path_calculator_helpers.py:26-68    (initial "almost straight" path)
//   fsd_path_planning/utils/math_utils.py:579-646                        (hyper circle fit)
// Every dense array (spline evaluations, chord lengths, curvature windows) is lane-strided; prefix
// sums are warp scans; argmin / first-index searches are warp reductions.
#pragma once

#include "lane.cuh"
#include "plan_types.cuh"
#include "spline.cuh"

namespace fsd {

constexpr int PCAP = 704;   // path points held per frame (fallback path: 62.8 m / 0.1 m + extension)
constexpr int GRID_CAP = 128;  // size of the last evaluation grid (P = 120 or 121 at run time)

enum { RC_OK = 0, RC_VALUE_ERROR = 1, RC_RAISES = 2, RC_UNSUPPORTED = 3 };

// per-frame point buffers live in per-CTA global scratch (L2 resident): PCAP x (16 + 8) bytes
constexpr size_t PATH_SCRATCH_BYTES = (size_t)PCAP * (sizeof(d2) + sizeof(double));

struct PathSmem {
  d2 *pts;    // [PCAP]
  double *u;  // [PCAP]
  SplineWork W;
  d2 centre[FSD_HORIZON];
  d2 prev_xy[FSD_HORIZON];
  double curv[GRID_CAP];
  int32_t si[4];
};

// ---- hyper circle fit -----------------------------------------------------------------------------

FSD_DEVFN void hyper_from_moments(double mx, double my, double Mxx, double Myy, double Mxy, double Mxz, double Myz,
                                double Mzz, double &cx, double &cy, double &r) {
  const double Mz = Mxx + Myy, Cov_xy = Mxx * Myy - Mxy * Mxy, Var_z = Mzz - Mz * Mz;
  const double A2 = 4.0 * Cov_xy - 3.0 * Mz * Mz - Mzz;
  const double A1 = Var_z * Mz + 4.0 * Cov_xy * Mz - Mxz * Mxz - Myz * Myz;
  const double A0 = Mxz * (Mxz * Myy - Myz * Mxy) + Myz * (Myz * Mxx - Mxz * Mxy) - Var_z * Cov_xy;
  const double A22 = A2 + A2;
  double y = A0, x = 0.0;
  for (int it = 0; it < 99; ++it) {
    double Dy = A1 + x * (A22 + 16.0 * x * x);
    double xn = x - fdiv(y, Dy);
    if (xn == x || !isfinite(xn)) break;
    double yn = A0 + xn * (A1 + xn * (A2 + 4.0 * xn * xn));
    if (fabs(yn) >= fabs(y)) break;
    x = xn;
    y = yn;
  }
  const double det = x * x - x * Mz + Cov_xy;
  const double Xc = fdiv(Mxz * (Myy - x) - Myz * Mxy, det) / 2.0;
  const double Yc = fdiv(Myz * (Mxx - x) - Mxz * Mxy, det) / 2.0;
  cx = Xc + mx;
  cy = Yc + my;
  r = fsqrt(fabs(Xc * Xc + Yc * Yc + Mz));
}

// one lane fits one window (curvature)
FSD_DEVFN double circle_radius_serial(const d2 *p, int n) {
  double mx = 0.0, my = 0.0;
  for (int i = 0; i < n; ++i) {
    mx += p[i].x;
    my += p[i].y;
  }
  const double inv_n = fdiv(1.0, (double)n);
  mx *= inv_n;
  my *= inv_n;
  double Mxy = 0, Mxx = 0, Myy = 0, Mxz = 0, Myz = 0, Mzz = 0;
  for (int i = 0; i < n; ++i) {
    double xi = p[i].x - mx, yi = p[i].y - my, zi = xi * xi + yi * yi;
    Mxy += xi * yi;
    Mxx += xi * xi;
    Myy += yi * yi;
    Mxz += xi * zi;
    Myz += yi * zi;
    Mzz += zi * zi;
  }
  double cx, cy, r;
  hyper_from_moments(mx, my, Mxx * inv_n, Myy * inv_n, Mxy * inv_n, Mxz * inv_n, Myz * inv_n, Mzz * inv_n, cx, cy, r);
  return r;
}

// the whole warp fits one point set (path extension)
FSD_DEVFN void circle_fit_warp(const d2 *p, int n, double &cx, double &cy, double &r) {
  double sx = 0.0, sy = 0.0;
  for (int i = fsd_lane(); i < n; i += FSD_LANES) {
    sx += p[i].x;
    sy += p[i].y;
  }
  const double inv_n = fdiv(1.0, (double)n);
  const double mx = wsum(sx) * inv_n, my = wsum(sy) * inv_n;
  double Mxy = 0, Mxx = 0, Myy = 0, Mxz = 0, Myz = 0, Mzz = 0;
  for (int i = fsd_lane(); i < n; i += FSD_LANES) {
    double xi = p[i].x - mx, yi = p[i].y - my, zi = xi * xi + yi * yi;
    Mxy += xi * yi;
    Mxx += xi * xi;
    Myy += yi * yi;
    Mxz += xi * zi;
    Myz += yi * zi;
    Mzz += zi * zi;
  }
  Mxy = wsum(Mxy) * inv_n;
  Mxx = wsum(Mxx) * inv_n;
  Myy = wsum(Myy) * inv_n;
  Mxz = wsum(Mxz) * inv_n;
  Myz = wsum(Myz) * inv_n;
  Mzz = wsum(Mzz) * inv_n;
  hyper_from_moments(mx, my, Mxx, Myy, Mxy, Mxz, Myz, Mzz, cx, cy, r);
}

FSD_DEV double orient(const d2 &p0, const d2 &p1, const d2 &p2) {
  // sign of det [[1, p0], [1, p1], [1, p2]] (np.linalg.det in the reference)
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

// ---- chord-length parameters: u[0] = 0, u[i] = u[i-1] + |p_i - p_{i-1}| (np.cumsum) -------------------

FSD_DEVFN void chord_params(const d2 *p, int m, double *u) {
  const int lane = fsd_lane();
  double carry = 0.0;
  if (lane == 0) u[0] = 0.0;
  for (int base = 1; base < m; base += FSD_LANES) {
    const int i = base + lane;
    double d = 0.0;
    if (i < m) {
      double ddx = p[i].x - p[i - 1].x, ddy = p[i].y - p[i - 1].y;
      d = fsqrt(ddx * ddx + ddy * ddy);
    }
    double incl = wscan_incl(d) + carry;
    if (i < m) u[i] = incl;
    carry = wlast(incl);
  }
  wsync();
}

// ---- SplineFitterFactory.fit(...).predict(der=0) ------------------------------------------------------
// fits src[0..m) and evaluates every `step` up to max_u (the fit's own when max_u_override <= 0) into
// dst[0..*n_out).  dst may alias src (the evaluation only needs the coefficients).

FSD_DEVFN int fit_predict(PathSmem &S, const d2 *src, double *u, int m, double smoothing, double step,
                          double max_u_override, d2 *dst, int dst_cap, int *n_out, unsigned *status) {
  if (m < 2) return RC_RAISES;  // NullSplineEvaluator
  chord_params(src, m, u);
  int ier = fit_curve(S.W, src, u, m, smoothing, status);
  if (ier == 10) return (*status & FSD_ST_UNSUPPORTED) ? RC_UNSUPPORTED : RC_VALUE_ERROR;
  const double mu = max_u_override > 0.0 ? max_u_override : S.W.max_u;
  const double q = ceil(fdiv(mu, step));  // len(np.arange(0, max_u, step))
  const int n = q > 0.0 ? (q > 1e6 ? 1000000 : (int)q) : 0;
  if (n > dst_cap) {
    *status |= FSD_ST_OVERFLOW;
    return RC_UNSUPPORTED;
  }
  wsync();
  for (int i = fsd_lane(); i < n; i += FSD_LANES) {
    double x, y;
    spline_point(S.W, (double)i * step, x, y);
    dst[i].x = x;
    dst[i].y = y;
  }
  wsync();
  *n_out = n;
  return RC_OK;
}

// ---- PathParameterizer.parameterize_path (path_parameterization.py:297-328) ----------------------------
// path = S.pts[0..n).  Writes the (40, 4) result to out (lane-strided) and P to *P_out.

FSD_DEVFN int parameterize(PathSmem &S, int n, int force_P, const DevParams &P, double *out, int *P_out,
                           unsigned *status) {
  const int lane = fsd_lane();
  if (n < 2) return RC_RAISES;
  // _refit_spline :125-161
  double len = 0.0, first10 = 0.0;
  for (int i = lane; i + 1 < n; i += FSD_LANES) {
    double ddx = S.pts[i + 1].x - S.pts[i].x, ddy = S.pts[i + 1].y - S.pts[i].y;
    double d = fsqrt(ddx * ddx + ddy * ddy);
    len += d;
    if (i < 10) first10 += d;
  }
  const double path_length = wsum(len);
  const int nm = n - 1 < 10 ? n - 1 : 10;
  const double mean_dist = fdiv(wsum(first10), (double)nm);
  const double predict_every = fdiv(fdiv(path_length, (double)FSD_HORIZON), 3.0);
  const double ratio = fdiv(predict_every, mean_dist);
  int skip = 1;
  if (isfinite(ratio) && ratio < 1e6 && (int)ratio > 1) skip = (int)ratio;
  int ms = (n + skip - 1) / skip;
  wsync();
  if (skip > 1) {
    if (lane == 0)
      for (int i = 1; i < ms; ++i) S.pts[i] = S.pts[i * skip];  // path[::skip]
    wsync();
  }
  chord_params(S.pts, ms, S.u);
  int ier = fit_curve(S.W, S.pts, S.u, ms, P.refit_smoothing, status);
  if (ier == 10) return (*status & FSD_ST_UNSUPPORTED) ? RC_UNSUPPORTED : RC_VALUE_ERROR;
  // size of the evaluation grid np.arange(0, max_u, predict_every): SURVEY.md Q13
  int Pn;
  if (force_P > 0) {
    Pn = force_P;
  } else {
    const double q = fdiv(S.W.max_u, predict_every);
    const double r = rint(q);
    if (fabs(q - r) < 1e-9) {
      Pn = (int)r;
      *status |= FSD_ST_TIE_P;
    } else {
      Pn = q < 1e6 ? (int)ceil(q) : 1000000;
    }
  }
  *P_out = Pn;
  if (Pn < FSD_HORIZON) return RC_VALUE_ERROR;  // repeated sample indices (:284-285)
  if (Pn > GRID_CAP) {
    *status |= FSD_ST_OVERFLOW;
    return RC_UNSUPPORTED;
  }
  wsync();
  for (int i = lane; i < Pn; i += FSD_LANES) {
    double x, y;
    spline_point(S.W, (double)i * predict_every, x, y);
    S.pts[i].x = x;
    S.pts[i].y = y;
  }
  wsync();
  // _calculate_path_curvature :163-193 / calculate_path_curvature :49-93 (open path)
  int window = Pn / 5 < 30 ? Pn / 5 : 30;
  if (window % 2 == 0) window += 1;
  const int hw = window / 2;
  for (int i = lane; i < Pn; i += FSD_LANES) {
    int lo = i - hw < 0 ? 0 : i - hw;
    int hi = i + hw > Pn - 1 ? Pn - 1 : i + hw;
    const int cnt = hi - lo + 1;
    double r = circle_radius_serial(S.pts + lo, cnt);
    r = fmin(fmax(r, 1.0), 3000.0);
    double sg = sgn(orient(S.pts[lo], S.pts[lo + cnt / 2], S.pts[hi]));
    S.curv[i] = fdiv(1.0, r) * sg;
  }
  wsync();
  // uniform_filter1d(size, mode="nearest") evaluated at the 40 sampled indices only;
  // indices np.linspace(0, P-1, 40, dtype=int) (:277-282)
  const int fs = window / 2 > 2 ? window / 2 : 2;
  const double stp = fdiv((double)(Pn - 1), (double)(FSD_HORIZON - 1));
  for (int j = lane; j < FSD_HORIZON; j += FSD_LANES) {
    const int idx = j == FSD_HORIZON - 1 ? Pn - 1 : (int)floor((double)j * stp);
    double acc = 0.0;
    for (int q = idx - fs / 2; q <= idx + fs - fs / 2 - 1; ++q) {
      int qq = q < 0 ? 0 : (q > Pn - 1 ? Pn - 1 : q);
      acc += S.curv[qq];
    }
    out[4 * j + 0] = (double)idx * predict_every;
    out[4 * j + 1] = S.pts[idx].x;
    out[4 * j + 2] = S.pts[idx].y;
    out[4 * j + 3] = fdiv(acc, (double)fs);
  }
  wsync();
  return RC_OK;
}

// ---- CalculatePath MPC tail (core_calculate_path.py:336-417) on path = S.pts[1 .. 1+n_in) -----------------

FSD_DEVFN int mpc_tail(PathSmem &S, int n_in, const FramePose &F, int force_P, const DevParams &P, double *out,
                       int *P_out, int *n_trim, unsigned *status) {
  const int lane = fsd_lane();
  if (n_in < 1) return RC_RAISES;
  d2 *path = S.pts + 1;
  int n = n_in;
  // connect_path_to_car :430-457
  {
    const double fx = path[0].x - F.px, fy = path[0].y - F.py;
    const double d = fsqrt(fx * fx + fy * fy);
    const bool behind = cos_between(fx, fy, F.dx, F.dy) < 0.0;  // angle > pi/2
    wsync();
    if (!(d < 0.5 || behind)) {
      if (lane == 0) {
        S.pts[0].x = F.px + fdiv(fx, d) * 0.2;
        S.pts[0].y = F.py + fdiv(fy, d) * 0.2;
      }
      path = S.pts;
      n = n_in + 1;
    }
    wsync();
  }
  // extend_path :261-334
  {
    int first = n;
    for (int i = lane; i < n; i += FSD_LANES)
      if ((path[i].x - F.px) * F.dx + (path[i].y - F.py) * F.dy > 0.0) {
        first = i;
        break;
      }
    first = wmin_i(first);
    int start = n - 20 < 0 ? 0 : n - 20;
    if (first < start) start = first;
    const int nf = n - start;
    if (nf < 2) return RC_RAISES;
    double part = 0.0;
    for (int i = start + lane; i + 1 < n; i += FSD_LANES) {
      double ddx = path[i + 1].x - path[i].x, ddy = path[i + 1].y - path[i].y;
      part += fsqrt(ddx * ddx + ddy * ddy);
    }
    const double plen = wsum(part);
    if (!(plen > P.mpc_len)) {
      const int nr = nf < 20 ? nf : 20;
      const d2 *rel = path + (n - nr);
      double cx, cy, radius;
      circle_fit_warp(rel, nr, cx, cy, radius);
      const double r_use = fmin(fmax(radius, 10.0), 100.0);
      const double lastx = path[n - 1].x, lasty = path[n - 1].y;
      const int room = PCAP - (int)(path - S.pts) - n;
      if (room < 49) {
        *status |= FSD_ST_OVERFLOW;
        return RC_UNSUPPORTED;
      }
      if (r_use < 80.0) {
        d2 p0 = {rel[0].x - cx, rel[0].y - cy}, p1 = {rel[nr / 2].x - cx, rel[nr / 2].y - cy},
           p2 = {rel[nr - 1].x - cx, rel[nr - 1].y - cy};
        const double sg = sgn(orient(p0, p1, p2));
        const double a0 = fsd_atan2(p0.y, p0.x), a1 = a0 + sg * PI;
        const double stepa = fdiv(a1 - a0, 49.0);  // np.linspace(a0, a1) has 50 samples; the first is dropped
        const double r0x = fsd_cos(a0) * r_use, r0y = fsd_sin(a0) * r_use;
        wsync();
        for (int i = 1 + lane; i < 50; i += FSD_LANES) {
          double ang = i == 49 ? a1 : (double)i * stepa + a0;
          path[n + i - 1].x = fsd_cos(ang) * r_use - r0x + lastx;
          path[n + i - 1].y = fsd_sin(ang) * r_use - r0y + lasty;
        }
        n += 49;
      } else {
        double ddx = lastx - path[n - 2].x, ddy = lasty - path[n - 2].y;
        const double nrm = fsqrt(ddx * ddx + ddy * ddy);
        ddx = fdiv(ddx, nrm);
        ddy = fdiv(ddy, nrm);
        wsync();
        for (int i = 1 + lane; i < 30; i += FSD_LANES) {
          path[n + i - 1].x = lastx + ddx * (double)i;
          path[n + i - 1].y = lasty + ddy * (double)i;
        }
        n += 29;
      }
      wsync();
    }
  }
  // remove_path_behind_car :459-465: first point of minimal distance to the car
  int i0;
  {
    double bv = 0.0;
    int bi = -1;
    for (int i = lane; i < n; i += FSD_LANES) {
      double ddx = F.px - path[i].x, ddy = F.py - path[i].y;
      double d = fsqrt(ddx * ddx + ddy * ddy);
      if (bi < 0 || d < bv) {
        bv = d;
        bi = i;
      }
    }
    wargmin(bv, bi);
    i0 = bi;
  }
  // refit_path_for_mpc_with_safety_factor :239-259: evaluate up to u = 1.5 * mpc_path_length
  int nfix = 0;
  const int off = (int)(path - S.pts) + i0;
  int rc = fit_predict(S, S.pts + off, S.u + off, n - i0, P.smoothing, P.predict_every, P.mpc_len * 1.5, S.pts, PCAP,
                       &nfix, status);
  if (rc == RC_RAISES) rc = RC_UNSUPPORTED;  // the reference re-parameterises a (40, 4) array here (latent bug)
  if (rc != RC_OK) return rc;
  // remove_path_not_in_prediction_horizon :467-499
  int keep;
  {
    if (nfix - 1 <= 1) return RC_UNSUPPORTED;
    int first_over = nfix;
    double carry = 0.0;
    for (int base = 0; base < nfix - 1; base += FSD_LANES) {
      const int i = base + lane;
      double d = 0.0;
      if (i < nfix - 1) {
        double ddx = S.pts[i + 1].x - S.pts[i].x, ddy = S.pts[i + 1].y - S.pts[i].y;
        d = fsqrt(ddx * ddx + ddy * ddy);
      }
      double incl = wscan_incl(d) + carry;
      if (i < nfix - 1 && incl > P.mpc_len && i < first_over) first_over = i;
      carry = wlast(incl);
    }
    first_over = wmin_i(first_over);
    keep = first_over >= nfix ? nfix - 1 : first_over;
  }
  *n_trim = keep;
  return parameterize(S, keep, force_P, P, out, P_out, status);
}

// ---- second half of run_path_calculation (core_calculate_path.py:555-575) -----------------------------------------
// The path update sits in S.pts[1 .. 1+nu), S.prev_xy holds the previous path's xy.  Validity check, MPC tail,
// fallbacks.  out: 40 x 4 fp64; grid[0] = P, grid[1] = points entering the last re-fit.

FSD_DEVFN unsigned path_from_update(PathSmem &S, int nu, const FramePose &F, int force_P, const double *prev,
                                    const DevParams &P, double *out, int *grid) {
  const int lane = fsd_lane();
  unsigned status = 0;
  int P_grid = 0, n_trim = 0;
  bool ok = true;
  int rc;
  {
    // overwrite_path_if_it_is_too_far_away :225-237
    double best = INFINITY;
    for (int i = lane; i < nu; i += FSD_LANES) {
      double ddx = F.px - S.pts[1 + i].x, ddy = F.py - S.pts[1 + i].y;
      best = fmin(best, fsqrt(ddx * ddx + ddy * ddy));
    }
    best = wmin_d(best);
    wsync();
    if (best > P.max_valid_dist) {
      status |= FSD_ST_PATH_TOO_FAR;
      for (int i = lane; i < FSD_HORIZON; i += FSD_LANES) S.pts[1 + i] = S.prev_xy[i];
      nu = FSD_HORIZON;
      wsync();
    }
    // do_all_mpc_parameter_calculations, ValueError -> redo with the previous path (:561-570)
    unsigned st = 0;
    rc = mpc_tail(S, nu, F, force_P, P, out, &P_grid, &n_trim, &st);
    if (rc == RC_VALUE_ERROR) {
      status |= FSD_ST_MPC_FAILED;
      st = 0;
      wsync();
      for (int i = lane; i < FSD_HORIZON; i += FSD_LANES) S.pts[1 + i] = S.prev_xy[i];
      wsync();
      rc = mpc_tail(S, FSD_HORIZON, F, force_P, P, out, &P_grid, &n_trim, &st);
    }
    status |= st;
    if (rc != RC_OK) {
      status |= rc == RC_UNSUPPORTED ? FSD_ST_UNSUPPORTED : FSD_ST_REF_RAISES;
      ok = false;
    }
  }
  if (!ok) {
    wsync();
    for (int i = lane; i < FSD_HORIZON * 4; i += FSD_LANES) out[i] = prev[i];
  }
  if (grid && lane == 0) {
    grid[0] = P_grid;
    grid[1] = n_trim;
  }
  wsync();
  return status;
}

// ---- CalculatePath.run_path_calculation (core_calculate_path.py:514-575), global_path is None ------------
// Inputs: the with-virtual cone lists and matches (any memory space).  prev: previous path (40 x 4 fp64).
// out: 40 x 4 fp64.  grid[0] = P, grid[1] = points entering the last re-fit.

FSD_DEVFN unsigned path_frame(PathSmem &S, const d2 *left, int nl, const d2 *right, int nr, const int16_t *l2r,
                              const int16_t *r2l, const FramePose &F, int force_P, const double *prev,
                              const DevParams &P, double *out, int *grid) {
  const int lane = fsd_lane();
  unsigned status = 0;
  for (int i = lane; i < FSD_HORIZON; i += FSD_LANES) {
    S.prev_xy[i].x = prev[4 * i + 1];
    S.prev_xy[i].y = prev[4 * i + 2];
  }
  wsync();
  const d2 *cl = S.prev_xy;
  int ncl = FSD_HORIZON;
  if (nl < 3 && nr < 3) {
    status |= FSD_ST_FEW_CONES;
  } else {
    if (lane == 0) {
      // select_side_to_use :165-183: max over (number of matches, sum of match indices), ties -> left
      int nml = 0, nmr = 0, sl = 0, sr = 0;
      for (int i = 0; i < nl; ++i)
        if (l2r[i] != -1) {
          ++nml;
          sl += l2r[i];
        }
      for (int i = 0; i < nr; ++i)
        if (r2l[i] != -1) {
          ++nmr;
          sr += r2l[i];
        }
      const bool use_left = !(nmr > nml || (nmr == nml && sr > sl));
      const d2 *a = use_left ? left : right, *b = use_left ? right : left;
      const int16_t *mt = use_left ? l2r : r2l;
      const int ns = use_left ? nl : nr;
      int nc = 0;
      // calculate_centerline_points_of_matches :185-205
      for (int i = 0; i < ns; ++i)
        if (mt[i] != -1 && nc < FSD_HORIZON) {
          S.centre[nc].x = (a[i].x + b[mt[i]].x) / 2.0;
          S.centre[nc].y = (a[i].y + b[mt[i]].y) / 2.0;
          ++nc;
        }
      S.si[0] = nc;
    }
    wsync();
    const int nc = S.si[0];
    if (nc < 2) {
      status |= FSD_ST_FEW_MATCHES;
    } else {
      cl = S.centre;
      ncl = nc;
    }
  }
  // fit_matches_as_spline :207-223 (path update lives in S.pts[1..], slot 0 is kept for connect_path_to_car)
  int nu = 0;
  int rc = fit_predict(S, cl, S.u, ncl, P.smoothing, P.predict_every, -1.0, S.pts + 1, PCAP - 1, &nu, &status);
  if (rc == RC_VALUE_ERROR) {
    status |= FSD_ST_FIT1_FAILED;
    rc = fit_predict(S, S.prev_xy, S.u, FSD_HORIZON, P.smoothing, P.predict_every, -1.0, S.pts + 1, PCAP - 1, &nu,
                     &status);
  }
  if (!(rc == RC_OK && nu >= 1)) {
    status |= rc == RC_UNSUPPORTED ? FSD_ST_UNSUPPORTED : FSD_ST_REF_RAISES;
    wsync();
    for (int i = lane; i < FSD_HORIZON * 4; i += FSD_LANES) out[i] = prev[i];
    if (grid && lane == 0) grid[0] = grid[1] = 0;
    wsync();
    return status;
  }
  return status | path_from_update(S, nu, F, force_P, prev, P, out, grid);
}

// ---- initial path of a fresh planner (core_calculate_path.py:103-121) -------------------------------------

FSD_DEVFN unsigned initial_path_frame(PathSmem &S, const DevParams &P, double *out) {
  unsigned status = 0;
  // calculate_almost_straight_path: 40 points of a chord, radius 1000 m, angle pi/50, turned by -pi/2
  const double max_angle = PI / 50.0, radius = 1000.0, stp = max_angle / (FSD_HORIZON - 1);
  const double c = fsd_cos(-PI / 2.0), s = fsd_sin(-PI / 2.0);
  for (int i = fsd_lane(); i < FSD_HORIZON; i += FSD_LANES) {
    double a = i == FSD_HORIZON - 1 ? max_angle : (double)i * stp;
    double px = (fsd_cos(a) - 1.0) * radius, py = fsd_sin(a) * radius;
    S.centre[i].x = px * c - py * s;
    S.centre[i].y = px * s + py * c;
  }
  wsync();
  int nd = 0, Pg = 0;
  int rc = fit_predict(S, S.centre, S.u, FSD_HORIZON, P.smoothing, P.predict_every, -1.0, S.pts, PCAP, &nd, &status);
  if (rc == RC_OK) rc = parameterize(S, nd, 0, P, out, &Pg, &status);
  if (rc != RC_OK) status |= FSD_ST_UNSUPPORTED;
  return status;
}

}  // namespace fsd
