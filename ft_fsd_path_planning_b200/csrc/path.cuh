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

#ifndef FSD_PCAP
#define FSD_PCAP 704
#endif
constexpr int PCAP = FSD_PCAP;  // path points held per frame (fallback path: 62.8 m / 0.1 m + extension)
constexpr int GRID_CAP = 128;  // size of the last evaluation grid (P = 120 or 121 at run time)

enum { RC_OK = 0, RC_VALUE_ERROR = 1, RC_RAISES = 2, RC_UNSUPPORTED = 3 };

// per-frame point buffers live in per-CTA global scratch (L2 resident): PCAP x (16 + 8) bytes
constexpr size_t PATH_SCRATCH_BYTES = (size_t)PCAP * (sizeof(d2) + sizeof(double));

// states and state of the path machine (the pipeline itself: further down)
enum {
  PS_FIT1 = 1,
  PS_FIT1_DONE = 2,  // alignment point
  PS_TAIL = 3,
  PS_FIT2 = 4,
  PS_FIT2_DONE = 5,  // alignment point
  PS_FIT3 = 6,
  PS_FIT3_DONE = 7,  // alignment point
  PS_DONE = 100
};

struct PathMachine {
  int state, mode;  // mode 0: planner frame, 1: initial path (fit #1 on the chord, then the parameterisation only)
  FitState fit;
  FramePose F;
  const double *prev;  // previous path, 40 x 4
  double *out;         // 40 x 4
  int force_P, nu, P_grid, n_trim;
  unsigned status, tail_status;
  bool fit1_retry, tail_retry;
  double predict_every;
};

struct PathSmem {
  d2 *pts;       // [pcap]
  double *u;     // [pcap]
  int32_t pcap;  // path points behind pts / u (PCAP unless the caller provides larger buffers)
  int32_t pad_[3];
  // The frame's path machine lives HERE, in the frame's shared-memory slot, not on the stack: passed by reference to the
  // out-of-line stages a stack object is local memory -- 32 identical copies per warp whose stores are written through to
  // L2 (most of the path kernel's HBM write-back before r2_zb).  One copy per lane group: every lane stores the same value.
  PathMachine M;
  SplineWork W;  // LAST member (its arena may extend past the struct, see SplineWork::cap)
};

// a frame's working memory: the point buffers (`pcap` points in one block: pts, then u) and the spline arena (`cap`
// records, at least NCAP of them inside S itself)
FSD_DEV void path_smem_bind(PathSmem &S, unsigned char *points, int pcap, int cap, int suspendable = 0) {
  if (PG::lane() == 0) {
    S.pts = reinterpret_cast<d2 *>(points);
    S.u = reinterpret_cast<double *>(points + (size_t)pcap * sizeof(d2));
    S.pcap = pcap;
    S.W.cap = cap;
    S.W.suspendable = suspendable;  // a fit that outgrows `cap` stops as FIT_SUSPENDED instead of being truncated
  }
  PG::sync();
}

// ---- hyper circle fit -----------------------------------------------------------------------------

// Returns the radius BY VALUE and writes the centre through `centre` only when asked to: reference outputs of an out-of-line
// function are local memory, and the 120 curvature windows of every frame need nothing but the radius.
FSD_DEVFN double hyper_from_moments(double mx, double my, double Mxx, double Myy, double Mxy, double Mxz, double Myz,
                                  double Mzz, d2 *centre) {
  const double Mz = Mxx + Myy, Cov_xy = Mxx * Myy - Mxy * Mxy, Var_z = Mzz - Mz * Mz;
  const double A2 = 4.0 * Cov_xy - 3.0 * Mz * Mz - Mzz;
  const double A1 = Var_z * Mz + 4.0 * Cov_xy * Mz - Mxz * Mxz - Myz * Myz;
  const double A0 = Mxz * (Mxz * Myy - Myz * Mxy) + Myz * (Myz * Mxx - Mxz * Mxy) - Var_z * Cov_xy;
  const double A22 = A2 + A2;
  double y = A0, x = 0.0;
#pragma unroll 1
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
  if (centre) {
    centre->x = Xc + mx;
    centre->y = Yc + my;
  }
  return fsqrt(fabs(Xc * Xc + Yc * Yc + Mz));
}

FSD_DEV double orient(const d2 &p0, const d2 &p1, const d2 &p2) {
  // sign of det [[1, p0], [1, p1], [1, p2]] (np.linalg.det in the reference)
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

// Curvature windows (calculate_path_curvature :49-93, open path): one hyper circle fit per path point over the points
// within +-hw of it.  Neighbouring windows share all but two points, so a lane takes a run of consecutive windows and
// SLIDES the raw moment sums (10 monomials up to degree 4, in coordinates local to the run: |x|, |y| < 3 m, so the
// shift to the window mean below cancels a few digits at most) instead of re-reading 25 points per window: 31 point
// visits per lane instead of 100.  curv[i] = sign / clamp(radius, 1, 3000).
FSD_DEVFN void curvature_windows(const d2 *p, int Pn, int hw, double *curv) {
  const int per = (Pn + PG::N - 1) / PG::N;
  const int i0 = PG::lane() * per, i1 = i0 + per < Pn ? i0 + per : Pn;
  if (i0 >= i1) return;
  const double ox = p[i0].x, oy = p[i0].y;
  double s1x = 0, s1y = 0, sxx = 0, sxy = 0, syy = 0, sxxx = 0, sxxy = 0, sxyy = 0, syyy = 0, sz2 = 0;
  int lo = i0 - hw < 0 ? 0 : i0 - hw, hi = lo - 1;  // the window [lo, hi] covered by the sums (empty)
#pragma unroll 1
  for (int i = i0; i < i1; ++i) {
    const int nlo = i - hw < 0 ? 0 : i - hw, nhi = i + hw > Pn - 1 ? Pn - 1 : i + hw;
#pragma unroll 1
    while (hi < nhi || lo < nlo) {
      // one point enters (sign +1) or leaves (sign -1)
      const bool enter = hi < nhi;
      const int q = enter ? ++hi : lo++;
      const double sg = enter ? 1.0 : -1.0;
      const double x = p[q].x - ox, y = p[q].y - oy;
      const double xx = x * x, yy = y * y, z = xx + yy;
      const double sx = sg * x, sy = sg * y;
      s1x += sx;
      s1y += sy;
      sxx += sx * x;
      sxy += sx * y;
      syy += sy * y;
      sxxx += sx * xx;
      sxxy += sy * xx;
      sxyy += sx * yy;
      syyy += sy * yy;
      sz2 += sg * z * z;
    }
    // central moments about the window mean from the raw sums
    const double n = (double)(nhi - nlo + 1), inv = frcp(n);
    const double mx = s1x * inv, my = s1y * inv;
    const double Mxx = sxx * inv - mx * mx, Myy = syy * inv - my * my, Mxy = sxy * inv - mx * my;
    const double n2 = n + n;
    const double c3x = sxxx - 3.0 * mx * sxx + n2 * mx * mx * mx;
    const double cxxy = sxxy - my * sxx - 2.0 * mx * sxy + n2 * mx * mx * my;
    const double cxyy = sxyy - mx * syy - 2.0 * my * sxy + n2 * mx * my * my;
    const double c3y = syyy - 3.0 * my * syy + n2 * my * my * my;
    const double Mxz = (c3x + cxyy) * inv, Myz = (cxxy + c3y) * inv;
    const double q2 = mx * mx + my * my;
    const double sl2 = mx * mx * sxx + 2.0 * mx * my * sxy + my * my * syy;
    const double szl = mx * (sxxx + sxyy) + my * (sxxy + syyy);
    const double Mzz = (sz2 + 4.0 * sl2 - 4.0 * szl + 2.0 * q2 * (sxx + syy) - 3.0 * n * q2 * q2) * inv;
    double r = hyper_from_moments(mx, my, Mxx, Myy, Mxy, Mxz, Myz, Mzz, nullptr);
    r = fmin(fmax(r, 1.0), 3000.0);
    const int cnt = nhi - nlo + 1;
    curv[i] = frcp(r) * sgn(orient(p[nlo], p[nlo + cnt / 2], p[nhi]));
  }
}

// the whole warp fits one point set (path extension)
FSD_DEVFN void circle_fit_warp(const d2 *p, int n, double &cx, double &cy, double &r) {
  double sx = 0.0, sy = 0.0;
#pragma unroll 1
  for (int i = PG::lane(); i < n; i += PG::N) {
    sx += p[i].x;
    sy += p[i].y;
  }
  const double inv_n = frcp((double)n);
  const double mx = PG::sum(sx) * inv_n, my = PG::sum(sy) * inv_n;
  double Mxy = 0, Mxx = 0, Myy = 0, Mxz = 0, Myz = 0, Mzz = 0;
#pragma unroll 1
  for (int i = PG::lane(); i < n; i += PG::N) {
    double xi = p[i].x - mx, yi = p[i].y - my, zi = xi * xi + yi * yi;
    Mxy += xi * yi;
    Mxx += xi * xi;
    Myy += yi * yi;
    Mxz += xi * zi;
    Myz += yi * zi;
    Mzz += zi * zi;
  }
  Mxy = PG::sum(Mxy) * inv_n;
  Mxx = PG::sum(Mxx) * inv_n;
  Myy = PG::sum(Myy) * inv_n;
  Mxz = PG::sum(Mxz) * inv_n;
  Myz = PG::sum(Myz) * inv_n;
  Mzz = PG::sum(Mzz) * inv_n;
  d2 c;
  r = hyper_from_moments(mx, my, Mxx, Myy, Mxy, Mxz, Myz, Mzz, &c);
  cx = c.x;
  cy = c.y;
}

// ---- chord-length parameters: u[0] = 0, u[i] = u[i-1] + |p_i - p_{i-1}| (np.cumsum) -------------------

FSD_DEVFN void chord_params(const d2 *p, int m, double *u) {
  const int lane = PG::lane();
  double carry = 0.0;
  if (lane == 0) u[0] = 0.0;
#pragma unroll 1
  for (int base = 1; base < m; base += PG::N) {
    const int i = base + lane;
    double d = 0.0;
    if (i < m) {
      double ddx = p[i].x - p[i - 1].x, ddy = p[i].y - p[i - 1].y;
      d = fsqrt(ddx * ddx + ddy * ddy);
    }
    double incl = PG::scan_incl(d) + carry;
    if (i < m) u[i] = incl;
    carry = PG::last(incl);
  }
  PG::sync();
}

// ---- the path pipeline as a resumable state machine -----------------------------------------------------------------
// One frame's path calculation = fit #1 -> validity / connect / extend / cut -> fit #2 -> trim -> fit #3 -> curvature
// and sampling, with the reference's fallbacks (re-fit of the previous path, redo of the tail with the previous
// path) as backward transitions.  pm_step advances ONE unit: a pass of a spline fit (spline.cuh) or one stage between
// fits.  The path kernel keeps the warps of a CTA in lockstep at this granularity (they wait for each other at the
// *_DONE states), the blocking wrappers below (path_frame, path_from_update, initial_path_frame) simply run the
// machine to completion.  Every lane holds an identical copy of the machine state.

FSD_DEV bool pm_is_alignment_state(int st) { return st == PS_FIT1_DONE || st == PS_FIT2_DONE || st == PS_FIT3_DONE; }

// x, y of the previous path (40 x 4, global memory) into dst[0 .. 40): only the fallbacks need them, so they are read
// when a fallback fires instead of being kept in shared memory
FSD_DEVFN void pm_prev_to(const PathMachine &M, d2 *dst) {
  PG::sync();
#pragma unroll 1
  for (int i = PG::lane(); i < FSD_HORIZON; i += PG::N) {
    dst[i].x = M.prev[4 * i + 1];
    dst[i].y = M.prev[4 * i + 2];
  }
  PG::sync();
}

// evaluate the fitted spline every `step` up to max_u into dst (len(np.arange(0, max_u, step)) points)
FSD_DEVFN int pm_evaluate(PathSmem &S, double max_u, double step, d2 *dst, int dst_cap) {
  const double q = ceil(fdiv(max_u, step));
  const int n = q > 0.0 ? (q > 1e6 ? 1000000 : (int)q) : 0;
  if (n > dst_cap) return -1;
  PG::sync();
#pragma unroll 1
  for (int i = PG::lane(); i < n; i += PG::N) {
    const d2 p = spline_point(S.W, (double)i * step);
    dst[i].x = p.x;
    dst[i].y = p.y;
  }
  PG::sync();
  return n;
}

FSD_DEVFN void pm_finish_with_prev(PathMachine &M, unsigned bits) {
  M.status |= bits;
  PG::sync();
#pragma unroll 1
  for (int i = PG::lane(); i < FSD_HORIZON * 4; i += PG::N) M.out[i] = M.prev[i];
  M.P_grid = 0;
  M.n_trim = 0;
  M.state = PS_DONE;
  PG::sync();
}

// the tail failed with return code rc: ValueError -> once more with the previous path (:561-570), else give up
FSD_DEVFN void pm_tail_failed(PathSmem &S, PathMachine &M, int rc) {
  if (M.mode == 1) {
    M.status |= FSD_ST_UNSUPPORTED;
    M.state = PS_DONE;
    return;
  }
  if (rc == RC_VALUE_ERROR && !M.tail_retry) {
    M.status |= FSD_ST_MPC_FAILED;
    M.tail_retry = true;
    M.tail_status = 0;
    pm_prev_to(M, S.pts + 1);
    M.nu = FSD_HORIZON;
    M.state = PS_TAIL;
    return;
  }
  pm_finish_with_prev(M, M.tail_status | (rc == RC_UNSUPPORTED ? FSD_ST_UNSUPPORTED : FSD_ST_REF_RAISES));
}

FSD_DEVFN void pm_start_fit1(PathSmem &S, PathMachine &M, const d2 *src, int m, const DevParams &P) {
  if (m >= 2) chord_params(src, m, S.u);
  fit_init(S.W, M.fit, src, S.u, m, P.smoothing);
  M.state = PS_FIT1;
}

// overwrite_path_if_it_is_too_far_away :225-237 on the path update S.pts[1 .. 1+nu), then the tail
FSD_DEVFN void pm_enter_tail(PathSmem &S, PathMachine &M, const DevParams &P) {
  const int lane = PG::lane();
  double best = INFINITY;
#pragma unroll 1
  for (int i = lane; i < M.nu; i += PG::N) {  // the smallest distance to the car, squared (only compared)
    const double ddx = M.F.px - S.pts[1 + i].x, ddy = M.F.py - S.pts[1 + i].y;
    best = fmin(best, ddx * ddx + ddy * ddy);
  }
  best = PG::min_d(best);
  PG::sync();
  if (best > P.max_valid_dist * P.max_valid_dist) {
    M.status |= FSD_ST_PATH_TOO_FAR;
    pm_prev_to(M, S.pts + 1);
    M.nu = FSD_HORIZON;
  }
  M.tail_status = 0;
  M.state = PS_TAIL;
}

// _refit_spline :125-161 on S.pts[0..n): path length, sub-sampling, start of fit #3
FSD_DEVFN void pm_start_fit3(PathSmem &S, PathMachine &M, int n, const DevParams &P) {
  const int lane = PG::lane();
  if (n < 2) {
    pm_tail_failed(S, M, RC_RAISES);
    return;
  }
  double len = 0.0, first10 = 0.0;
#pragma unroll 1
  for (int i = lane; i + 1 < n; i += PG::N) {
    const double d = fnorm(S.pts[i + 1].x - S.pts[i].x, S.pts[i + 1].y - S.pts[i].y);
    len += d;
    if (i < 10) first10 += d;
  }
  const double path_length = PG::sum(len);
  const int nm = n - 1 < 10 ? n - 1 : 10;
  const double mean_dist = fdiv(PG::sum(first10), (double)nm);
  M.predict_every = fdiv(fdiv(path_length, (double)FSD_HORIZON), 3.0);
  const double ratio = fdiv(M.predict_every, mean_dist);
  int skip = 1;
  if (isfinite(ratio) && ratio < 1e6 && (int)ratio > 1) skip = (int)ratio;
  const int ms = (n + skip - 1) / skip;
  PG::sync();
  if (skip > 1) {
    if (lane == 0)
      for (int i = 1; i < ms; ++i) S.pts[i] = S.pts[i * skip];  // path[::skip]
    PG::sync();
  }
  chord_params(S.pts, ms, S.u);
  fit_init(S.W, M.fit, S.pts, S.u, ms, P.refit_smoothing);
  M.state = PS_FIT3;
}

// connect_path_to_car, extend_path, remove_path_behind_car (core_calculate_path.py:430-465, 261-334), start of fit #2
FSD_DEVFN void pm_stage_tail(PathSmem &S, PathMachine &M, const DevParams &P) {
  const int lane = PG::lane();
  const FramePose &F = M.F;
  if (M.nu < 1) {
    pm_tail_failed(S, M, RC_RAISES);
    return;
  }
  d2 *path = S.pts + 1;
  int n = M.nu;
  {
    const double fx = path[0].x - F.px, fy = path[0].y - F.py;
    const double d = fnorm(fx, fy);
    const bool behind = cos_between(fx, fy, F.dx, F.dy) < 0.0;  // angle > pi/2
    PG::sync();
    if (!(d < 0.5 || behind)) {
      if (lane == 0) {
        S.pts[0].x = F.px + fdiv(fx, d) * 0.2;
        S.pts[0].y = F.py + fdiv(fy, d) * 0.2;
      }
      path = S.pts;
      n = M.nu + 1;
    }
    PG::sync();
  }
  {
    int first = n;
#pragma unroll 1
    for (int i = lane; i < n; i += PG::N)
      if ((path[i].x - F.px) * F.dx + (path[i].y - F.py) * F.dy > 0.0) {
        first = i;
        break;
      }
    first = PG::min_i(first);
    int start = n - 20 < 0 ? 0 : n - 20;
    if (first < start) start = first;
    const int nf = n - start;
    if (nf < 2) {
      pm_tail_failed(S, M, RC_RAISES);
      return;
    }
    double part = 0.0;
#pragma unroll 1
    for (int i = start + lane; i + 1 < n; i += PG::N) part += fnorm(path[i + 1].x - path[i].x, path[i + 1].y - path[i].y);
    const double plen = PG::sum(part);
    if (!(plen > P.mpc_len)) {
      const int nr = nf < 20 ? nf : 20;
      const d2 *rel = path + (n - nr);
      double cx, cy, radius;
      circle_fit_warp(rel, nr, cx, cy, radius);
      const double r_use = fmin(fmax(radius, 10.0), 100.0);
      const double lastx = path[n - 1].x, lasty = path[n - 1].y;
      const int room = S.pcap - (int)(path - S.pts) - n;
      if (room < 49) {
        M.tail_status |= FSD_ST_OVERFLOW;
        pm_tail_failed(S, M, RC_UNSUPPORTED);
        return;
      }
      if (r_use < 80.0) {
        d2 p0 = {rel[0].x - cx, rel[0].y - cy}, p1 = {rel[nr / 2].x - cx, rel[nr / 2].y - cy},
           p2 = {rel[nr - 1].x - cx, rel[nr - 1].y - cy};
        const double sg = sgn(orient(p0, p1, p2));
        const double a0 = fsd_atan2(p0.y, p0.x), a1 = a0 + sg * PI;
        const double stepa = fdiv(a1 - a0, 49.0);  // np.linspace(a0, a1) has 50 samples; the first is dropped
        const double r0x = fsd_cos(a0) * r_use, r0y = fsd_sin(a0) * r_use;
        PG::sync();
#pragma unroll 1
        for (int i = 1 + lane; i < 50; i += PG::N) {
          const double ang = i == 49 ? a1 : (double)i * stepa + a0;
          path[n + i - 1].x = fsd_cos(ang) * r_use - r0x + lastx;
          path[n + i - 1].y = fsd_sin(ang) * r_use - r0y + lasty;
        }
        n += 49;
      } else {
        double ddx = lastx - path[n - 2].x, ddy = lasty - path[n - 2].y;
        const double nrm = fnorm(ddx, ddy);
        ddx = fdiv(ddx, nrm);
        ddy = fdiv(ddy, nrm);
        PG::sync();
#pragma unroll 1
        for (int i = 1 + lane; i < 30; i += PG::N) {
          path[n + i - 1].x = lastx + ddx * (double)i;
          path[n + i - 1].y = lasty + ddy * (double)i;
        }
        n += 29;
      }
      PG::sync();
    }
  }
  // first point of minimal distance to the car
  double bv = 0.0;
  int bi = -1;
#pragma unroll 1
  for (int i = lane; i < n; i += PG::N) {
    const double ddx = F.px - path[i].x, ddy = F.py - path[i].y;
    const double d = ddx * ddx + ddy * ddy;  // arg-min of the distance: squared
    if (bi < 0 || d < bv) {
      bv = d;
      bi = i;
    }
  }
  PG::argmin(bv, bi);
  // refit_path_for_mpc_with_safety_factor :239-259
  const int off = (int)(path - S.pts) + bi, m2 = n - bi;
  if (m2 < 2) {
    pm_tail_failed(S, M, RC_UNSUPPORTED);  // the reference re-parameterises a (40, 4) array here (latent bug)
    return;
  }
  chord_params(S.pts + off, m2, S.u + off);
  fit_init(S.W, M.fit, S.pts + off, S.u + off, m2, P.smoothing);
  M.state = PS_FIT2;
}

// fit #2 is done: evaluate up to 1.5 x the MPC length, keep the first mpc_path_length metres (:467-499), start fit #3
FSD_DEVFN void pm_stage_after_fit2(PathSmem &S, PathMachine &M, const DevParams &P) {
  const int lane = PG::lane();
  if (M.fit.ier == 10) {
    pm_tail_failed(S, M, (M.tail_status & FSD_ST_UNSUPPORTED) ? RC_UNSUPPORTED : RC_VALUE_ERROR);
    return;
  }
  const int nfix = pm_evaluate(S, P.mpc_len * 1.5, P.predict_every, S.pts, S.pcap);
  if (nfix < 0 || nfix - 1 <= 1) {
    if (nfix < 0) M.tail_status |= FSD_ST_OVERFLOW;
    pm_tail_failed(S, M, RC_UNSUPPORTED);
    return;
  }
  int first_over = nfix;
  double carry = 0.0;
#pragma unroll 1
  for (int base = 0; base < nfix - 1; base += PG::N) {
    const int i = base + lane;
    double d = 0.0;
    if (i < nfix - 1) d = fnorm(S.pts[i + 1].x - S.pts[i].x, S.pts[i + 1].y - S.pts[i].y);
    const double incl = PG::scan_incl(d) + carry;
    if (i < nfix - 1 && incl > P.mpc_len && i < first_over) first_over = i;
    carry = PG::last(incl);
  }
  first_over = PG::min_i(first_over);
  M.n_trim = first_over >= nfix ? nfix - 1 : first_over;
  pm_start_fit3(S, M, M.n_trim, P);
}

// fit #3 is done: evaluation grid, curvature, 40 samples (path_parameterization.py:163-295)
FSD_DEVFN void pm_stage_after_fit3(PathSmem &S, PathMachine &M, const DevParams &P) {
  const int lane = PG::lane();
  if (M.fit.ier == 10) {
    pm_tail_failed(S, M, (M.tail_status & FSD_ST_UNSUPPORTED) ? RC_UNSUPPORTED : RC_VALUE_ERROR);
    return;
  }
  const double predict_every = M.predict_every;
  // size of the evaluation grid np.arange(0, max_u, predict_every): SURVEY.md Q13
  int Pn;
  if (M.force_P > 0) {
    Pn = M.force_P;
  } else {
    const double q = fdiv(S.W.max_u, predict_every);
    const double r = rint(q);
    if (fabs(q - r) < 1e-9) {
      Pn = (int)r;
      M.tail_status |= FSD_ST_TIE_P;
    } else {
      Pn = q < 1e6 ? (int)ceil(q) : 1000000;
    }
  }
  M.P_grid = Pn;
  if (Pn < FSD_HORIZON) {  // repeated sample indices (:284-285)
    pm_tail_failed(S, M, RC_VALUE_ERROR);
    return;
  }
  if (Pn > GRID_CAP) {
    M.tail_status |= FSD_ST_OVERFLOW;
    pm_tail_failed(S, M, RC_UNSUPPORTED);
    return;
  }
  PG::sync();
#pragma unroll 1
  for (int i = lane; i < Pn; i += PG::N) {
    const d2 p = spline_point(S.W, (double)i * predict_every);
    S.pts[i].x = p.x;
    S.pts[i].y = p.y;
  }
  PG::sync();
  // _calculate_path_curvature :163-193 / calculate_path_curvature :49-93 (open path).  The curvature samples live in the
  // chord-parameter buffer: the fits are over and their data abscissae dead.
  static_assert(PCAP >= GRID_CAP, "curvature samples alias the chord-parameter buffer");
  double *curv = S.u;
  int window = Pn / 5 < 30 ? Pn / 5 : 30;
  if (window % 2 == 0) window += 1;
  const int hw = window / 2;
  curvature_windows(S.pts, Pn, hw, curv);
  PG::sync();
  // uniform_filter1d(size, mode="nearest") evaluated at the 40 sampled indices only;
  // indices np.linspace(0, P-1, 40, dtype=int) (:277-282)
  const int fs = window / 2 > 2 ? window / 2 : 2;
  const double stp = fdiv((double)(Pn - 1), (double)(FSD_HORIZON - 1));
#pragma unroll 1
  for (int j = lane; j < FSD_HORIZON; j += PG::N) {
    const int idx = j == FSD_HORIZON - 1 ? Pn - 1 : (int)floor((double)j * stp);
    double acc = 0.0;
#pragma unroll 1
    for (int q = idx - fs / 2; q <= idx + fs - fs / 2 - 1; ++q) {
      const int qq = q < 0 ? 0 : (q > Pn - 1 ? Pn - 1 : q);
      acc += curv[qq];
    }
    M.out[4 * j + 0] = (double)idx * predict_every;
    M.out[4 * j + 1] = S.pts[idx].x;
    M.out[4 * j + 2] = S.pts[idx].y;
    M.out[4 * j + 3] = fdiv(acc, (double)fs);
  }
  PG::sync();
  M.status |= M.tail_status;
  M.state = PS_DONE;
}

// fit #1 is done (fit_matches_as_spline :207-223): evaluate the path update every predict_every metres
FSD_DEVFN void pm_stage_after_fit1(PathSmem &S, PathMachine &M, const DevParams &P) {
  if (M.fit.ier == 10) {
    const bool unsupported = (M.status & FSD_ST_UNSUPPORTED) != 0;
    if (!unsupported && !M.fit1_retry && M.mode == 0) {
      M.status |= FSD_ST_FIT1_FAILED;  // ValueError -> the previous path is fitted instead
      M.fit1_retry = true;
      pm_prev_to(M, S.pts);  // the point buffer is free until the fit has been evaluated
      pm_start_fit1(S, M, S.pts, FSD_HORIZON, P);
      return;
    }
    if (M.mode == 1) {
      M.status |= FSD_ST_UNSUPPORTED;
      M.state = PS_DONE;
    } else {
      pm_finish_with_prev(M, unsupported ? FSD_ST_UNSUPPORTED : FSD_ST_REF_RAISES);
    }
    return;
  }
  if (M.mode == 1) {
    const int nd = pm_evaluate(S, S.W.max_u, P.predict_every, S.pts, S.pcap);
    if (nd < 0) {
      M.status |= FSD_ST_UNSUPPORTED | FSD_ST_OVERFLOW;
      M.state = PS_DONE;
      return;
    }
    M.tail_status = 0;
    pm_start_fit3(S, M, nd, P);
    return;
  }
  // the path update lives in S.pts[1..], slot 0 is kept for connect_path_to_car
  const int nu = pm_evaluate(S, S.W.max_u, P.predict_every, S.pts + 1, S.pcap - 1);
  if (nu < 1) {
    pm_finish_with_prev(M, nu < 0 ? (FSD_ST_UNSUPPORTED | FSD_ST_OVERFLOW) : FSD_ST_REF_RAISES);
    return;
  }
  M.nu = nu;
  pm_enter_tail(S, M, P);
}

// advance the machine by one unit
FSD_DEVFN void pm_step(PathSmem &S, PathMachine &M, const DevParams &P) {
  switch (M.state) {
    case PS_FIT1:
    case PS_FIT2:
    case PS_FIT3:
#ifdef FSD_LOCKSTEP_PER_STEP
      if (M.fit.phase != FIT_DONE) fit_step(S.W, M.fit, M.state == PS_FIT1 ? &M.status : &M.tail_status);
#else
      // the whole fit in one machine step (the kernel aligns its warps at the fit boundaries only)
      if (M.fit.phase != FIT_DONE) fit_run(S.W, M.fit, M.state == PS_FIT1 ? &M.status : &M.tail_status);
#endif
      if (M.fit.phase == FIT_DONE) {
        // (the machine may live in shared memory, one copy for all lanes: every lane reads before any lane writes)
        const int next = M.state + 1;
        PG::sync();
        M.state = next;
      }
      break;
    case PS_FIT1_DONE:
      pm_stage_after_fit1(S, M, P);
      break;
    case PS_TAIL:
      pm_stage_tail(S, M, P);
      break;
    case PS_FIT2_DONE:
      pm_stage_after_fit2(S, M, P);
      break;
    case PS_FIT3_DONE:
      pm_stage_after_fit3(S, M, P);
      break;
    default:
      break;
  }
}

FSD_DEVFN void pm_init(PathSmem &S, PathMachine &M, int mode, const FramePose &F, int force_P, const double *prev,
                       double *out) {
  M.mode = mode;
  M.F = F;
  M.force_P = force_P;
  M.prev = prev;
  M.out = out;
  M.status = 0;
  M.tail_status = 0;
  M.fit1_retry = false;
  M.tail_retry = false;
  M.nu = 0;
  M.P_grid = 0;
  M.n_trim = 0;
  M.predict_every = 0.0;
  M.fit.phase = FIT_DONE;
  M.fit.ier = 10;
  M.state = PS_DONE;
}

// ---- CalculatePath.run_path_calculation (core_calculate_path.py:514-575), global_path is None ------------
// Inputs: the with-virtual cone lists and matches (any memory space).  prev: previous path (40 x 4 fp64).
// Sets the machine up to the start of fit #1.

FSD_DEVFN void pm_begin_frame(PathSmem &S, PathMachine &M, const d2 *left, int nl, const d2 *right, int nr,
                              const int16_t *l2r, const int16_t *r2l, const FramePose &F, int force_P,
                              const double *prev, const DevParams &P, double *out) {
  const int lane = PG::lane();
  pm_init(S, M, 0, F, force_P, prev, out);
  // the data of fit #1 (centre line of the matches, or the previous path) lives at the start of the point buffer, which
  // is free until the fit has been evaluated
  int ncl = 0;
  if (nl < 3 && nr < 3) {
    M.status |= FSD_ST_FEW_CONES;
  } else {
    // select_side_to_use :165-183: max over (number of matches, sum of match indices), ties -> left; one cone per lane
    int nml = 0, nmr = 0, sl = 0, sr = 0;
#pragma unroll 1
    for (int base = 0; base < nl || base < nr; base += PG::N) {
      const int i = base + lane;
      const int ml = i < nl ? (int)l2r[i] : -1, mr = i < nr ? (int)r2l[i] : -1;
      nml += FSD_POPC(PG::ballot(ml != -1));
      nmr += FSD_POPC(PG::ballot(mr != -1));
      sl += PG::sum_i(ml != -1 ? ml : 0);
      sr += PG::sum_i(mr != -1 ? mr : 0);
    }
    const bool use_left = !(nmr > nml || (nmr == nml && sr > sl));
    const d2 *a = use_left ? left : right, *b = use_left ? right : left;
    const int16_t *mt = use_left ? l2r : r2l;
    const int ns = use_left ? nl : nr;
    // calculate_centerline_points_of_matches :185-205: the matched cones in order, compacted by ballot
    int nc = 0;
#pragma unroll 1
    for (int base = 0; base < ns; base += PG::N) {
      const int i = base + lane;
      const int m = i < ns ? (int)mt[i] : -1;
      const unsigned mask = PG::ballot(m != -1);
      const int slot = nc + FSD_POPC(mask & ((1u << lane) - 1u));
      if (m != -1 && slot < FSD_HORIZON) {
        S.pts[slot].x = (a[i].x + b[m].x) / 2.0;
        S.pts[slot].y = (a[i].y + b[m].y) / 2.0;
      }
      nc += FSD_POPC(mask);
    }
    if (nc > FSD_HORIZON) nc = FSD_HORIZON;
    PG::sync();
    if (nc < 2)
      M.status |= FSD_ST_FEW_MATCHES;
    else
      ncl = nc;
  }
  if (ncl == 0) {
    pm_prev_to(M, S.pts);
    ncl = FSD_HORIZON;
  }
  pm_start_fit1(S, M, S.pts, ncl, P);
}

// run_path_calculation with a GLOBAL PATH (core_calculate_path.py:516-528; PathPlanner.set_global_path, and the
// acceleration mission's known map): the centre line is the part of the global path within 30 m of the car, in the order
// of np.roll(path, -argmin + M / 3), i.e. starting M / 3 points before the closest one.  Distances lane-strided, the
// closest point by a warp arg-min (first minimum, like np.argmin), the kept points compacted by ballot in rolled order.
FSD_DEVFN void pm_begin_global(PathSmem &S, PathMachine &M, const double *gpath, int Mn, const FramePose &F, int force_P,
                               const double *prev, const DevParams &P, double *out) {
  const int lane = PG::lane();
  pm_init(S, M, 0, F, force_P, prev, out);
  double bv = 0.0;
  int bi = -1;
#pragma unroll 1
  for (int i = lane; i < Mn; i += PG::N) {
    const double d = fnorm(F.px - gpath[2 * i], F.py - gpath[2 * i + 1]);
    if (bi < 0 || d < bv) {
      bv = d;
      bi = i;
    }
  }
  PG::argmin(bv, bi);
  int ncl = 0;
  bool overflow = false;
  const int third = Mn / 3;
#pragma unroll 1
  for (int base = 0; base < Mn; base += PG::N) {
    const int i = base + lane;
    bool keep = false;
    double x = 0.0, y = 0.0;
    if (i < Mn) {
      int src = (i + bi - third) % Mn;
      if (src < 0) src += Mn;
      x = gpath[2 * src];
      y = gpath[2 * src + 1];
      keep = fnorm(F.px - x, F.py - y) < 30.0;
    }
    const unsigned mask = PG::ballot(keep);
    const int slot = ncl + FSD_POPC(mask & ((1u << lane) - 1u));
    if (keep) {
      if (slot < S.pcap) {
        S.pts[slot].x = x;
        S.pts[slot].y = y;
      } else {
        overflow = true;
      }
    }
    ncl += FSD_POPC(mask);
  }
  if (PG::any(overflow)) {
    // more points within 30 m than the point buffer holds: flagged, planned with the previous path
    pm_finish_with_prev(M, FSD_ST_OVERFLOW | FSD_ST_UNSUPPORTED);
    return;
  }
  PG::sync();
  pm_start_fit1(S, M, S.pts, ncl, P);
}

// second half of run_path_calculation only: the path update already sits in S.pts[1 .. 1+nu) (skidpad)
FSD_DEVFN void pm_begin_update(PathSmem &S, PathMachine &M, int nu, const FramePose &F, int force_P, const double *prev,
                               const DevParams &P, double *out) {
  pm_init(S, M, 0, F, force_P, prev, out);
  M.nu = nu;
  pm_enter_tail(S, M, P);
}

// the constant path of a fresh planner (core_calculate_path.py:103-121): fit of the "almost straight" chord
// (path_calculator_helpers.py:26-68: 40 points, radius 1000 m, angle pi/50, turned by -pi/2), then the parameterisation
FSD_DEVFN void pm_begin_initial(PathSmem &S, PathMachine &M, const DevParams &P, double *out) {
  FramePose F = {0, 0, 1, 0, 1, 0};
  pm_init(S, M, 1, F, 0, nullptr, out);
  const double max_angle = PI / 50.0, radius = 1000.0, stp = max_angle / (FSD_HORIZON - 1);
  const double c = fsd_cos(-PI / 2.0), s = fsd_sin(-PI / 2.0);
#pragma unroll 1
  for (int i = PG::lane(); i < FSD_HORIZON; i += PG::N) {
    const double a = i == FSD_HORIZON - 1 ? max_angle : (double)i * stp;
    const double px = (fsd_cos(a) - 1.0) * radius, py = fsd_sin(a) * radius;
    S.pts[i].x = px * c - py * s;
    S.pts[i].y = px * s + py * c;
  }
  PG::sync();
  pm_start_fit1(S, M, S.pts, FSD_HORIZON, P);
}

// a fit of the machine is waiting for a larger arena (only when the caller set SplineWork::suspendable)
FSD_DEV bool pm_suspended(const PathMachine &M) { return M.state != PS_DONE && M.fit.phase == FIT_SUSPENDED; }

FSD_DEVFN void pm_run(PathSmem &S, PathMachine &M, const DevParams &P) {
#pragma unroll 1
  while (M.state != PS_DONE && !pm_suspended(M)) pm_step(S, M, P);
}

// ---- blocking wrappers -------------------------------------------------------------------------------------------

// a suspended fit (SplineWork::suspendable) that nobody resumes: flagged like a truncated one, output = previous path
FSD_DEVFN void pm_give_up_suspended(PathMachine &M) {
  if (pm_suspended(M)) pm_finish_with_prev(M, FSD_ST_OVERFLOW | FSD_ST_UNSUPPORTED);
}

// xcap: knot records really available behind S.W.r when the caller set S.W.suspendable and holds a larger arena in
// reserve than S.W.cap says (the host-check build does, to exercise the suspend / resume path of path_kernel): a
// suspended fit is resumed with them.
FSD_DEVFN unsigned path_frame(PathSmem &S, const d2 *left, int nl, const d2 *right, int nr, const int16_t *l2r,
                              const int16_t *r2l, const FramePose &F, int force_P, const double *prev,
                              const DevParams &P, double *out, int *grid, int xcap = 0) {
  PathMachine &M = S.M;
  pm_begin_frame(S, M, left, nl, right, nr, l2r, r2l, F, force_P, prev, P, out);
  pm_run(S, M, P);
  if (pm_suspended(M) && xcap > S.W.cap) {
    const int cap0 = S.W.cap;
    PG::sync();
    if (PG::lane() == 0) {
      S.W.cap = xcap;
      S.W.suspendable = 0;
    }
    PG::sync();
    fit_resume(S.W, M.fit);
    pm_run(S, M, P);
    PG::sync();
    if (PG::lane() == 0) {
      S.W.cap = cap0;
      S.W.suspendable = 1;
    }
    PG::sync();
  }
  pm_give_up_suspended(M);
  if (grid && PG::lane() == 0) {
    grid[0] = M.P_grid;
    grid[1] = M.n_trim;
  }
  PG::sync();
  return M.status;
}

FSD_DEVFN unsigned path_global(PathSmem &S, const double *gpath, int Mn, const FramePose &F, int force_P,
                               const double *prev, const DevParams &P, double *out, int *grid) {
  PathMachine &M = S.M;
  pm_begin_global(S, M, gpath, Mn, F, force_P, prev, P, out);
  pm_run(S, M, P);
  if (grid && PG::lane() == 0) {
    grid[0] = M.P_grid;
    grid[1] = M.n_trim;
  }
  PG::sync();
  return M.status;
}

FSD_DEVFN unsigned path_from_update(PathSmem &S, int nu, const FramePose &F, int force_P, const double *prev,
                                    const DevParams &P, double *out, int *grid) {
  PathMachine &M = S.M;
  pm_begin_update(S, M, nu, F, force_P, prev, P, out);
  pm_run(S, M, P);
  if (grid && PG::lane() == 0) {
    grid[0] = M.P_grid;
    grid[1] = M.n_trim;
  }
  PG::sync();
  return M.status;
}

FSD_DEVFN unsigned initial_path_frame(PathSmem &S, const DevParams &P, double *out) {
  PathMachine &M = S.M;
  pm_begin_initial(S, M, P, out);
  pm_run(S, M, P);
  return M.status;
}

}  // namespace fsd
