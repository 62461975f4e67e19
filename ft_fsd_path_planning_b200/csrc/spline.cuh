// Smoothing-spline curve fit (the reference's scipy.interpolate.splprep / splev call sites,
// fsd_path_planning/utils/spline_fit.py:46-128) for ONE frame by ONE warp.
//
// Same fit as Dierckx' FITPACK parcur/fppara (knot strategy, smoothing-parameter iteration and
// acceptance tests follow SURVEY.md Appendix A step by step), but the linear algebra is laid out
// for a warp instead of FITPACK's row-by-row Givens sweep over the m data points:
//
//   * least-squares spline for a knot set:   N = B^T B  (banded, half-bandwidth k),  r = B^T x,
//     assembled lane-parallel over the data points, one knot interval at a time (all points of an
//     interval touch the same (k+1)^2 block), then ONE banded Cholesky N = G^T G of the tiny
//     (n-k-1)^2 system.  G equals FITPACK's triangular factor (positive diagonal), so
//     p0 = nk1 / sum(diag G) and every later decision see the same numbers up to rounding;
//   * smoothing step F(p) = s:   (N + D^T D / p^2) c = r  with D the discontinuity-jump matrix
//     (fpdisc), again one banded Cholesky per p, and a lane-parallel residual sweep for F(p).
//
// Conditioning measured on the reference's call sites: cond(G) <= 12, cond([G; D/p]) <= 22, so
// the normal equations lose nothing at fp64 (knot vectors identical to scipy's, coefficients to
// 4e-13 on 260 call-site fits; tests/test_hostcheck.py and tests/test_gpu_parity.py repeat this).
#pragma once

#include "lane.cuh"
#include "plan_types.cuh"

namespace fsd {

#ifndef FSD_NCAP
#define FSD_NCAP 34
#endif
constexpr int NCAP = FSD_NCAP;  // knots handled per fit (the reference's own data stays below 20)
constexpr int BW = 5;     // k + 2 for cubic splines

// One record per knot / coefficient index: every per-knot quantity of the fit lives in record i, so the address of any
// entry is  base + i * sizeof(KnotRec) + constant  whatever the capacity of the arena behind it.  The planner kernels
// give every frame an arena of NCAP records in shared memory; a frame whose fit wants more knots than that is planned
// again by the same code over a larger arena (SplineWork::cap records: the whole CTA's shared memory, kernels.cu).
struct KnotRec {
  // (the order of the fields and the odd record length of 29 doubles keep the warp's accesses in the elimination step of
  // chol_solve -- rows i .. i + 4 of G and c at once -- on distinct shared-memory banks)
  union {
    double G[BW];   // banded system being solved (factorised in place)
    double bd[BW];  // discontinuity jumps: dead once D^T D is built, before the first smoothing solve
  };
  double N[BW];    // row of the banded normal matrix B^T B (upper band)
  double t;        // knot
  double c[2];     // B-spline coefficients (chol_solve works on them in place)
  double rhs[2];
  double rpiv;     // reciprocal pivot of chol_solve
  union {
    struct {  // knot-selection phase only
      double fpint;     // residual of knot interval i
      int32_t nrdata;   // data points strictly inside knot interval i
      int32_t start;    // first data index of knot interval i
    };
    double c0[2];  // smoothing phase only: least-squares coefficients of the final knot set (F(p) as a quadratic form)
  };
  double DtD[BW];
  double rk[6];  // reciprocal knot differences of the B-spline recursion for knot interval i
};
constexpr int RS = sizeof(KnotRec) / sizeof(double);  // record stride in doubles
static_assert(sizeof(KnotRec) == 29 * sizeof(double), "KnotRec layout");

struct SplineWork {
  int32_t n, k;   // result: knot count, degree
  int32_t cap;    // records behind r[] (NCAP unless the caller provides a larger arena)
  int32_t suspendable;  // != 0: a fit that outgrows `cap` is suspended (FIT_SUSPENDED), not truncated
  double max_u;   // last parameter value of the fitted data
  KnotRec r[NCAP];  // LAST member: r[i] with NCAP <= i < cap runs on into the memory the caller provides behind it
};

// Reciprocal knot differences 1 / (t[l+i] - t[l+i-j]) of the de Boor recursion for every knot interval
// (0 where the difference vanishes, which reproduces fpbspl's "h(i+1) = 0" branch).  Entry order: (j,i) =
// (1,1) (2,1) (2,2) (3,1) (3,2) (3,3).  One division per table entry instead of six per evaluated point.
FSD_DEVFN void knot_reciprocals(SplineWork &W, int n, int k) {
  const int nrint = n - 2 * k - 1;
#pragma unroll 1
  for (int e = PG::lane(); e < nrint * 6; e += PG::N) {
    const int ii = e / 6, q = e % 6;
    const int j = q == 0 ? 1 : (q < 3 ? 2 : 3);
    const int i = q == 0 ? 1 : (q < 3 ? q : q - 2);
    double r = 0.0;
    if (j <= k) {
      const int l = k + ii;
      const double d = W.r[l + i].t - W.r[l + i - j].t;
      r = d == 0.0 ? 0.0 : frcp(d);
    }
    W.r[ii].rk[q] = r;
  }
  PG::sync();
}

// B-spline basis values of degree k at x for knot interval ii (t[k+ii] <= x < t[k+ii+1]); fully unrolled,
// division-free, h stays in registers
template <int K>
FSD_DEV void bspl_k(const SplineWork &W, double x, int ii, double (&h)[4]) {
  const int l = K + ii;
  const double *rk = W.r[ii].rk;
  double hh[4];
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= K; ++j) {
#pragma unroll
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
#pragma unroll
    for (int i = 1; i <= j; ++i) {
      const double tl = W.r[l + i].t, tr = W.r[l + i - j].t;
      const double f = hh[i - 1] * rk[(j * (j - 1)) / 2 + i - 1];
      h[i - 1] += f * (tl - x);
      h[i] = f * (x - tr);
    }
  }
}

// degrees below 3 only occur for fits of 2 or 3 points: one small out-of-line loop version
FSD_DEVFN void bspl_low(const SplineWork &W, int k, double x, int ii, double *h) {
  const int l = k + ii;
  const double *rk = W.r[ii].rk;
  double hh[4];
  h[0] = 1.0;
#pragma unroll 1
  for (int j = 1; j <= k; ++j) {
#pragma unroll 1
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
#pragma unroll 1
    for (int i = 1; i <= j; ++i) {
      const double f = hh[i - 1] * rk[(j * (j - 1)) / 2 + i - 1];
      h[i - 1] += f * (W.r[l + i].t - x);
      h[i] = f * (x - W.r[l + i - j].t);
    }
  }
}

FSD_DEV void bspl(const SplineWork &W, int k, double x, int ii, double (&h)[4]) {
  if (k == 3) {
    bspl_k<3>(W, x, ii, h);
  } else {
    // the out-of-line version gets a buffer of its own: handing it `h` would pin the caller's h to local memory
    double low[4] = {0, 0, 0, 0};
    bspl_low(W, k, x, ii, low);
    h[0] = low[0];
    h[1] = low[1];
    h[2] = low[2];
    h[3] = low[3];
  }
}

// (the point BY VALUE: reference outputs of an out-of-line function are local memory)
FSD_DEVFN d2 spline_point(const SplineWork &W, double x) {
  // splev with ext=0: the end polynomial pieces extrapolate
  const int k = W.k, nk1 = W.n - k - 1;
  int l = k;
#pragma unroll 1
  while (l < nk1 - 1 && x >= W.r[l + 1].t) ++l;
  double h[4] = {0, 0, 0, 0};
  bspl(W, k, x, l - k, h);
  double sx = 0.0, sy = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j <= k) {
      sx += W.r[l - k + j].c[0] * h[j];
      sy += W.r[l - k + j].c[1] * h[j];
    }
  d2 out;
  out.x = sx;
  out.y = sy;
  return out;
}

// Task (a, b) of lane/task index e in the elimination step of chol_solve: e < npairs -> band update (a, b), 1 <= a <= b <=
// kbm, in row-major order; otherwise right-hand-side update (a, column)
FSD_DEV void chol_task(int e, int kbm, int npairs, int &ta, int &tb) {
  if (e < npairs) {
    // (a, b) of the e-th pair, one nibble each, for kbm = 1 .. 4 (row-major over 1 <= a <= b <= kbm)
    const unsigned long long lut_a = kbm == 4 ? 0x4332221111ull : (kbm == 3 ? 0x322111ull : (kbm == 2 ? 0x211ull : 0x1ull));
    const unsigned long long lut_b = kbm == 4 ? 0x4434324321ull : (kbm == 3 ? 0x332321ull : (kbm == 2 ? 0x221ull : 0x1ull));
    ta = (int)((lut_a >> (4 * e)) & 15ull);
    tb = (int)((lut_b >> (4 * e)) & 15ull);
  } else {
    ta = 1 + ((e - npairs) >> 1);
    tb = (e - npairs) & 1;
  }
}

// Banded symmetric solve G c = rhs (upper band of G in the records, factorised in place; two right-hand sides),
// cooperative over the warp, as the square-root-free factorisation G = L D L^T: right-looking elimination, one matrix row
// per step; the <= 10 trailing updates of the band and the <= 8 updates of the two right-hand sides are ONE task per lane
// (decoded once, the task of a lane never changes: three pointers that advance by one record per row), one reciprocal per
// row and one warp barrier per row.  On return r[i].G[0] holds the pivot d_i (the diagonal of the Cholesky factor of G is
// sqrt(d_i): fppara's p0 needs it), r[i].G[1..] the unscaled rows d_i L[i+a][i], r[i].rpiv = 1 / d_i.  Back substitution
// c = L^-T D^-1 L^-1 rhs as a column sweep (<= 10 tasks per row).  All lanes return the same flag: false on a non-positive
// pivot.
FSD_DEVFN bool chol_solve(SplineWork &W, int nk1, int kb) {
  const int lane = PG::lane();
  const int kbm = kb - 1, npairs = kbm * (kbm + 1) / 2, ntasks = npairs + 2 * kbm;
  KnotRec *R = W.r;
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) {
    R[i].c[0] = R[i].rhs[0];
    R[i].c[1] = R[i].rhs[1];
  }
#ifdef FSD_DEVICE_BUILD
  // this lane's tasks as running pointers (factor, multiplier, target) and the first row without a target; a group of 16
  // lanes has up to 18 tasks, so lanes 0 and 1 carry a second one
  int ta = 0, tb = 0;
  chol_task(lane, kbm, npairs, ta, tb);
  const bool is_pair = lane < npairs, has_task = lane < ntasks;
  const double *pa = &R[0].G[has_task ? ta : 0];
  const double *pb = is_pair ? &R[0].G[tb] : &R[0].c[tb & 1];
  double *pt = is_pair ? &R[ta].G[tb - ta] : &R[ta].c[tb & 1];
  const int ilim = has_task ? nk1 - (is_pair ? tb : ta) : 0;
#if FSD_PATH_LANES < 18
  // second task of lanes 0 and 1 (groups narrower than the task list)
  const int e2 = lane + PG::N;
  const bool has2 = e2 < ntasks;
  int ta2 = 1, tb2 = 0;
  if (has2) chol_task(e2, kbm, npairs, ta2, tb2);
  const bool is_pair2 = e2 < npairs;
  const double *pa2 = &R[0].G[has2 ? ta2 : 0];
  const double *pb2 = is_pair2 ? &R[0].G[tb2] : &R[0].c[tb2 & 1];
  double *pt2 = is_pair2 ? &R[ta2].G[tb2 - ta2] : &R[ta2].c[tb2 & 1];
  const int ilim2 = has2 ? nk1 - (is_pair2 ? tb2 : ta2) : 0;
#endif
  const double *pd = &R[0].G[0];  // the pivot of the current row
  double *pr = &R[0].rpiv;
  PG::sync();
#pragma unroll 1
  for (int i = 0; i < nk1; ++i) {
    const double s = *pd;
    if (!(s > 0.0)) return false;
    const double rs = frcp(s);
    if (i < ilim) *pt -= *pa * rs * *pb;
#if FSD_PATH_LANES < 18
    if (i < ilim2) *pt2 -= *pa2 * rs * *pb2;
    pa2 += RS;
    pb2 += RS;
    pt2 += RS;
#endif
    if (lane == 0) *pr = rs;
    pa += RS;
    pb += RS;
    pt += RS;
    pd += RS;
    pr += RS;
    PG::sync();
  }
#else
  PG::sync();
#pragma unroll 1
  for (int i = 0; i < nk1; ++i) {
    const double s = R[i].G[0];
    if (!(s > 0.0)) return false;
    const double rs = frcp(s);
#pragma unroll 1
    for (int e = 0; e < ntasks; ++e) {
      int ta, tb;
      chol_task(e, kbm, npairs, ta, tb);
      const double f = R[i].G[ta] * rs;
      if (e < npairs) {
        if (i + tb < nk1) R[i + ta].G[tb - ta] -= f * R[i].G[tb];
      } else {
        if (i + ta < nk1) R[i + ta].c[tb] -= f * R[i].c[tb];
      }
    }
    R[i].rpiv = rs;
  }
#endif
  const int nback = 2 * kbm;
#pragma unroll 1
  for (int i = nk1 - 1; i >= 0; --i) {
    FSD_FOR_PTASKS(e, nback) {
      const int l = 1 + (e >> 1), col = e & 1;
      // row i is final from here on (the sweep only touches the rows above it); it is scaled by 1 / d_i after the sweep
      const double ci = R[i].c[col] * R[i].rpiv;
      if (i - l >= 0) R[i - l].c[col] -= R[i - l].G[l] * ci;
    }
    PG::sync();
  }
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) {
    const double rp = R[i].rpiv;
    R[i].c[0] *= rp;
    R[i].c[1] *= rp;
  }
  PG::sync();
  return true;
}

// first data index of every knot interval.  Interior knots are data abscissae (fpknot puts every new knot on a data
// point) and nrdata[ii] counts the data points strictly inside interval ii, so start[ii+1] = start[ii] + nrdata[ii] + 1
// -- no search over the data.
FSD_DEVFN void interval_starts(SplineWork &W, int m, int n, int k) {
  const int nrint = n - 2 * k - 1;
  if (PG::lane() == 0) {
    int s = 0;
    W.r[0].start = 0;
#pragma unroll 1
    for (int ii = 0; ii + 1 < nrint; ++ii) {
      s += W.r[ii].nrdata + 1;
      W.r[ii + 1].start = s;
    }
    W.r[nrint].start = m;
  }
  PG::sync();
}

// N = B^T B and r = B^T x for the current knots (k == 3 uses all 4 x 4 entries; lower degrees fewer)
FSD_DEVFN void assemble_normal(SplineWork &W, const d2 *pts, const double *u, int n, int k) {
  const int lane = PG::lane();
  const int nk1 = n - k - 1, nrint = n - 2 * k - 1, k1 = k + 1;
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) {
#pragma unroll
    for (int d = 0; d < BW; ++d) W.r[i].N[d] = 0.0;
    W.r[i].rhs[0] = 0.0;
    W.r[i].rhs[1] = 0.0;
  }
#ifdef FSD_DEVICE_BUILD
  // the sum this lane owns after the transposed reduction below and where it goes (relative to knot interval 0)
  double *own_base = &W.r[0].N[0];
  bool own_ok = PG::owner16();
  {
    const int e = PG::owned16();
    if (e < 10) {
      int a = 0, rem = e;
#pragma unroll 1
      while (rem >= 4 - a) {
        rem -= 4 - a;
        ++a;
      }
      own_base = &W.r[a].N[rem];  // entry (a, b = a + rem) of the 4 x 4 block -> N[ii + a][b - a]
      own_ok = own_ok && a + rem < k1;
    } else {
      const int a = e < 14 ? e - 10 : e - 14;
      own_base = &W.r[a].rhs[e < 14 ? 0 : 1];
      own_ok = own_ok && a < k1;
    }
  }
#endif
  PG::sync();
#pragma unroll 1
  for (int ii = 0; ii < nrint; ++ii) {
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double rx[4] = {0, 0, 0, 0}, ry[4] = {0, 0, 0, 0};
    const int lo = W.r[ii].start, hi = W.r[ii + 1].start;
#pragma unroll 1
    for (int i = lo + lane; i < hi; i += PG::N) {
      double h[4] = {0, 0, 0, 0};
      bspl(W, k, u[i], ii, h);
      const double x = pts[i].x, y = pts[i].y;
      int e = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = a; b < 4; ++b) acc[e++] += h[a] * h[b];
        rx[a] += h[a] * x;
        ry[a] += h[a] * y;
      }
    }
#ifdef FSD_DEVICE_BUILD
    // 16 of the 18 sums through the transposed reduction (the lane pair 2e, 2e+1 ends up with sum e), ry[2..3] through
    // a plain butterfly; every sum is added to the matrix by the lane that owns it
    double red[16], tail[2] = {ry[2], ry[3]};
#pragma unroll
    for (int e = 0; e < 10; ++e) red[e] = acc[e];
#pragma unroll
    for (int a = 0; a < 4; ++a) red[10 + a] = rx[a];
    red[14] = ry[0];
    red[15] = ry[1];
    PG::sum16_transposed(red);
    PG::sum_vec(tail);
    if (own_ok) own_base[ii * RS] += red[0];
    if (lane == 1 && 2 < k1) W.r[ii + 2].rhs[1] += tail[0];  // (every lane holds the totals; any two distinct lanes do)
    if (lane == 3 && 3 < k1) W.r[ii + 3].rhs[1] += tail[1];
#else
    double red[18];
#pragma unroll
    for (int e = 0; e < 10; ++e) red[e] = acc[e];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      red[10 + a] = rx[a];
      red[14 + a] = ry[a];
    }
    if (lane == 0) {
      int e = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = a; b < 4; ++b) {
          if (b < k1) W.r[ii + a].N[b - a] += red[e];
          ++e;
        }
        if (a < k1) {
          W.r[ii + a].rhs[0] += red[10 + a];
          W.r[ii + a].rhs[1] += red[14 + a];
        }
      }
    }
#endif
    PG::sync();
  }
}

// squared residuals: fp (returned) and, when `per_interval`, fpint[] with FITPACK's half/half split
// of a data point that coincides with a knot (fppara's residual walk)
FSD_DEVFN double residuals(SplineWork &W, const d2 *pts, const double *u, int n, int k, bool per_interval) {
  const int lane = PG::lane();
  const int nrint = n - 2 * k - 1;
  double fp = 0.0;
#pragma unroll 1
  for (int ii = 0; ii < nrint; ++ii) {
    const int lo = W.r[ii].start, hi = W.r[ii + 1].start;
    const int last = ii < nrint - 1 ? hi : hi - 1;  // the next interval's first point is shared
    double part = 0.0, full = 0.0;
#pragma unroll 1
    for (int i = lo + lane; i <= last; i += PG::N) {
      const int li = i >= hi ? ii + 1 : ii;
      double h[4] = {0, 0, 0, 0};
      bspl(W, k, u[i], li, h);
      double sx = 0.0, sy = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j <= k) {
          sx += W.r[li + j].c[0] * h[j];
          sy += W.r[li + j].c[1] * h[j];
        }
      double ex = sx - pts[i].x, ey = sy - pts[i].y;
      double term = ex * ex + ey * ey;
      double wgt = ((i == lo && ii > 0) || i >= hi) ? 0.5 : 1.0;
      part += wgt * term;
      if (i < hi) full += term;
    }
    double red[2] = {part, full};
    PG::sum_vec(red);
    fp += red[1];
    if (per_interval && lane == 0) W.r[ii].fpint = red[0];
  }
  PG::sync();
  return fp;
}

// F(p) - F(inf) of the smoothing iterations without touching the data: the least-squares spline c0 of the knot set is
// the orthogonal projection of the data onto the spline space, so for any coefficients c
//     sum |B c - x|^2 = sum |B c0 - x|^2 + (c - c0)^T N (c - c0),      N = B^T B (already assembled, banded)
// -- a sum of two non-negative terms, no cancellation.  One row of N per lane; the factor in G is dead after chol_solve and lends its first two entries.
FSD_DEVFN double smoothing_excess(SplineWork &W, int nk1, int k) {
  const int lane = PG::lane();
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) {
    W.r[i].G[0] = W.r[i].c[0] - W.r[i].c0[0];  // (the factor in G is dead after the solve)
    W.r[i].G[1] = W.r[i].c[1] - W.r[i].c0[1];
  }
  PG::sync();
  double part = 0.0;
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) {
    const double dx = W.r[i].G[0], dy = W.r[i].G[1];
    double ax = W.r[i].N[0] * dx, ay = W.r[i].N[0] * dy;
#pragma unroll 1
    for (int d = 1; d <= k; ++d)
      if (i + d < nk1) {
        const double w2 = W.r[i].N[d] + W.r[i].N[d];
        ax += w2 * W.r[i + d].G[0];
        ay += w2 * W.r[i + d].G[1];
      }
    part += dx * ax + dy * ay;
  }
  PG::sync();
  return PG::sum(part);
}

// fpknot: split the interval with the largest residual at its middle data point (lane 0)
FSD_DEVFN void add_knot(SplineWork &W, const double *u, int n, int nrint) {
  const int k = (n - nrint - 1) / 2;
  double fpmax = 0.0;
  int jbegin = 1, number = 1, maxpt = 0, maxbeg = 1;
#pragma unroll 1
  for (int j = 1; j <= nrint; ++j) {
    int jpoint = W.r[j - 1].nrdata;
    if (!(fpmax >= W.r[j - 1].fpint || jpoint == 0)) {
      fpmax = W.r[j - 1].fpint;
      number = j;
      maxpt = jpoint;
      maxbeg = jbegin;
    }
    jbegin += jpoint + 1;
  }
  const int ihalf = maxpt / 2 + 1, nrx = maxbeg + ihalf, next = number + 1;
#pragma unroll 1
  for (int j = nrint; j >= next; --j) {
    W.r[j].fpint = W.r[j - 1].fpint;
    W.r[j].nrdata = W.r[j - 1].nrdata;
    W.r[j + k].t = W.r[j + k - 1].t;
  }
  W.r[number - 1].nrdata = ihalf - 1;
  W.r[next - 1].nrdata = maxpt - ihalf;
  const double am = maxpt > 0 ? (double)maxpt : 1.0;
  W.r[number - 1].fpint = fdiv(fpmax * (double)W.r[number - 1].nrdata, am);
  W.r[next - 1].fpint = fdiv(fpmax * (double)W.r[next - 1].nrdata, am);
  W.r[next + k - 1].t = u[nrx - 1];
}

// discontinuity jumps of the k-th derivative at the interior knots (fpdisc), rows lane-strided
FSD_DEVFN void disc_jumps(SplineWork &W, int n, int k) {
  const int k1 = k + 1, k2 = k + 2, nk1 = n - k1, nrint = nk1 - k;
  const double fac = fdiv((double)nrint, W.r[nk1].t - W.r[k].t);
#pragma unroll 1
  for (int l = k2 + PG::lane(); l <= nk1; l += PG::N) {  // 1-based row index of FITPACK
    const int lmk = l - k1;
    double h[10];
#pragma unroll 1
    for (int j = 1; j <= k1; ++j) {
      h[j - 1] = W.r[l - 1].t - W.r[l + j - k2 - 1].t;
      h[j + k1 - 1] = W.r[l - 1].t - W.r[l + j - 1].t;
    }
    int lp = lmk;
#pragma unroll 1
    for (int j = 1; j <= k2; ++j) {
      int jk = j;
      double prod = h[j - 1];
#pragma unroll 1
      for (int i = 1; i <= k; ++i) {
        ++jk;
        prod = prod * h[jk - 1] * fac;
      }
      W.r[lmk - 1].bd[j - 1] = fdiv(W.r[lp + k1 - 1].t - W.r[lp - 1].t, prod);
      ++lp;
    }
  }
  PG::sync();
}

// ---- the fit as a resumable state machine -----------------------------------------------------------------------
// fit_init + repeated fit_step == FITPACK's fppara.  One step is ONE least-squares pass over the data (knot-selection
// phase) or ONE evaluation of F(p) (smoothing phase): the unit at which the warps of a CTA are kept in lockstep by
// the path kernel, so that they execute the same code at the same time (instruction-cache sharing).  Every lane
// holds an identical copy of the state.

// FIT_SUSPENDED: the knot-selection phase wants more knots than the arena holds and the caller asked to be told
// (SplineWork::suspendable) instead of getting a truncated fit: nothing has been truncated yet, the state is intact, and
// fit_resume continues the very same fit once the caller has provided a larger arena (W.cap).
enum { FIT_KNOTS = 0, FIT_SMOOTH_SETUP = 1, FIT_SMOOTH = 2, FIT_DONE = 3, FIT_SUSPENDED = 4 };

struct FitState {
  const d2 *pts;
  const double *u;
  int m, k, n, nest, nmax, nplus, ier, nk1, phase, iter, ich1, ich3, left;
  bool capped, suspendable;
  double s, acc, fp, fpold, fp0, fpms, p, p1, f1, p3, f3, fp_ls;
};

// pts/u: m data points and their (strictly increasing) parameters.  ier = 10 (phase FIT_DONE) on invalid input,
// the reference's ValueError.
FSD_DEVFN void fit_init(SplineWork &W, FitState &F, const d2 *pts, const double *u, int m, double s) {
  const int lane = PG::lane();
  F.pts = pts;
  F.u = u;
  F.m = m;
  F.s = s;
  F.phase = FIT_DONE;
  F.ier = 10;
  F.n = 0;
  if (m < 2) return;
  const int k = m - 1 < 1 ? 1 : (m - 1 > 3 ? 3 : m - 1);
  F.k = k;
  // u strictly increasing (parcur's input check)
  int bad = 0;
#pragma unroll 1
  for (int i = 1 + lane; i < m; i += PG::N) bad |= !(u[i - 1] < u[i]);
  if (PG::any(bad != 0)) return;
  F.acc = 1e-3 * s;
  F.nest = m + 2 * k;
  F.capped = false;
  const int cap = W.cap;  // records of this frame's arena
  if (F.nest > cap) {
    F.nest = cap;
    F.capped = true;
  }
  F.suspendable = W.suspendable != 0;
  F.left = 0;
  F.nmax = m + k + 1;
  F.n = 2 * (k + 1);
  F.nplus = 0;
  F.ier = 0;
  F.nk1 = 0;
  F.fp = F.fpold = F.fp0 = F.fpms = 0.0;
  F.iter = 0;
  F.phase = FIT_KNOTS;
  if (lane == 0) {
    W.r[0].nrdata = m - 2;
    W.k = k;
    W.max_u = u[m - 1];
  }
  PG::sync();
}

FSD_DEVFN void fit_finish(SplineWork &W, FitState &F, int ier) {
  F.ier = ier;
  F.phase = FIT_DONE;
  if (PG::lane() == 0) W.n = F.n;
  PG::sync();
}

// the new knots of a knot-selection pass (fppara's inner loop over fpknot), F.nplus of them unless a bound is reached
// (inlined into fit_step_knots, the hot caller; fit_resume holds the second, cold copy)
FSD_DEV void fit_add_knots(SplineWork &W, FitState &F, int count) {
  const int lane = PG::lane();
  const int k = F.k, k1 = k + 1, k2 = k + 2, nmin = 2 * k1, m = F.m;
  const double *u = F.u;
  int n = F.n;
  int nrint = n - nmin + 1;
#pragma unroll 1
  for (int l = 1; l <= count; ++l) {
    if (lane == 0) add_knot(W, u, n, nrint);
    ++n;
    ++nrint;
    if (n == F.nmax) {
      // every data abscissa becomes a knot (interpolating curve); k is odd whenever interior knots exist
      if (lane == 0) {
        const int k3 = k / 2;
        int i = k2, j = k3 + 2;
#pragma unroll 1
        for (int l2 = 0; l2 < m - k1; ++l2) {
          W.r[i - 1].t = (k3 * 2 != k) ? u[j - 1] : (u[j - 1] + u[j - 2]) * 0.5;
          ++i;
          ++j;
        }
        // data points strictly inside each interval of the interpolating knot set (k is odd here): the first
        // interior knot is data point k3 + 1 (0-based), consecutive knots are consecutive data points
        const int nint = m - k1 + 1;
#pragma unroll 1
        for (int q = 0; q < nint; ++q) W.r[q].nrdata = 0;
        W.r[0].nrdata = k3;
        W.r[nint - 1].nrdata = m - 1 - (k3 + 1 + (m - k1 - 1)) - 1;
      }
      break;
    }
    if (n == F.nest) {
      if (l < count && F.capped && F.suspendable) {
        // the arena is full and this round of knots is not complete: suspended with the remaining count, NOT truncated
        F.left = count - l;
        F.n = n;
        F.phase = FIT_SUSPENDED;
        PG::sync();
        return;
      }
      break;
    }
  }
  F.n = n;
  // (FitState may live in shared memory, one copy for all lanes: every lane reads the counter before any lane writes it)
  const int iter = F.iter + 1;
  PG::sync();
  F.iter = iter;
  if (iter >= m) fit_finish(W, F, F.ier);  // fppara's outer loop bound (never reached in practice)
  PG::sync();
}

// how many knots the next round adds (fppara's nplus rule)
FSD_DEV void fit_choose_nplus(FitState &F) {
  if (F.ier == 0) {
    int npl1 = F.nplus * 2;
    const double rn = (double)F.nplus;
    if (F.fpold - F.fp > F.acc) npl1 = (int)fdiv(rn * F.fpms, F.fpold - F.fp);
    int mx = npl1 > F.nplus / 2 ? npl1 : F.nplus / 2;
    if (mx < 1) mx = 1;
    F.nplus = F.nplus * 2 < mx ? F.nplus * 2 : mx;
  } else {
    F.nplus = 1;
    F.ier = 0;
  }
  F.fpold = F.fp;
}

// continue a suspended fit: the caller has made the arena larger (W.cap) in the meantime
FSD_DEVFN void fit_resume(SplineWork &W, FitState &F) {
  F.nest = F.m + 2 * F.k;
  F.capped = false;
  if (F.nest > W.cap) {
    F.nest = W.cap;
    F.capped = true;
  }
  F.suspendable = W.suspendable != 0;
  F.phase = FIT_KNOTS;
  if (F.left < 0) {  // suspended after a pass, before the size of the next round of knots was chosen
    fit_choose_nplus(F);
    PG::sync();
    fit_add_knots(W, F, F.nplus);
  } else {  // suspended in the middle of a round of knots
    fit_add_knots(W, F, F.left);
  }
}

// one least-squares pass for the current knots + FITPACK's decision what to do next
FSD_DEVFN void fit_step_knots(SplineWork &W, FitState &F, unsigned *status) {
  const int lane = PG::lane();
  PG::sync();  // (lanes aligned where a step starts: see fit_step_smooth)
  const int k = F.k, k1 = k + 1, k2 = k + 2, nmin = 2 * k1, m = F.m;
  const double *u = F.u;
  int n = F.n;
  if (n == nmin) F.ier = -2;
  int nrint = n - nmin + 1;
  F.nk1 = n - k1;
  if (lane == 0)
    for (int j = 0; j < k1; ++j) {
      W.r[j].t = u[0];
      W.r[n - 1 - j].t = u[m - 1];
    }
  PG::sync();
  interval_starts(W, m, n, k);
  knot_reciprocals(W, n, k);
  assemble_normal(W, F.pts, u, n, k);
#pragma unroll 1
  for (int i = lane; i < F.nk1; i += PG::N) {
#pragma unroll
    for (int d = 0; d < BW; ++d) W.r[i].G[d] = W.r[i].N[d];
  }
  PG::sync();
  if (!chol_solve(W, F.nk1, k1)) {
    *status |= FSD_ST_UNSUPPORTED;
    fit_finish(W, F, 10);
    return;
  }
  F.fp = residuals(W, F.pts, u, n, k, true);
  if (F.ier == -2) F.fp0 = F.fp;
  if (lane == 0) {
    W.r[n - 1].fpint = F.fp0;
    W.r[n - 2].fpint = F.fpold;
    W.r[n - 1].nrdata = F.nplus;
  }
  F.fpms = F.fp - F.s;
  if (fabs(F.fpms) < F.acc) {
    fit_finish(W, F, F.ier);
    return;
  }
  if (F.fpms < 0.0) {
    if (F.ier == -2)
      fit_finish(W, F, -2);  // the polynomial already satisfies fp <= s
    else
      F.phase = FIT_SMOOTH_SETUP;
    return;
  }
  if (n == F.nmax) {
    fit_finish(W, F, -1);
    return;
  }
  if (n == F.nest) {
    if (F.capped && F.suspendable) {
      // the arena is full and the fit wants more knots: suspended with its state intact (fit_resume), NOT truncated
      F.left = -1;
      F.phase = FIT_SUSPENDED;
      return;
    }
    if (F.capped) *status |= FSD_ST_OVERFLOW;
    fit_finish(W, F, 1);
    return;
  }
  fit_choose_nplus(F);
  PG::sync();
  fit_add_knots(W, F, F.nplus);
}

// smoothing phase, set-up: discontinuity jumps, D^T D, initial p
FSD_DEVFN void fit_step_smooth_setup(SplineWork &W, FitState &F) {
  const int lane = PG::lane();
  const int k = F.k, k2 = k + 2, nmin = 2 * (k + 1), n = F.n, nk1 = F.nk1;
  // p0 = nk1 / trace of the Cholesky factor of N (chol_solve leaves the pivots d_i = G_ii^2 on the diagonal); read
  // before the jump matrix overwrites G (they share storage)
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) W.r[i].rpiv = fsqrt(W.r[i].G[0]);
  PG::sync();
  double p = 0.0;
#pragma unroll 1
  for (int i = 0; i < nk1; ++i) p += W.r[i].rpiv;
  F.p = fdiv((double)nk1, p);
  PG::sync();
  disc_jumps(W, n, k);
  // D^T D, one band entry per lane: (D^T D)[i][i+d] = sum over the jump rows r = i - a of bd[r][a] bd[r][a+d]
  const int n8 = n - nmin;
#pragma unroll 1
  for (int e = lane; e < nk1 * BW; e += PG::N) {
    const int i = e / BW, d = e % BW;
    double acc = 0.0;
    if (i + d < nk1)
#pragma unroll 1
      for (int a = k2 - 1 - d; a >= 0; --a) {  // ascending r
        const int r = i - a;
        if (r >= 0 && r < n8) acc += W.r[r].bd[a] * W.r[r].bd[a + d];
      }
    W.r[i].DtD[d] = acc;
  }
  PG::sync();
  F.p1 = 0.0;
  F.f1 = F.fp0 - F.s;
  F.p3 = -1.0;
  F.f3 = F.fpms;
  // the least-squares spline of this knot set: its coefficients and residual anchor F(p) below
  F.fp_ls = F.fp;
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) {
    W.r[i].c0[0] = W.r[i].c[0];
    W.r[i].c0[1] = W.r[i].c[1];
  }
  F.ich1 = F.ich3 = 0;
  F.iter = 0;
  F.phase = FIT_SMOOTH;
  PG::sync();
}

// smoothing phase, one evaluation of F(p) and the next p (fppara's iteration incl. fprati)
FSD_DEVFN void fit_step_smooth(SplineWork &W, FitState &F, unsigned *status) {
  const int lane = PG::lane();
  const int k = F.k, k2 = k + 2, nk1 = F.nk1;
  const double con1 = 0.1, con9 = 0.9, con4 = 0.04;
  // FitState may live in shared memory, ONE copy for all lanes (path_kernel): the lanes are aligned where a step starts, and
  // where a field is updated from its own value every lane reads before any lane writes
  PG::sync();
  const int iter = F.iter + 1;
  PG::sync();
  F.iter = iter;
  const double pinv = frcp(F.p), pinv2 = pinv * pinv;
#pragma unroll 1
  for (int i = lane; i < nk1; i += PG::N) {
#pragma unroll
    for (int d = 0; d < BW; ++d) W.r[i].G[d] = W.r[i].N[d] + W.r[i].DtD[d] * pinv2;
  }
  PG::sync();
  if (!chol_solve(W, nk1, k2)) {
    *status |= FSD_ST_UNSUPPORTED;
    fit_finish(W, F, 10);
    return;
  }
  F.fp = F.fp_ls + smoothing_excess(W, nk1, k);
  F.fpms = F.fp - F.s;
  if (fabs(F.fpms) < F.acc) {
    fit_finish(W, F, 0);
    return;
  }
  if (F.iter == 20) {
    fit_finish(W, F, 3);
    return;
  }
  const double p2 = F.p, f2 = F.fpms;
  PG::sync();
  if (F.ich3 == 0) {
    if (!((f2 - F.f3) > F.acc)) {
      F.p3 = p2;
      F.f3 = f2;
      F.p = p2 * con4;
      if (F.p <= F.p1) F.p = F.p1 * con9 + p2 * con1;
      return;
    }
    if (f2 < 0.0) F.ich3 = 1;
  }
  if (F.ich1 == 0) {
    if (!((F.f1 - f2) > F.acc)) {
      F.p1 = p2;
      F.f1 = f2;
      F.p = fdiv(p2, con4);
      if (F.p3 < 0.0) return;
      if (F.p >= F.p3) F.p = p2 * con1 + F.p3 * con9;
      return;
    }
    if (f2 > 0.0) F.ich1 = 1;
  }
  if (f2 >= F.f1 || f2 <= F.f3) {
    fit_finish(W, F, 2);
    return;
  }
  // fprati: rational interpolation through (p1,f1), (p2,f2), (p3,f3)
  double pn;
  if (F.p3 > 0.0) {
    const double h1 = F.f1 * (f2 - F.f3), h2 = f2 * (F.f3 - F.f1), h3 = F.f3 * (F.f1 - f2);
    pn = -fdiv(F.p1 * p2 * h3 + p2 * F.p3 * h1 + F.p3 * F.p1 * h2, F.p1 * h1 + p2 * h2 + F.p3 * h3);
  } else {
    pn = fdiv(F.p1 * (F.f1 - F.f3) * f2 - p2 * (f2 - F.f3) * F.f1, (F.f1 - f2) * F.f3);
  }
  if (f2 < 0.0) {
    F.p3 = p2;
    F.f3 = f2;
  } else {
    F.p1 = p2;
    F.f1 = f2;
  }
  F.p = pn;
}

FSD_DEVFN void fit_step(SplineWork &W, FitState &F, unsigned *status) {
  if (F.phase == FIT_KNOTS)
    fit_step_knots(W, F, status);
  else if (F.phase == FIT_SMOOTH_SETUP)
    fit_step_smooth_setup(W, F);
  else if (F.phase == FIT_SMOOTH)
    fit_step_smooth(W, F, status);
}

// the rest of the fit, phase by phase (no dispatch per pass)
FSD_DEVFN void fit_run(SplineWork &W, FitState &F, unsigned *status) {
  while (F.phase == FIT_KNOTS) fit_step_knots(W, F, status);
  if (F.phase == FIT_SMOOTH_SETUP) fit_step_smooth_setup(W, F);
  while (F.phase == FIT_SMOOTH) fit_step_smooth(W, F, status);
}

// blocking form.  Result in W.t, W.c, W.n, W.k, W.max_u.  Returns FITPACK's ier.
FSD_DEVFN int fit_curve(SplineWork &W, const d2 *pts, const double *u, int m, double s, unsigned *status) {
  FitState F;
  fit_init(W, F, pts, u, m, s);
#pragma unroll 1
  while (F.phase != FIT_DONE) fit_step(W, F, status);
  return F.ier;
}

}  // namespace fsd
