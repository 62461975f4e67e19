// Cone sorting for ONE frame by ONE warp (S1-S8 of SURVEY.md section 8a).
//
// Behaviour follows the reference's TraceSorter
//   fsd_path_planning/sorting_cones/trace_sorter/core_trace_sorter.py:148-465   (seeds, per-side driver)
//   .../adjacency_matrix.py:60-128, common.py:36-67                             (k-NN graph, reachability)
//   .../end_configurations.py:108-520                                           (exhaustive search + filter)
//   .../cost_function.py, cone_distance_cost.py, nearby_cone_search.py          (7-term cost)
//   .../combine_traces.py:21-275                                                (left/right conflict)
// but is laid out for a warp: the O(N^2) distance/k-NN step and all O(N) masks are lane-strided,
// the per-pop admissibility test runs one candidate neighbour per lane, the sparse serial
// decisions run on lane 0.  Pure sign / threshold tests on angles are evaluated on cosines and
// cross products instead of atan2/acos (same predicate, no trigonometry); angles that enter sums
// or are compared against each other keep atan2.
#pragma once

#include "lane.cuh"
#include "plan_types.cuh"

namespace fsd {

// Loops of the serial (lane 0) and rarely executed parts of the sort stage are kept ROLLED: fully unrolled they were 60 KB
// of straight-line code that every frame streams through the instruction cache once (the free-running sort kernel spent
// 41 % of its stall samples waiting for instructions, profiles/r2_e_sort_kernel_ncu_raw.txt).
#ifdef FSD_SORT_UNROLL_COLD
#define FSD_ROLLED
#else
#define FSD_ROLLED _Pragma("unroll 1")
#endif

constexpr int MAX_LEAVES = 24;  // raw leaves of one side's search (the reference's own data: <= 12); more -> FSD_ST_OVERFLOW
constexpr int STACK_CAP = 64;   // depth <= 12, <= 5 pushes per level

// The two sides of a frame are searched at the same time by the two half-warps (lanes 0-15: LEFT, lanes 16-31: RIGHT;
// the host-check build runs them one after the other on its single lane): every lane group has its own scratch.
#ifdef FSD_DEVICE_BUILD
using SG = Grp<16>;
#else
using SG = Grp<1>;
#endif

// per-side scratch of the search (seeds, reachability, exhaustive search, filter, cost)
struct SideScratch {
  uint8_t idxs[FSD_MAX_CONES];   // cone indices used by any configuration (cost term) / BFS queue
  uint8_t close[FSD_MAX_CONES];  // nearby cones (cost term)
  int16_t n_good[MAX_LEAVES], n_bad[MAX_LEAVES];
  uint8_t flag[FSD_MAX_CONES];  // seed mask / in-configuration mask
  uint8_t flag2[FSD_MAX_CONES];
  uint8_t stack_node[STACK_CAP], stack_pos[STACK_CAP];
  int16_t attempt[16];
  int16_t leaves[MAX_LEAVES][FSD_MAX_SORTED];
  double costs[MAX_LEAVES];
  int32_t scratch[4];
};

// Shared-memory image of one frame.  The fp32 staging area of the TMA copy, the k-NN lists (build_knn) and the two sides'
// search scratch are overlaid: each is dead when the next is first written.
struct SortSmem {
  d2 xy[FSD_MAX_CONES];
  uint8_t type[FSD_MAX_CONES];
  int16_t best[2][FSD_MAX_SORTED];
  int32_t nbest[2];
  uint8_t nbr[2][FSD_MAX_CONES][5];  // mutual-edge adjacency lists, ascending
  uint8_t deg[2][FSD_MAX_CONES];
  union {
    alignas(16) float raw[2 * (FSD_MAX_CONES + 2)];
    uint8_t knn[2][FSD_MAX_CONES][5];  // unused slots hold the row's own index
    SideScratch side[2];
  };
};

// ---- k-NN graph: adjacency_matrix.py:60-110 ------------------------------------------------
// Both sides in one sweep: the LEFT graph ignores yellow cones, the RIGHT graph ignores blue.

// The k <= 5 nearest neighbours of one cone for one side, ascending distance, held in registers (no indexing by a
// run-time value, so nothing lands in local memory).  Empty slots hold +inf.
struct KnnList {
  double d0, d1, d2, d3, d4;
  int i0, i1, i2, i3, i4;
};

FSD_DEV void knn_clear(KnnList &L) {
  L.d0 = L.d1 = L.d2 = L.d3 = L.d4 = INFINITY;
  L.i0 = L.i1 = L.i2 = L.i3 = L.i4 = -1;
}

// insert (v, j); ties keep the earlier, i.e. lower, index first (candidates arrive in ascending j)
FSD_DEV void knn_insert(KnnList &L, double v, int j) {
  if (!(v < L.d4)) return;
  const bool c3 = v < L.d3, c2 = v < L.d2, c1 = v < L.d1, c0 = v < L.d0;
  L.d4 = c3 ? L.d3 : v;
  L.i4 = c3 ? L.i3 : j;
  L.d3 = c3 ? (c2 ? L.d2 : v) : L.d3;
  L.i3 = c3 ? (c2 ? L.i2 : j) : L.i3;
  L.d2 = c2 ? (c1 ? L.d1 : v) : L.d2;
  L.i2 = c2 ? (c1 ? L.i1 : j) : L.i2;
  L.d1 = c1 ? (c0 ? L.d0 : v) : L.d1;
  L.i1 = c1 ? (c0 ? L.i0 : j) : L.i1;
  L.d0 = c0 ? v : L.d0;
  L.i0 = c0 ? j : L.i0;
}

// the first k entries of the list (fewer when the list is shorter); unused slots get the row's own index
FSD_DEV void knn_store(const KnnList &L, int k, int own, uint8_t *out) {
  out[0] = (uint8_t)(k > 0 && L.i0 >= 0 ? L.i0 : own);
  out[1] = (uint8_t)(k > 1 && L.i1 >= 0 ? L.i1 : own);
  out[2] = (uint8_t)(k > 2 && L.i2 >= 0 ? L.i2 : own);
  out[3] = (uint8_t)(k > 3 && L.i3 >= 0 ? L.i3 : own);
  out[4] = (uint8_t)(k > 4 && L.i4 >= 0 ? L.i4 : own);
}

// k of the k-NN graph of a frame with n cones (adjacency_matrix.py:60-75)
FSD_DEV int knn_k(int n, const DevParams &P) {
  int k = n - 1 < P.max_n_neighbors ? n - 1 : P.max_n_neighbors;
  return k > 5 ? 5 : k;
}

// row i of the distance matrix -> the row's LEFT and RIGHT 5-nearest lists (S.knn)
FSD_DEV void knn_row(SortSmem &S, int n, int k, int i, const DevParams &P) {
  KnnList KL, KR;
  knn_clear(KL);
  knn_clear(KR);
  const double xi = S.xy[i].x, yi = S.xy[i].y;
  const int ti = S.type[i];
  const bool li = ti != FSD_CONE_RIGHT, ri = ti != FSD_CONE_LEFT;
  // Edges longer than max_dist are removed after the k-NN selection in the reference (:102-107) and a longer edge can
  // never displace a shorter one, so they are dropped before the selection.  The frame is swept in chunks of 32
  // cones: a branch-free distance loop leaves the chunk's in-range cones as a bit mask, then the handful of set bits
  // is offered to the row's lists in ascending j (the tie order of the selection).  Lanes stay converged through
  // the sweep and diverge only by the number of in-range cones per chunk.
#pragma unroll 1
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int jn = n - j0 < 32 ? n - j0 : 32;
    unsigned mask = 0;
#pragma unroll 4
    for (int jj = 0; jj < jn; ++jj) {
      const double ddx = S.xy[j0 + jj].x - xi, ddy = S.xy[j0 + jj].y - yi;
      mask |= (ddx * ddx + ddy * ddy <= P.max_dist2 ? 1u : 0u) << jj;
    }
    if ((unsigned)(i - j0) < 32u) mask &= ~(1u << (i - j0));
    while (mask) {
      const int jj = FSD_FFS(mask) - 1;
      mask &= mask - 1;
      const int j = j0 + jj;
      const double ddx = S.xy[j].x - xi, ddy = S.xy[j].y - yi;
      const double dd = ddx * ddx + ddy * ddy;
      const int tj = S.type[j];
      if (li && tj != FSD_CONE_RIGHT) knn_insert(KL, dd, j);
      if (ri && tj != FSD_CONE_LEFT) knn_insert(KR, dd, j);
    }
  }
  // unused slots hold the row's own index: a cone is never its own neighbour, so they match nothing below
  knn_store(KL, k, i, S.knn[0][i]);
  knn_store(KR, k, i, S.knn[1][i]);
}

// row i: keep edges present in both directions (:110); neighbour lists in ascending index order, the order np.where
// gives the CSR lists of end_configurations.py:28-71.  All rows of the frame must have passed knn_row.
FSD_DEV void knn_mutual_row(SortSmem &S, int i) {
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    int cnt = 0;
    int t0 = 256, t1 = 256, t2 = 256, t3 = 256, t4 = 256;  // ascending, 256 = empty
#pragma unroll 1
    for (int q = 0; q < 5; ++q) {
      const int j = S.knn[s][i][q];
      if (j == i) break;  // end of the list
      const uint8_t *kj = S.knn[s][j];
      const bool back = (kj[0] == i) | (kj[1] == i) | (kj[2] == i) | (kj[3] == i) | (kj[4] == i);
      if (back) {
        ++cnt;
        const bool c3 = j < t3, c2 = j < t2, c1 = j < t1, c0 = j < t0;
        t4 = c3 ? t3 : j;
        t3 = c3 ? (c2 ? t2 : j) : t3;
        t2 = c2 ? (c1 ? t1 : j) : t2;
        t1 = c1 ? (c0 ? t0 : j) : t1;
        t0 = c0 ? j : t0;
      }
    }
    uint8_t *ni = S.nbr[s][i];
    ni[0] = (uint8_t)t0;
    ni[1] = (uint8_t)t1;
    ni[2] = (uint8_t)t2;
    ni[3] = (uint8_t)t3;
    ni[4] = (uint8_t)t4;
    S.deg[s][i] = (uint8_t)cnt;
  }
}

// one warp, one frame (the sort kernel pools the rows of the frames of a CTA instead, kernels.cu)
FSD_DEVFN void build_knn(SortSmem &S, int n, const DevParams &P) {
  const int k = knn_k(n, P);
#pragma unroll 1
  for (int i = fsd_lane(); i < n; i += FSD_LANES) knn_row(S, n, k, i, P);
  wsync();
#pragma unroll 1
  for (int i = fsd_lane(); i < n; i += FSD_LANES) knn_mutual_row(S, i);
  wsync();
}

// ---- seeds: core_trace_sorter.py:344-465 ----------------------------------------------------

FSD_DEVFN int select_first_k(SortSmem &S, SideScratch &Q, int n, const FramePose &F, int side, const DevParams &P, int *fk) {
  const int opp = side == FSD_CONE_LEFT ? FSD_CONE_RIGHT : FSD_CONE_LEFT;
  const double c = F.ux, s = F.uy;  // rotation by -yaw
  const double cos_max = P.cos_seed_max, cos_min = P.cos_seed_min, max_first2 = P.max_dist_to_first * P.max_dist_to_first;
  double bv = 0.0;
  int bi = -1;
#pragma unroll 1
  for (int i = SG::lane(); i < n; i += SG::N) {
    double px = S.xy[i].x - F.px, py = S.xy[i].y - F.py;
    double rx = px * c + py * s, ry = -px * s + py * c;
    // distances to the car are compared squared (monotone, no square root)
    double r = rx * rx + ry * ry;
    bool in_ellipse = (rx * rx * P.seed_inv_major2 + ry * ry * P.seed_inv_minor2) < 1.0;
    // sign(bearing) == +-1, pi/10 < |bearing| < 4pi/5  (:395-399), on the cosine rx / |r| of the bearing
    bool side_ok = side == FSD_CONE_LEFT ? ry > 0.0 : ry < 0.0;
    bool ang_ok = gt_scaled(rx, cos_max, r) && lt_scaled(rx, cos_min, r);
    int t = S.type[i];
    bool valid = in_ellipse && ((side_ok && ang_ok) || t == side) && t != opp;
    Q.flag[i] = valid ? 1 : 0;
    Q.flag2[i] = rx > 0.0 ? 1 : 0;  // in front of the car: |angle to heading| < pi/2 (:433)
    if (valid && (bi < 0 || r < bv)) {
      bv = r;
      bi = i;
    }
  }
  SG::argmin(bv, bi);
  SG::sync();
  if (bi < 0 || bv > max_first2) return 0;
  int i1 = bi;
  bv = 0.0;
  bi = -1;
#pragma unroll 1
  for (int i = SG::lane(); i < n; i += SG::N) {
    if (!Q.flag[i] || Q.flag2[i] || i == i1) continue;
    double px = S.xy[i].x - F.px, py = S.xy[i].y - F.py;
    double rx = px * c + py * s, ry = -px * s + py * c;
    double r = rx * rx + ry * ry;
    if (bi < 0 || r < bv) {
      bv = r;
      bi = i;
    }
  }
  SG::argmin(bv, bi);
  SG::sync();
  if (bi < 0 || bv > max_first2) {
    fk[0] = i1;
    return 1;
  }
  int i2 = bi;
  double ex = S.xy[i1].x - S.xy[i2].x, ey = S.xy[i1].y - S.xy[i2].y;
  // angle(c1 - c2, heading) > angle(c2 - c1, heading)  <=>  (c1 - c2) . heading < 0  (:450-457)
  if (ex * F.dx + ey * F.dy < 0.0) {
    int t = i1;
    i1 = i2;
    i2 = t;
  }
  const double d2 = ex * ex + ey * ey, dmax = P.max_dist * 1.1;
  if (d2 > dmax * dmax || d2 < 1.4 * 1.4) {
    fk[0] = i1;
    return 1;
  }
  fk[0] = i2;
  fk[1] = i1;
  return 2;
}

// ---- reachability: common.py:36-67; only min(#reachable, max_length) is used ----------------

FSD_DEVFN int reachable_count(SortSmem &S, SideScratch &Q, int n, int sidx, int start, int cap) {
  const int lane = SG::lane();
  // flag2 doubles as the visited mask, idxs as the queue; breadth first, the <= 5 neighbours of a node one per lane
#pragma unroll 1
  for (int i = lane; i < n; i += SG::N) Q.flag2[i] = 0;
  SG::sync();
  if (lane == 0) {
    Q.idxs[0] = (uint8_t)start;
    Q.flag2[start] = 1;
  }
  SG::sync();
  int head = 0, tail = 1;
  while (head < tail && tail < cap) {
    const int node = Q.idxs[head++];
    const int deg = S.deg[sidx][node];
    unsigned fresh = 0;
#pragma unroll 1
    for (int base = 0; base < deg; base += SG::N) {
      const int q = base + lane;
      const int j = q < deg ? (int)S.nbr[sidx][node][q] : -1;
      const bool f = j >= 0 && !Q.flag2[j];
      const unsigned m = SG::ballot(f);
      if (f) {
        Q.idxs[tail + FSD_POPC(fresh) + FSD_POPC(m & ((1u << lane) - 1u))] = (uint8_t)j;
        Q.flag2[j] = 1;
      }
      fresh |= m << base;
    }
    tail += FSD_POPC(fresh);
    SG::sync();
  }
  return tail < cap ? tail : cap;
}

// ---- admissibility of one candidate: end_configurations.py:108-278 --------------------------

FSD_DEV bool segments_intersect(double a0x, double a0y, double a1x, double a1y, double b0x, double b0y, double b1x,
                                double b1y) {
  // lines_segments_intersect_indicator, line_segment_intersection.py:136-200
  const double eps = 1e-6;
  double lax = a0y - a1y, lay = a1x - a0x, laz = a0x * a1y - a0y * a1x;
  double lbx = b0y - b1y, lby = b1x - b0x, lbz = b0x * b1y - b0y * b1x;
  double ix = lay * lbz - laz * lby;
  double iy = laz * lbx - lax * lbz;
  double iz = lax * lby - lay * lbx;
  if (fabs(iz) < eps) {
    // parallel case :34-72 (`difference[0] < epsilon` has no abs in the reference)
    double ddx = a1x - a0x, ddy = a1y - a0y;
    bool overlap;
    double slope;
    if (ddx < eps) {
      overlap = fabs(a0x - b0x) < eps;
      slope = INFINITY;
    } else {
      slope = fdiv(ddy, ddx);
      overlap = fabs((a0y - slope * a0x) - (b0y - slope * b0x)) < eps;
    }
    if (!overlap) return false;
    bool use_y = slope > 1.0;
    double a0 = use_y ? a0y : a0x, a1 = use_y ? a1y : a1x, b0 = use_y ? b0y : b0x, b1 = use_y ? b1y : b1x;
    double left_end, right_start;
    if (a0 < b0) {
      left_end = a1;
      right_start = fmin(b0, b1);
    } else {
      left_end = b1;
      right_start = fmin(a0, a1);
    }
    return left_end >= right_start;
  }
  const double inv = frcp(iz);
  double x = ix * inv, y = iy * inv;
  return (fmin(a0x, a1x) - eps <= x && x <= fmax(a0x, a1x) + eps) &&
         (fmin(b0x, b1x) - eps <= x && x <= fmax(b0x, b1x) + eps) &&
         (fmin(a0y, a1y) - eps <= y && y <= fmax(a0y, a1y) + eps) &&
         (fmin(b0y, b1y) - eps <= y && y <= fmax(b0y, b1y) + eps);
}

FSD_DEVFN bool can_be_added(const SortSmem &S, const SideScratch &Q, const FramePose &F, int side, int sidx, int pos, int i,
                            const DevParams &P) {
  const int last = Q.attempt[pos];
  const uint8_t *nb = S.nbr[sidx][last];
  const int nnb = S.deg[sidx][last];
  const int cand = nb[i];
  if (Q.flag[cand]) return false;  // already in the attempt (:126); find_leaves keeps the membership flags
  const double lx = S.xy[last].x, ly = S.xy[last].y;
  const double cx = S.xy[cand].x, cy = S.xy[cand].y;
  const double bx = cx - lx, by = cy - ly;  // last -> candidate
  double ax = 0.0, ay = 0.0;                // previous -> last
  if (pos >= 1) {
    const int prev = Q.attempt[pos - 1];
    ax = lx - S.xy[prev].x;
    ay = ly - S.xy[prev].y;
    // ellipse around `last`, major axis 6 m along (last - previous), minor 3 m (:281-300):
    // (b.a)^2 / 36 + (b x a)^2 / 9 < |a|^2  -- the rotated-frame test without the normalisation
    const double dt = bx * ax + by * ay, cr = by * ax - bx * ay;
    if (!(dt * dt * (1.0 / 36.0) + cr * cr * (1.0 / 9.0) < ax * ax + ay * ay)) return false;
  } else {
    // second cone of the attempt must lie on the expected side of the car, 5 deg tolerance (:260-278)
    double vx = cx - F.px, vy = cy - F.py;
    double cross = F.dx * vy - F.dy * vx;
    bool expected = side == FSD_CONE_LEFT ? cross > 0.0 : cross < 0.0;
    if (!expected && !(cos_between(F.dx, F.dy, vx, vy) > P.cos_5deg)) return false;
  }
  // another neighbour of `last` lying between `last` and the candidate (:226-257)
  for (int q = 0; q < nnb; ++q) {
    int o = nb[q];
    if (o == cand) continue;
    double v1x = lx - S.xy[o].x, v1y = ly - S.xy[o].y;
    double v2x = cx - S.xy[o].x, v2y = cy - S.xy[o].y;
    double d1 = v1x * v1x + v1y * v1y, d2 = v2x * v2x + v2y * v2y;
    // angle(v1, v2) > 150 deg  <=>  v1 . v2 < cos(150 deg) |v1| |v2|, without the square root
    if (d2 < 36.0 && d1 < 36.0 && lt_scaled(v1x * v2x + v1y * v2y, P.cos_150deg, d1 * d2)) return false;
  }
  if (pos >= 1) {
    // turn angle at `last` (:173-191): difference = wrap(atan2(b) - atan2(a)) has sine cr / v and cosine dt / v with
    // v = |a| |b|, so the three threshold tests are evaluated on cr, dt and v^2 -- no arctangent, no square root:
    //   |difference| > t        <=>  dt < cos(t) v
    //   difference >= t (t > 0) <=>  cr >= 0 and dt <= cos(t) v
    const double cr = ax * by - ay * bx, dt = ax * bx + ay * by, v2 = cr * cr + dt * dt;
    if (lt_scaled(dt, P.cos_thr_abs, v2)) return false;
    const bool short_edge = bx * bx + by * by < 16.0;
    const bool beyond = (side == FSD_CONE_LEFT ? cr >= 0.0 : cr <= 0.0) && !gt_scaled(dt, P.cos_thr_dir, v2);
    if (beyond && !short_edge) return false;
    if (pos >= 2) {
      // change of turning direction (:193-205): sign(difference) != sign(difference_2) and |difference -
      // difference_2| = |difference| + |difference_2| > 1.3, on the sine and cosine of the sum of the two magnitudes
      const int prev = Q.attempt[pos - 1], pp = Q.attempt[pos - 2];
      const double zx = S.xy[prev].x - S.xy[pp].x, zy = S.xy[prev].y - S.xy[pp].y;
      const double cr2 = zx * ay - zy * ax, dt2 = zx * ax + zy * ay, w2 = cr2 * cr2 + dt2 * dt2;
      if (isgn(cr) != isgn(cr2)) {
        const double sa = fabs(cr), sb = fabs(cr2);
        const double sn = sa * dt2 + dt * sb, cs = dt * dt2 - sa * sb;
        if (sn < 0.0 || lt_scaled(cs, P.cos_1p3, v2 * w2)) return false;
      }
    }
  }
  if (pos == 1) {
    // angle(heading, candidate - first) < pi/2 (:207-211)
    const int first = Q.attempt[0];
    if (!(F.dx * (cx - S.xy[first].x) + F.dy * (cy - S.xy[first].y) > 0.0)) return false;
  }
  // the new edge must not cross the car (:213-221)
  double csx = F.px - F.ux * P.car_size / 2.0, csy = F.py - F.uy * P.car_size / 2.0;
  double cex = F.px + F.ux * P.car_size, cey = F.py + F.uy * P.car_size;
  return !segments_intersect(lx, ly, cx, cy, csx, csy, cex, cey);
}

// ---- exhaustive search: end_configurations.py:320-431 ----------------------------------------
// returns the number of raw leaves in Q.leaves (rows padded with -1 up to FSD_MAX_SORTED)

FSD_DEVFN int find_leaves(SortSmem &S, SideScratch &Q, int n, const FramePose &F, int side, int sidx, const int *fk, int nfk, int L,
                          const DevParams &P, int *pops_out, unsigned *status) {
  const int lane = SG::lane();
  int sp = 0, n_leaves = 0, pops = 0;
  // Q.flag[c] == 1 <=> cone c is in the current attempt
#pragma unroll 1
  for (int i = lane; i < n; i += SG::N) Q.flag[i] = 0;
  SG::sync();
  if (lane == 0) {
    for (int q = 0; q < 16; ++q) Q.attempt[q] = -1;
    if (nfk > 1) {
      Q.attempt[0] = (int16_t)fk[0];
      Q.flag[fk[0]] = 1;
      Q.stack_node[0] = (uint8_t)fk[1];
      Q.stack_pos[0] = 1;
    } else {
      Q.stack_node[0] = (uint8_t)fk[0];
      Q.stack_pos[0] = 0;
    }
  }
  SG::sync();
  while (sp >= 0) {
    if (++pops > P.max_dfs_pops) {
      *status |= FSD_ST_OVERFLOW;
      break;
    }
    const int node = Q.stack_node[sp], pos = Q.stack_pos[sp];
    --sp;
    SG::sync();
    // the popped node becomes entry `pos` of the attempt, everything behind it is cleared (one entry per lane)
#pragma unroll 1
    for (int q = pos + lane; q < L; q += SG::N) {
      const int old = Q.attempt[q];
      if (old >= 0) Q.flag[old] = 0;
      Q.attempt[q] = (int16_t)(q == pos ? node : -1);
    }
    SG::sync();
    if (lane == 0) Q.flag[node] = 1;
    SG::sync();
    // one candidate neighbour per lane; the admissible ones as a bit mask
    const int nnb = S.deg[sidx][node];
    unsigned ok_mask = 0;
#pragma unroll 1
    for (int base = 0; base < nnb; base += SG::N) {
      const int i = base + lane;
      const bool ok = i < nnb && can_be_added(S, Q, F, side, sidx, pos, i, P);
      ok_mask |= SG::ballot(ok) << base;
    }
    const int n_ok = FSD_POPC(ok_mask);
    if (pos < L - 1 && n_ok > 0) {
      // push the admissible neighbours in list order (each lane writes its own slot)
#pragma unroll 1
      for (int i = lane; i < nnb; i += SG::N)
        if ((ok_mask >> i) & 1u) {
          const int w = sp + 1 + FSD_POPC(ok_mask & ((1u << i) - 1u));
          if (w < STACK_CAP) {
            Q.stack_node[w] = S.nbr[sidx][node][i];
            Q.stack_pos[w] = (uint8_t)(pos + 1);
          }
        }
      sp += n_ok;
      if (sp >= STACK_CAP) {
        sp = STACK_CAP - 1;
        *status |= FSD_ST_OVERFLOW;
      }
    } else {
      if (n_leaves < MAX_LEAVES) {
#pragma unroll 1
        for (int q = lane; q < FSD_MAX_SORTED; q += SG::N) Q.leaves[n_leaves][q] = q < L ? Q.attempt[q] : (int16_t)-1;
        ++n_leaves;
      } else {
        *status |= FSD_ST_OVERFLOW;
      }
    }
    SG::sync();
  }
  *pops_out = pops;
  return n_leaves;
}

// ---- post-filter: end_configurations.py:484-518 (lane 0; a handful of rows) -------------------

#ifdef FSD_SORT_UNROLL_COLD
#define FSD_ROWFN FSD_DEV
#else
#define FSD_ROWFN FSD_DEVFN
#endif
FSD_ROWFN int row_len(const int16_t *row) {
  int n = 0;
  FSD_ROLLED
  for (int q = 0; q < FSD_MAX_SORTED; ++q) n += row[q] != -1;
  return n;
}

FSD_ROWFN int row_cmp(const int16_t *a, const int16_t *b) {
  FSD_ROLLED
  for (int q = 0; q < FSD_MAX_SORTED; ++q)
    if (a[q] != b[q]) return a[q] < b[q] ? -1 : 1;
  return 0;
}

FSD_DEVFN int post_filter(SortSmem &S, SideScratch &Q, int n_leaves, int side, const int *fk, int nfk) {
  if (SG::lane() == 0) {
    int kept = 0;
    FSD_ROLLED
    for (int r = 0; r < n_leaves; ++r) {
      int16_t *row = Q.leaves[r];
      int len = row_len(row);
      if (len <= 2) continue;
      bool ok = true;
      if (nfk > 1)
        FSD_ROLLED
        for (int q = 0; q < nfk; ++q) ok &= row[q] == fk[q];
      if (!ok) continue;
      // a trailing cone that does not have the side's colour is dropped (:492-500)
      if (S.type[row[len - 1]] != side) {
        row[len - 1] = -1;
        --len;
      }
      if (len < 3) continue;
      // np.unique(axis=0): sorted insertion, duplicates skipped
      int p = kept;
      bool dup = false;
      int16_t tmp[FSD_MAX_SORTED];
      FSD_ROLLED
      for (int q = 0; q < FSD_MAX_SORTED; ++q) tmp[q] = row[q];
      FSD_ROLLED
      while (p > 0) {
        int c = row_cmp(Q.leaves[p - 1], tmp);
        if (c == 0) dup = true;
        if (c <= 0) break;
        --p;
      }
      if (dup) continue;
      FSD_ROLLED
      for (int m = kept; m > p; --m)
        FSD_ROLLED
        for (int q = 0; q < FSD_MAX_SORTED; ++q) Q.leaves[m][q] = Q.leaves[m - 1][q];
      FSD_ROLLED
      for (int q = 0; q < FSD_MAX_SORTED; ++q) Q.leaves[p][q] = tmp[q];
      ++kept;
    }
    // rows that are a strict prefix of another row are removed (:509-515)
    int out = 0;
    FSD_ROLLED
    for (int j = 0; j < kept; ++j) {
      int covered = 0;
      FSD_ROLLED
      for (int i = 0; i < kept; ++i) {
        bool all = true;
        FSD_ROLLED
        for (int q = 0; q < FSD_MAX_SORTED; ++q) all &= (Q.leaves[i][q] == Q.leaves[j][q]) || (Q.leaves[j][q] == -1);
        covered += all;
      }
      Q.flag2[j] = covered > 1 ? 1 : 0;
    }
    FSD_ROLLED
    for (int j = 0; j < kept; ++j)
      if (!Q.flag2[j]) {
        if (out != j)
          FSD_ROLLED
          for (int q = 0; q < FSD_MAX_SORTED; ++q) Q.leaves[out][q] = Q.leaves[j][q];
        ++out;
      }
    Q.scratch[0] = out;
  }
  SG::sync();
  int c = Q.scratch[0];
  SG::sync();
  return c;
}

// NOTE on the unique step above: rows kept so far are sorted; the incoming row may alias a slot
// that is about to be overwritten (r >= kept always holds, and slots < kept are only shifted up to
// index kept <= r), so it is copied to `tmp` first.

// ---- cost: cost_function.py:213-304 -------------------------------------------------------------

FSD_DEV void search_dir(const SortSmem &S, int a, int b, int side, double &ox, double &oy) {
  // normal of the track direction, +90 deg for RIGHT / -90 deg for LEFT (match_directions.py:7-20)
  double tx = S.xy[b].x - S.xy[a].x, ty = S.xy[b].y - S.xy[a].y;
  // not normalised: only the angle to this direction is used
  ox = side == FSD_CONE_RIGHT ? -ty : ty;
  oy = side == FSD_CONE_RIGHT ? tx : -tx;
}

// lane-strided stream compaction of the indices i < n with flag[i] != 0 (ascending order)
FSD_DEVFN int compact_flags(const uint8_t *flag, int n, uint8_t *out) {
  int count = 0;
  const int lane = SG::lane();
  FSD_ROLLED
  for (int base = 0; base < n; base += SG::N) {
    int i = base + lane;
    bool p = i < n && flag[i];
    unsigned m = SG::ballot(p);
    if (p) out[count + FSD_POPC(m & ((1u << lane) - 1u))] = (uint8_t)i;
    count += FSD_POPC(m);
  }
  SG::sync();
  return count;
}

FSD_DEVFN void cones_on_either_side(SortSmem &S, SideScratch &Q, int n, int C, int side) {
  // nearby_cone_search.py:212-297, search distance 6 m, search angle 120 deg
  const int lane = SG::lane();
  const double range2 = 36.0;
#pragma unroll 1
  for (int i = lane; i < n; i += SG::N) Q.flag[i] = 0;
  SG::sync();
  if (lane == 0)
    FSD_ROLLED
    for (int r = 0; r < C; ++r)
      FSD_ROLLED
      for (int q = 0; q < FSD_MAX_SORTED; ++q)
        if (Q.leaves[r][q] != -1) Q.flag[Q.leaves[r][q]] = 1;
  SG::sync();
  int nidx = compact_flags(Q.flag, n, Q.idxs);
  // cones within 6 m of any configuration cone (:97-103)
#pragma unroll 1
  for (int j = lane; j < n; j += SG::N) {
    bool near = false;
    FSD_ROLLED
    for (int q = 0; q < nidx && !near; ++q) {
      int i = Q.idxs[q];
      if (i == j) continue;
      double ddx = S.xy[i].x - S.xy[j].x, ddy = S.xy[i].y - S.xy[j].y;
      near = ddx * ddx + ddy * ddy < range2;
    }
    Q.flag2[j] = near ? 1 : 0;
  }
  SG::sync();
  // `close` first holds all nearby cones, then loses the entries hit by the reference's
  // sorted_set_diff (:88-94): mask[searchsorted(all, idxs)] = False, no membership test
  int nall = compact_flags(Q.flag2, n, Q.close);
#pragma unroll 1
  for (int j = lane; j < n; j += SG::N) Q.flag2[j] = 0;
  SG::sync();
#pragma unroll 1
  for (int q = lane; q < nidx; q += SG::N) {
    int v = Q.idxs[q], lo = 0, hi = nall;
    FSD_ROLLED
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (Q.close[mid] < v)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo < nall) Q.flag2[Q.close[lo]] = 1;  // removed
  }
  SG::sync();
  // One configuration per round, ONE CONE OF THE CONFIGURATION PER LANE (<= 12 of the group's lanes); every lane walks the
  // candidate lists itself -- all lanes read the same entries, i.e. broadcasts.  (The first version ran the (configuration,
  // cone) pairs one after the other with the candidates lane-strided: 84 short passes with a reduction-sized tail each for
  // 7 configurations; frames with >= 2 configurations, 10 % of the stream, cost 4 x the others and make the kernel's tail.)
  FSD_ROLLED
  for (int r = 0; r < C; ++r) {
    const int16_t *c = Q.leaves[r];
    const int len = row_len(c);
    int good = 0, bad = 0;
    // the configuration's own cones as a bit mask (cone indices < 256)
    unsigned long long m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    FSD_ROLLED
    for (int w = 0; w < len; ++w) {
      const int v = c[w];
      const unsigned long long bit = 1ull << (v & 63);
      const int k = v >> 6;
      m0 |= k == 0 ? bit : 0ull;
      m1 |= k == 1 ? bit : 0ull;
      m2 |= k == 2 ? bit : 0ull;
      m3 |= k == 3 ? bit : 0ull;
    }
#pragma unroll 1
    for (int j = lane; j < len; j += SG::N) {
      double sx, sy;
      if (j == 0)
        search_dir(S, c[0], c[1], side, sx, sy);
      else if (j == len - 1)
        search_dir(S, c[j - 1], c[j], side, sx, sy);
      else
        search_dir(S, c[j - 1], c[j + 1], side, sx, sy);
      const int cj = c[j];
      const double x0 = S.xy[cj].x, y0 = S.xy[cj].y, s2 = sx * sx + sy * sy;
      // other = close (minus removed) ++ configuration cones of OTHER configurations
      FSD_ROLLED
      for (int q = 0; q < nall + nidx; ++q) {
        int o;
        if (q < nall) {
          o = Q.close[q];
          if (Q.flag2[o]) continue;
        } else {
          o = Q.idxs[q - nall];
          const int k = o >> 6;
          const unsigned long long mk = k == 0 ? m0 : (k == 1 ? m1 : (k == 2 ? m2 : m3));
          if ((mk >> (o & 63)) & 1ull) continue;
        }
        if (o == cj) continue;
        const double vx = S.xy[o].x - x0, vy = S.xy[o].y - y0;
        const double v2 = vx * vx + vy * vy;
        if (!(v2 < range2)) continue;
        const double dot = vx * sx + vy * sy, w2 = v2 * s2;
        good += gt_scaled(dot, 0.5, w2);   // angle to the search direction < 60 deg
        bad += lt_scaled(dot, -0.5, w2);   // angle to the opposite direction < 60 deg
      }
    }
    good = SG::sum_i(good);
    bad = SG::sum_i(bad);
    if (lane == 0) {
      Q.n_good[r] = good;
      Q.n_bad[r] = bad;
    }
  }
  SG::sync();
}

FSD_DEVFN int best_configuration(SortSmem &S, SideScratch &Q, int n, int C, int side, const FramePose &F) {
  if (C == 1) return 0;
  cones_on_either_side(S, Q, n, C, side);
  const double wsum_ = 9200.0;
  int mn = 0;
  FSD_ROLLED
  for (int r = 0; r < C; ++r) {
    int d = Q.n_good[r] - Q.n_bad[r];
    if (r == 0 || d < mn) mn = d;
  }
#pragma unroll 1
  for (int r = SG::lane(); r < C; r += SG::N) {
    const int16_t *c = Q.leaves[r];
    const int len = row_len(c);
#define px(q) S.xy[c[q]].x
#define py(q) S.xy[c[q]].y
    // angle cost (:41-79): mean of (pi - theta)/pi over interior angles, times (1 + #{theta < 40 deg})
    double asum = 0.0;
    int under = 0;
    FSD_ROLLED
    for (int q = 0; q + 2 < len; ++q) {
      double th = fsd_acos(cos_between(px(q + 1) - px(q + 2), py(q + 1) - py(q + 2), px(q + 1) - px(q), py(q + 1) - py(q)));
      asum += (PI - th) / PI;
      under += th < 40.0 * PI / 180.0;
    }
    double angle_cost = asum / (double)(len - 2) * (double)(under + 1);
    // residual distance (cone_distance_cost.py:14-32)
    double resid = 0.0;
    FSD_ROLLED
    for (int q = 0; q + 1 < len; ++q) {
      double ddx = px(q + 1) - px(q), ddy = py(q + 1) - py(q);
      double d = fsqrt(ddx * ddx + ddy * ddy) - 3.0;
      resid += d > 0.0 ? d : 0.0;
    }
    double ncones = 1.0 / (double)len;
    double init_dir = fsd_acos(cos_between(px(1) - px(0), py(1) - py(0), F.dx, F.dy));
    double either = 1.0 / (double)(Q.n_good[r] - Q.n_bad[r] + (mn < 0 ? -mn : mn) + 1);
    // wrong direction (:149-188)
    double wrong = 0.0;
    if (len != 3) {
      double unwanted = side == FSD_CONE_LEFT ? 1.0 : -1.0, sum = 0.0;
      double prev = fsd_atan2(py(1) - py(0), px(1) - px(0));
      FSD_ROLLED
      for (int q = 1; q + 1 < len; ++q) {
        double cur = fsd_atan2(py(q + 1) - py(q), px(q + 1) - px(q));
        double d = angle_difference(prev, cur);
        if (sgn(d) == unwanted && fabs(d) > 40.0 * PI / 180.0) sum += d;
        prev = cur;
      }
      wrong = fabs(sum);
    }
    // sum of w_i term_i / sum(w) in the reference's order (weights 1000, 200, 5000, 1000, 0, 1000, 1000; the
    // change-of-direction term has weight 0)
    double total = 0.0;
    total += angle_cost * (1000.0 / wsum_);
    total += resid * (200.0 / wsum_);
    total += ncones * (5000.0 / wsum_);
    total += init_dir * (1000.0 / wsum_);
    total += 0.0 * (0.0 / wsum_);
    total += either * (1000.0 / wsum_);
    total += wrong * (1000.0 / wsum_);
    Q.costs[r] = total;
#undef px
#undef py
  }
  SG::sync();
  int arg = 0;
  FSD_ROLLED
  for (int r = 1; r < C; ++r)
    if (Q.costs[r] < Q.costs[arg]) arg = r;
  SG::sync();
  return arg;
}

// ---- one side: core_trace_sorter.py:252-327 -----------------------------------------------------

// One side's search by one lane group: seeds and search depth, exhaustive search, filter + cost; the side's result goes
// to S.best[sidx], its length is returned (0: no configuration).  Every lane of the group holds the same values.
FSD_DEVFN int sort_one_side(SortSmem &S, SideScratch &Q, int n, const FramePose &F, int side, const DevParams &P,
                            int16_t *dbg, unsigned *status) {
  const int sidx = side == FSD_CONE_LEFT ? 0 : 1;
  int fk[2] = {-1, -1}, L = 0, n_leaves = 0, pops = 0, len = 0, n_cfg = 0;
  const int nfk = n < 3 ? 0 : select_first_k(S, Q, n, F, side, P, fk);
  if (nfk > 0) {
    const int R = reachable_count(S, Q, n, sidx, fk[0], P.max_length);
    L = R < P.max_length ? R : P.max_length;  // find_configs_and_scores.py:76
  }
  if (nfk > 0 && L >= 3) {
    n_leaves = find_leaves(S, Q, n, F, side, sidx, fk, nfk, L, P, &pops, status);
    n_cfg = post_filter(S, Q, n_leaves, side, fk, nfk);
    if (n_cfg > 0) {
      const int arg = best_configuration(S, Q, n, n_cfg, side, F);
      len = row_len(Q.leaves[arg]);
      if (SG::lane() == 0)
        FSD_ROLLED
        for (int q = 0; q < FSD_MAX_SORTED; ++q) S.best[sidx][q] = Q.leaves[arg][q];
      SG::sync();
    }
  }
  if (dbg && SG::lane() == 0) {
    dbg[2 * sidx] = (int16_t)fk[0];
    dbg[2 * sidx + 1] = (int16_t)(nfk > 1 ? fk[1] : -1);
    dbg[4 + sidx] = (int16_t)n_cfg;
    dbg[6 + sidx] = (int16_t)(pops > 32767 ? 32767 : pops);
  }
  return len;
}

// Both sides of a frame.  Device: the LEFT search on lanes 0-15, the RIGHT search on lanes 16-31, at the same time (the
// two searches share nothing but the read-only frame: coordinates, types, adjacency lists); the half-warps meet again
// at the end and exchange their results.  Host-check build: one after the other.
FSD_DEVFN void sort_sides(SortSmem &S, int n, const FramePose &F, const DevParams &P, int16_t *dbg, unsigned *status,
                          int &nl, int &nr) {
#ifdef FSD_DEVICE_BUILD
  __syncwarp();
  const int sidx = SG::index();
  unsigned st = 0;
  const int len = sort_one_side(S, S.side[sidx], n, F, sidx == 0 ? FSD_CONE_LEFT : FSD_CONE_RIGHT, P, dbg, &st);
  __syncwarp();
  nl = __shfl_sync(FULL, len, 0);
  nr = __shfl_sync(FULL, len, 16);
  *status |= __shfl_sync(FULL, st, 0) | __shfl_sync(FULL, st, 16);
#else
  nl = sort_one_side(S, S.side[0], n, F, FSD_CONE_LEFT, P, dbg, status);
  nr = sort_one_side(S, S.side[1], n, F, FSD_CONE_RIGHT, P, dbg, status);
#endif
}

// ---- left/right conflict: combine_traces.py:115-257 (lane 0) --------------------------------------

FSD_DEV double angle_change_at(const SortSmem &S, const int16_t *cfg, int p) {
  int a = cfg[p - 1], b = cfg[p], c = cfg[p + 1];
  double an = fsd_atan2(S.xy[c].y - S.xy[b].y, S.xy[c].x - S.xy[b].x);
  double ap = fsd_atan2(S.xy[a].y - S.xy[b].y, S.xy[a].x - S.xy[b].x);
  return angle_difference(an, ap);
}

FSD_DEVFN void combine_sides(const SortSmem &S, int &nl, int &nr) {
  const int16_t *left = S.best[0], *right = S.best[1];
  // first entry of each side that also appears in the other one (one entry per lane)
  unsigned ml = 0, mr = 0;
#pragma unroll 1
  for (int base = 0; base < nl || base < nr; base += FSD_LANES) {
    const int e = base + fsd_lane();
    bool hl = false, hr = false;
    if (e < nl)
      FSD_ROLLED
      for (int b = 0; b < nr; ++b) hl |= left[e] == right[b];
    if (e < nr)
      FSD_ROLLED
      for (int a = 0; a < nl; ++a) hr |= right[e] == left[a];
    ml |= wballot(hl) << base;
    mr |= wballot(hr) << base;
  }
  if (!ml) return;
  const int li = FSD_FFS(ml) - 1, ri = FSD_FFS(mr) - 1;
  int ls = -1, rs = -1;
  bool have = false;
  if (li > 0 && ri > 0) {
    int pl = left[li - 1], pr = right[ri - 1], ic = left[li];
    double dlx = S.xy[ic].x - S.xy[pl].x, dly = S.xy[ic].y - S.xy[pl].y;
    double drx = S.xy[ic].x - S.xy[pr].x, dry = S.xy[ic].y - S.xy[pr].y;
    bool l_low = dlx * dlx + dly * dly < 9.0, r_low = drx * drx + dry * dry < 9.0;
    if ((l_low || r_low) && !(l_low && r_low)) {
      have = true;
      if (l_low) {
        ls = nl;
        rs = ri;
      } else {
        ls = li;
        rs = nr;
      }
    }
  }
  if (!have && left[li] == right[ri] && li >= 1 && li <= nl - 2 && ri >= 1 && ri <= nr - 2) {
    double al = angle_change_at(S, left, li), ar = angle_change_at(S, right, ri);
    double sl = sgn(al), sr = sgn(ar);
    int ndiff = nl > nr ? nl - nr : nr - nl;
    bool prefer_left;
    bool both = false;
    if (sl == sr)
      prefer_left = sl == 1.0;
    else if (ndiff > 2)
      prefer_left = nl > nr;
    else if (fabs(fabs(al) - fabs(ar)) > 5.0 * PI / 180.0)
      prefer_left = fabs(al) > fabs(ar);
    else {
      both = true;
      prefer_left = false;
    }
    if (both) {
      ls = li;
      rs = ri;
    } else if (prefer_left) {
      ls = nl;
      rs = ri;
    } else {
      ls = li;
      rs = nr;
    }
  } else if (!have) {
    bool l_end = li == nl - 1, r_end = ri == nr - 1;
    if (l_end && r_end) {
      ls = nl - 1;
      rs = nr - 1;
    } else if (l_end) {
      rs = nr;
      ls = li;
    } else if (r_end) {
      ls = nl;
      rs = ri;
    } else {
      ls = li;
      rs = ri;
    }
  }
  nl = ls;
  nr = rs;
}

// ---- whole frame: TraceSorter.sort_left_right, core_trace_sorter.py:148-216 -----------------------
// S.xy / S.type must hold the frame's n cones.  On return S.best[0..1] / S.nbest hold the sort indices.

// last part of sort_left_right: conflict resolution and the final index lists
FSD_DEVFN unsigned sort_finish(SortSmem &S, int nl, int nr) {
  unsigned status = 0;
  if (nl == 0) status |= FSD_ST_NO_LEFT;
  if (nr == 0) status |= FSD_ST_NO_RIGHT;
  if (nl > 0 && nr > 0) combine_sides(S, nl, nr);
  if (fsd_lane() == 0) {
    FSD_ROLLED
    for (int q = nl; q < FSD_MAX_SORTED; ++q) S.best[0][q] = -1;
    FSD_ROLLED
    for (int q = nr; q < FSD_MAX_SORTED; ++q) S.best[1][q] = -1;
    S.nbest[0] = nl;
    S.nbest[1] = nr;
  }
  wsync();
  return status;
}

FSD_DEVFN unsigned sort_frame(SortSmem &S, int n, const FramePose &F, const DevParams &P, int16_t *dbg) {
  unsigned status = 0;
  if (n >= 3) build_knn(S, n, P);
  int nl = 0, nr = 0;
  sort_sides(S, n, F, P, dbg, &status, nl, nr);
  return status | sort_finish(S, nl, nr);
}

}  // namespace fsd
