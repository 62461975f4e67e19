// Left/right cone matching with virtual cones for ONE frame by ONE warp (M1-M6 of SURVEY.md 8a).
//
// Behaviour follows the reference's
//   fsd_path_planning/cone_matching/functional_cone_matching.py:73-588
//   fsd_path_planning/cone_matching/match_directions.py:7-44
// with the parameters of core_cone_matching.py:101-117 / config.py:124-129, 162 (non-monotonic).
// The (M x N) candidate test runs one row per lane; tests on angles are evaluated on cosines
// (|atan2(ry, rx)| / 2 > 50 deg  <=>  rx / r < cos(100 deg);  angle(n_i, n_j) < 90 deg  <=>  n_i . n_j > 0).
// The splice of virtual cones into the real ones is a short serial edit list and runs on lane 0.
#pragma once

#include "lane.cuh"
#include "plan_types.cuh"

namespace fsd {

constexpr int WV_CAP = FSD_MAX_WV;

struct MatchSmem {
  d2 side[2][FSD_MAX_SORTED];  // sorted cones: [0] left, [1] right
  d2 wv[2][WV_CAP];            // with virtual cones: [0] left, [1] right
  d2 dirs[WV_CAP];             // search directions of the side being matched
  d2 odirs[WV_CAP];            // search directions of the other side
  d2 virt[WV_CAP];
  d2 ex[WV_CAP + 1];
  d2 ins[WV_CAP];
  double key[WV_CAP];
  int16_t match[2][WV_CAP];  // [0] l2r, [1] r2l
  int16_t tmp_match[WV_CAP];
  int32_t nside[2], nwv[2];
  int32_t scratch[4];
};

FSD_DEVFN void match_directions(const d2 *c, int n, int side, d2 *out) {
  // calculate_match_search_direction, match_directions.py:23-44
#pragma unroll 1
  for (int i = fsd_lane(); i < n; i += FSD_LANES) {
    int a = i == 0 ? 0 : (i == n - 1 ? n - 2 : i - 1);
    int b = i == 0 ? 1 : (i == n - 1 ? n - 1 : i + 1);
    double tx = c[b].x - c[a].x, ty = c[b].y - c[a].y;
    double rx = side == FSD_CONE_RIGHT ? -ty : ty, ry = side == FSD_CONE_RIGHT ? tx : -tx;
    const double inv = frsqrt(rx * rx + ry * ry);  // unit vector by one reciprocal square root
    out[i].x = rx * inv;
    out[i].y = ry * inv;
  }
}

// calculate_matches_for_side :340-384.  Leaves S.dirs = search directions of `cones`.
// Returns true when the reference raises (other side has exactly one cone, :130).
FSD_DEVFN bool matches_for_side(MatchSmem &S, const d2 *cones, int n, int side, const d2 *other, int m,
                                int16_t *match, const DevParams &P) {
  const int lane = fsd_lane();
  if (n <= 1) {
#pragma unroll 1
    for (int i = lane; i < n; i += FSD_LANES) match[i] = -1;
    wsync();
    return false;
  }
  match_directions(cones, n, side, S.dirs);
  if (m > 1) match_directions(other, m, side == FSD_CONE_RIGHT ? FSD_CONE_LEFT : FSD_CONE_RIGHT, S.odirs);
  wsync();
  if (m <= 1) {
#pragma unroll 1
    for (int i = lane; i < n; i += FSD_LANES) match[i] = -1;
    wsync();
    return m == 1;
  }
  const double inv_major2 = P.match_inv_major2, inv_minor2 = P.match_inv_minor2;
  const double cos_limit = P.cos_match_limit;
#pragma unroll 1
  for (int i = lane; i < n; i += FSD_LANES) {
    const double dxi = S.dirs[i].x, dyi = S.dirs[i].y;
    bool any = false;
    int best = 0;
    double best_d = 0.0;
    for (int j = 0; j < m; ++j) {
      double vx = other[j].x - cones[i].x, vy = other[j].y - cones[i].y;
      double rx = vx * dxi + vy * dyi, ry = dxi * vy - dyi * vx;
      double r2 = rx * rx + ry * ry;
      bool ok = (rx * rx * inv_major2 + ry * ry * inv_minor2) < 1.0;
      if (lt_scaled(rx, cos_limit, r2)) ok = false;                          // cos(angle) < limit :125
      if (dxi * S.odirs[j].x + dyi * S.odirs[j].y > 0.0) ok = false;  // :127
      any |= ok;
      // the match is the nearest cone of the other side, masked or not (:162)
      double dd = vx * vx + vy * vy;
      if (j == 0 || dd < best_d) {
        best_d = dd;
        best = j;
      }
    }
    match[i] = (int16_t)(any ? best : -1);
  }
  wsync();
  return false;
}

// insert_virtual_cones_to_existing :195-261 (lane 0).  Result in S.ex, returns its length.
FSD_DEVFN int insert_virtual(MatchSmem &S, const d2 *other, int no, int nv, const FramePose &F, const DevParams &P) {
  int ne, ni;
  if (no > nv) {
    for (int i = 0; i < no; ++i) S.ex[i] = other[i];
    for (int i = 0; i < nv; ++i) S.ins[i] = S.virt[i];
    ne = no;
    ni = nv;
  } else {
    for (int i = 0; i < nv; ++i) S.ex[i] = S.virt[i];
    for (int i = 0; i < no; ++i) S.ins[i] = other[i];
    ne = nv;
    ni = no;
  }
  // insertion order: ascending distance to the nearest existing cone (:212)
  int order[WV_CAP];
  for (int i = 0; i < ni; ++i) {
    double mn = 0.0;
    for (int j = 0; j < ne; ++j) {
      double ddx = S.ins[i].x - S.ex[j].x, ddy = S.ins[i].y - S.ex[j].y;
      double dd = ddx * ddx + ddy * ddy;
      if (j == 0 || dd < mn) mn = dd;
    }
    S.key[i] = mn;
    int p = i;
    while (p > 0 && S.key[order[p - 1]] > mn) {
      order[p] = order[p - 1];
      --p;
    }
    order[p] = i;
  }
  for (int oi = 0; oi < ni; ++oi) {
    const double cx = S.ins[order[oi]].x, cy = S.ins[order[oi]].y;
    int index;
    if (ne == 1) {
      // calculate_insert_index_for_one_cone :264-282
      double dvx = cx - F.px, dvy = cy - F.py, dex = S.ex[0].x - F.px, dey = S.ex[0].y - F.py;
      index = dvx * dvx + dvy * dvy < dex * dex + dey * dey ? 0 : 1;  // distances compared squared
    } else {
      int c1 = -1, c2 = -1;
      double d1 = 0.0, d2v = 0.0;
      for (int j = 0; j < ne; ++j) {
        double ddx = S.ex[j].x - cx, ddy = S.ex[j].y - cy;
        double d = ddx * ddx + ddy * ddy;  // only compared: squared
        if (c1 < 0 || d < d1) {
          c2 = c1;
          d2v = d1;
          c1 = j;
          d1 = d;
        } else if (c2 < 0 || d < d2v) {
          c2 = j;
          d2v = d;
        }
      }
      int gap = c1 - c2;
      if (gap != 1 && gap != -1) continue;  // virtual cone skipped (:226-227)
      // angle(closest - v, second - v) > 90 deg  <=>  the cone lies between the two (:229-241)
      bool between = (S.ex[c1].x - cx) * (S.ex[c2].x - cx) + (S.ex[c1].y - cy) * (S.ex[c2].y - cy) < 0.0;
      if (between)
        index = (c1 < c2 ? c1 : c2) + 1;
      else
        index = c1 < c2 ? c1 : c1 + 1;
    }
    if (ne >= WV_CAP) continue;
    for (int j = ne; j > index; --j) S.ex[j] = S.ex[j - 1];
    S.ex[index].x = cx;
    S.ex[index].y = cy;
    ++ne;
  }
  // interior points whose polyline angle is below 85 deg are removed, all at once (:252-259)
  const double cos85 = P.cos_85deg;
  bool drop[WV_CAP + 1];
  for (int i = 0; i < ne; ++i) drop[i] = false;
  for (int i = 1; i + 1 < ne; ++i) {
    // angle < 85 deg  <=>  a . b > cos(85 deg) |a| |b|, without the square root
    const double ax = S.ex[i + 1].x - S.ex[i].x, ay = S.ex[i + 1].y - S.ex[i].y;
    const double bx = S.ex[i - 1].x - S.ex[i].x, by = S.ex[i - 1].y - S.ex[i].y;
    drop[i] = gt_scaled(ax * bx + ay * by, cos85, (ax * ax + ay * ay) * (bx * bx + by * by));
  }
  int w = 0;
  for (int i = 0; i < ne; ++i)
    if (!drop[i]) S.ex[w++] = S.ex[i];
  return w;
}

// calculate_cones_for_other_side :387-440: the cones of `side` generate the OTHER side with virtual cones
FSD_DEVFN int cones_for_other_side(MatchSmem &S, const d2 *cones, int n, int side, const d2 *other, int m, d2 *out,
                                   const FramePose &F, const DevParams &P, bool *raises) {
  *raises |= matches_for_side(S, cones, n, side, other, m, S.tmp_match, P);
  if (fsd_lane() == 0) {
    int nv = 0;
    for (int i = 0; i < n; ++i)
      if (S.tmp_match[i] == -1) {
        // calculate_positions_of_virtual_cones :178-192
        S.virt[nv].x = cones[i].x + S.dirs[i].x * P.min_track_width;
        S.virt[nv].y = cones[i].y + S.dirs[i].y * P.min_track_width;
        ++nv;
      }
    int no;
    // combine_and_sort_virtual_with_real :306-337
    if (m == 0) {
      for (int i = 0; i < nv; ++i) out[i] = S.virt[i];
      no = nv;
    } else if (nv == 0) {
      for (int i = 0; i < m; ++i) out[i] = other[i];
      no = m;
    } else {
      no = insert_virtual(S, other, m, nv, F, P);
      for (int i = 0; i < no; ++i) out[i] = S.ex[i];
    }
    if (no < 2) {
      for (int i = 0; i < m; ++i) out[i] = other[i];
      no = m;
    }
    S.scratch[0] = no;
  }
  wsync();
  int no = S.scratch[0];
  wsync();
  return no;
}

// calculate_virtual_cones_for_both_sides :479-588.  Input S.side / S.nside; output S.wv, S.nwv, S.match.
FSD_DEVFN unsigned match_frame(MatchSmem &S, const FramePose &F, const DevParams &P) {
  int nl = S.nside[0], nr = S.nside[1];
  unsigned status = 0;
  if (nl < 2 && nr < 2) {
    if (fsd_lane() == 0) S.nwv[0] = S.nwv[1] = 0;
    wsync();
    return 0;
  }
  int mn = nl < nr ? nl : nr, mx = nl < nr ? nr : nl;
  if (mn == 0 || mx > 2 * mn) {
    if (nl < nr)
      nl = 0;
    else
      nr = 0;
  }
  bool raises = false;
  int nrw, nlw;
  if (nl >= 2) {
    nrw = cones_for_other_side(S, S.side[0], nl, FSD_CONE_LEFT, S.side[1], nr, S.wv[1], F, P, &raises);
  } else {
#pragma unroll 1
    for (int i = fsd_lane(); i < nr; i += FSD_LANES) S.wv[1][i] = S.side[1][i];
    nrw = nr;
  }
  if (nr >= 2) {
    nlw = cones_for_other_side(S, S.side[1], nr, FSD_CONE_RIGHT, S.side[0], nl, S.wv[0], F, P, &raises);
  } else {
#pragma unroll 1
    for (int i = fsd_lane(); i < nl; i += FSD_LANES) S.wv[0][i] = S.side[0][i];
    nlw = nl;
  }
  wsync();
  // match_both_sides_with_virtual_cones :443-476
  raises |= matches_for_side(S, S.wv[0], nlw, FSD_CONE_LEFT, S.wv[1], nrw, S.match[0], P);
  raises |= matches_for_side(S, S.wv[1], nrw, FSD_CONE_RIGHT, S.wv[0], nlw, S.match[1], P);
  if (fsd_lane() == 0) {
    S.nwv[0] = nlw;
    S.nwv[1] = nrw;
  }
  wsync();
  if (raises) status |= FSD_ST_REF_RAISES;
  return status;
}

}  // namespace fsd
