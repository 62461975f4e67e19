// Skidpad mission for ONE trajectory / ONE step by ONE warp (rows K1, K2 of SURVEY.md section 8a).
//
// Behaviour follows the reference's
//   fsd_path_planning/relocalization/skidpad/skidpad_relocalizer.py:31-240   (relocalization, once per trajectory)
//   fsd_path_planning/calculate_path/skidpad_calculate_path.py:49-71         (stateful window tracker)
//   fsd_path_planning/full_pipeline/full_pipeline.py:122-140, 178-194        (pose into / path out of the map frame)
// K1: the C(20,3) = 1140 hyper circle fits run one triple per lane; sklearn's DBSCAN(eps=3, min_samples=1) is
// single-linkage connected components, computed by lane-parallel label propagation; per-cluster medians by
// lane-parallel rank selection.  K2: the window arg-min is a lane-strided scan + warp reduction; steps of one
// trajectory are sequential (the window is centred on the previous index), trajectories are independent.
#pragma once

#include "lane.cuh"
#include "path.cuh"
#include "plan_types.cuh"

namespace fsd {

constexpr int SKID_NEAR = 20;
constexpr int SKID_TRIPLES = 1140;  // C(20, 3)

// relocalization result, 8 doubles per trajectory
struct SkidReloc {
  double tx, ty, rotation, rrx, rry, rcx, rcy, ok;  // translation, rotation, right reference / calculated centre, flag
};

struct SkidSmem {
  d2 near[SKID_NEAR];
  d2 cen[SKID_TRIPLES];
  double val[SKID_TRIPLES];
  int16_t label[SKID_TRIPLES];
  uint8_t tri[SKID_TRIPLES][3];
  uint8_t used[FSD_MAX_CONES];
  double sc[8];
  int32_t si[8];
};

// value of rank r (0-based, ascending; ties by index) among val[i] with label[i] == lab
FSD_DEVFN double select_rank(SkidSmem &S, int n, int lab, int coord, int r) {
  double out = 0.0;
  int found = 0;
#pragma unroll 1
  for (int i = fsd_lane(); i < n; i += FSD_LANES) {
    if (S.label[i] != lab) continue;
    const double v = coord == 0 ? S.cen[i].x : S.cen[i].y;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      if (S.label[j] != lab) continue;
      const double w = coord == 0 ? S.cen[j].x : S.cen[j].y;
      rank += (w < v) || (w == v && j < i);
    }
    if (rank == r) {
      out = v;
      found = 1;
    }
  }
  // exactly one lane found it
  double total = wsum(found ? out : 0.0);
  return total;
}

FSD_DEVFN double cluster_median(SkidSmem &S, int n, int lab, int coord, int count) {
  if (count % 2) return select_rank(S, n, lab, coord, count / 2);
  return 0.5 * (select_rank(S, n, lab, coord, count / 2 - 1) + select_rank(S, n, lab, coord, count / 2));
}

// SkidpadRelocalizer.do_relocalization_once.  cones: n points (fp64, any memory space).
FSD_DEVFN void skidpad_relocalize(SkidSmem &S, const double *cones, int n, double px, double py, double opx, double opy,
                                  double odx, double ody, const double *jitter, const double *ref, SkidReloc *out,
                                  int *n_accepted) {
  const int lane = fsd_lane();
  if (n > FSD_MAX_CONES) n = FSD_MAX_CONES;
  const int m = n < SKID_NEAR ? n : SKID_NEAR;
#pragma unroll 1
  for (int i = lane; i < n; i += FSD_LANES) S.used[i] = 0;
  wsync();
  // the 20 cones nearest to the vehicle, nearest first (:207-212)
  for (int q = 0; q < m; ++q) {
    double bv = 0.0;
    int bi = -1;
#pragma unroll 1
    for (int i = lane; i < n; i += FSD_LANES) {
      if (S.used[i]) continue;
      double d = fnorm(cones[2 * i] - px, cones[2 * i + 1] - py);
      if (bi < 0 || d < bv) {
        bv = d;
        bi = i;
      }
    }
    wargmin(bv, bi);
    if (lane == 0) {
      S.used[bi] = 1;
      S.near[q].x = cones[2 * bi];
      S.near[q].y = cones[2 * bi + 1];
    }
    wsync();
  }
  // triples in itertools.combinations order (only 3-subsets are ever fitted)
  int ntri = 0;
  if (lane == 0) {
    for (int a = 0; a < m; ++a)
      for (int b = a + 1; b < m; ++b)
        for (int c = b + 1; c < m; ++c) {
          S.tri[ntri][0] = (uint8_t)a;
          S.tri[ntri][1] = (uint8_t)b;
          S.tri[ntri][2] = (uint8_t)c;
          ++ntri;
        }
    S.si[0] = ntri;
  }
  wsync();
  ntri = S.si[0];
  int nacc = 0;
  for (int base = 0; base < ntri; base += FSD_LANES) {
    const int t = base + lane;
    bool accept = false;
    double cx = 0.0, cy = 0.0;
    if (t < ntri) {
      d2 p[3];
      for (int q = 0; q < 3; ++q) p[q] = S.near[S.tri[t][q]];
      // mean distance to the closest other point of the subset (:44-47)
      double mean_d = 0.0;
      for (int q = 0; q < 3; ++q) {
        double mn = INFINITY;
        for (int r = 0; r < 3; ++r) {
          if (r == q) continue;
          double d = fnorm(p[r].x - p[q].x, p[r].y - p[q].y);
          if (d < mn) mn = d;
        }
        mean_d += mn;
      }
      mean_d *= (1.0 / 3.0);
      for (int q = 0; q < 3; ++q) {
        p[q].x += jitter[6 * t + 2 * q] * 1e-3;
        p[q].y += jitter[6 * t + 2 * q + 1] * 1e-3;
      }
      // hyper circle fit of three points
      const double mx = (p[0].x + p[1].x + p[2].x) * (1.0 / 3.0), my = (p[0].y + p[1].y + p[2].y) * (1.0 / 3.0);
      double Mxy = 0, Mxx = 0, Myy = 0, Mxz = 0, Myz = 0, Mzz = 0;
      for (int q = 0; q < 3; ++q) {
        double xi = p[q].x - mx, yi = p[q].y - my, zi = xi * xi + yi * yi;
        Mxy += xi * yi;
        Mxx += xi * xi;
        Myy += yi * yi;
        Mxz += xi * zi;
        Myz += yi * zi;
        Mzz += zi * zi;
      }
      d2 ctr;
      const double r = hyper_from_moments(mx, my, Mxx * (1.0 / 3.0), Myy * (1.0 / 3.0), Mxy * (1.0 / 3.0),
                                          Mxz * (1.0 / 3.0), Myz * (1.0 / 3.0), Mzz * (1.0 / 3.0), &ctr);
      cx = ctr.x;
      cy = ctr.y;
      double resid = 0.0;
      for (int q = 0; q < 3; ++q) resid += fabs(fnorm(cx - p[q].x, cy - p[q].y) - r);
      resid *= (1.0 / 3.0);
      accept = fabs(r - 7.625) < 1.0 && fabs(mean_d - 2.4) < 1.5 && resid < 0.4;
    }
    const unsigned mask = wballot(accept);
    if (accept) {
      const int slot = nacc + FSD_POPC(mask & ((1u << lane) - 1u));
      S.cen[slot].x = cx;
      S.cen[slot].y = cy;
    }
    nacc += FSD_POPC(mask);
  }
  wsync();
  if (lane == 0) {
    *n_accepted = nacc;
    out->ok = 0.0;
  }
  if (nacc < 3) return;
  // DBSCAN(eps=3, min_samples=1) == connected components; label = smallest member index, i.e. clusters are
  // numbered in order of first appearance like sklearn's
#pragma unroll 1
  for (int i = lane; i < nacc; i += FSD_LANES) S.label[i] = (int16_t)i;
  wsync();
  for (int iter = 0; iter < nacc; ++iter) {
    bool changed = false;
#pragma unroll 1
    for (int i = lane; i < nacc; i += FSD_LANES) {
      int best = S.label[i];
      for (int j = 0; j < nacc; ++j) {
        double ddx = S.cen[i].x - S.cen[j].x, ddy = S.cen[i].y - S.cen[j].y;
        if (ddx * ddx + ddy * ddy <= 9.0 && S.label[j] < best) best = S.label[j];
      }
      if (best < S.label[i]) {
        S.label[i] = (int16_t)best;
        changed = true;
      }
    }
    wsync();
    if (!wany(changed)) break;
  }
  // cluster representatives (label[i] == i), ascending, with sizes and medians
  int nlab = 0;
  if (lane == 0) {
    for (int i = 0; i < nacc; ++i)
      if (S.label[i] == i) ++nlab;
    S.si[1] = nlab;
  }
  wsync();
  nlab = S.si[1];
  if (nlab < 2) return;
  // medians of every cluster -> val[2*l], val[2*l+1] (cluster order = ascending representative)
  int l = 0;
  for (int rep = 0; rep < nacc; ++rep) {
    if (S.label[rep] != rep) continue;
    int count = 0;
#pragma unroll 1
    for (int i = lane; i < nacc; i += FSD_LANES) count += S.label[i] == rep;
    count = wsum_i(count);
    const double mxv = cluster_median(S, nacc, rep, 0, count);
    const double myv = cluster_median(S, nacc, rep, 1, count);
    if (lane == 0) {
      S.val[2 * l] = mxv;
      S.val[2 * l + 1] = myv;
    }
    ++l;
  }
  wsync();
  if (lane == 0) {
    // pair of cluster medians closest to 18.25 m apart (:77-91)
    double best = 1000.0;
    int b0 = -1, b1 = -1;
    for (int a = 0; a < nlab; ++a)
      for (int b = a + 1; b < nlab; ++b) {
        double dist = fabs(18.25 - fnorm(S.val[2 * a] - S.val[2 * b], S.val[2 * a + 1] - S.val[2 * b + 1]));
        if (dist < best) {
          best = dist;
          b0 = a;
          b1 = b;
        }
      }
    if (!(best > 0.5)) {
      // calculate_transformation :101-169 (centres in the frame of the pose of the first attempt)
      const double c[2][2] = {{S.val[2 * b0], S.val[2 * b0 + 1]}, {S.val[2 * b1], S.val[2 * b1 + 1]}};
      const double on = fnorm(odx, ody), ux = fdiv(odx, on), uy = fdiv(ody, on);
      int right = -1, left = -1;
      for (int q = 0; q < 2; ++q) {
        const double vx = c[q][0] - opx, vy = c[q][1] - opy;
        const double ry = -vx * uy + vy * ux;
        if (ry < 0.0) {
          if (right < 0) right = q;
        } else if (left < 0) {
          left = q;
        }
      }
      if (right >= 0 && left >= 0) {
        out->tx = ref[0] - c[right][0];
        out->ty = ref[1] - c[right][1];
        out->rotation = fsd_atan2(ref[3] - ref[1], ref[2] - ref[0]) -
                        fsd_atan2(c[left][1] - c[right][1], c[left][0] - c[right][0]);
        out->rrx = ref[0];
        out->rry = ref[1];
        out->rcx = c[right][0];
        out->rcy = c[right][1];
        out->ok = 1.0;
      }
    }
  }
  wsync();
}

FSD_DEV void skid_to_known(const SkidReloc &R, double px, double py, double &ox, double &oy) {
  const double c = fsd_cos(R.rotation), s = fsd_sin(R.rotation);
  const double x = px + R.tx - R.rrx, y = py + R.ty - R.rry;
  ox = x * c - y * s + R.rrx;
  oy = x * s + y * c + R.rry;
}

FSD_DEV void skid_to_original(const SkidReloc &R, double px, double py, double &ox, double &oy) {
  const double c = fsd_cos(-R.rotation), s = fsd_sin(-R.rotation);
  const double x = px - R.tx - R.rcx, y = py - R.ty - R.rcy;
  ox = x * c - y * s + R.rcx;
  oy = x * s + y * c + R.rcy;
}

// K2, one trajectory: for every step the pose in the map frame and the index of the nearest path point inside the
// window around the previous index.  known[s] = (x, y, dir_x, dir_y); index[s]; *state = index_along_path (in/out).
FSD_DEVFN void skidpad_track(const SkidReloc &R, const double *table, int n_table, const double *pos, const double *dir,
                             int n_steps, int *state, double *known, int *index) {
  const int lane = fsd_lane();
  double mean = 0.0;
  for (int i = 0; i < 9; ++i) mean += fnorm(table[2 * i + 2] - table[2 * i], table[2 * i + 3] - table[2 * i + 1]);
  mean = fdiv(mean, 9.0);
  const int mac = (int)fdiv(20.0, mean);
  int cur = *state;
  for (int s = 0; s < n_steps; ++s) {
    double kx = pos[2 * s], ky = pos[2 * s + 1], dx = dir[2 * s], dy = dir[2 * s + 1];
    if (R.ok != 0.0) {
      const double yaw = fsd_atan2(dy, dx) + R.rotation;
      skid_to_known(R, pos[2 * s], pos[2 * s + 1], kx, ky);
      dx = fsd_cos(yaw);
      dy = fsd_sin(yaw);
      const int lo = cur - mac < 0 ? 0 : cur - mac, hi = cur + mac > n_table ? n_table : cur + mac;
      double bv = 0.0;
      int bi = -1;
#pragma unroll 1
      for (int i = lo + lane; i < hi; i += FSD_LANES) {
        double d = fnorm(kx - table[2 * i], ky - table[2 * i + 1]);
        if (bi < 0 || d < bv) {
          bv = d;
          bi = i;
        }
      }
      wargmin(bv, bi);
      cur = bi < 0 ? lo : bi;
    }
    if (lane == 0) {
      known[4 * s] = kx;
      known[4 * s + 1] = ky;
      known[4 * s + 2] = dx;
      known[4 * s + 3] = dy;
      index[s] = R.ok != 0.0 ? cur : -1;
    }
  }
  if (lane == 0) *state = cur;
  wsync();
}

// one step: path update = 25 m of the canonical path from `index` (or the trivial straight path before
// relocalization), MPC tail in the map frame, path back to the SLAM frame.  out_internal keeps the map-frame path
// (the next step's previous path).
FSD_DEVFN unsigned skidpad_step(PathSmem &S, const SkidReloc &R, const double *table, int n_table, int index,
                                const double *known, int force_P, const double *prev, const DevParams &P, double *out,
                                double *out_internal, int *grid) {
  const int lane = PG::lane();
  const FramePose F = make_pose(known[0], known[1], known[2], known[3]);
  int nu;
  if (index >= 0) {
    double mean = 0.0;
    for (int i = 0; i < 9; ++i) mean += fnorm(table[2 * i + 2] - table[2 * i], table[2 * i + 3] - table[2 * i + 1]);
    mean = fdiv(mean, 9.0);
    int fin = index + (int)fdiv(25.0, mean);
    if (fin > n_table) fin = n_table;
    nu = fin - index;
    if (nu > S.pcap - 64) nu = S.pcap - 64;
#pragma unroll 1
    for (int i = lane; i < nu; i += PG::N) {
      S.pts[1 + i].x = table[2 * (index + i)];
      S.pts[1 + i].y = table[2 * (index + i) + 1];
    }
  } else {
    // calculate_trivial_path, core_calculate_path.py:127-134
    const double max_angle = PI / 50.0, radius = 1000.0, stp = max_angle / (FSD_HORIZON - 1);
    const double c0 = fsd_cos(-PI / 2.0), s0 = fsd_sin(-PI / 2.0);
    const double yaw = fsd_atan2(F.dy, F.dx), cy = fsd_cos(yaw), sy = fsd_sin(yaw);
    nu = FSD_HORIZON - 1;
#pragma unroll 1
    for (int i = lane; i < nu; i += PG::N) {
      const int k = i + 1;
      const double a = k == FSD_HORIZON - 1 ? max_angle : (double)k * stp;
      const double px = (fsd_cos(a) - 1.0) * radius, py = fsd_sin(a) * radius;
      const double qx = px * c0 - py * s0, qy = px * s0 + py * c0;
      S.pts[1 + i].x = qx * cy - qy * sy + F.px;
      S.pts[1 + i].y = qx * sy + qy * cy + F.py;
    }
  }
  PG::sync();
  unsigned status = path_from_update(S, nu, F, force_P, prev, P, out_internal, grid);
  PG::sync();
#pragma unroll 1
  for (int i = lane; i < FSD_HORIZON; i += PG::N) {
    double x = out_internal[4 * i + 1], y = out_internal[4 * i + 2];
    if (index >= 0) skid_to_original(R, x, y, x, y);
    out[4 * i] = out_internal[4 * i];
    out[4 * i + 1] = x;
    out[4 * i + 2] = y;
    out[4 * i + 3] = out_internal[4 * i + 3];
  }
  PG::sync();
  return status;
}

}  // namespace fsd
