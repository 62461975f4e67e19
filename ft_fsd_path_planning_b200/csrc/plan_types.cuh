// Device-side view of fsd_params (include/fsdplan.h) with the derived constants the kernels use.
#pragma once

#include "../../include/fsdplan.h"
#include "lane.cuh"

namespace fsd {

struct FramePose {
  double px, py, dx, dy;  // vehicle position and direction (not normalised)
  double ux, uy;          // unit direction
};

FSD_DEV FramePose make_pose(double px, double py, double dx, double dy) {
  FramePose F;
  F.px = px;
  F.py = py;
  F.dx = dx;
  F.dy = dy;
  const double inv = frcp(fsqrt(dx * dx + dy * dy));
  F.ux = dx * inv;
  F.uy = dy * inv;
  return F;
}

struct DevParams {
  int max_n_neighbors, max_length, max_dfs_pops;
  double max_dist, max_dist2, max_dist_to_first, thr_dir, thr_abs, car_size;
  double cos_5deg, cos_150deg, cos_seed_max, cos_seed_min, cos_match_limit, cos_85deg;
  double cos_thr_dir, cos_thr_abs, cos_1p3;
  double min_track_width, match_major, match_minor, max_search_angle;
  double seed_inv_major2, seed_inv_minor2, match_inv_major2, match_inv_minor2;
  double smoothing, predict_every, max_valid_dist, mpc_len, refit_smoothing;
};

static inline DevParams make_dev_params(const fsd_params &p) {
  DevParams d;
  d.max_n_neighbors = p.max_n_neighbors;
  d.max_length = p.max_length > FSD_MAX_SORTED ? FSD_MAX_SORTED : p.max_length;
  d.max_dfs_pops = p.max_dfs_pops;
  d.max_dist = p.max_dist;
  d.max_dist2 = p.max_dist * p.max_dist;
  d.max_dist_to_first = p.max_dist_to_first;
  d.thr_dir = p.threshold_directional_angle;
  d.thr_abs = p.threshold_absolute_angle;
  d.car_size = p.car_size;
  d.cos_5deg = cos(5.0 * PI / 180.0);
  d.cos_150deg = cos(150.0 * PI / 180.0);
  d.cos_seed_max = cos(PI - PI / 5.0);   // |bearing| < 4 pi / 5
  d.cos_seed_min = cos(PI / 10.0);        // |bearing| > pi / 10
  d.cos_match_limit = cos(2.0 * p.max_search_angle);
  d.cos_85deg = cos(85.0 * PI / 180.0);
  d.cos_thr_dir = cos(p.threshold_directional_angle);
  d.cos_thr_abs = cos(p.threshold_absolute_angle);
  d.cos_1p3 = cos(1.3);
  d.min_track_width = p.min_track_width;
  d.match_major = p.max_search_range * 1.5;
  d.match_minor = p.min_track_width;
  d.max_search_angle = p.max_search_angle;
  {
    const double major = p.max_dist_to_first * 1.5, minor = p.max_dist_to_first / 1.5;  // core_trace_sorter.py:388-394
    d.seed_inv_major2 = 1.0 / (major * major);
    d.seed_inv_minor2 = 1.0 / (minor * minor);
    d.match_inv_major2 = 1.0 / (d.match_major * d.match_major);
    d.match_inv_minor2 = 1.0 / (d.match_minor * d.match_minor);
  }
  d.smoothing = p.smoothing;
  d.predict_every = p.predict_every;
  d.max_valid_dist = p.maximal_distance_for_valid_path;
  d.mpc_len = p.mpc_path_length;
  d.refit_smoothing = p.refit_smoothing;
  return d;
}

#ifdef FSD_DEVICE_BUILD
#define FSD_POPC(x) __popc(x)
#define FSD_FFS(x) __ffs((int)(x))
#else
#define FSD_POPC(x) __builtin_popcount(x)
#define FSD_FFS(x) __builtin_ffs((int)(x))
#endif

}  // namespace fsd
