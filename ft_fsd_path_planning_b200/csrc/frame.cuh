// Per-frame drivers shared by the CUDA kernels (kernels.cu) and the host-check build
// (hostcheck.cpp): what one warp does with one frame, between "cones are in shared memory" and
// "results are in the output tensors".
#pragma once

#include "lane.cuh"
#include "match.cuh"
#include "path.cuh"
#include "plan_types.cuh"
#include "sort.cuh"

namespace fsd {

// output tensors of the sort(+match) stage; every pointer may be null except status
struct StageOut {
  int16_t *left_idx, *right_idx;  // [B][12]
  int16_t *sort_dbg;              // [B][8]
  int16_t *n_wv;                  // [B][2]
  double *left_wv, *right_wv;     // [B][FSD_MAX_WV][2]
  int16_t *l2r, *r2l;             // [B][FSD_MAX_WV]
  uint32_t *status;               // [B]
};

// plain (non-TMA) frame load with widening to fp64; the CUDA sort kernel stages xy with a bulk copy instead
template <typename T>
FSD_DEVFN void load_frame_plain(SortSmem &S, const T *xy, const uint8_t *type, int n) {
#pragma unroll 1
  for (int i = fsd_lane(); i < n; i += FSD_LANES) {
    S.xy[i].x = (double)xy[2 * i];
    S.xy[i].y = (double)xy[2 * i + 1];
    S.type[i] = type[i];
  }
  wsync();
}

FSD_DEVFN void store_sort(const SortSmem &S, int b, const StageOut &O) {
#pragma unroll 1
  for (int q = fsd_lane(); q < 2 * FSD_MAX_SORTED; q += FSD_LANES) {
    const int s = q / FSD_MAX_SORTED, j = q % FSD_MAX_SORTED;
    int16_t *dst = s == 0 ? O.left_idx : O.right_idx;
    if (dst) dst[(size_t)b * FSD_MAX_SORTED + j] = S.best[s][j];
  }
}

// S holds the sorted frame; the matching state M is filled from it (host builds: the CUDA path matches in a kernel of
// its own, which gathers the sorted cones from global memory)
FSD_DEVFN unsigned match_from_sort(const SortSmem &S, MatchSmem &M, const FramePose &F, const DevParams &P) {
  wsync();
#pragma unroll 1
  for (int q = fsd_lane(); q < 2 * FSD_MAX_SORTED; q += FSD_LANES) {
    const int s = q / FSD_MAX_SORTED, j = q % FSD_MAX_SORTED;
    if (j < S.nbest[s]) M.side[s][j] = S.xy[S.best[s][j]];
  }
  if (fsd_lane() == 0) {
    M.nside[0] = S.nbest[0];
    M.nside[1] = S.nbest[1];
  }
  wsync();
  return match_frame(M, F, P);
}

FSD_DEVFN void store_match(const MatchSmem &M, int b, const StageOut &O) {
  const int lane = fsd_lane();
  if (lane == 0 && O.n_wv) {
    O.n_wv[2 * (size_t)b] = (int16_t)M.nwv[0];
    O.n_wv[2 * (size_t)b + 1] = (int16_t)M.nwv[1];
  }
#pragma unroll 1
  for (int q = lane; q < 2 * WV_CAP; q += FSD_LANES) {
    const int s = q / WV_CAP, j = q % WV_CAP;
    double *wv = s == 0 ? O.left_wv : O.right_wv;
    int16_t *mt = s == 0 ? O.l2r : O.r2l;
    const bool live = j < M.nwv[s];
    if (wv) {
      wv[((size_t)b * WV_CAP + j) * 2] = live ? M.wv[s][j].x : 0.0;
      wv[((size_t)b * WV_CAP + j) * 2 + 1] = live ? M.wv[s][j].y : 0.0;
    }
    if (mt) mt[(size_t)b * WV_CAP + j] = live ? M.match[s][j] : (int16_t)-2;
  }
}

// path stage for frame b, reading the matching tensors written by store_match
FSD_DEVFN void path_from_tensors(PathSmem &S, int b, const StageOut &O, const FramePose &F, int force_P,
                                 const double *prev, const DevParams &P, double *out_f64, float *out_f32,
                                 int16_t *grid_out, int xcap = 0) {
  const int nl = O.n_wv[2 * (size_t)b], nr = O.n_wv[2 * (size_t)b + 1];
  const d2 *left = reinterpret_cast<const d2 *>(O.left_wv + (size_t)b * WV_CAP * 2);
  const d2 *right = reinterpret_cast<const d2 *>(O.right_wv + (size_t)b * WV_CAP * 2);
  int grid[2] = {0, 0};
  double *out = out_f64 + (size_t)b * FSD_HORIZON * 4;
  unsigned st = path_frame(S, left, nl, right, nr, O.l2r + (size_t)b * WV_CAP, O.r2l + (size_t)b * WV_CAP, F, force_P,
                           prev, P, out, grid, xcap);
  wsync();
  if (out_f32)
#pragma unroll 1
    for (int i = fsd_lane(); i < FSD_HORIZON * 4; i += FSD_LANES)
      out_f32[(size_t)b * FSD_HORIZON * 4 + i] = (float)out[i];
  if (fsd_lane() == 0) {
    O.status[b] |= st;
    if (grid_out) {
      grid_out[2 * (size_t)b] = (int16_t)grid[0];
      grid_out[2 * (size_t)b + 1] = (int16_t)grid[1];
    }
  }
}

}  // namespace fsd
