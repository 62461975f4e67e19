// fsd_plan_batch_cpu: the planner on the HOST, from the same per-frame sources as the CUDA kernels.
//
// sort.cuh / match.cuh / spline.cuh / path.cuh are written once against lane.cuh; this translation unit compiles them
// for the host as a warp of ONE lane (FSD_HOSTCHECK: every warp primitive degenerates to the identity), one frame after
// the other per thread.  It exists for BASELINE config 1 ("CPU plumbing, no GPU") and for machines without a GPU, and
// is an EXPLICIT entry point: nothing in the CUDA path falls back to it, the Python host layer uses it only when the
// caller asks for device="cpu".  It never touches oracle/ (test infrastructure).
#define FSD_HOSTCHECK 1
// no shared-memory budget on the host: the large static bounds of kernels_big.cu from the start
#define FSD_PCAP 2048
#define FSD_NCAP 64
#include <cstring>
#include <thread>
#include <vector>

#include "frame.cuh"

using namespace fsd;

namespace {

struct FrameOut {
  int16_t li[FSD_MAX_SORTED], ri[FSD_MAX_SORTED], n_wv[2], l2r[FSD_MAX_WV], r2l[FSD_MAX_WV], grid[2], dbg[8];
  double lw[FSD_MAX_WV * 2], rw[FSD_MAX_WV * 2], path[FSD_HORIZON * 4];
  uint32_t status;
};

void plan_range(const DevParams &P, int lo_frame, int hi_frame, const double *xy, const uint8_t *type,
                const int32_t *offsets, const double *pos, const double *dir, float *out_path, int16_t *out_li,
                int16_t *out_ri, const fsd_intermediate *inter, const int16_t *force_P, const double *prev, int stride,
                uint32_t *out_status) {
  SortSmem *S = new SortSmem();
  MatchSmem *MS = new MatchSmem();
  PathSmem *Q = new PathSmem();
  Q->pts = new d2[PCAP];
  Q->u = new double[PCAP];
  Q->pcap = PCAP;
  Q->W.cap = NCAP;
  FrameOut F0;
  StageOut O = {F0.li, F0.ri, F0.dbg, F0.n_wv, F0.lw, F0.rw, F0.l2r, F0.r2l, &F0.status};
  for (int b = lo_frame; b < hi_frame; ++b) {
    const int lo = offsets[b];
    int n = offsets[b + 1] - lo;
    unsigned st = 0;
    if (n > FSD_MAX_CONES) {
      n = FSD_MAX_CONES;
      st |= FSD_ST_OVERFLOW;
    }
    if (n < 0) n = 0;
    const FramePose F = make_pose(pos[2 * b], pos[2 * b + 1], dir[2 * b], dir[2 * b + 1]);
    load_frame_plain(*S, xy + 2 * (size_t)lo, type + lo, n);
    st |= sort_frame(*S, n, F, P, F0.dbg);
    store_sort(*S, 0, O);
    st |= match_from_sort(*S, *MS, F, P);
    store_match(*MS, 0, O);
    F0.status = st;
    F0.grid[0] = F0.grid[1] = 0;
    path_from_tensors(*Q, 0, O, F, force_P ? force_P[b] : 0, prev + (size_t)b * stride, P, F0.path, nullptr, F0.grid);
    // copy the frame's results to the caller's tensors
    const size_t B = (size_t)b;
    if (out_li) std::memcpy(out_li + B * FSD_MAX_SORTED, F0.li, sizeof(F0.li));
    if (out_ri) std::memcpy(out_ri + B * FSD_MAX_SORTED, F0.ri, sizeof(F0.ri));
    if (out_path)
      for (int i = 0; i < FSD_HORIZON * 4; ++i) out_path[B * FSD_HORIZON * 4 + i] = (float)F0.path[i];
    out_status[b] = F0.status;
    if (inter) {
      if (inter->path_f64) std::memcpy(inter->path_f64 + B * FSD_HORIZON * 4, F0.path, sizeof(F0.path));
      if (inter->n_wv) std::memcpy(inter->n_wv + B * 2, F0.n_wv, sizeof(F0.n_wv));
      if (inter->left_wv) std::memcpy(inter->left_wv + B * FSD_MAX_WV * 2, F0.lw, sizeof(F0.lw));
      if (inter->right_wv) std::memcpy(inter->right_wv + B * FSD_MAX_WV * 2, F0.rw, sizeof(F0.rw));
      if (inter->l2r) std::memcpy(inter->l2r + B * FSD_MAX_WV, F0.l2r, sizeof(F0.l2r));
      if (inter->r2l) std::memcpy(inter->r2l + B * FSD_MAX_WV, F0.r2l, sizeof(F0.r2l));
      if (inter->grid) std::memcpy(inter->grid + B * 2, F0.grid, sizeof(F0.grid));
      if (inter->sort_dbg) std::memcpy(inter->sort_dbg + B * 8, F0.dbg, sizeof(F0.dbg));
    }
  }
  delete[] Q->pts;
  delete[] Q->u;
  delete Q;
  delete MS;
  delete S;
}

}  // namespace

extern "C" int fsd_plan_batch_cpu(const fsd_params *params, int mission, int n_frames, const double *cones_xy,
                                  const uint8_t *cones_type, const int32_t *offsets, const double *pos,
                                  const double *dir, float *out_path, int16_t *out_left_idx, int16_t *out_right_idx,
                                  const fsd_intermediate *inter, const int16_t *force_P, const double *prev_path,
                                  int prev_path_stride, uint32_t *out_status, int n_threads) {
  if (!params || n_frames < 0) return FSD_ERR_ARG;
  if (mission != FSD_MISSION_AUTOCROSS && mission != FSD_MISSION_TRACKDRIVE) return FSD_ERR_MISSION;
  if (n_frames == 0) return FSD_OK;
  if (!offsets || !pos || !dir || !out_status || (offsets[n_frames] > offsets[0] && (!cones_xy || !cones_type)))
    return FSD_ERR_ARG;
  if (!out_path && !(inter && inter->path_f64)) return FSD_ERR_ARG;
  if (prev_path && prev_path_stride != 0 && prev_path_stride != FSD_HORIZON * 4) return FSD_ERR_ARG;
  const DevParams P = make_dev_params(*params);
  double initial[FSD_HORIZON * 4];
  const double *prev = prev_path;
  int stride = prev_path_stride;
  if (!prev) {  // the constant path of a fresh planner, computed by the same spline code
    PathSmem *Q = new PathSmem();
    Q->pts = new d2[PCAP];
    Q->u = new double[PCAP];
    Q->pcap = PCAP;
    Q->W.cap = NCAP;
    initial_path_frame(*Q, P, initial);
    delete[] Q->pts;
    delete[] Q->u;
    delete Q;
    prev = initial;
    stride = 0;
  }
  int T = n_threads < 1 ? 1 : n_threads;
  if (T > n_frames) T = n_frames;
  if (T == 1) {
    plan_range(P, 0, n_frames, cones_xy, cones_type, offsets, pos, dir, out_path, out_left_idx, out_right_idx, inter,
               force_P, prev, stride, out_status);
    return FSD_OK;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < T; ++t) {
    const int lo = (int)((long long)n_frames * t / T), hi = (int)((long long)n_frames * (t + 1) / T);
    pool.emplace_back(plan_range, std::cref(P), lo, hi, cones_xy, cones_type, offsets, pos, dir, out_path, out_left_idx,
                      out_right_idx, inter, force_P, prev, stride, out_status);
  }
  for (auto &th : pool) th.join();
  return FSD_OK;
}

// the constant initial path of a fresh planner on the host (fsd_initial_path's counterpart); out: [40][4] fp64
extern "C" int fsd_initial_path_cpu(const fsd_params *params, double *out) {
  if (!params || !out) return FSD_ERR_ARG;
  const DevParams P = make_dev_params(*params);
  PathSmem *Q = new PathSmem();
  Q->pts = new d2[PCAP];
  Q->u = new double[PCAP];
  Q->pcap = PCAP;
  Q->W.cap = NCAP;
  initial_path_frame(*Q, P, out);
  delete[] Q->pts;
  delete[] Q->u;
  delete Q;
  return FSD_OK;
}
