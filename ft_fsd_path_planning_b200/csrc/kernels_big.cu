// Kernels built with LARGE static bounds: the per-frame sources (spline.cuh, path.cuh) compiled a second time, in a
// namespace of their own, with room for 2 048 path points and 64 knots per fit.  A centre line taken from a GLOBAL PATH
// (PathPlanner.set_global_path, the acceleration mission's map; core_calculate_path.py:516-528) spans up to 60 m + 60 m of
// path -- 1 250 evaluation points at 0.1 m -- where the planner kernels of kernels.cu are sized for the <= 45 m a centre
// line of <= 12 matched cones can have.  Not on the batched hot path: one lane group per pose, 4 warps per CTA.
#define FSD_PCAP 2048
#define FSD_NCAP 64
#define fsd fsd_big  // the same sources, a second namespace (distinct symbols next to kernels.cu's instantiation)
#include <cstring>

#include "big_kernels.h"
#include "path.cuh"

using namespace fsd_big;

namespace {

constexpr int BIG_WPC = 4;                          // warps per CTA
constexpr int BIG_FPC = BIG_WPC * (32 / PG::N);     // lane groups (= poses / frames in flight) per CTA
constexpr size_t BIG_STRIDE = (sizeof(PathSmem) + 15) / 16 * 16;

// this lane group's slice of the CTA's shared memory and its point buffers in `scratch`
__device__ PathSmem &group_state(unsigned char *smem_raw, unsigned char *scratch) {
  const int grp = (int)threadIdx.x / PG::N;
  PathSmem &S = *reinterpret_cast<PathSmem *>(smem_raw + (size_t)grp * BIG_STRIDE);
  path_smem_bind(S, scratch + ((size_t)blockIdx.x * BIG_FPC + grp) * PATH_SCRATCH_BYTES, PCAP, NCAP);
  return S;
}

__device__ void store_result(const double *out, unsigned st, const int *grid, int b, double *out_f64, float *out_f32,
                             int16_t *grid_out, const fsd_gather *G = nullptr) {
  for (int i = PG::lane(); i < FSD_HORIZON * 4; i += PG::N) {
    const double v = out[i];
    if (out_f32) out_f32[(size_t)b * FSD_HORIZON * 4 + i] = (float)v;
    if (out_f64) out_f64[(size_t)b * FSD_HORIZON * 4 + i] = v;
    if (G) fsd_store_peers(*G, b, i, (float)v);
  }
  if (PG::lane() == 0 && grid_out) {
    grid_out[2 * (size_t)b] = (int16_t)grid[0];
    grid_out[2 * (size_t)b + 1] = (int16_t)grid[1];
  }
}

// one lane group per pose, free-running with dynamic fetch
__global__ void __launch_bounds__(32 * BIG_WPC)
    global_path_kernel(DevParams P, int n_poses, const double *pos, const double *dir, const double *gpath, int n_points,
                       const int16_t *force_P, const double *prev, int prev_stride, double *out_f64, float *out_f32,
                       int16_t *grid_out, uint32_t *status, unsigned char *scratch, int *counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PathSmem &S = group_state(smem_raw, scratch);
  for (;;) {
    int b = 0;
    if (PG::lane() == 0) b = atomicAdd(counter, 1);
    b = PG::bcast0(b);
    if (b >= n_poses) break;
    const FramePose F = make_pose(pos[2 * b], pos[2 * b + 1], dir[2 * b], dir[2 * b + 1]);
    int grid[2] = {0, 0};
    double *out = reinterpret_cast<double *>(S.W.r);  // the 40 x 4 result is assembled in shared memory (the fits' factor storage is dead)
    const unsigned st = path_global(S, gpath, n_points, F, force_P ? (int)force_P[b] : 0, prev + (size_t)b * prev_stride, P,
                                    out, grid);
    store_result(out, st, grid, b, out_f64, out_f32, grid_out);
    if (PG::lane() == 0) status[b] = st;
    PG::sync();
  }
}

// Second chance for frames on which a static bound of the batched path kernel overflowed (more than 32 knots in one fit,
// more than 704 path points): the path stage again, with the large bounds.  path_kernel marks such frames (bit 31 of the
// status word, the sort / match stage's own status bits parked in bits 16-30); every lane group of this kernel scans
// its share of the batch and re-plans the marked frames.  Almost always there is nothing to do.
__global__ void __launch_bounds__(32 * BIG_WPC)
    path_fixup_kernel(DevParams P, int n_frames, int coords_f64, const void *pos_v, const void *dir_v, const int16_t *n_wv,
                      const double *left_wv, const double *right_wv, const int16_t *l2r, const int16_t *r2l,
                      const int16_t *force_P, const double *prev, int prev_stride, double *out_f64, float *out_f32,
                      int16_t *grid_out, uint32_t *status, unsigned char *scratch, fsd_gather G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PathSmem &S = group_state(smem_raw, scratch);
  const int n_groups = (int)gridDim.x * BIG_FPC, g = (int)blockIdx.x * BIG_FPC + (int)threadIdx.x / PG::N;
  for (int base = g * PG::N; base < n_frames; base += n_groups * PG::N) {
    const int mine = base + PG::lane();
    unsigned marked = PG::ballot(mine < n_frames && (status[mine] >> 31) != 0u);
    while (marked) {
      const int b = base + __ffs((int)marked) - 1;
      marked &= marked - 1;
      double px, py, dx, dy;
      if (coords_f64) {
        const double *pos = static_cast<const double *>(pos_v), *dir = static_cast<const double *>(dir_v);
        px = pos[2 * b], py = pos[2 * b + 1], dx = dir[2 * b], dy = dir[2 * b + 1];
      } else {
        const float *pos = static_cast<const float *>(pos_v), *dir = static_cast<const float *>(dir_v);
        px = pos[2 * b], py = pos[2 * b + 1], dx = dir[2 * b], dy = dir[2 * b + 1];
      }
      const FramePose F = make_pose(px, py, dx, dy);
      int grid[2] = {0, 0};
      double *out = reinterpret_cast<double *>(S.W.r);
      const unsigned st = path_frame(S, reinterpret_cast<const d2 *>(left_wv + (size_t)b * FSD_MAX_WV * 2), n_wv[2 * (size_t)b],
                                     reinterpret_cast<const d2 *>(right_wv + (size_t)b * FSD_MAX_WV * 2), n_wv[2 * (size_t)b + 1],
                                     l2r + (size_t)b * FSD_MAX_WV, r2l + (size_t)b * FSD_MAX_WV, F,
                                     force_P ? (int)force_P[b] : 0, prev + (size_t)b * prev_stride, P, out, grid);
      store_result(out, st, grid, b, out_f64, out_f32, grid_out, (G.n_peers > 0 || G.multicast_out_path) ? &G : nullptr);
      if (PG::lane() == 0) status[b] = ((status[b] >> 16) & 0x7fffu) | st;  // the sort / match stage's bits + this run's
      PG::sync();
    }
  }
}

constexpr int FIXUP_CTAS = 8;

int big_grid(int n_poses, int sm_count) {
  const long need = ((long)n_poses + BIG_FPC - 1) / BIG_FPC, cap = (long)sm_count;
  return (int)(need < cap ? need : cap);
}

}  // namespace

size_t fsd_big_path_fixup_scratch_bytes() { return (size_t)FIXUP_CTAS * BIG_FPC * PATH_SCRATCH_BYTES; }

int fsd_big_path_fixup(const fsd_params *params, int n_frames, int coords_f64, const void *pos, const void *dir,
                       const int16_t *n_wv, const double *left_wv, const double *right_wv, const int16_t *l2r,
                       const int16_t *r2l, const int16_t *force_P, const double *prev, int prev_stride, double *out_f64,
                       float *out_f32, int16_t *grid_out, uint32_t *status, unsigned char *scratch, cudaStream_t stream,
                       const fsd_gather *gather) {
  const size_t smem = BIG_FPC * BIG_STRIDE;
  fsd_gather G;
  memset(&G, 0, sizeof(G));
  if (gather) G = *gather;
  if (cudaFuncSetAttribute(path_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return FSD_ERR_LAUNCH;
  }
  const int need = (n_frames + PG::N * BIG_FPC - 1) / (PG::N * BIG_FPC);
  path_fixup_kernel<<<need < FIXUP_CTAS ? need : FIXUP_CTAS, 32 * BIG_WPC, smem, stream>>>(
      make_dev_params(*params), n_frames, coords_f64, pos, dir, n_wv, left_wv, right_wv, l2r, r2l, force_P, prev,
      prev_stride, out_f64, out_f32, grid_out, status, scratch, G);
  return cudaGetLastError() == cudaSuccess ? FSD_OK : FSD_ERR_LAUNCH;
}

size_t fsd_big_global_path_scratch_bytes(int n_poses, int sm_count) {
  return (size_t)big_grid(n_poses > 0 ? n_poses : 1, sm_count) * BIG_FPC * PATH_SCRATCH_BYTES;
}

int fsd_big_global_path(const fsd_params *params, int n_poses, const double *pos, const double *dir, const double *gpath,
                        int n_points, const int16_t *force_P, const double *prev, int prev_stride, double *out_f64,
                        float *out_f32, int16_t *grid_out, uint32_t *status, unsigned char *scratch, int *counter,
                        int sm_count, cudaStream_t stream) {
  static_assert(sizeof(SplineWork::r) >= FSD_HORIZON * 4 * sizeof(double), "the result aliases the spline arena");
  const size_t smem = BIG_FPC * BIG_STRIDE;
  if (cudaFuncSetAttribute(global_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return FSD_ERR_LAUNCH;
  }
  global_path_kernel<<<big_grid(n_poses, sm_count), 32 * BIG_WPC, smem, stream>>>(
      make_dev_params(*params), n_poses, pos, dir, gpath, n_points, force_P, prev, prev_stride, out_f64, out_f32, grid_out,
      status, scratch, counter);
  return cudaGetLastError() == cudaSuccess ? FSD_OK : FSD_ERR_LAUNCH;
}
