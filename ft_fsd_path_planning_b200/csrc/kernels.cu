// libfsdplan.so: CUDA kernels (sm_100a) and the C-ABI of include/fsdplan.h.
//
// Execution model: ONE WARP PLANS ONE FRAME.  CTAs hold 8 warps, each with its frame's state in its slice of the CTA's
// shared memory; grids are  min(ceil(B / 8), #SM x resident CTAs per SM)  persistent CTAs -- a multiple of the SM count
// whenever B allows -- whose warps (sort / match / k-NN) or CTAs (path) take frames from a counter in global memory.
//
//   sort_kernel         S1-S8.  The frame's cone coordinates are staged global -> shared with a TMA bulk copy
//                       (cp.async.bulk ... mbarrier::complete_tx, SASS: UBLKCP) and widened to fp64 in shared
//                       memory; the k-NN cost matrix never leaves registers/shared memory; the two sides are
//                       searched at the same time on the two half-warps.  Free-running warps, dynamic frame fetch.
//   match_kernel        M1-M6 on the sort indices.  Free-running warps, dynamic frame fetch.
//   knn_kernel          the cost-matrix step alone (stage entry point fsd_knn_batch).
//   path_kernel         P1-P4: three smoothing-spline fits, extension, curvature, 40 samples.
//   initial_path_kernel the constant path of a fresh planner, one warp.
//
// Per-frame algorithms live in sort.cuh / match.cuh / spline.cuh / path.cuh (see frame.cuh).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "big_kernels.h"
#include "frame.cuh"
#include "skidpad.cuh"

using namespace fsd;

namespace {

// Warps (= frames in flight) per CTA.
#ifndef FSD_WARPS_PER_CTA
#define FSD_WARPS_PER_CTA 8
#endif
constexpr int WPC = FSD_WARPS_PER_CTA;
#ifndef FSD_DEFAULT_PLAN_MODE
#define FSD_DEFAULT_PLAN_MODE 29 /* 1 + 4 + 8 + 16, see plan_mode() and profiles/r2_plan_mode_ab.txt */
#endif
constexpr int CTA_THREADS = 32 * WPC;
constexpr int CTAS_PER_SM = 16 / WPC;
// ---- TMA bulk copy helpers (raw PTX) ------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// shared-memory image of one sort/match CTA
struct SortCta {
  SortSmem S;  // includes the fp32 staging area of the bulk copy (S.raw)
  alignas(8) uint64_t mbar;
};
constexpr size_t SORT_CTA_STRIDE = (sizeof(SortCta) + 15) / 16 * 16;
constexpr size_t PATH_CTA_STRIDE = (sizeof(PathSmem) + 15) / 16 * 16;

// Stage the frame's coordinates into S.xy (fp64).  The 16-byte aligned interior of the frame's slice goes
// through the TMA bulk copy; a leading / trailing cone that is not 16-byte aligned (fp32 input, odd offset)
// is loaded directly.
template <typename T>
__device__ void stage_frame(SortCta &C, const T *xy, const uint8_t *type, int n, uint32_t &phase) {
  const int lane = fsd_lane();
  SortSmem &S = C.S;
  for (int i = lane; i < n; i += 32) S.type[i] = type[i];
  if constexpr (sizeof(T) == 8) {
    const bool aligned = (reinterpret_cast<uintptr_t>(xy) & 15u) == 0;
    if (aligned && n > 0) {
      if (lane == 0) {
        mbar_expect_tx(&C.mbar, (uint32_t)n * 16u);
        bulk_g2s(S.xy, xy, (uint32_t)n * 16u, &C.mbar);
      }
      mbar_wait(&C.mbar, phase);
      phase ^= 1u;
    } else {
      for (int i = lane; i < n; i += 32) {
        S.xy[i].x = (double)xy[2 * i];
        S.xy[i].y = (double)xy[2 * i + 1];
      }
    }
  } else {
    // fp32: cone i lives at byte 8 i of the slice
    const int head = (int)((reinterpret_cast<uintptr_t>(xy) >> 3) & 1u);  // 1 -> first cone is not 16 B aligned
    const int h = head < n ? head : n;
    const int interior = ((n - h) / 2) * 2;
    float *store = S.raw + 2 * head;  // keeps the interior 16 B aligned in shared memory
    if (interior > 0) {
      if (lane == 0) {
        mbar_expect_tx(&C.mbar, (uint32_t)interior * 8u);
        bulk_g2s(store + 2 * h, xy + 2 * h, (uint32_t)interior * 8u, &C.mbar);
      }
    }
    // leading / trailing cones outside the aligned interior
    if (lane == 0 && h == 1) {
      S.xy[0].x = (double)xy[0];
      S.xy[0].y = (double)xy[1];
    }
    if (lane == 1 && h + interior < n) {
      const int i = n - 1;
      S.xy[i].x = (double)xy[2 * i];
      S.xy[i].y = (double)xy[2 * i + 1];
    }
    if (interior > 0) {
      mbar_wait(&C.mbar, phase);
      phase ^= 1u;
      for (int i = h + lane; i < h + interior; i += 32) {
        S.xy[i].x = (double)store[2 * i];
        S.xy[i].y = (double)store[2 * i + 1];
      }
    }
  }
  __syncwarp();
}

// next frame of a free-running warp: frames are handed out in order from a counter in global memory (zeroed before the
// launch), so a warp that finishes early simply takes more frames -- no static partition, no ragged last wave
__device__ __forceinline__ int next_frame(int *counter) {
  int b = 0;
  if (fsd_lane() == 0) b = atomicAdd(counter, 1);
  return __shfl_sync(FULL, b, 0);
}

// Cone sorting: free-running warps, dynamic frame fetch.  Per frame: TMA staging, the k-NN graph of both sides by the
// whole warp, then the LEFT and the RIGHT search at the same time on the two half-warps (sort.cuh: sort_sides), conflict
// resolution, the sort indices to global memory.  Matching is a kernel of its own (match_kernel): with its code out of
// this kernel the SM's instruction cache holds the sorting code, and the warps need not be kept in lockstep to share it.
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS, CTAS_PER_SM)
    sort_kernel(DevParams P, int n_frames, const T *cones_xy, const uint8_t *cones_type, const int32_t *offsets,
                const T *pos, const T *dir, StageOut O, int *counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortCta &C = *reinterpret_cast<SortCta *>(smem_raw + (size_t)(threadIdx.x >> 5) * SORT_CTA_STRIDE);
  if (fsd_lane() == 0) mbar_init(&C.mbar, 1);
  __syncwarp();
  uint32_t phase = 0;
  for (;;) {
    const int b = next_frame(counter);
    if (b >= n_frames) break;
    unsigned st = 0;
    const int lo = offsets[b];
    int n = offsets[b + 1] - lo;
    if (n > FSD_MAX_CONES) {
      n = FSD_MAX_CONES;
      st |= FSD_ST_OVERFLOW;
    }
    if (n < 0) n = 0;
    const FramePose F = make_pose((double)pos[2 * b], (double)pos[2 * b + 1], (double)dir[2 * b], (double)dir[2 * b + 1]);
    int16_t *dbg = O.sort_dbg ? O.sort_dbg + 8 * (size_t)b : nullptr;
#ifdef FSD_FRAME_CYCLES
    const long long fsd_t0 = clock64();
#endif
    stage_frame<T>(C, cones_xy + 2 * (size_t)lo, cones_type + lo, n, phase);
    if (n >= 3) build_knn(C.S, n, P);
    int nl = 0, nr = 0;
    sort_sides(C.S, n, F, P, dbg, &st, nl, nr);
    st |= sort_finish(C.S, nl, nr);
    store_sort(C.S, b, O);
    if (fsd_lane() == 0) O.status[b] = st;
#ifdef FSD_FRAME_CYCLES
    // measurement build only: the frame's sort time in units of 64 cycles instead of the right side's DFS pop count
    if (dbg && fsd_lane() == 0) dbg[7] = (int16_t)min((clock64() - fsd_t0) >> 6, 32767ll);
#endif
    __syncwarp();
  }
}

// The cost-matrix step in isolation (stage entry point fsd_knn_batch): stage the frame, build both sides' k-NN graphs and
// write the mutual-edge adjacency lists.  Free-running warps, dynamic frame fetch; 6 KB of code, the N x N distances
// never leave registers.
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS, CTAS_PER_SM)
    knn_kernel(DevParams P, int n_frames, const T *cones_xy, const uint8_t *cones_type, const int32_t *offsets,
               uint8_t *out_nbr, uint8_t *out_deg, int *counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortCta &C = *reinterpret_cast<SortCta *>(smem_raw + (size_t)(threadIdx.x >> 5) * SORT_CTA_STRIDE);
  const int lane = fsd_lane();
  if (lane == 0) mbar_init(&C.mbar, 1);
  __syncwarp();
  uint32_t phase = 0;
  for (;;) {
    const int b = next_frame(counter);
    if (b >= n_frames) break;
    const int lo = offsets[b];
    int n = offsets[b + 1] - lo;
    n = n > FSD_MAX_CONES ? FSD_MAX_CONES : (n < 0 ? 0 : n);
    stage_frame<T>(C, cones_xy + 2 * (size_t)lo, cones_type + lo, n, phase);
    if (n >= 3) build_knn(C.S, n, P);
    // [cone][side][5] neighbour indices (ascending, frame-local; entries >= degree are unspecified) and [cone][side] degrees
    for (int e = lane; e < n * 10; e += 32) {
      const int i = e / 10, s = (e % 10) / 5, q = e % 5;
      out_nbr[(size_t)lo * 10 + e] = n >= 3 ? C.S.nbr[s][i][q] : (uint8_t)0;
    }
    for (int e = lane; e < n * 2; e += 32) out_deg[(size_t)lo * 2 + e] = n >= 3 ? C.S.deg[e & 1][e >> 1] : (uint8_t)0;
    __syncwarp();
  }
}

// matching on given sort indices (stage entry point fsd_match_batch; second kernel of the split sort stage): free-running
// warps, dynamic frame fetch, the <= 24 sorted cones gathered straight from global memory
constexpr size_t MATCH_CTA_STRIDE = (sizeof(MatchSmem) + 15) / 16 * 16;
// one frame's matching by one warp
template <typename T>
__device__ __forceinline__ void match_one(MatchSmem &M, const DevParams &P, int b, const T *cones_xy, const int32_t *offsets,
                                          const T *pos, const T *dir, const int16_t *left_idx, const int16_t *right_idx,
                                          const StageOut &O, int or_status) {
  const T *xy = cones_xy + 2 * (size_t)offsets[b];
  const FramePose F = make_pose((double)pos[2 * b], (double)pos[2 * b + 1], (double)dir[2 * b], (double)dir[2 * b + 1]);
  int nl = 0, nr = 0;
  {
    const int q = fsd_lane();
    const int s = q / FSD_MAX_SORTED, j = q % FSD_MAX_SORTED;
    const int idx = q < 2 * FSD_MAX_SORTED ? (int)(s == 0 ? left_idx : right_idx)[(size_t)b * FSD_MAX_SORTED + j] : -1;
    if (idx >= 0) {
      M.side[s][j].x = (double)xy[2 * idx];
      M.side[s][j].y = (double)xy[2 * idx + 1];
    }
    const unsigned have = __ballot_sync(FULL, idx >= 0);
    nl = __popc(have & ((1u << FSD_MAX_SORTED) - 1u));
    nr = __popc(have >> FSD_MAX_SORTED);
  }
  if (fsd_lane() == 0) {
    M.nside[0] = nl;
    M.nside[1] = nr;
  }
  __syncwarp();
  unsigned st = match_frame(M, F, P);
  store_match(M, b, O);
  if (fsd_lane() == 0) O.status[b] = or_status ? (O.status[b] | st) : st;
  __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(CTA_THREADS, CTAS_PER_SM)
    match_kernel(DevParams P, int n_frames, const T *cones_xy, const int32_t *offsets, const T *pos, const T *dir,
                 const int16_t *left_idx, const int16_t *right_idx, StageOut O, int or_status, int *counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MatchSmem &M = *reinterpret_cast<MatchSmem *>(smem_raw + (size_t)(threadIdx.x >> 5) * MATCH_CTA_STRIDE);
  for (;;) {
    const int b = next_frame(counter);
    if (b >= n_frames) break;
    match_one<T>(M, P, b, cones_xy, offsets, pos, dir, left_idx, right_idx, O, or_status);
  }
}

// Path stage.  One frame per lane group (PG, lane.cuh: the whole warp by default; two frames per warp is a build option
// that lost its A/B).  The PATH_FPC frames of a CTA
// run their path machines (path.cuh) in LOCKSTEP at the granularity of a spline fit: a group that finishes a fit waits
// (CTA barrier + vote in shared memory) until no group of the CTA is behind it, so all of them are inside the same fit --
// the same ~40 KB of code -- at the same time and share every instruction-cache fill (identical frames per CTA run 1.7x
// faster than distinct ones, profiles/r1_lockstep_probe.txt; waiting itself is free, profiles/r2_plan_mode_ab.txt).
// Rounds of PATH_FPC frames are handed out to the CTAs from a counter in global memory.
#ifndef FSD_PATH_WPC
#define FSD_PATH_WPC 8
#endif
#ifndef FSD_PATH_CTAS_PER_SM
#define FSD_PATH_CTAS_PER_SM 2
#endif
constexpr int PATH_WPC = FSD_PATH_WPC;             // warps per path CTA
constexpr int PATH_FPW = 32 / PG::N;               // frames per warp
constexpr int PATH_FPC = PATH_WPC * PATH_FPW;      // frames per CTA and round
constexpr int PATH_THREADS = 32 * PATH_WPC;
#ifndef FSD_PATH_SMEM_PAD
#define FSD_PATH_SMEM_PAD 0 /* measurement builds only: extra dynamic shared memory that lowers the CTAs resident per SM */
#endif
// (the frames' path machines live in their slots, PathSmem::M: shared memory, not the stack)
constexpr size_t PATH_KERNEL_SMEM = PATH_FPC * PATH_CTA_STRIDE + FSD_PATH_SMEM_PAD;
#if FSD_PATH_SMEM_PAD == 0 && FSD_PATH_WPC == 8 && FSD_PATH_CTAS_PER_SM == 2
// two CTAs (+ 1 KB each that the driver reserves) under the 132 KB shared-memory carve-out: what is left of the SM's 256 KB is
// the L1 that serves the point buffers -- a slot that grows past this costs 32 KB of it (section 3 of DESIGN.md)
static_assert(2 * (PATH_KERNEL_SMEM + 1024) <= 132 * 1024, "the path kernel's frame slots outgrew the 132 KB carve-out");
#endif
// knot records behind frame slot 0 when its arena is extended over the shared memory of all the CTA's slots
constexpr int PATH_XCAP_RAW = (int)((PATH_FPC * PATH_CTA_STRIDE - (sizeof(PathSmem) - sizeof(SplineWork::r))) / sizeof(KnotRec));
constexpr int PATH_XCAP = PATH_XCAP_RAW < 192 ? PATH_XCAP_RAW : 192;

// drop a frame's point-buffer lines from L2 (they are dead; written back they were 5x the algorithmic bytes of the step)
__device__ __forceinline__ void discard_points(unsigned char *mine, bool aligned) {
  static_assert(PATH_SCRATCH_BYTES % 128 == 0, "point buffers must be whole L2 lines");
  if (aligned)
    for (size_t o = (size_t)PG::lane() * 128; o + 128 <= PATH_SCRATCH_BYTES; o += (size_t)PG::N * 128)
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(mine + o) : "memory");
}

// ---- work-ordered scheduling of the path stage --------------------------------------------------------------------------
// The frames of a round run their fits in lockstep, so a round costs what its slowest frame costs, and a heavy frame taken
// late is what the whole kernel ends up waiting for.  How heavy a frame is can be guessed from its centre line before any
// spline is fitted (path_key_kernel; tools/dump_cycles.py, profiles/r2_session2_ab.txt): frames are taken in descending
// order of that guess -- rounds become homogeneous (less waiting at the fit boundaries, more shared instruction-cache
// fills) and the heavy frames come first (short tail).  10 240 frames: path kernel 1.37 -> 1.19 ms, 1.08 ms with a perfect
// predictor.  The order changes WHEN a frame is planned, never its result.
constexpr int ORDER_BINS = 64;  // (= COUNTER_STRIDE - 8: the histogram lives in the counter ring)

// one warp per frame: how far the centre line of the matches (the points fit #1 will get, path.cuh pm_begin_frame) is from
// a cubic polynomial in its chord length -- the squared residual of the least-squares cubic, which is what the first knot
// pass of the spline fits sees --, on a logarithmic scale, combined with how the centre line turns and how many points it
// has; ORDER_BINS bins; histogram of the bins.  (Spearman correlation with the measured per-frame time: residual alone 0.66,
// total turning angle alone 0.51, the combination 0.73.)
// (fp32 arithmetic: only the ORDER of the keys matters -- with the parameter centred on [-1, 1] the cubic's normal equations
// are well conditioned, and the fp32 keys fall into the same bins as fp64 ones on all 28 432 frames of the three bench
// streams; in fp64 -- divisions, log1p, atan2 -- this function cost the matching kernel 0.05 ms per 10 240 frames.)
// left / right / l2r / r2l: the frame's with-virtual cones and matches, in global OR shared memory.
__device__ __noinline__ int path_key_bin(const d2 *left, int nl, const d2 *right, int nr, const int16_t *l2r,
                                         const int16_t *r2l) {
  const int lane = (int)(threadIdx.x & 31u);
  const int ml = lane < nl ? (int)l2r[lane] : -1;
  const int mr = lane < nr ? (int)r2l[lane] : -1;
  const int nml = __popc(__ballot_sync(FULL, ml >= 0)), nmr = __popc(__ballot_sync(FULL, mr >= 0));
  const int sl = __reduce_add_sync(FULL, ml >= 0 ? ml : 0), sr = __reduce_add_sync(FULL, mr >= 0 ? mr : 0);
  const bool use_left = !(nmr > nml || (nmr == nml && sr > sl));  // select_side_to_use (core_calculate_path.py:165-183)
  const d2 *a = use_left ? left : right, *o = use_left ? right : left;
  const int m = use_left ? ml : mr;
  const unsigned have = __ballot_sync(FULL, m >= 0);
  const int nc = __popc(have);
  double cx = 0.0, cy = 0.0;
  if (m >= 0) {
    cx = 0.5 * (a[lane].x + o[m].x);
    cy = 0.5 * (a[lane].y + o[m].y);
  }
  // compaction: lane j fetches the point of the lane that holds the j-th match; relative to the first point (the
  // subtraction in fp64: SLAM coordinates are large), then fp32
  const int src = __fns(have, 0, lane + 1);  // lane of the (lane+1)-th set bit, or -1
  const double gx = __shfl_sync(FULL, cx, src < 0 ? 0 : src), gy = __shfl_sync(FULL, cy, src < 0 ? 0 : src);
  const bool on = lane < nc;
  const double x0 = __shfl_sync(FULL, gx, 0), y0 = __shfl_sync(FULL, gy, 0);
  const float px = on ? (float)(gx - x0) : 0.0f, py = on ? (float)(gy - y0) : 0.0f;
  // chord-length parameter, scaled to [-1, 1]
  const float qx = __shfl_up_sync(FULL, px, 1), qy = __shfl_up_sync(FULL, py, 1);
  float u = (on && lane > 0) ? sqrtf((px - qx) * (px - qx) + (py - qy) * (py - qy)) : 0.0f;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float v = __shfl_up_sync(FULL, u, off);
    if (lane >= off) u += v;
  }
  const float total = __shfl_sync(FULL, u, nc > 0 ? nc - 1 : 0);
  float res = 0.0f;
  if (nc >= 5 && total > 0.0f) {
    const float t = on ? 2.0f * __fdividef(u, total) - 1.0f : 0.0f, wgt = on ? 1.0f : 0.0f;
    // normal equations of the cubic: moments s[k] = sum t^k (k = 0..6), bx[k] = sum t^k x, by[k] = sum t^k y (k = 0..3)
    float v[15];
    float tk = wgt;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      v[k] = tk;
      if (k < 4) {
        v[7 + k] = tk * px;
        v[11 + k] = tk * py;
      }
      tk *= t;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int e = 0; e < 15; ++e) v[e] += __shfl_xor_sync(FULL, v[e], off);
    }
    // every lane solves the same 4 x 4 system with two right-hand sides (Gaussian elimination; SPD, no pivoting)
    float A[4][4], rx[4], ry[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) A[i][j] = v[i + j];
      rx[i] = v[7 + i];
      ry[i] = v[11 + i];
    }
    bool ok = true;
    float inv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok = ok && A[i][i] > 1e-30f;
      inv[i] = __frcp_rn(ok ? A[i][i] : 1.0f);
#pragma unroll
      for (int r = i + 1; r < 4; ++r) {
        const float f = A[r][i] * inv[i];
#pragma unroll
        for (int c = i; c < 4; ++c) A[r][c] -= f * A[i][c];
        rx[r] -= f * rx[i];
        ry[r] -= f * ry[i];
      }
    }
    float kx[4], ky[4];
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      float sx = rx[i], sy = ry[i];
#pragma unroll
      for (int c = i + 1; c < 4; ++c) {
        sx -= A[i][c] * kx[c];
        sy -= A[i][c] * ky[c];
      }
      kx[i] = sx * inv[i];
      ky[i] = sy * inv[i];
    }
    if (ok && on) {
      const float ex = ((kx[3] * t + kx[2]) * t + kx[1]) * t + kx[0] - px;
      const float ey = ((ky[3] * t + ky[2]) * t + ky[1]) * t + ky[0] - py;
      res = ex * ex + ey * ey;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) res += __shfl_xor_sync(FULL, res, off);
  }
  // turning of the centre line: total absolute turning angle minus its largest single term (many moderate bends cost more
  // knots than one sharp corner)
  float ang = 0.0f;
  {
    const float nx = __shfl_down_sync(FULL, px, 1), ny = __shfl_down_sync(FULL, py, 1);
    if (lane >= 1 && lane + 1 < nc) {
      const float ax = px - qx, ay = py - qy, bx = nx - px, by = ny - py;
      ang = fabsf(atan2f(ax * by - ay * bx, ax * bx + ay * by));
    }
  }
  float turn = ang, mx = ang;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    turn += __shfl_xor_sync(FULL, turn, off);
    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, off));
  }
  // least-squares combination of the features against the measured per-frame times of 10 240 bench frames (spearman 0.73,
  // the same on the mixed stream it was not fitted on); only the ORDER of the keys matters
  const float k = (res > 0.0f && res == res ? log1pf(100.0f * res) : 0.0f) + (turn - mx) - 0.43f * (float)nc;
  int bin = (int)((k + 5.5f) * 4.0f);
  bin = bin < 0 ? 0 : (bin > ORDER_BINS - 1 ? ORDER_BINS - 1 : bin);
  return __shfl_sync(FULL, bin, 0);
}

// one warp per frame: the frame's bin; the frame is appended to its bin's list bin_items [ORDER_BINS][n_frames] (the path
// kernel looks its rounds up in the lists: no counting-sort launch), or -- key != nullptr, FSD_PLAN_MODE bit 8 -- the bin goes
// to key[b] for path_order_kernel.  hist [ORDER_BINS] zero at launch.
// (Measured and rejected, profiles/r2_session4_ab.txt: the keys computed by the matching kernel's warps right after a frame's
// matching -- no key launch at all, but the persistent 16-warps-per-SM kernel pays 0.026 ms for what this kernel's 10 240
// independent warps do in 0.01 ms; and sorting + matching as ONE kernel whose warps match the sorted frames when there is
// nothing left to sort, to fill the idle tail of the sort kernel: 0.19 ms SLOWER, the matching code evicts the sorting code
// from the SM's instruction cache exactly while the last heavy frames, the kernel's critical path, are being sorted.)
__global__ void __launch_bounds__(256) path_key_kernel(int n_frames, StageOut O, uint8_t *key, int *hist, int *bin_items) {
  const int b = (int)blockIdx.x * 8 + (int)(threadIdx.x >> 5);
  if (b >= n_frames) return;
  const int bin = path_key_bin(reinterpret_cast<const d2 *>(O.left_wv + (size_t)b * WV_CAP * 2), O.n_wv[2 * (size_t)b],
                               reinterpret_cast<const d2 *>(O.right_wv + (size_t)b * WV_CAP * 2), O.n_wv[2 * (size_t)b + 1],
                               O.l2r + (size_t)b * WV_CAP, O.r2l + (size_t)b * WV_CAP);
  if ((threadIdx.x & 31u) == 0) {
    if (key) {
      key[b] = (uint8_t)bin;
      atomicAdd(&hist[bin], 1);
    } else {
      bin_items[(size_t)bin * n_frames + atomicAdd(&hist[bin], 1)] = b;
    }
  }
}

// one CTA: counting sort of the frames by bin, heaviest bin first (the order inside a bin is whatever the atomics give)
__global__ void __launch_bounds__(1024) path_order_kernel(int n_frames, const uint8_t *key, const int *hist, int *order) {
  __shared__ int s_cursor[ORDER_BINS];
  if (threadIdx.x == 0) {
    int at = 0;
    for (int bin = ORDER_BINS - 1; bin >= 0; --bin) {
      s_cursor[bin] = at;
      at += hist[bin];
    }
  }
  __syncthreads();
  for (int b = (int)threadIdx.x; b < n_frames; b += (int)blockDim.x) order[atomicAdd(&s_cursor[key[b]], 1)] = b;
}

// a finished frame: the 40 x 4 result from shared memory to the caller's buffers (and, multi-GPU, to every peer's gathered
// buffer: the all-gather of the paths happens HERE), the status word, the grid sizes
__device__ __noinline__ void store_path_frame(const double *out, unsigned st, const int *gr, int b, uint32_t *status,
                                                 double *out_f64, float *out_f32, int16_t *grid_out, int flags,
                                                 const fsd_gather &G) {
  const int lane = PG::lane();
  PG::sync();
  for (int i = lane; i < FSD_HORIZON * 4; i += PG::N) {
    const double v = out[i];
    if (out_f32) out_f32[(size_t)b * FSD_HORIZON * 4 + i] = (float)v;
    if (out_f64) out_f64[(size_t)b * FSD_HORIZON * 4 + i] = v;  // only when the caller asked for the fp64 path
    if (flags & 4) fsd_store_peers(G, b, i, (float)v);
  }
  if (lane == 0) {
    const unsigned before = status[b];
    unsigned after = before | st;
    // a static bound overflowed even with the CTA's whole memory: marked for the large-bounds kernel (kernels_big.cu), the
    // earlier stages' bits parked in bits 16-30 meanwhile
    if ((flags & 2) && (st & FSD_ST_OVERFLOW)) after |= 0x80000000u | ((before & 0x7fffu) << 16);
    status[b] = after;
    if (grid_out) {
      grid_out[2 * (size_t)b] = (int16_t)gr[0];
      grid_out[2 * (size_t)b + 1] = (int16_t)gr[1];
    }
  }
  PG::sync();
}

template <typename T>
__global__ void __launch_bounds__(PATH_THREADS, FSD_PATH_CTAS_PER_SM)
    path_kernel(DevParams P, int n_frames, const T *pos, const T *dir, StageOut O, const int16_t *force_P,
                const double *prev, int prev_stride, double *out_f64, float *out_f32, int16_t *grid_out,
                unsigned char *scratch, int *counter, int flags, int cap0, const int *order, const int *bin_hist,
                const int *bin_items, const __grid_constant__ fsd_gather G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_state[PATH_FPC];
  __shared__ int s_base;
  __shared__ int s_bin_end[ORDER_BINS];  // frames in the bins up to and including q, heaviest bin (ORDER_BINS - 1) = q 0
  if (bin_hist) {
    if (threadIdx.x < 32) {
      // inclusive scan of the 64 bin counts in processing order, two per lane
      const int l = (int)threadIdx.x;
      const int c0 = bin_hist[ORDER_BINS - 1 - l], c1 = bin_hist[ORDER_BINS - 1 - (l + 32)];
      int a = c0, bsum = c1;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int va = __shfl_up_sync(FULL, a, off), vb = __shfl_up_sync(FULL, bsum, off);
        if (l >= off) {
          a += va;
          bsum += vb;
        }
      }
      const int first_half = __shfl_sync(FULL, a, 31);
      s_bin_end[l] = a;
      s_bin_end[l + 32] = first_half + bsum;
    }
    __syncthreads();
  }
  const int grp = (int)threadIdx.x / PG::N, lane = PG::lane();  // this frame's slot in the CTA, lane within the frame
  PathSmem &S = *reinterpret_cast<PathSmem *>(smem_raw + (size_t)grp * PATH_CTA_STRIDE);
  unsigned char *mine = scratch + ((size_t)blockIdx.x * PATH_FPC + grp) * PATH_SCRATCH_BYTES;
  const bool aligned = (reinterpret_cast<uintptr_t>(scratch) & 127u) == 0;  // never touch a line shared with a neighbour
  path_smem_bind(S, mine, PCAP, cap0, (flags & 8) ? 1 : 0);  // cap0 = NCAP (less only in tests of the suspend / resume path)
  for (int base = (int)blockIdx.x * PATH_FPC;; base += (int)gridDim.x * PATH_FPC) {
    if (counter) {
      // rounds handed out from a counter (zeroed before the launch): a CTA whose frames were cheap takes more rounds, so
      // the kernel ends when the WORK ends, not when the unluckiest static share ends
      __syncthreads();
      if (threadIdx.x == 0) s_base = atomicAdd(counter, PATH_FPC);
      __syncthreads();
      base = s_base;
    }
    if (base >= n_frames) break;
    const bool active = base + grp < n_frames;
    int b = base + grp;
    if (active && order) b = order[b];  // (work-ordered: see path_key_kernel)
    if (bin_hist) {
      // the (base + grp)-th frame of the heaviest-first order: the first bin whose running count exceeds it
      // (binary search; j < n_frames = s_bin_end[ORDER_BINS - 1])
      const int j = active ? base + grp : 0;
      int q = 0;
#pragma unroll
      for (int step = ORDER_BINS / 2; step > 0; step >>= 1)
        if (s_bin_end[q + step - 1] <= j) q += step;
      if (active) b = bin_items[(size_t)(ORDER_BINS - 1 - q) * n_frames + (j - (q ? s_bin_end[q - 1] : 0))];
    }
#ifdef FSD_FRAME_CYCLES
    const long long fsd_t0 = clock64();
    long long fsd_t1 = fsd_t0;
#endif
    PathMachine &M = S.M;
    M.state = PS_DONE;
    M.status = 0;
    M.P_grid = M.n_trim = 0;
    // the frame's 40 x 4 result is assembled in shared memory (the fits' factor storage is dead by the time it is written)
    static_assert(sizeof(S.W.r) >= FSD_HORIZON * 4 * sizeof(double), "the result aliases the spline arena");
    double *out = reinterpret_cast<double *>(S.W.r);
    if (active) {
      const FramePose F =
          make_pose((double)pos[2 * b], (double)pos[2 * b + 1], (double)dir[2 * b], (double)dir[2 * b + 1]);
      const int nl = O.n_wv[2 * (size_t)b], nr = O.n_wv[2 * (size_t)b + 1];
      const d2 *left = reinterpret_cast<const d2 *>(O.left_wv + (size_t)b * WV_CAP * 2);
      const d2 *right = reinterpret_cast<const d2 *>(O.right_wv + (size_t)b * WV_CAP * 2);
      pm_begin_frame(S, M, left, nl, right, nr, O.l2r + (size_t)b * WV_CAP, O.r2l + (size_t)b * WV_CAP, F,
                     force_P ? (int)force_P[b] : 0, prev + (size_t)b * prev_stride, P, out);
    }
#if defined(FSD_NO_LOCKSTEP)
    while (M.state != PS_DONE && !pm_suspended(M)) pm_step(S, M, P);  // A/B and measurement builds only: free-running groups
#ifdef FSD_FRAME_CYCLES
    fsd_t1 = clock64();
#endif
    __syncthreads();
#else
    for (;;) {
      // free-run to the next alignment point (the end of a spline fit): the groups of the CTA are then inside the same
      // fit at the same time without waiting for each other after every pass
      while (M.state != PS_DONE && !pm_is_alignment_state(M.state) && !pm_suspended(M)) pm_step(S, M, P);
      if (lane == 0) s_state[grp] = pm_suspended(M) ? (int)PS_DONE : M.state;  // nobody waits for a suspended frame
      __syncthreads();
      int behind = PS_DONE;
#pragma unroll
      for (int g = 0; g < PATH_FPC; ++g) behind = min(behind, s_state[g]);
      __syncthreads();
      if (behind == PS_DONE) break;
      if (M.state != PS_DONE && !pm_suspended(M) && !(pm_is_alignment_state(M.state) && behind < M.state)) pm_step(S, M, P);
    }
#endif
    // A frame whose fit wants more knots than its arena holds is SUSPENDED, not truncated (spline.cuh: FIT_SUSPENDED); the
    // others of its round do not wait for it.  When the round is over -- every other frame stored, their shared memory
    // dead -- the frame's own warp moves its state to frame slot 0, whose arena now extends over the shared memory of ALL
    // the CTA's slots (PATH_XCAP knot records), and carries on from where it stopped: same code, warm instruction caches,
    // nothing recomputed, while the other CTAs keep taking rounds from the counter.  Rare (about one frame in 10^5), but a
    // frame that is planned again from scratch by a kernel of its own after this one costs a whole single-warp path
    // calculation of latency at the end of the step.
    const bool resume = active && pm_suspended(M);
    if (active && !resume) {
      int gr[2] = {M.P_grid, M.n_trim};
#ifdef FSD_FRAME_CYCLES
      gr[1] = (int)min((fsd_t1 - fsd_t0) >> 8, 32767ll);  // measurement build only: the frame's path-machine time / 256 cycles
#endif
      store_path_frame(out, M.status, gr, b, O.status, out_f64, out_f32, grid_out, flags, G);
    }
    if (flags & 8) {
      if (lane == 0) s_state[grp] = resume ? 1 : 0;
      __syncthreads();
      int any = 0;
#pragma unroll
      for (int g = 0; g < PATH_FPC; ++g) any |= s_state[g] << g;
      if (any) {
        PathSmem &X = *reinterpret_cast<PathSmem *>(smem_raw);  // frame slot 0
        const int first = __ffs(any) - 1;
        // Further suspended frames of the same round (their images would not survive the first one's resume) are flagged
        // like a truncated fit -- with fixup scratch (flags & 2) the large-bounds kernel plans them again --, and they are
        // dealt with FIRST: their slots, path machines included, are intact now, not after the resume.
        if (resume && grp != first) {
          pm_give_up_suspended(M);
          int gr[2] = {0, 0};
          store_path_frame(out, M.status, gr, b, O.status, out_f64, out_f32, grid_out, flags, G);
        }
        __syncthreads();
        if (grp == first) {
          if (first != 0) {
            // the frame's image (pointers to ITS point buffers, path machine, spline header, knot records) moves to slot 0
            const double *src = reinterpret_cast<const double *>(&S);
            double *dst = reinterpret_cast<double *>(&X);
            for (int i = lane; i < (int)(sizeof(PathSmem) / sizeof(double)); i += PG::N) dst[i] = src[i];
            PG::sync();
          }
          if (lane == 0) {
            X.W.cap = PATH_XCAP;
            X.W.suspendable = 0;  // whatever does not fit now is flagged (FSD_ST_OVERFLOW) as ever
          }
          PG::sync();
          PathMachine &MX = X.M;  // (the machine moved with the image; this frame's own slot is about to be run over)
          MX.out = reinterpret_cast<double *>(X.W.r);
          fit_resume(X.W, MX.fit);
          pm_run(X, MX, P);
          int gr[2] = {MX.P_grid, MX.n_trim};
          store_path_frame(MX.out, MX.status, gr, b, O.status, out_f64, out_f32, grid_out, flags, G);
        }
        __syncthreads();
        // The slots belong to their warps again.  EVERY slot is bound afresh: the extended arena of the resumed fit runs over
        // the neighbours' slots, and a fit that ends up with >= 39 knots writes band rows (record 34 and up) on top of slot
        // 1's header -- its point-buffer pointers and arena capacity (fewer knots only touch the neighbours' dead records).
#ifdef FSD_REBIND_SLOT0_ONLY /* the behaviour before r2_zc, kept for tools/resume_probe.py */
        if (grp == 0)
#endif
        path_smem_bind(S, mine, PCAP, cap0, 1);
      }
      if (!counter) __syncthreads();  // (static stride: no barrier at the top of the next round before s_state is rewritten)
    }
    if (flags & 1) {
      // the frame's points are dead: drop the buffer's lines from L2 now, before the cache gets round to writing the
      // dirty data back to HBM (the next frame re-allocates the lines by writing them)
      PG::sync();
      discard_points(mine, aligned);
    }
  }
  PG::sync();
  discard_points(mine, aligned);
}

__global__ void __launch_bounds__(32) initial_path_kernel(DevParams P, double *out) {
  // one lane group (launched with PG::N threads), once per device: the point buffers sit in shared memory behind the
  // PathSmem image
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PathSmem &S = *reinterpret_cast<PathSmem *>(smem_raw);
  path_smem_bind(S, smem_raw + ((sizeof(PathSmem) + 15) / 16) * 16, PCAP, NCAP);
  initial_path_frame(S, P, out);
}

// ---- skidpad mission -----------------------------------------------------------------------

__global__ void __launch_bounds__(32) skid_reloc_kernel(int n_traj, const double *cones_xy, const int32_t *offsets,
                                                        const double *pos, const double *orig_pos,
                                                        const double *orig_dir, const double *jitter,
                                                        const double *ref, double *reloc, int32_t *n_accepted) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SkidSmem &S = *reinterpret_cast<SkidSmem *>(smem_raw);
  for (int t = blockIdx.x; t < n_traj; t += gridDim.x) {
    const int lo = offsets[t], n = offsets[t + 1] - lo;
    SkidReloc *R = reinterpret_cast<SkidReloc *>(reloc + 8 * (size_t)t);
    int nacc = 0;
    skidpad_relocalize(S, cones_xy + 2 * (size_t)lo, n, pos[2 * t], pos[2 * t + 1], orig_pos[2 * t], orig_pos[2 * t + 1],
                       orig_dir[2 * t], orig_dir[2 * t + 1], jitter, ref, R, &nacc);
    if (fsd_lane() == 0 && n_accepted) n_accepted[t] = nacc;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32) skid_track_kernel(int n_traj, const int32_t *step_offsets, const double *pos,
                                                        const double *dir, const double *reloc, int32_t *index_state,
                                                        const double *table, int n_table, double *known,
                                                        int32_t *index, int32_t *traj_of_step) {
  for (int t = blockIdx.x; t < n_traj; t += gridDim.x) {
    const int s0 = step_offsets[t], n = step_offsets[t + 1] - s0;
    const SkidReloc R = *reinterpret_cast<const SkidReloc *>(reloc + 8 * (size_t)t);
    int state = index_state[t];
    skidpad_track(R, table, n_table, pos + 2 * (size_t)s0, dir + 2 * (size_t)s0, n, &state, known + 4 * (size_t)s0,
                  index + s0);
    for (int s = fsd_lane(); s < n; s += 32) traj_of_step[s0 + s] = t;
    if (fsd_lane() == 0) index_state[t] = state;
    __syncwarp();
  }
}

// one lane group per pose step (two steps per warp), all steps of all trajectories in parallel, free-running
__global__ void __launch_bounds__(PATH_THREADS, FSD_PATH_CTAS_PER_SM)
    skid_step_kernel(DevParams P, int n_steps, const double *reloc, const int32_t *traj_of_step, const double *table,
                     int n_table, const int32_t *index, const double *known, const int16_t *force_P,
                     const double *prev, int prev_stride, double *out_f64, double *out_internal, float *out_f32,
                     int16_t *grid_out, uint32_t *status, unsigned char *scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int grp = (int)threadIdx.x / PG::N;
  PathSmem &S = *reinterpret_cast<PathSmem *>(smem_raw + (size_t)grp * PATH_CTA_STRIDE);
  path_smem_bind(S, scratch + ((size_t)blockIdx.x * PATH_FPC + grp) * PATH_SCRATCH_BYTES, PCAP, NCAP);
  for (int s = (int)blockIdx.x * PATH_FPC + grp; s < n_steps; s += (int)gridDim.x * PATH_FPC) {
    const SkidReloc R = *reinterpret_cast<const SkidReloc *>(reloc + 8 * (size_t)traj_of_step[s]);
    int grid[2] = {0, 0};
    double *out = out_f64 + (size_t)s * FSD_HORIZON * 4;
    unsigned st = skidpad_step(S, R, table, n_table, index[s], known + 4 * (size_t)s, force_P ? (int)force_P[s] : 0,
                               prev + (size_t)s * prev_stride, P, out, out_internal + (size_t)s * FSD_HORIZON * 4, grid);
    if (out_f32)
      for (int i = PG::lane(); i < FSD_HORIZON * 4; i += PG::N) out_f32[(size_t)s * FSD_HORIZON * 4 + i] = (float)out[i];
    if (PG::lane() == 0) {
      status[s] = st;
      if (grid_out) {
        grid_out[2 * (size_t)s] = (int16_t)grid[0];
        grid_out[2 * (size_t)s + 1] = (int16_t)grid[1];
      }
    }
    PG::sync();
  }
}

// Sequential semantics inside a trajectory: a step whose previous-path fallback fired (path too far from the car,
// failed tail, ...) must see the PREVIOUS STEP's path (core_calculate_path.py:573), which the step-parallel launch
// cannot know.  One lane group per trajectory (CTAs of PG::N threads) walks its steps in order and recomputes only the
// flagged ones.
__global__ void __launch_bounds__(32) skid_fixup_kernel(DevParams P, int n_traj, const int32_t *step_offsets,
                                                        const double *reloc, const double *table, int n_table,
                                                        const int32_t *index, const double *known,
                                                        const int16_t *force_P, double *out_f64, double *out_internal,
                                                        float *out_f32, int16_t *grid_out, uint32_t *status,
                                                        unsigned char *scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PathSmem &S = *reinterpret_cast<PathSmem *>(smem_raw);
  path_smem_bind(S, scratch + (size_t)blockIdx.x * PATH_SCRATCH_BYTES, PCAP, NCAP);
  const unsigned uses_prev = FSD_ST_FEW_CONES | FSD_ST_FEW_MATCHES | FSD_ST_FIT1_FAILED | FSD_ST_PATH_TOO_FAR |
                             FSD_ST_MPC_FAILED | FSD_ST_REF_RAISES | FSD_ST_UNSUPPORTED;
  for (int t = blockIdx.x; t < n_traj; t += gridDim.x) {
    const SkidReloc R = *reinterpret_cast<const SkidReloc *>(reloc + 8 * (size_t)t);
    for (int s = step_offsets[t] + 1; s < step_offsets[t + 1]; ++s) {
      if (!(status[s] & uses_prev)) continue;
      int grid[2] = {0, 0};
      double *out = out_f64 + (size_t)s * FSD_HORIZON * 4;
      unsigned st = skidpad_step(S, R, table, n_table, index[s], known + 4 * (size_t)s, force_P ? (int)force_P[s] : 0,
                                 out_internal + (size_t)(s - 1) * FSD_HORIZON * 4, P, out,
                                 out_internal + (size_t)s * FSD_HORIZON * 4, grid);
      if (out_f32)
        for (int i = PG::lane(); i < FSD_HORIZON * 4; i += PG::N) out_f32[(size_t)s * FSD_HORIZON * 4 + i] = (float)out[i];
      if (PG::lane() == 0) {
        status[s] = st;
        if (grid_out) {
          grid_out[2 * (size_t)s] = (int16_t)grid[0];
          grid_out[2 * (size_t)s + 1] = (int16_t)grid[1];
        }
      }
      PG::sync();
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------

constexpr int MAX_DEVICES = 64;
constexpr size_t INITIAL_SMEM = ((sizeof(PathSmem) + 15) / 16) * 16 + PATH_SCRATCH_BYTES;
__device__ double g_initial_path[FSD_HORIZON * 4];  // cache of the default-parameter initial path

// a side stream with its fork / join events: the second chunk of a large batch runs on it (plan_batch_impl)
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool busy = false;
};
constexpr int SIDE_STREAMS = 8;

struct DeviceInfo {
  int sm_count = 0;
  int sort_ctas = 0, path_ctas = 0;  // resident CTAs per SM
  bool initial_ready = false;
  double key[3] = {0, 0, 0};
  SideStream side[SIDE_STREAMS];
  int *counter_ring = nullptr;  // COUNTER_SLOTS x COUNTER_STRIDE ints of device memory: frame counters of the free-running kernels
  unsigned counter_next = 0;
};
constexpr int COUNTER_SLOTS = 128;
constexpr int COUNTER_STRIDE = 8 + 64;  // eight counters + the histogram of the path keys (ORDER_BINS), zeroed by ONE memset
static_assert(COUNTER_STRIDE == 8 + ORDER_BINS, "the histogram of the path keys lives behind the eight counters");
DeviceInfo g_dev[MAX_DEVICES];
std::mutex g_mutex;

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename K>
void set_smem(K kernel, size_t bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// How a batch is planned (FSD_PLAN_MODE in the environment, read once; a measurement knob, not an API):
//   (bit 0, the sort stage as two free-running kernels instead of one lockstep kernel, is always on since r2_h)
//   (bit 1, the path stage as three free-running kernels cut at the fit boundaries, lost its A/B by 8 % and was removed)
//   bit 2: the lockstep path kernel takes its rounds of frames from a counter instead of a static stride
//   bit 3: fsd_plan_batch never splits a batch into two chunks on two streams
//   bit 4: the path kernel discards a frame's point-buffer lines from L2 when the frame ends
//   (bit 5, the point buffers as a persisting L2 access-policy window, was measured and removed: 4x MORE write-back)
//   bit 7: NO work-ordered scheduling of the path stage (frames in batch order)
//   bit 6: fits that outgrow their arena are truncated and flagged (kernels_big.cu plans the frame again afterwards) instead
//          of being suspended and resumed inside the path kernel with the CTA's whole shared memory
int plan_mode() {
  static const int mode = [] {
    const char *e = std::getenv("FSD_PLAN_MODE");
    return e ? std::atoi(e) : FSD_DEFAULT_PLAN_MODE;
  }();
  return mode;
}

// Knot records a frame's fits start with in path_kernel: NCAP.  FSD_TEST_CAP (8 .. NCAP, read once) lowers it so that
// ordinary frames outgrow their arena and exercise the suspend / resume path in the GPU tests; results do not change.
int start_cap() {
  static const int cap = [] {
    const char *e = std::getenv("FSD_TEST_CAP");
    const int v = e ? std::atoi(e) : NCAP;
    return v >= 8 && v <= NCAP ? v : NCAP;
  }();
  return cap;
}

int device_info(DeviceInfo **out) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) {
    cudaGetLastError();
    return FSD_ERR_NO_DEVICE;
  }
  DeviceInfo &D = g_dev[dev];
  std::lock_guard<std::mutex> lock(g_mutex);
  if (D.sm_count == 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      cudaGetLastError();
      return FSD_ERR_NO_DEVICE;
    }
    int a = 0, b = 0;
    set_smem(sort_kernel<float>, WPC * SORT_CTA_STRIDE);
    set_smem(sort_kernel<double>, WPC * SORT_CTA_STRIDE);
    set_smem(knn_kernel<float>, WPC * SORT_CTA_STRIDE);
    set_smem(knn_kernel<double>, WPC * SORT_CTA_STRIDE);
    set_smem(match_kernel<float>, WPC * MATCH_CTA_STRIDE);
    set_smem(match_kernel<double>, WPC * MATCH_CTA_STRIDE);
    cudaFuncSetAttribute(path_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PATH_KERNEL_SMEM);
    cudaFuncSetAttribute(path_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PATH_KERNEL_SMEM);
    cudaFuncSetAttribute(skid_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PATH_KERNEL_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sort_kernel<float>, CTA_THREADS, WPC * SORT_CTA_STRIDE);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, path_kernel<float>, PATH_THREADS, PATH_KERNEL_SMEM);
    cudaFuncSetAttribute(initial_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INITIAL_SMEM);
    D.sort_ctas = a > 0 ? a : 1;
    D.path_ctas = b > 0 ? b : 1;
    // The constant initial path of a fresh planner with the default spline parameters, computed ONCE per device here --
    // on a private stream with its own synchronisation -- so that no later call has to synchronise the caller's stream
    // (the batch entry points stay asynchronous and capturable; the first call on a device must not be made under
    // stream capture, see fsdplan.h).
    {
      fsd_params dp;
      fsd_params_default(&dp);
      double *cached = nullptr;
      cudaStream_t st = nullptr;
      bool ok = cudaGetSymbolAddress(reinterpret_cast<void **>(&cached), g_initial_path) == cudaSuccess &&
                cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
      if (ok) {
        initial_path_kernel<<<1, PG::N, INITIAL_SMEM, st>>>(make_dev_params(dp), cached);
        ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
      }
      if (st) cudaStreamDestroy(st);
      if (!ok) {
        cudaGetLastError();
        return FSD_ERR_LAUNCH;
      }
      D.key[0] = dp.smoothing;
      D.key[1] = dp.predict_every;
      D.key[2] = dp.refit_smoothing;
      D.initial_ready = true;
    }
    if (cudaMalloc(reinterpret_cast<void **>(&D.counter_ring), COUNTER_SLOTS * COUNTER_STRIDE * sizeof(int)) != cudaSuccess) {
      cudaGetLastError();
      D.counter_ring = nullptr;  // the free-running kernels are not used without it
    }
    D.sm_count = prop.multiProcessorCount;
  }
  *out = &D;
  return FSD_OK;
}

// Eight zeroed frame counters (+ a zeroed histogram of ORDER_BINS ints behind them) for one sequence of free-running kernels on `stream` (a slot of the per-device ring; a slot
// comes round again after COUNTER_SLOTS sequences, far more than can be in flight).  nullptr: not available.
int *take_counters(DeviceInfo &D, cudaStream_t stream) {
  if (!D.counter_ring) return nullptr;
  unsigned slot;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    slot = D.counter_next++ % COUNTER_SLOTS;
  }
  int *p = D.counter_ring + COUNTER_STRIDE * slot;
  if (cudaMemsetAsync(p, 0, COUNTER_STRIDE * sizeof(int), stream) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

int grid_for(int n_frames, int sm_count, int ctas_per_sm, int frames_per_cta = WPC) {
  long cap = (long)sm_count * ctas_per_sm;
  long need = ((long)n_frames + frames_per_cta - 1) / frames_per_cta;
  return (int)(need < cap ? need : cap);
}

struct Carve {
  unsigned char *base;
  size_t used, cap;
  template <typename T>
  T *take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    T *p = reinterpret_cast<T *>(base + used);
    used += bytes;
    return p;
  }
};

// upper bound of resident path CTAs (each owns PATH_SCRATCH_BYTES of point buffers in the workspace)
size_t path_grid_bound(int n_frames) {
  size_t cap = 160 * 32;  // no device visible: any current part
  DeviceInfo *D = nullptr;
  if (device_info(&D) == FSD_OK) cap = (size_t)D->sm_count * D->path_ctas * PATH_FPC;
  const size_t B = ((size_t)(n_frames > 0 ? n_frames : 0) + PATH_FPC - 1) / PATH_FPC * PATH_FPC;
  return B < cap ? B : cap;
}

// A batch of at least two full waves of resident warps is planned as two chunks on two streams (the caller's and a side
// stream): the ragged last wave of one kernel then overlaps the first wave of the next launch instead of leaving SMs
// idle (persistent CTAs stride statically over their chunk).  Returns the size of the first chunk (n_frames: no split).
int first_chunk(int n_frames) {
  DeviceInfo *D = nullptr;
  if (device_info(&D) != FSD_OK || (plan_mode() & 8)) return n_frames;
  const long wave = (long)D->sm_count * D->path_ctas * PATH_FPC;
  if ((long)n_frames < 2 * wave) return n_frames;
  return (n_frames / 2 + PATH_FPC - 1) / PATH_FPC * PATH_FPC;
}

// work-ordered scheduling of the path stage: worth it from two full waves of resident frames on; FSD_PLAN_MODE bit 7 = off
bool order_wanted(int n_frames, const DeviceInfo &D) {
  return !(plan_mode() & 128) && (long)n_frames >= 2L * D.sm_count * D.path_ctas * PATH_FPC;
}

size_t path_scratch_bytes(int n_frames) {
  const int a = first_chunk(n_frames);
  return align_up(path_grid_bound(a) * PATH_SCRATCH_BYTES, 256) + align_up(path_grid_bound(n_frames - a) * PATH_SCRATCH_BYTES, 256);
}

// memory of the work-ordered scheduling of n frames: order [n] int32, key [n] uint8, histogram [ORDER_BINS] int32
// + the bins' frame lists [ORDER_BINS][n] int32 (path_key_kernel appends, path_kernel looks its rounds up in them)
size_t order_ws_bytes(int n_frames) {
  const size_t n = (size_t)(n_frames > 0 ? n_frames : 0);
  return align_up(n * sizeof(int), 256) + align_up(n, 256) + align_up(ORDER_BINS * sizeof(int), 256) +
         align_up((size_t)ORDER_BINS * n * sizeof(int), 256);
}

struct OrderWs {
  int *ord;       // [n] order of the frames (path_order_kernel; stage entry points only)
  uint8_t *key;   // [n]
  int *hist;      // [ORDER_BINS]
  int *items;     // [ORDER_BINS][n]
  size_t head_bytes;  // ord .. hist
};

OrderWs order_ws_view(unsigned char *ws, int n_frames) {
  const size_t n = (size_t)n_frames;
  OrderWs v;
  v.ord = reinterpret_cast<int *>(ws);
  v.key = ws + align_up(n * sizeof(int), 256);
  v.hist = reinterpret_cast<int *>(ws + align_up(n * sizeof(int), 256) + align_up(n, 256));
  v.head_bytes = align_up(n * sizeof(int), 256) + align_up(n, 256) + align_up(ORDER_BINS * sizeof(int), 256);
  v.items = reinterpret_cast<int *>(ws + v.head_bytes);
  return v;
}

size_t workspace_bytes(int n_frames) {
  const size_t B = (size_t)(n_frames > 0 ? n_frames : 0);
  size_t total = align_up(path_scratch_bytes(n_frames), 256);
  total += align_up(B * 2 * sizeof(int16_t), 256);                      // n_wv
  total += 2 * align_up(B * FSD_MAX_WV * 2 * sizeof(double), 256);      // left_wv, right_wv
  total += 2 * align_up(B * FSD_MAX_WV * sizeof(int16_t), 256);         // l2r, r2l
  total += align_up(B * 2 * sizeof(int16_t), 256);                      // grid
  total += align_up(FSD_HORIZON * 4 * sizeof(double), 256);             // initial path (non-default params)
  total += align_up(B * 2 * FSD_MAX_SORTED * sizeof(int16_t), 256);     // sort indices when the caller wants none
  total += 2 * align_up(fsd_big_path_fixup_scratch_bytes(), 256);       // point buffers of the large-bounds second chance
  total += 2 * align_up(order_ws_bytes(n_frames), 256);                 // work-ordered scheduling (one block per chunk)
  return total;
}

// resolve every intermediate tensor to user memory or workspace
struct Extra {
  int16_t *idx = nullptr;
  unsigned char *fixup[2] = {nullptr, nullptr};
  unsigned char *order[2] = {nullptr, nullptr};
};

int resolve(const fsd_intermediate *inter, int n_frames, void *workspace, size_t workspace_bytes_given,
            fsd_intermediate *out, double **initial_slot, unsigned char **path_scratch, Extra *extra = nullptr) {
  fsd_intermediate r;
  std::memset(&r, 0, sizeof(r));
  if (inter) r = *inter;
  if (!workspace || workspace_bytes_given < workspace_bytes(n_frames)) return FSD_ERR_WORKSPACE;
  Carve cv = {static_cast<unsigned char *>(workspace), 0, workspace_bytes_given};
  const size_t B = (size_t)n_frames;
  *path_scratch = cv.take<unsigned char>(path_scratch_bytes(n_frames));
  int16_t *w_nwv = cv.take<int16_t>(B * 2);
  double *w_lwv = cv.take<double>(B * FSD_MAX_WV * 2);
  double *w_rwv = cv.take<double>(B * FSD_MAX_WV * 2);
  int16_t *w_l2r = cv.take<int16_t>(B * FSD_MAX_WV);
  int16_t *w_r2l = cv.take<int16_t>(B * FSD_MAX_WV);
  int16_t *w_grid = cv.take<int16_t>(B * 2);
  double *w_init = cv.take<double>(FSD_HORIZON * 4);
  int16_t *w_idx = cv.take<int16_t>(B * 2 * FSD_MAX_SORTED);
  unsigned char *w_fix0 = cv.take<unsigned char>(fsd_big_path_fixup_scratch_bytes());
  unsigned char *w_fix1 = cv.take<unsigned char>(fsd_big_path_fixup_scratch_bytes());
  unsigned char *w_ord0 = cv.take<unsigned char>(order_ws_bytes(n_frames));
  unsigned char *w_ord1 = cv.take<unsigned char>(order_ws_bytes(n_frames));
  if (extra) {
    extra->idx = w_idx;
    extra->fixup[0] = w_fix0;
    extra->fixup[1] = w_fix1;
    extra->order[0] = w_ord0;
    extra->order[1] = w_ord1;
  }
  if (!r.n_wv) r.n_wv = w_nwv;
  if (!r.left_wv) r.left_wv = w_lwv;
  if (!r.right_wv) r.right_wv = w_rwv;
  if (!r.l2r) r.l2r = w_l2r;
  if (!r.r2l) r.r2l = w_r2l;
  if (!r.grid) r.grid = w_grid;
  if (initial_slot) *initial_slot = w_init;
  *out = r;
  return FSD_OK;
}

int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? FSD_OK : FSD_ERR_LAUNCH;
}

// the previous path used when the caller passes none: the initial path of a fresh planner.  Default spline parameters:
// the per-device constant computed in device_info; otherwise computed into `scratch` on the caller's stream
// (asynchronous, no synchronisation).
int default_prev_path(const fsd_params *params, const DevParams &P, DeviceInfo &D, double *scratch,
                      cudaStream_t stream, const double **prev) {
  const double key[3] = {params->smoothing, params->predict_every, params->refit_smoothing};
  if (D.initial_ready && std::memcmp(D.key, key, sizeof(key)) == 0) {
    double *cached = nullptr;
    if (cudaGetSymbolAddress(reinterpret_cast<void **>(&cached), g_initial_path) != cudaSuccess) {
      cudaGetLastError();
      return FSD_ERR_LAUNCH;
    }
    *prev = cached;
    return FSD_OK;
  }
  if (!scratch) return FSD_ERR_ARG;  // non-default spline parameters: the caller must pass prev_path
  initial_path_kernel<<<1, PG::N, INITIAL_SMEM, stream>>>(P, scratch);
  *prev = scratch;
  return check_launch();
}

// idx_scratch: room for the sort indices (2 x n_frames x 12) when the caller passes no output tensors for them
template <typename T>
int sort_match_impl(const fsd_params *params, int n_frames, const T *cones_xy, const uint8_t *cones_type,
                    const int32_t *offsets, const T *pos, const T *dir, int16_t *out_left_idx, int16_t *out_right_idx,
                    const fsd_intermediate *inter, uint32_t *out_status, cudaStream_t stream,
                    int16_t *idx_scratch = nullptr) {
  if (!params || n_frames < 0 || !offsets || !pos || !dir || !out_status || !inter) return FSD_ERR_ARG;
  if (!inter->n_wv || !inter->left_wv || !inter->right_wv || !inter->l2r || !inter->r2l) return FSD_ERR_ARG;
  if (n_frames == 0) return FSD_OK;
  if (!cones_xy || !cones_type) return FSD_ERR_ARG;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  StageOut O = {out_left_idx, out_right_idx, inter->sort_dbg, inter->n_wv, inter->left_wv, inter->right_wv,
                inter->l2r,   inter->r2l,    out_status};
  // two free-running kernels: sorting (writes the sort indices), then matching on them
  if (!(out_left_idx || idx_scratch) || !(out_right_idx || idx_scratch)) return FSD_ERR_ARG;
  int *counters = take_counters(*D, stream);
  if (!counters) return FSD_ERR_LAUNCH;
  if (!O.left_idx) O.left_idx = idx_scratch;
  if (!O.right_idx) O.right_idx = idx_scratch + (size_t)n_frames * FSD_MAX_SORTED;
  const DevParams P = make_dev_params(*params);
  sort_kernel<T><<<grid_for(n_frames, D->sm_count, D->sort_ctas), CTA_THREADS, WPC * SORT_CTA_STRIDE, stream>>>(
      P, n_frames, cones_xy, cones_type, offsets, pos, dir, O, counters);
  rc = check_launch();
  if (rc != FSD_OK) return rc;
  match_kernel<T><<<grid_for(n_frames, D->sm_count, CTAS_PER_SM), CTA_THREADS, WPC * MATCH_CTA_STRIDE, stream>>>(
      P, n_frames, cones_xy, offsets, pos, dir, O.left_idx, O.right_idx, O, 1, counters + 1);
  return check_launch();
}

template <typename T>
int path_impl(const fsd_params *params, int n_frames, const T *pos, const T *dir, const fsd_intermediate *inter,
              const int16_t *force_P, const double *prev_path, int prev_path_stride, double *init_scratch,
              unsigned char *path_scratch, float *out_path, uint32_t *out_status, cudaStream_t stream,
              unsigned char *fixup_scratch = nullptr, const fsd_gather *gather = nullptr,
              unsigned char *order_ws = nullptr) {
  if (!path_scratch) return FSD_ERR_WORKSPACE;
  if (!params || n_frames < 0 || !pos || !dir || !out_status || !inter) return FSD_ERR_ARG;
  if (!inter->n_wv || !inter->left_wv || !inter->right_wv || !inter->l2r || !inter->r2l) return FSD_ERR_ARG;
  fsd_gather G;
  std::memset(&G, 0, sizeof(G));
  if (gather) {
    if (gather->n_peers < 0 || gather->n_peers > FSD_MAX_PEERS || gather->first_row < 0) return FSD_ERR_ARG;
    for (int r = 0; r < gather->n_peers; ++r)
      if (!gather->peer_out_path[r]) return FSD_ERR_ARG;
    G = *gather;
  }
  const bool peers = G.n_peers > 0 || G.multicast_out_path;
  if (!inter->path_f64 && !out_path && !peers) return FSD_ERR_ARG;  // at least one path output
  if (prev_path && prev_path_stride != 0 && prev_path_stride != FSD_HORIZON * 4) return FSD_ERR_ARG;
  if (n_frames == 0) return FSD_OK;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  const DevParams P = make_dev_params(*params);
  const double *prev = prev_path;
  int stride = prev_path_stride;
  if (!prev) {
    rc = default_prev_path(params, P, *D, init_scratch, stream, &prev);
    if (rc != FSD_OK) return rc;
    stride = 0;
  }
  StageOut O = {nullptr,    nullptr,    nullptr,   inter->n_wv, inter->left_wv, inter->right_wv,
                inter->l2r, inter->r2l, out_status};
  const int grid = grid_for(n_frames, D->sm_count, D->path_ctas, PATH_FPC);
  // work-ordered scheduling: path_key_kernel files every frame in the list of its bin, the path kernel takes its rounds from
  // the lists, heaviest bin first (FSD_PLAN_MODE bit 8: key array + counting sort into an order array, as before r2_zb).
  // The histogram lives behind the round counter in the counter ring: one memset for both.
  const int *order = nullptr, *bin_hist = nullptr, *bin_items = nullptr;
  const bool ordered = order_ws && order_wanted(n_frames, *D);
  int *slot = ((plan_mode() & 4) || ordered) ? take_counters(*D, stream) : nullptr;
  if (ordered && slot) {
    const OrderWs W = order_ws_view(order_ws, n_frames);
    const bool lists = !(plan_mode() & 256);
    int *hist = slot + 8;
    path_key_kernel<<<(n_frames + 7) / 8, 256, 0, stream>>>(n_frames, O, lists ? nullptr : W.key, hist, W.items);
    if (lists) {
      bin_hist = hist;
      bin_items = W.items;
    } else {
      path_order_kernel<<<1, 1024, 0, stream>>>(n_frames, W.key, hist, W.ord);
      order = W.ord;
    }
    rc = check_launch();
    if (rc != FSD_OK) return rc;
  }
  int *round_counter = (plan_mode() & 4) ? slot : nullptr;
  const int flags = ((plan_mode() & 16) ? 1 : 0) | (fixup_scratch ? 2 : 0) | (peers ? 4 : 0) |
                    (((plan_mode() & 64) || PATH_FPW != 1) ? 0 : 8);
  path_kernel<T><<<grid, PATH_THREADS, PATH_KERNEL_SMEM, stream>>>(
      P, n_frames, pos, dir, O, force_P, prev, stride, inter->path_f64, out_path, inter->grid, path_scratch,
      round_counter, flags, start_cap(), order, bin_hist, bin_items, G);
  rc = check_launch();
  if (rc != FSD_OK || !fixup_scratch) return rc;
  // frames on which a static bound of path_kernel overflowed get a second chance with the large bounds (kernels_big.cu)
  return fsd_big_path_fixup(params, n_frames, sizeof(T) == 8, pos, dir, inter->n_wv, inter->left_wv, inter->right_wv,
                            inter->l2r, inter->r2l, force_P, prev, stride, inter->path_f64, out_path, inter->grid,
                            out_status, fixup_scratch, stream, peers ? &G : nullptr);
}

// take / return a side stream of the current device (created on first use)
SideStream *acquire_side(DeviceInfo &D) {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (SideStream &s : D.side) {
    if (s.busy) continue;
    if (!s.stream) {
      if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
    }
    s.busy = true;
    return &s;
  }
  return nullptr;
}

void release_side(SideStream *s) {
  std::lock_guard<std::mutex> lock(g_mutex);
  s->busy = false;
}

fsd_intermediate shift_intermediate(const fsd_intermediate &R, size_t h) {
  fsd_intermediate r = R;
  if (r.path_f64) r.path_f64 += h * FSD_HORIZON * 4;
  if (r.n_wv) r.n_wv += h * 2;
  if (r.left_wv) r.left_wv += h * FSD_MAX_WV * 2;
  if (r.right_wv) r.right_wv += h * FSD_MAX_WV * 2;
  if (r.l2r) r.l2r += h * FSD_MAX_WV;
  if (r.r2l) r.r2l += h * FSD_MAX_WV;
  if (r.grid) r.grid += h * 2;
  if (r.sort_dbg) r.sort_dbg += h * 8;
  return r;
}

template <typename T>
int plan_batch_impl(const fsd_params *params, int mission, int n_frames, const T *cones_xy, const uint8_t *cones_type,
                    const int32_t *offsets, const T *pos, const T *dir, float *out_path, int16_t *out_left_idx,
                    int16_t *out_right_idx, const fsd_intermediate *inter, const int16_t *force_P,
                    const double *prev_path, int prev_path_stride, uint32_t *out_status, void *workspace,
                    size_t workspace_bytes_given, void *stream_v, void *chunk_ready_v = nullptr,
                    const fsd_gather *gather = nullptr) {
  if (!params || n_frames < 0) return FSD_ERR_ARG;
  if (mission != FSD_MISSION_AUTOCROSS && mission != FSD_MISSION_TRACKDRIVE) return FSD_ERR_MISSION;
  if (n_frames == 0) return FSD_OK;
  if (!offsets || !pos || !dir || !out_status) return FSD_ERR_ARG;
  if (prev_path && prev_path_stride != 0 && prev_path_stride != FSD_HORIZON * 4) return FSD_ERR_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  fsd_intermediate R;
  double *init_slot = nullptr;
  unsigned char *path_scratch = nullptr;
  Extra X;
  int rc = resolve(inter, n_frames, workspace, workspace_bytes_given, &R, &init_slot, &path_scratch, &X);
  if (rc != FSD_OK) return rc;
  DeviceInfo *D = nullptr;
  rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  // the previous path both chunks fall back to (resolved once, on the caller's stream, before the fork)
  const double *prev = prev_path;
  int stride = prev_path_stride;
  if (!prev) {
    rc = default_prev_path(params, make_dev_params(*params), *D, init_slot, stream, &prev);
    if (rc != FSD_OK) return rc;
    stride = 0;
  }
  const int na = first_chunk(n_frames), nb = n_frames - na;
  SideStream *side = nb > 0 ? acquire_side(*D) : nullptr;
  if (!side) {
    rc = sort_match_impl<T>(params, n_frames, cones_xy, cones_type, offsets, pos, dir, out_left_idx, out_right_idx, &R,
                            out_status, stream, X.idx);
    if (rc != FSD_OK) return rc;
    rc = path_impl<T>(params, n_frames, pos, dir, &R, force_P, prev, stride, init_slot, path_scratch, out_path,
                      out_status, stream, X.fixup[0], gather, X.order[0]);
    if (rc == FSD_OK && chunk_ready_v && cudaEventRecord(static_cast<cudaEvent_t>(chunk_ready_v), stream) != cudaSuccess) {
      cudaGetLastError();
      rc = FSD_ERR_LAUNCH;
    }
    return rc;
  }
  // chunk A = frames [0, na) on the caller's stream, chunk B = [na, n_frames) on the side stream; the CSR offsets are
  // absolute, so chunk B simply starts further into the same arrays
  const size_t h = (size_t)na;
  const fsd_intermediate RB = shift_intermediate(R, h);
  unsigned char *scratch_b = path_scratch + align_up(path_grid_bound(na) * PATH_SCRATCH_BYTES, 256);
  bool ok = cudaEventRecord(side->fork, stream) == cudaSuccess &&
            cudaStreamWaitEvent(side->stream, side->fork, 0) == cudaSuccess;
  if (ok) {
    // (the sort-index scratch of chunk A is [0, 2 na x 12), chunk B's follows it)
    rc = sort_match_impl<T>(params, na, cones_xy, cones_type, offsets, pos, dir, out_left_idx, out_right_idx, &R,
                            out_status, stream, X.idx);
    if (rc == FSD_OK)
      rc = sort_match_impl<T>(params, nb, cones_xy, cones_type, offsets + h, pos + 2 * h, dir + 2 * h,
                              out_left_idx ? out_left_idx + h * FSD_MAX_SORTED : nullptr,
                              out_right_idx ? out_right_idx + h * FSD_MAX_SORTED : nullptr, &RB, out_status + h,
                              side->stream, X.idx + 2 * h * FSD_MAX_SORTED);
    if (rc == FSD_OK)
      rc = path_impl<T>(params, na, pos, dir, &R, force_P, prev, stride, init_slot, path_scratch, out_path, out_status,
                        stream, X.fixup[0], gather, X.order[0]);
    // the outputs of frames [0, na) are final here: a caller that passed an event can start consuming them (e.g. an
    // all-gather on a communication stream) while chunk B is still being planned
    if (rc == FSD_OK && chunk_ready_v && cudaEventRecord(static_cast<cudaEvent_t>(chunk_ready_v), stream) != cudaSuccess) {
      cudaGetLastError();
      rc = FSD_ERR_LAUNCH;
    }
    if (rc == FSD_OK) {
      fsd_gather GB;  // chunk B's rows follow chunk A's in the gathered buffers
      if (gather) {
        GB = *gather;
        GB.first_row += (int64_t)h;
      }
      rc = path_impl<T>(params, nb, pos + 2 * h, dir + 2 * h, &RB, force_P ? force_P + h : nullptr,
                        prev + h * (size_t)stride, stride, init_slot, scratch_b,
                        out_path ? out_path + h * FSD_HORIZON * 4 : nullptr, out_status + h, side->stream, X.fixup[1],
                        gather ? &GB : nullptr, X.order[1]);
    }
  }
  // always join, so that the caller's stream never runs ahead of work queued on the side stream
  ok = cudaEventRecord(side->join, side->stream) == cudaSuccess && ok;
  ok = cudaStreamWaitEvent(stream, side->join, 0) == cudaSuccess && ok;
  release_side(side);
  if (!ok) {
    cudaGetLastError();
    return FSD_ERR_LAUNCH;
  }
  return rc;
}

}  // namespace

// ---- C-ABI ------------------------------------------------------------------------------------------------

extern "C" {

int fsd_abi_version(void) { return FSD_ABI_VERSION; }

const char *fsd_strerror(int code) {
  switch (code) {
    case FSD_OK: return "ok";
    case FSD_ERR_ARG: return "invalid argument";
    case FSD_ERR_WORKSPACE: return "workspace missing or smaller than fsd_workspace_bytes()";
    case FSD_ERR_LAUNCH: return "CUDA launch failed";
    case FSD_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case FSD_ERR_MISSION: return "mission not handled by the batched planner (trackdrive / autocross only)";
    default: return "unknown error";
  }
}

int fsd_params_default(fsd_params *p) {
  if (!p) return FSD_ERR_ARG;
  std::memset(p, 0, sizeof(*p));
  p->max_n_neighbors = 5;
  p->max_length = 12;
  p->max_dist = 6.5;
  p->max_dist_to_first = 6.0;
  p->threshold_directional_angle = 40.0 * PI / 180.0;
  p->threshold_absolute_angle = 65.0 * PI / 180.0;
  p->car_size = 2.1;
  p->max_dfs_pops = 1 << 16;
  p->min_track_width = 3.0;
  p->max_search_range = 5.0;
  p->max_search_angle = 50.0 * PI / 180.0;
  p->smoothing = 0.2;
  p->predict_every = 0.1;
  p->maximal_distance_for_valid_path = 5.0;
  p->mpc_path_length = 20.0;
  p->refit_smoothing = 0.01;
  return FSD_OK;
}

size_t fsd_workspace_bytes(int n_frames, int total_cones) {
  (void)total_cones;
  return workspace_bytes(n_frames);
}

int fsd_plan_launches(int n_frames) {
  if (n_frames <= 0) return 0;
  // kernels of this library per chunk: sort, match, path, the large-bounds second chance of the path stage, and -- batches
  // large enough for the work-ordered path stage -- the path keys (FSD_PLAN_MODE bit 8: + the counting sort of the keys)
  DeviceInfo *D = nullptr;
  const bool dev = device_info(&D) == FSD_OK;
  auto chunk = [&](int n) { return n <= 0 ? 0 : 4 + (dev && order_wanted(n, *D) ? ((plan_mode() & 256) ? 2 : 1) : 0); };
  const int na = first_chunk(n_frames);
  return chunk(na) + chunk(n_frames - na);
}

int fsd_plan_first_chunk(int n_frames) { return n_frames <= 0 ? 0 : first_chunk(n_frames); }

int fsd_plan_batch_ex(const fsd_params *params, int mission, int n_frames, int coords_f64, const void *cones_xy,
                      const uint8_t *cones_type, const int32_t *offsets, const void *pos, const void *dir,
                      float *out_path, int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                      const int16_t *force_P, const double *prev_path, int prev_path_stride, uint32_t *out_status,
                      void *workspace, size_t workspace_bytes_given, void *stream, void *chunk_ready_event) {
  if (coords_f64)
    return plan_batch_impl<double>(params, mission, n_frames, static_cast<const double *>(cones_xy), cones_type, offsets,
                                   static_cast<const double *>(pos), static_cast<const double *>(dir), out_path,
                                   out_left_idx, out_right_idx, inter, force_P, prev_path, prev_path_stride, out_status,
                                   workspace, workspace_bytes_given, stream, chunk_ready_event);
  return plan_batch_impl<float>(params, mission, n_frames, static_cast<const float *>(cones_xy), cones_type, offsets,
                                static_cast<const float *>(pos), static_cast<const float *>(dir), out_path, out_left_idx,
                                out_right_idx, inter, force_P, prev_path, prev_path_stride, out_status, workspace,
                                workspace_bytes_given, stream, chunk_ready_event);
}

int fsd_plan_batch_gather(const fsd_params *params, int mission, int n_frames, int coords_f64, const void *cones_xy,
                          const uint8_t *cones_type, const int32_t *offsets, const void *pos, const void *dir,
                          float *out_path, int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                          const int16_t *force_P, const double *prev_path, int prev_path_stride, uint32_t *out_status,
                          void *workspace, size_t workspace_bytes_given, void *stream, void *chunk_ready_event,
                          const fsd_gather *gather) {
  if (coords_f64)
    return plan_batch_impl<double>(params, mission, n_frames, static_cast<const double *>(cones_xy), cones_type, offsets,
                                   static_cast<const double *>(pos), static_cast<const double *>(dir), out_path,
                                   out_left_idx, out_right_idx, inter, force_P, prev_path, prev_path_stride, out_status,
                                   workspace, workspace_bytes_given, stream, chunk_ready_event, gather);
  return plan_batch_impl<float>(params, mission, n_frames, static_cast<const float *>(cones_xy), cones_type, offsets,
                                static_cast<const float *>(pos), static_cast<const float *>(dir), out_path, out_left_idx,
                                out_right_idx, inter, force_P, prev_path, prev_path_stride, out_status, workspace,
                                workspace_bytes_given, stream, chunk_ready_event, gather);
}

int fsd_initial_path(const fsd_params *params, double *out_prev_path, void *stream) {
  if (!params || !out_prev_path) return FSD_ERR_ARG;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  initial_path_kernel<<<1, PG::N, INITIAL_SMEM, static_cast<cudaStream_t>(stream)>>>(make_dev_params(*params),
                                                                                      out_prev_path);
  return check_launch();
}

int fsd_plan_batch(const fsd_params *params, int mission, int n_frames, const float *cones_xy,
                   const uint8_t *cones_type, const int32_t *offsets, const float *pos, const float *dir,
                   float *out_path, int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                   const int16_t *force_P, const double *prev_path, int prev_path_stride, uint32_t *out_status,
                   void *workspace, size_t workspace_bytes_given, void *stream) {
  return plan_batch_impl<float>(params, mission, n_frames, cones_xy, cones_type, offsets, pos, dir, out_path,
                                out_left_idx, out_right_idx, inter, force_P, prev_path, prev_path_stride, out_status,
                                workspace, workspace_bytes_given, stream);
}

int fsd_plan_batch_f64(const fsd_params *params, int mission, int n_frames, const double *cones_xy,
                       const uint8_t *cones_type, const int32_t *offsets, const double *pos, const double *dir,
                       float *out_path, int16_t *out_left_idx, int16_t *out_right_idx,
                       const fsd_intermediate *inter, const int16_t *force_P, const double *prev_path,
                       int prev_path_stride, uint32_t *out_status, void *workspace, size_t workspace_bytes_given,
                       void *stream) {
  return plan_batch_impl<double>(params, mission, n_frames, cones_xy, cones_type, offsets, pos, dir, out_path,
                                 out_left_idx, out_right_idx, inter, force_P, prev_path, prev_path_stride, out_status,
                                 workspace, workspace_bytes_given, stream);
}

int fsd_sort_batch(const fsd_params *params, int n_frames, const float *cones_xy, const uint8_t *cones_type,
                   const int32_t *offsets, const float *pos, const float *dir, int16_t *out_left_idx,
                   int16_t *out_right_idx, int16_t *sort_dbg, uint32_t *out_status, void *stream) {
  if (!params || n_frames < 0 || !offsets || !pos || !dir || !out_status || !out_left_idx || !out_right_idx)
    return FSD_ERR_ARG;
  if (n_frames == 0) return FSD_OK;
  if (!cones_xy || !cones_type) return FSD_ERR_ARG;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  StageOut O = {out_left_idx, out_right_idx, sort_dbg, nullptr, nullptr, nullptr, nullptr, nullptr, out_status};
  int *counters = take_counters(*D, static_cast<cudaStream_t>(stream));
  if (!counters) return FSD_ERR_LAUNCH;
  sort_kernel<float><<<grid_for(n_frames, D->sm_count, D->sort_ctas), CTA_THREADS, WPC * SORT_CTA_STRIDE,
                       static_cast<cudaStream_t>(stream)>>>(make_dev_params(*params), n_frames, cones_xy, cones_type,
                                                            offsets, pos, dir, O, counters);
  return check_launch();
}

int fsd_knn_batch(const fsd_params *params, int n_frames, int coords_f64, const void *cones_xy,
                  const uint8_t *cones_type, const int32_t *offsets, uint8_t *out_nbr, uint8_t *out_deg, void *stream) {
  if (!params || n_frames < 0 || !offsets || !out_nbr || !out_deg) return FSD_ERR_ARG;
  if (n_frames == 0) return FSD_OK;
  if (!cones_xy || !cones_type) return FSD_ERR_ARG;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int *counters = take_counters(*D, st);
  if (!counters) return FSD_ERR_LAUNCH;
  const int grid = grid_for(n_frames, D->sm_count, D->sort_ctas);
  const DevParams P = make_dev_params(*params);
  if (coords_f64)
    knn_kernel<double><<<grid, CTA_THREADS, WPC * SORT_CTA_STRIDE, st>>>(
        P, n_frames, static_cast<const double *>(cones_xy), cones_type, offsets, out_nbr, out_deg, counters);
  else
    knn_kernel<float><<<grid, CTA_THREADS, WPC * SORT_CTA_STRIDE, st>>>(
        P, n_frames, static_cast<const float *>(cones_xy), cones_type, offsets, out_nbr, out_deg, counters);
  return check_launch();
}

int fsd_match_batch(const fsd_params *params, int n_frames, const float *cones_xy, const int32_t *offsets,
                    const float *pos, const float *dir, const int16_t *left_idx, const int16_t *right_idx,
                    const fsd_intermediate *inter, uint32_t *out_status, void *stream) {
  if (!params || n_frames < 0 || !offsets || !pos || !dir || !out_status || !left_idx || !right_idx || !inter)
    return FSD_ERR_ARG;
  if (!inter->n_wv || !inter->left_wv || !inter->right_wv || !inter->l2r || !inter->r2l) return FSD_ERR_ARG;
  if (n_frames == 0) return FSD_OK;
  if (!cones_xy) return FSD_ERR_ARG;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  StageOut O = {nullptr,        nullptr,    nullptr,   inter->n_wv, inter->left_wv, inter->right_wv,
                inter->l2r,     inter->r2l, out_status};
  int *counters = take_counters(*D, static_cast<cudaStream_t>(stream));
  if (!counters) return FSD_ERR_LAUNCH;
  match_kernel<float><<<grid_for(n_frames, D->sm_count, CTAS_PER_SM), CTA_THREADS, WPC * MATCH_CTA_STRIDE,
                        static_cast<cudaStream_t>(stream)>>>(make_dev_params(*params), n_frames, cones_xy, offsets, pos,
                                                             dir, left_idx, right_idx, O, 0, counters);
  return check_launch();
}

int fsd_sort_match_batch(const fsd_params *params, int n_frames, int coords_f64, const void *cones_xy,
                         const uint8_t *cones_type, const int32_t *offsets, const void *pos, const void *dir,
                         int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                         uint32_t *out_status, void *stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (coords_f64)
    return sort_match_impl<double>(params, n_frames, static_cast<const double *>(cones_xy), cones_type, offsets,
                                   static_cast<const double *>(pos), static_cast<const double *>(dir), out_left_idx,
                                   out_right_idx, inter, out_status, st);
  return sort_match_impl<float>(params, n_frames, static_cast<const float *>(cones_xy), cones_type, offsets,
                                static_cast<const float *>(pos), static_cast<const float *>(dir), out_left_idx,
                                out_right_idx, inter, out_status, st);
}

int fsd_path_batch_gather(const fsd_params *params, int n_frames, int coords_f64, const void *pos, const void *dir,
                   const fsd_intermediate *inter, const int16_t *force_P, const double *prev_path,
                   int prev_path_stride, float *out_path, uint32_t *out_status, void *workspace,
                   size_t workspace_bytes_given, void *stream, const fsd_gather *gather) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_frames > 0 && (!workspace || workspace_bytes_given < path_grid_bound(n_frames) * PATH_SCRATCH_BYTES))
    return FSD_ERR_WORKSPACE;
  unsigned char *scratch = static_cast<unsigned char *>(workspace);
  // the point buffers of the large-bounds second chance sit behind the kernel's own when the workspace has room for them
  // (a workspace of fsd_workspace_bytes() always has)
  unsigned char *fixup = nullptr;
  const size_t fix_at = align_up(path_grid_bound(n_frames) * PATH_SCRATCH_BYTES, 256);
  if (n_frames > 0 && workspace_bytes_given >= fix_at + fsd_big_path_fixup_scratch_bytes()) fixup = scratch + fix_at;
  // ... and the memory of the work-ordered scheduling behind those
  unsigned char *order_ws = nullptr;
  const size_t ord_at = fix_at + align_up(fsd_big_path_fixup_scratch_bytes(), 256);
  if (fixup && workspace_bytes_given >= ord_at + order_ws_bytes(n_frames)) order_ws = scratch + ord_at;
  if (coords_f64)
    return path_impl<double>(params, n_frames, static_cast<const double *>(pos), static_cast<const double *>(dir), inter,
                             force_P, prev_path, prev_path_stride, nullptr, scratch, out_path, out_status, st, fixup, gather,
                             order_ws);
  return path_impl<float>(params, n_frames, static_cast<const float *>(pos), static_cast<const float *>(dir), inter,
                          force_P, prev_path, prev_path_stride, nullptr, scratch, out_path, out_status, st, fixup, gather,
                          order_ws);
}

int fsd_path_batch(const fsd_params *params, int n_frames, int coords_f64, const void *pos, const void *dir,
                   const fsd_intermediate *inter, const int16_t *force_P, const double *prev_path,
                   int prev_path_stride, float *out_path, uint32_t *out_status, void *workspace,
                   size_t workspace_bytes_given, void *stream) {
  return fsd_path_batch_gather(params, n_frames, coords_f64, pos, dir, inter, force_P, prev_path, prev_path_stride,
                               out_path, out_status, workspace, workspace_bytes_given, stream, nullptr);
}

size_t fsd_global_path_workspace_bytes(int n_poses) {
  DeviceInfo *D = nullptr;
  const int sms = device_info(&D) == FSD_OK ? D->sm_count : 160;
  return fsd_big_global_path_scratch_bytes(n_poses, sms);
}

int fsd_global_path_batch(const fsd_params *params, int n_poses, const double *pos, const double *dir,
                          const double *global_path, int n_points, const int16_t *force_P, const double *prev_path,
                          int prev_path_stride, float *out_path, double *out_path_f64, int16_t *out_grid,
                          uint32_t *out_status, void *workspace, size_t workspace_bytes_given, void *stream_v) {
  if (!params || n_poses < 0) return FSD_ERR_ARG;
  if (n_poses == 0) return FSD_OK;
  if (!pos || !dir || !global_path || n_points < 1 || !out_status || (!out_path && !out_path_f64)) return FSD_ERR_ARG;
  if (prev_path && prev_path_stride != 0 && prev_path_stride != FSD_HORIZON * 4) return FSD_ERR_ARG;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  if (!workspace || workspace_bytes_given < fsd_big_global_path_scratch_bytes(n_poses, D->sm_count)) return FSD_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const double *prev = prev_path;
  int stride = prev_path_stride;
  if (!prev) {
    rc = default_prev_path(params, make_dev_params(*params), *D, nullptr, stream, &prev);
    if (rc != FSD_OK) return rc;
    stride = 0;
  }
  int *counters = take_counters(*D, stream);
  if (!counters) return FSD_ERR_LAUNCH;
  return fsd_big_global_path(params, n_poses, pos, dir, global_path, n_points, force_P, prev, stride, out_path_f64, out_path,
                             out_grid, out_status, static_cast<unsigned char *>(workspace), counters, D->sm_count, stream);
}

size_t fsd_skidpad_workspace_bytes(int n_steps) {
  const size_t S = (size_t)(n_steps > 0 ? n_steps : 0);
  return align_up(path_grid_bound(n_steps) * PATH_SCRATCH_BYTES, 256) + align_up(S * 4 * sizeof(double), 256) +
         align_up(S * sizeof(int32_t), 256);
}

int fsd_skidpad_relocalize_batch(const fsd_params *params, int n_traj, const double *cones_xy, const int32_t *offsets,
                                 const double *pos, const double *orig_pos, const double *orig_dir,
                                 const double *jitter, const double *ref_centers, double *reloc, int32_t *n_accepted,
                                 void *stream) {
  if (!params || n_traj < 0) return FSD_ERR_ARG;
  if (n_traj == 0) return FSD_OK;
  if (!cones_xy || !offsets || !pos || !orig_pos || !orig_dir || !jitter || !ref_centers || !reloc) return FSD_ERR_ARG;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  const int grid = n_traj < D->sm_count * 4 ? n_traj : D->sm_count * 4;
  skid_reloc_kernel<<<grid, 32, sizeof(SkidSmem), static_cast<cudaStream_t>(stream)>>>(
      n_traj, cones_xy, offsets, pos, orig_pos, orig_dir, jitter, ref_centers, reloc, n_accepted);
  return check_launch();
}

int fsd_skidpad_plan_batch(const fsd_params *params, int n_traj, int n_steps, const int32_t *step_offsets,
                           const double *pos, const double *dir, const double *reloc, int32_t *index_state,
                           const double *path_table, int n_table, const int16_t *force_P, const double *prev_path,
                           int prev_path_stride, float *out_path, double *out_path_f64, double *out_internal_f64,
                           int32_t *out_index, int16_t *out_grid, uint32_t *out_status, void *workspace,
                           size_t workspace_bytes_given, void *stream_v) {
  if (!params || n_traj < 0 || n_steps < 0) return FSD_ERR_ARG;
  if (n_traj == 0 || n_steps == 0) return FSD_OK;
  if (!step_offsets || !pos || !dir || !reloc || !index_state || !path_table || n_table < 16 || !out_path_f64 ||
      !out_internal_f64 || !out_index || !out_status)
    return FSD_ERR_ARG;
  if (prev_path && prev_path_stride != 0 && prev_path_stride != FSD_HORIZON * 4) return FSD_ERR_ARG;
  if (!workspace || workspace_bytes_given < fsd_skidpad_workspace_bytes(n_steps)) return FSD_ERR_WORKSPACE;
  DeviceInfo *D = nullptr;
  int rc = device_info(&D);
  if (rc != FSD_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const DevParams P = make_dev_params(*params);
  Carve cv = {static_cast<unsigned char *>(workspace), 0, workspace_bytes_given};
  unsigned char *scratch = cv.take<unsigned char>(path_grid_bound(n_steps) * PATH_SCRATCH_BYTES);
  double *known = cv.take<double>((size_t)n_steps * 4);
  int32_t *traj_of_step = cv.take<int32_t>((size_t)n_steps);
  const double *prev = prev_path;
  int stride = prev_path_stride;
  if (!prev) {
    rc = default_prev_path(params, P, *D, nullptr, stream, &prev);
    if (rc != FSD_OK) return rc;
    stride = 0;
  }
  const int tgrid = n_traj < D->sm_count * 8 ? n_traj : D->sm_count * 8;
  skid_track_kernel<<<tgrid, 32, 0, stream>>>(n_traj, step_offsets, pos, dir, reloc, index_state, path_table, n_table,
                                              known, out_index, traj_of_step);
  rc = check_launch();
  if (rc != FSD_OK) return rc;
  skid_step_kernel<<<grid_for(n_steps, D->sm_count, D->path_ctas, PATH_FPC), PATH_THREADS, PATH_KERNEL_SMEM, stream>>>(
      P, n_steps, reloc, traj_of_step, path_table, n_table, out_index, known, force_P, prev, stride, out_path_f64,
      out_internal_f64, out_path, out_grid, out_status, scratch);
  rc = check_launch();
  if (rc != FSD_OK || stride != 0) return rc;  // per-step previous paths given: the caller owns the chaining
  const long bound = (long)path_grid_bound(n_steps);
  const int fgrid = (int)(n_traj < bound ? n_traj : bound);
  skid_fixup_kernel<<<fgrid, PG::N, PATH_CTA_STRIDE, stream>>>(P, n_traj, step_offsets, reloc, path_table, n_table,
                                                            out_index, known, force_P, out_path_f64, out_internal_f64,
                                                            out_path, out_grid, out_status, scratch);
  return check_launch();
}

}  // extern "C"
