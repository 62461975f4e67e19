// Warp-cooperative primitives, written once for two builds:
//
//   * device build (nvcc, sm_100a): one warp plans one frame; FSD_LANES == 32; the primitives are
//     shuffles / ballots / __syncwarp.
//   * host-check build (g++, -DFSD_HOSTCHECK): FSD_LANES == 1; every primitive degenerates to the
//     identity.  It exists only so that `pytest -m "not gpu"` can exercise the kernels' control
//     flow and numerics on the CPU (tests/test_hostcheck.py).  It is never loaded by the product
//     package; the product has no CPU path.
//
// Coding discipline that makes both builds correct from one source:
//   - data-parallel loops are lane-strided:   for (int i = fsd_lane(); i < n; i += FSD_LANES)
//   - values produced by one lane for all lanes travel through per-warp shared memory followed
//     by wsync(); reductions go through wsum / wargmin / ... which return the result to ALL lanes;
//   - serial sections are guarded by `if (fsd_lane() == 0)` and followed by wsync().
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(FSD_HOSTCHECK)
#define FSD_DEVICE_BUILD 1
#define FSD_DEV __device__ __forceinline__
#define FSD_DEVFN __device__ __noinline__
#else
#define FSD_DEV static inline
#define FSD_DEVFN static
#endif

namespace fsd {

struct alignas(16) d2 {
  double x, y;
};

#ifdef FSD_DEVICE_BUILD

constexpr int FSD_LANES = 32;
constexpr unsigned FULL = 0xffffffffu;
// `count` (<= 32) independent tasks, one per lane
#define FSD_FOR_TASKS(e, count) if (const int e = (int)(threadIdx.x & 31u); e < (count))

FSD_DEV int fsd_lane() { return (int)(threadIdx.x & 31u); }
FSD_DEV void wsync() { __syncwarp(); }

FSD_DEVFN double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
FSD_DEV int wsum_i(int v) { return __reduce_add_sync(FULL, v); }
FSD_DEV int wmin_i(int v) { return __reduce_min_sync(FULL, v); }
FSD_DEV int wmax_i(int v) { return __reduce_max_sync(FULL, v); }
FSD_DEVFN double wmin_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
FSD_DEV unsigned wballot(bool p) { return __ballot_sync(FULL, p); }
FSD_DEV bool wany(bool p) { return __any_sync(FULL, p); }
// argmin with ties -> smallest index; lanes with idx < 0 do not take part
FSD_DEVFN void wargmin(double &v, int &idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(FULL, v, o);
    int oi = __shfl_xor_sync(FULL, idx, o);
    bool take = (oi >= 0) && (idx < 0 || ov < v || (ov == v && oi < idx));
    if (take) {
      v = ov;
      idx = oi;
    }
  }
}
// inclusive prefix sum across the lanes
FSD_DEVFN double wscan_incl(double v) {
  int lane = fsd_lane();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// value of the last lane
FSD_DEV double wlast(double v) { return __shfl_sync(FULL, v, 31); }
// N sums at once, butterfly steps interleaved across the N values (independent shuffles issue back to back, so
// the latency is that of ONE reduction instead of N)
template <int N>
FSD_DEV void wsum_vec(double (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] += __shfl_xor_sync(FULL, v[e], o);
  }
}

// Warp totals of 16 values with 16 shuffles instead of 80: in every butterfly step a lane hands over the half of the
// values its partner becomes responsible for and adds the partner's copy of its own half.  On return v[0] of lane L
// holds the total of value (L >> 1) & 15 (both lanes of a pair hold the same total).
FSD_DEV void wsum16_transposed(double (&v)[16]) {
  const int lane = fsd_lane();
#pragma unroll
  for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < h; ++j) {
      const double send = up ? v[j] : v[j + h];
      const double keep = up ? v[j + h] : v[j];
      v[j] = keep + __shfl_xor_sync(FULL, send, o);
    }
  }
  v[0] += __shfl_xor_sync(FULL, v[0], 1);
}

// Lane GROUPS: G consecutive lanes (G = 32, 16 or 8, aligned) that work on one task while the other groups of the warp work
// on other tasks -- the sort stage runs the LEFT and the RIGHT search of a frame on the two half-warps at the same time.
// All collectives name the group's own lanes in their member mask, so the groups may diverge (different trip counts,
// different branches); where their control flow coincides the hardware executes them together.
template <int G>
struct Grp {
  static_assert(G == 32 || G == 16 || G == 8, "lane groups are aligned powers of two");
  static constexpr int N = G;
  FSD_DEV static int lane() { return (int)(threadIdx.x & (unsigned)(G - 1)); }
  FSD_DEV static int index() { return (int)((threadIdx.x & 31u) / (unsigned)G); }  // which group of the warp
  FSD_DEV static unsigned base() { return threadIdx.x & 31u & ~(unsigned)(G - 1); }
  FSD_DEV static unsigned mask() { return G == 32 ? FULL : (((1u << (G & 31)) - 1u) << base()); }
  FSD_DEV static void sync() { __syncwarp(mask()); }
  FSD_DEV static unsigned ballot(bool p) { return __ballot_sync(mask(), p) >> base(); }  // bit i = lane i of the group
  FSD_DEV static int sum_i(int v) { return __reduce_add_sync(mask(), v); }
  FSD_DEV static int min_i(int v) { return __reduce_min_sync(mask(), v); }
  FSD_DEV static int bcast0(int v) { return __shfl_sync(mask(), v, 0, G); }  // the value of the group's lane 0
  FSD_DEV static bool any(bool p) { return __any_sync(mask(), p); }
  __device__ __noinline__ static double sum(double v) {
    const unsigned m = mask();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
  }
  __device__ __noinline__ static double min_d(double v) {
    const unsigned m = mask();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(m, v, o));
    return v;
  }
  // inclusive prefix sum across the lanes of the group
  __device__ __noinline__ static double scan_incl(double v) {
    const unsigned m = mask();
    const int l = lane();
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      const double t = __shfl_up_sync(m, v, o, G);
      if (l >= o) v += t;
    }
    return v;
  }
  FSD_DEV static double last(double v) { return __shfl_sync(mask(), v, G - 1, G); }  // value of the group's last lane
  // N sums at once, butterfly steps interleaved across the N values
  template <int N>
  FSD_DEV static void sum_vec(double (&v)[N]) {
    const unsigned m = mask();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int e = 0; e < N; ++e) v[e] += __shfl_xor_sync(m, v[e], o);
    }
  }
  // Group totals of 16 values with 15 (G = 16) / 16 (G = 32) shuffles instead of 64 / 80: in every butterfly step a lane
  // hands over the half of the values its partner becomes responsible for and adds the partner's copy of its own half.
  // On return v[0] of lane L holds the total of value `owned16()`.
  FSD_DEV static void sum16_transposed(double (&v)[16]) {
    static_assert(G == 32 || G == 16, "16 values need at least 16 lanes");
    const unsigned m = mask();
    const int l = lane();
#pragma unroll
    for (int h = 8, o = G / 2; h >= 1; h >>= 1, o >>= 1) {
      const bool up = (l & o) != 0;
#pragma unroll
      for (int j = 0; j < h; ++j) {
        const double send = up ? v[j] : v[j + h];
        const double keep = up ? v[j + h] : v[j];
        v[j] = keep + __shfl_xor_sync(m, send, o);
      }
    }
    if (G == 32) v[0] += __shfl_xor_sync(m, v[0], 1);
  }
  // which of the 16 values this lane owns after sum16_transposed, and whether it is the lane that adds it to memory
  FSD_DEV static int owned16() { return G == 32 ? (lane() >> 1) : lane(); }
  FSD_DEV static bool owner16() { return G == 32 ? (lane() & 1) == 0 : true; }
  // argmin with ties -> smallest index; lanes with idx < 0 do not take part; the result reaches every lane of the group
  __device__ __noinline__ static void argmin(double &v, int &idx) {
    const unsigned m = mask();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(m, v, o);
      const int oi = __shfl_xor_sync(m, idx, o);
      const bool take = (oi >= 0) && (idx < 0 || ov < v || (ov == v && oi < idx));
      if (take) {
        v = ov;
        idx = oi;
      }
    }
  }
};

#else  // host-check build: a warp of one lane

template <int G>
struct Grp {
  static constexpr int N = 1;
  static inline int lane() { return 0; }
  static inline int index() { return 0; }
  static inline void sync() {}
  static inline unsigned ballot(bool p) { return p ? 1u : 0u; }
  static inline int sum_i(int v) { return v; }
  static inline int min_i(int v) { return v; }
  static inline int bcast0(int v) { return v; }
  static inline bool any(bool p) { return p; }
  static inline double sum(double v) { return v; }
  static inline double min_d(double v) { return v; }
  static inline double scan_incl(double v) { return v; }
  static inline double last(double v) { return v; }
  template <int N>
  static inline void sum_vec(double (&)[N]) {}
  static inline void argmin(double &, int &) {}
};

constexpr int FSD_LANES = 1;
#define FSD_FOR_TASKS(e, count) for (int e = 0; e < (count); ++e)

FSD_DEV int fsd_lane() { return 0; }
FSD_DEV void wsync() {}
FSD_DEV double wsum(double v) { return v; }
FSD_DEV int wsum_i(int v) { return v; }
FSD_DEV int wmin_i(int v) { return v; }
FSD_DEV int wmax_i(int v) { return v; }
FSD_DEV double wmin_d(double v) { return v; }
FSD_DEV unsigned wballot(bool p) { return p ? 1u : 0u; }
FSD_DEV bool wany(bool p) { return p; }
FSD_DEV void wargmin(double &, int &) {}
FSD_DEV double wscan_incl(double v) { return v; }
FSD_DEV double wlast(double v) { return v; }
template <int N>
FSD_DEV void wsum_vec(double (&)[N]) {}

#endif

// The path stage (spline.cuh, path.cuh) runs one frame per lane group of FSD_PATH_LANES lanes: the whole warp by default,
// one lane on the host.  (-DFSD_PATH_LANES=16 plans TWO frames per warp.  Measured, profiles/r2_plan_mode_ab.txt: the two
// half-warps run diverged almost all the time -- warp instructions per frame fall by only 7 %, not the ~30 % their shared
// control flow and 18-task solves promise -- while the doubled shared-memory footprint costs a quarter of the resident
// warps and half of L1: 1.68 ms against 1.39 ms.  Kept as a build option, off.)
#ifndef FSD_PATH_LANES
#define FSD_PATH_LANES 32
#endif
#ifdef FSD_DEVICE_BUILD
using PG = Grp<FSD_PATH_LANES>;
#else
using PG = Grp<1>;
#endif
// `count` (<= 32) independent tasks, one per lane of the path lane group (in rounds when the group is narrower)
#if defined(FSD_DEVICE_BUILD) && FSD_PATH_LANES >= 32
#define FSD_FOR_PTASKS(e, count) if (const int e = PG::lane(); e < (count))
#elif defined(FSD_DEVICE_BUILD)
#define FSD_FOR_PTASKS(e, count) for (int e = PG::lane(); e < (count); e += PG::N)
#else
#define FSD_FOR_PTASKS(e, count) for (int e = 0; e < (count); ++e)
#endif

constexpr double PI = 3.14159265358979323846;

// double-precision transcendental functions are large once inlined: one out-of-line copy per kernel
FSD_DEVFN double fsd_atan2(double y, double x) { return atan2(y, x); }
FSD_DEVFN double fsd_acos(double x) { return acos(x); }
FSD_DEVFN double fsd_cos(double x) { return cos(x); }
FSD_DEVFN double fsd_sin(double x) { return sin(x); }
// fp64 division / square root expand to ~35 / ~25 instructions inline; the kernels are bound by instruction
// fetch, so every call site shares one out-of-line copy
FSD_DEVFN double fdiv(double a, double b) { return a / b; }
FSD_DEVFN double fsqrt(double a) { return sqrt(a); }
FSD_DEVFN double fnorm(double x, double y) { return sqrt(x * x + y * y); }
#ifdef FSD_DEVICE_BUILD
FSD_DEVFN double frsqrt(double a) { return rsqrt(a); }
FSD_DEV double frcp(double a) { return __drcp_rn(a); }
#else
FSD_DEVFN double frsqrt(double a) { return 1.0 / sqrt(a); }
FSD_DEV double frcp(double a) { return 1.0 / a; }
#endif

FSD_DEV double sgn(double v) { return (double)((v > 0.0) - (v < 0.0)); }
FSD_DEV int isgn(double v) { return (v > 0.0) - (v < 0.0); }

// x < c sqrt(v2) and x > c sqrt(v2) for v2 >= 0, without the square root
FSD_DEV bool lt_scaled(double x, double c, double v2) {
  const double x2 = x * x, cv = c * c * v2;
  return c >= 0.0 ? (x < 0.0 || x2 < cv) : (x < 0.0 && x2 > cv);
}
FSD_DEV bool gt_scaled(double x, double c, double v2) {
  const double x2 = x * x, cv = c * c * v2;
  return c >= 0.0 ? (x > 0.0 && x2 > cv) : (x >= 0.0 || x2 < cv);
}

// (a1 - a2 + 3 pi) mod 2 pi - pi with Python's modulo
// (reference: fsd_path_planning/utils/math_utils.py:663-676)
FSD_DEV double angle_difference(double a1, double a2) {
  double v = fmod(a1 - a2 + 3.0 * PI, 2.0 * PI);
  if (v < 0.0) v += 2.0 * PI;
  return v - PI;
}

// cosine of the angle between two vectors, clipped like vec_angle_between
// (fsd_path_planning/utils/math_utils.py:70-100); comparisons `angle < a` become `cos > cos(a)`
FSD_DEVFN double cos_between(double ax, double ay, double bx, double by) {
  // one reciprocal square root instead of two square roots and a division
  double c = (ax * bx + ay * by) * frsqrt((ax * ax + ay * ay) * (bx * bx + by * by));
  if (c < -1.0) c = -1.0;  // NaN (zero-length vector) stays NaN: every comparison is then false,
  if (c > 1.0) c = 1.0;    // as with the reference's arccos(NaN)
  return c;
}

}  // namespace fsd
