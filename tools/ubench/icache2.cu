// Developer microbenchmark 2: warps of one SM running DIFFERENT regions of a large straight-line code block at the same
// time (no instruction-cache line sharing between warps), against all warps running the same region.
#include <cstdio>
#include <cuda_runtime.h>
#define I4 asm volatile("fma.rn.f32 %0, %0, %4, %5;\n fma.rn.f32 %1, %1, %4, %5;\n fma.rn.f32 %2, %2, %4, %5;\n fma.rn.f32 %3, %3, %4, %5;" : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3) : "f"(a), "f"(b));
#define I16 I4 I4 I4 I4
#define I64 I16 I16 I16 I16
#define I256 I64 I64 I64 I64
#define I1K I256 I256 I256 I256
// 16 regions of REG instructions each
template <int REGK>  // region size in units of 256 instructions (4 KB)
__global__ void k(float *out, int iters, long long *cyc, float a, float b, int mode) {
  float x0 = threadIdx.x, x1 = 1, x2 = 2, x3 = 3;
  const int w = threadIdx.x >> 5;
  // mode 0: every warp walks regions 0..15 in the same order; mode 1: warp w starts at region w (all 16 warps of the SM
  // are in different regions at any time); mode 2: warps of one scheduler (w % 4 equal) differ, schedulers agree
  const int start = mode == 0 ? 0 : (mode == 1 ? w : (w >> 2) * 4);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
      switch ((p + start) & 15) {
#define REGION if constexpr (REGK >= 4) { I1K } else if constexpr (REGK >= 2) { I256 I256 } else { I256 }
        case 0: REGION break; case 1: REGION break; case 2: REGION break; case 3: REGION break;
        case 4: REGION break; case 5: REGION break; case 6: REGION break; case 7: REGION break;
        case 8: REGION break; case 9: REGION break; case 10: REGION break; case 11: REGION break;
        case 12: REGION break; case 13: REGION break; case 14: REGION break; default: REGION break;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
template <int REGK>
void run(int warps, int mode, float *out, long long *cyc) {
  const int n_ins = REGK * 256 * 16, iters = (1 << 21) / n_ins;
  k<REGK><<<148, 32 * warps>>>(out, 1, cyc, 1.0f, 0.5f, mode);
  k<REGK><<<148, 32 * warps>>>(out, iters, cyc, 1.0f, 0.5f, mode);
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%4d KB code (16 regions), %2d warps/SM, mode %d: %.3f cycles per instr per warp, SM IPC %.2f\n", REGK * 64, warps, mode,
         (double)c / ((double)n_ins * iters), (double)n_ins * iters * warps / (double)c);
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16})
    for (int mode : {0, 1, 2}) { run<1>(w, mode, out, cyc); run<2>(w, mode, out, cyc); run<4>(w, mode, out, cyc); }
  return 0;
}
