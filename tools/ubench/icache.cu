// Developer microbenchmark: instruction-cache capacities seen by straight-line code on sm_100a.
// One or more warps per SM loop over a block of N independent FFMA instructions (16 B each); cycles per
// instruction vs. code footprint shows where the L0 / L1 (ICC) / GPC-level caches end.
#include <cstdio>
#include <cuda_runtime.h>
#define I4 asm volatile("fma.rn.f32 %0, %0, %4, %5;\n fma.rn.f32 %1, %1, %4, %5;\n fma.rn.f32 %2, %2, %4, %5;\n fma.rn.f32 %3, %3, %4, %5;" : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3) : "f"(a), "f"(b));
#define I16 I4 I4 I4 I4
#define I64 I16 I16 I16 I16
#define I256 I64 I64 I64 I64
#define I1K I256 I256 I256 I256
#define I4K I1K I1K I1K I1K
template <int KB>
__global__ void k(float *out, int iters, long long *cyc, float a, float b) {
  float x0 = threadIdx.x, x1 = 1, x2 = 2, x3 = 3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (KB >= 256) { I4K I4K I4K I4K }
    else if constexpr (KB >= 128) { I4K I4K }
    else if constexpr (KB >= 96) { I4K I1K I1K }
    else if constexpr (KB >= 64) { I4K }
    else if constexpr (KB >= 48) { I1K I1K I1K }
    else if constexpr (KB >= 32) { I1K I1K }
    else if constexpr (KB >= 24) { I1K I256 I256 }
    else if constexpr (KB >= 16) { I1K }
    else if constexpr (KB >= 8) { I256 I256 }
    else { I256 }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
template <int KB>
void run(int warps, float *out, long long *cyc) {
  const int n_ins = KB * 64, iters = (1 << 22) / n_ins;
  k<KB><<<148, 32 * warps>>>(out, 2, cyc, 1.0f, 0.5f);
  k<KB><<<148, 32 * warps>>>(out, iters, cyc, 1.0f, 0.5f);
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%4d KB code, %2d warps/SM: %.3f cycles per warp-instruction (per warp), %.3f issue cycles/SM-instr\n", KB, warps,
         (double)c / ((double)n_ins * iters), (double)c / ((double)n_ins * iters * warps) * 4);
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16}) {
    run<4>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc); run<24>(w, out, cyc); run<32>(w, out, cyc); run<48>(w, out, cyc);
    run<64>(w, out, cyc); run<96>(w, out, cyc); run<128>(w, out, cyc); run<256>(w, out, cyc);
  }
  return 0;
}
