// Launchers of the kernels built with LARGE static bounds (kernels_big.cu: 2 048 path points, 64 knots per fit).  They
// serve the inputs the planner kernels' own bounds do not cover -- a centre line taken from a global path is several times
// longer than one built from <= 12 matched cones -- and are not on the batched hot path.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/fsdplan.h"

size_t fsd_big_global_path_scratch_bytes(int n_poses, int sm_count);
int fsd_big_global_path(const fsd_params *params, int n_poses, const double *pos, const double *dir, const double *gpath,
                        int n_points, const int16_t *force_P, const double *prev, int prev_stride, double *out_f64,
                        float *out_f32, int16_t *grid_out, uint32_t *status, unsigned char *scratch, int *counter,
                        int sm_count, cudaStream_t stream);

// second chance for the frames path_kernel marked (status bit 31): the path stage with the large bounds
size_t fsd_big_path_fixup_scratch_bytes();
int fsd_big_path_fixup(const fsd_params *params, int n_frames, int coords_f64, const void *pos, const void *dir,
                       const int16_t *n_wv, const double *left_wv, const double *right_wv, const int16_t *l2r,
                       const int16_t *r2l, const int16_t *force_P, const double *prev, int prev_stride, double *out_f64,
                       float *out_f32, int16_t *grid_out, uint32_t *status, unsigned char *scratch, cudaStream_t stream,
                       const fsd_gather *gather = nullptr);

#ifdef __CUDACC__
// The all-gather of the output paths fused into the kernels that produce them (fsdplan.h: fsd_gather): value i of frame b
// goes to row first_row + b of every GPU's gathered buffer -- one multimem.st through the NVSwitch multicast address when
// the caller has one (the switch replicates the store to all GPUs), else one plain store per peer-mapped pointer (NVLink
// peer memory).  Posted writes: they overlap the planning of the next frame.
__device__ __forceinline__ void fsd_store_peers(const fsd_gather &G, int b, int i, float v) {
  const size_t o = (size_t)(G.first_row + b) * (FSD_HORIZON * 4) + (size_t)i;
  if (G.multicast_out_path) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(G.multicast_out_path + o), "f"(v) : "memory");
  } else {
    for (int r = 0; r < G.n_peers; ++r) G.peer_out_path[r][o] = v;
  }
}
#endif
