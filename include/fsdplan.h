/*
 * libfsdplan.so -- C-ABI of the B200-native batched cone-track path planner.
 *
 * Drop-in boundary for the reference's per-frame planner
 *   PathPlanner.calculate_path_in_global_frame            fsd_path_planning/full_pipeline/full_pipeline.py:84-207
 * i.e. its three stage calls
 *   ConeSorting.run_cone_sorting                          fsd_path_planning/sorting_cones/core_cone_sorting.py:117-136
 *   ConeMatching.run_cone_matching                        fsd_path_planning/cone_matching/core_cone_matching.py:87-124
 *   CalculatePath.run_path_calculation                    fsd_path_planning/calculate_path/core_calculate_path.py:514-575
 * executed for a whole batch of independent frames ("fresh PathPlanner per frame") per call.
 * The reference has no FFI layer of its own (it is pure Python); INTEGRATION.md shows the ctypes
 * binding a maintainer adds to route PathPlanner through this library.
 *
 * Conventions
 *   - every pointer except `params`, `inter` (the struct itself) is a DEVICE pointer; the caller
 *     owns all buffers (e.g. torch tensors); the library never allocates or frees caller memory,
 *     keeps no pointer after return and is asynchronous on `stream`;
 *   - return value 0 = launched; negative = argument / launch error (fsd_strerror); per-frame soft
 *     failures are reported in out_status bits, never as errors;
 *   - packed frame batch: frame b owns cones offsets[b] .. offsets[b+1]-1; at most
 *     FSD_MAX_CONES cones per frame (more -> FSD_ST_OVERFLOW, frame planned on the first 256);
 *   - sort indices index the frame's own cone list (reference: flatten_cones_by_type_array,
 *     fsd_path_planning/sorting_cones/trace_sorter/core_trace_sorter.py:37-54);
 *   - the library owns a few internal CUDA streams / events per device (created on first use): a large fsd_plan_batch
 *     call may run its second half on one of them, forked from and joined back into `stream` with events, so the call
 *     is still ONE asynchronous operation ordered on `stream` (see fsd_plan_launches);
 *   - no CPU fallback exists: without a CUDA device every CUDA entry point returns FSD_ERR_NO_DEVICE (the host planner
 *     fsd_plan_batch_cpu is a separate, explicit entry point);
 *   - besides its internal streams / events the library owns 36 KB of device memory per device (frame counters of its
 *     free-running kernels and the histograms of the path keys, allocated on first use).
 */
#ifndef FSDPLAN_H
#define FSDPLAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSD_ABI_VERSION 1

#define FSD_MAX_CONES 256 /* cones per frame */
#define FSD_MAX_SORTED 12 /* max_length, fsd_path_planning/config.py:37 */
#define FSD_MAX_WV 32     /* cones per side after virtual-cone insertion (12 real + 12 virtual, rounded up) */
#define FSD_HORIZON 40    /* mpc_prediction_horizon, fsd_path_planning/config.py:58 */

/* ConeTypes, fsd_path_planning/utils/cone_types.py:10-19 */
#define FSD_CONE_UNKNOWN 0
#define FSD_CONE_RIGHT 1 /* yellow */
#define FSD_CONE_LEFT 2  /* blue */
#define FSD_CONE_ORANGE_SMALL 3
#define FSD_CONE_ORANGE_BIG 4

/* MissionTypes, fsd_path_planning/utils/mission_types.py:11-25 (trackdrive/autocross share one path) */
#define FSD_MISSION_AUTOCROSS 3
#define FSD_MISSION_TRACKDRIVE 4

/* out_status bits */
#define FSD_ST_NO_LEFT (1u << 0)      /* no left configuration (NoPathError / no start cone) */
#define FSD_ST_NO_RIGHT (1u << 1)
#define FSD_ST_FEW_CONES (1u << 2)    /* both sides < 3 cones: previous path used as centre line  (core_calculate_path.py:531-536) */
#define FSD_ST_FEW_MATCHES (1u << 3)  /* < 2 centre points: previous path                          (:201-203) */
#define FSD_ST_FIT1_FAILED (1u << 4)  /* first fit invalid (ValueError): previous path re-fitted   (:214-221) */
#define FSD_ST_PATH_TOO_FAR (1u << 5) /* path > 5 m from the car: previous path                    (:232-237) */
#define FSD_ST_MPC_FAILED (1u << 6)   /* tail raised ValueError: redone with the previous path     (:561-570) */
#define FSD_ST_TIE_P (1u << 7)        /* size of the last evaluation grid decided by the tie rule */
#define FSD_ST_OVERFLOW (1u << 8)     /* a static bound was exceeded (cones, DFS leaves, knots, path points) */
#define FSD_ST_REF_RAISES (1u << 9)   /* the reference raises an exception on this input.  Raised by the path stage: output =
                                         previous path.  Raised by the matching stage (a sorted side with exactly one cone,
                                         functional_cone_matching.py:130): that side is matched as all -1 and the path is
                                         computed from what remains -- a usable path, but not one the reference produces;
                                         the drop-in PathPlanner raises on this bit like the reference does */
#define FSD_ST_UNSUPPORTED (1u << 10) /* reference takes a latent-bug path that is not reproduced; output = previous path */

#define FSD_OK 0
#define FSD_ERR_ARG (-1)
#define FSD_ERR_WORKSPACE (-2)
#define FSD_ERR_LAUNCH (-3)
#define FSD_ERR_NO_DEVICE (-4)
#define FSD_ERR_MISSION (-5)

/* Every tunable of the path, default-initialised to the reference's values by fsd_params_default. */
typedef struct fsd_params {
  /* cone sorting: fsd_path_planning/config.py:33-41 and inline constants of trace_sorter/ */
  int32_t max_n_neighbors;             /* 5 */
  int32_t max_length;                  /* 12 */
  double max_dist;                     /* 6.5 m */
  double max_dist_to_first;            /* 6.0 m */
  double threshold_directional_angle;  /* 40 deg, rad */
  double threshold_absolute_angle;     /* 65 deg, rad */
  double car_size;                     /* 2.1 m, find_configs_and_scores.py:93 */
  int32_t max_dfs_pops;                /* guard for the exhaustive search (the reference has none) */
  int32_t reserved0;
  /* cone matching: config.py:124-129, core_cone_matching.py:101-102 */
  double min_track_width;   /* 3 m */
  double max_search_range;  /* 5 m (major radius = 1.5 x) */
  double max_search_angle;  /* 50 deg, rad */
  /* path calculation: config.py:48, 55-59 */
  double smoothing;                        /* 0.2 */
  double predict_every;                    /* 0.1 m */
  double maximal_distance_for_valid_path;  /* 5 m */
  double mpc_path_length;                  /* 20 m */
  double refit_smoothing;                  /* 0.01, path_parameterization.py:157-159 */
} fsd_params;

/* Optional per-frame intermediates (device pointers, each nullable).  When a pointer is NULL the
 * library keeps that tensor in `workspace`. */
typedef struct fsd_intermediate {
  double *path_f64;   /* [B][40][4] u, x, y, curvature in fp64 (out_path is the fp32 copy) */
  int16_t *n_wv;      /* [B][2]   cones per side after matching: left, right */
  double *left_wv;    /* [B][FSD_MAX_WV][2] left cones with virtual cones */
  double *right_wv;   /* [B][FSD_MAX_WV][2] */
  int16_t *l2r;       /* [B][FSD_MAX_WV] match index into right_wv or -1; entries >= n are -2 */
  int16_t *r2l;       /* [B][FSD_MAX_WV] */
  int16_t *grid;      /* [B][2]   P (size of the last evaluation grid), points entering the last re-fit */
  int16_t *sort_dbg;  /* [B][8]   first_k left[2], right[2], n_configs[2], dfs pops[2] */
} fsd_intermediate;

int fsd_abi_version(void);
const char *fsd_strerror(int code);
int fsd_params_default(fsd_params *params);

/* bytes of scratch needed by the batch entry points below for B frames */
size_t fsd_workspace_bytes(int n_frames, int total_cones);

/* Kernel launches one fsd_plan_batch call makes for n_frames frames on the current device: 4 (sort, match, path, the
 * large-bounds second chance of the path stage), 5 for batches large enough for the work-ordered path stage (+ the path
 * keys), twice that when the plan mode splits the batch into two chunks -- the second one on an internal side stream that is
 * forked from and joined back into the caller's stream with events, so the call stays asynchronous and ordered on
 * the caller's stream (and capturable in a CUDA graph -- after one warm-up call: the FIRST call on a device initialises
 * per-device constants, among them the initial path of a fresh planner, on a private stream with its own
 * synchronisation, and must not be made under stream capture). */
int fsd_plan_launches(int n_frames);

/* The constant initial path of a fresh planner (core_calculate_path.py:103-107), computed on the
 * device with the path kernels; out_prev_path: device, [40][4] fp64. */
int fsd_initial_path(const fsd_params *params, double *out_prev_path, void *stream);

/*
 * Full planner: sort -> match -> path for n_frames independent frames.
 *   cones_xy [total][2], pos [B][2], dir [B][2] : fp32 (fsd_plan_batch) or fp64 (fsd_plan_batch_f64)
 *   out_path      [B][40][4] fp32  (u, x, y, curvature).  Write-only for the library: it may be a device pointer
 *                 or PINNED (mapped) HOST memory -- the path kernel then stores every finished frame straight
 *                 into host memory (posted PCIe writes that overlap the planning; no device->host copy afterwards)
 *   out_left_idx  [B][12] int16, -1 padded;  out_right_idx likewise  (the "sort indices")
 *   inter         nullable
 *   force_P       nullable, [B]; > 0 forces the size of the last evaluation grid (parity mode)
 *   prev_path     nullable; fp64 [40][4] (prev_path_stride = 0, shared) or per frame
 *                 (prev_path_stride = 160); NULL = the initial path of a fresh planner
 *   out_status    [B] uint32
 */
int fsd_plan_batch(const fsd_params *params, int mission, int n_frames, const float *cones_xy,
                   const uint8_t *cones_type, const int32_t *offsets, const float *pos, const float *dir,
                   float *out_path, int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                   const int16_t *force_P, const double *prev_path, int prev_path_stride, uint32_t *out_status,
                   void *workspace, size_t workspace_bytes, void *stream);

int fsd_plan_batch_f64(const fsd_params *params, int mission, int n_frames, const double *cones_xy,
                       const uint8_t *cones_type, const int32_t *offsets, const double *pos, const double *dir,
                       float *out_path, int16_t *out_left_idx, int16_t *out_right_idx,
                       const fsd_intermediate *inter, const int16_t *force_P, const double *prev_path,
                       int prev_path_stride, uint32_t *out_status, void *workspace, size_t workspace_bytes,
                       void *stream);

/*
 * fsd_plan_batch_cpu: the same planner on the HOST (BASELINE config 1, "CPU plumbing"; machines without a GPU).  All
 * pointers are HOST pointers, coordinates are fp64, no workspace, synchronous; n_threads worker threads share the batch.
 * It is compiled from the very sources the CUDA kernels are compiled from (csrc/sort.cuh, match.cuh, spline.cuh,
 * path.cuh as a warp of one lane, csrc/cpu_backend.cpp) -- not from the test oracle -- and it is an explicit entry point:
 * no CUDA entry point ever falls back to it.  inter (nullable) holds HOST pointers; at least one of out_path /
 * inter->path_f64 must be given.
 */
int fsd_plan_batch_cpu(const fsd_params *params, int mission, int n_frames, const double *cones_xy,
                       const uint8_t *cones_type, const int32_t *offsets, const double *pos, const double *dir,
                       float *out_path, int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                       const int16_t *force_P, const double *prev_path, int prev_path_stride, uint32_t *out_status,
                       int n_threads);
int fsd_initial_path_cpu(const fsd_params *params, double *out_prev_path /* host, [40][4] */);

/* fsd_plan_batch with the coordinate type as a flag (coords_f64 != 0: fp64) and an optional CUDA event:
 * `chunk_ready_event` (cudaEvent_t, nullable) is recorded on `stream` as soon as the outputs of the first
 * fsd_plan_first_chunk(n_frames) frames are final -- a caller can start consuming them (e.g. the all-gather of their paths
 * on a communication stream, ft_fsd_path_planning_b200/distributed.py) while the rest of the batch is still being planned.
 * The whole call remains ordered on `stream`. */
int fsd_plan_first_chunk(int n_frames);
int fsd_plan_batch_ex(const fsd_params *params, int mission, int n_frames, int coords_f64, const void *cones_xy,
                      const uint8_t *cones_type, const int32_t *offsets, const void *pos, const void *dir,
                      float *out_path, int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                      const int16_t *force_P, const double *prev_path, int prev_path_stride, uint32_t *out_status,
                      void *workspace, size_t workspace_bytes, void *stream, void *chunk_ready_event);

/*
 * Multi-GPU: the all-gather of the output paths FUSED into the path kernel (replaces the
 * all_gather_into_tensor that follows fsd_plan_batch when a frame batch is sharded over the GPUs of one box).
 * Every GPU plans its shard and the path kernel stores each finished frame's 40 x 4 fp32 path, besides out_path, into
 * row (first_row + b) of EVERY peer's gathered [n_global][40][4] buffer: plain stores through peer-mapped pointers
 * (NVLink / NVSwitch peer memory: CUDA IPC or torch symmetric memory, ft_fsd_path_planning_b200/distributed.py
 * PeerGather), or -- when multicast_out_path is non-NULL -- ONE multimem.st per value through the NVSwitch multicast
 * address of the buffer, which the switch replicates to all GPUs.  The transfer overlaps the planning frame by frame;
 * no collective runs afterwards, only a cross-GPU barrier before the gathered buffer is read (the caller's:
 * PeerGather.finish).  peer_out_path[r] may equal the local buffer for r = own rank.  Frames re-planned by the
 * large-bounds second chance are stored to the peers as well.
 */
#define FSD_MAX_PEERS 16
typedef struct fsd_gather {
  int32_t n_peers;                       /* entries used in peer_out_path (0: no peer stores) */
  int32_t reserved0;
  int64_t first_row;                     /* row of this call's frame 0 in the gathered buffers */
  float *peer_out_path[FSD_MAX_PEERS];   /* device-accessible pointers to every GPU's [n_global][40][4] fp32 buffer */
  float *multicast_out_path;             /* nullable: multicast (multimem) address of the same buffer */
} fsd_gather;

/* fsd_plan_batch_ex + gather (nullable: exactly fsd_plan_batch_ex).  out_path may be NULL when gather is given. */
int fsd_plan_batch_gather(const fsd_params *params, int mission, int n_frames, int coords_f64, const void *cones_xy,
                          const uint8_t *cones_type, const int32_t *offsets, const void *pos, const void *dir,
                          float *out_path, int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                          const int16_t *force_P, const double *prev_path, int prev_path_stride, uint32_t *out_status,
                          void *workspace, size_t workspace_bytes, void *stream, void *chunk_ready_event,
                          const fsd_gather *gather);

/* Stage entry points (same conventions).  fsd_sort_batch: ConeSorting only. */
int fsd_sort_batch(const fsd_params *params, int n_frames, const float *cones_xy, const uint8_t *cones_type,
                   const int32_t *offsets, const float *pos, const float *dir, int16_t *out_left_idx,
                   int16_t *out_right_idx, int16_t *sort_dbg, uint32_t *out_status, void *stream);

/* fsd_knn_batch: the cost-matrix step in isolation (SURVEY.md 8a row S3, 8d "cost-matrix step"): both sides' k-NN graphs
 * of create_adjacency_matrix, fsd_path_planning/sorting_cones/trace_sorter/adjacency_matrix.py:60-128 (squared distances,
 * opposite colour excluded per side, k = min(5, N-1) nearest, edges longer than max_dist cut, A & A^T).  Per cone (packed
 * order, frame-local indices): out_nbr [total][2][5] uint8 = the cone's neighbours in the LEFT / RIGHT graph in ascending
 * index order (the CSR order of end_configurations.py:28-71; entries past the degree are unspecified),
 * out_deg [total][2] uint8 = the degrees.  Frames with fewer than 3 cones get degree 0 (the sorter does not build a
 * graph for them).  Algorithmic bytes per frame: 9 N + 16 read, 12 N written (SURVEY's B_cm = 19.25 N + 16 counts the
 * 10 N bytes of lists + two N/8-byte masks). */
int fsd_knn_batch(const fsd_params *params, int n_frames, int coords_f64, const void *cones_xy,
                  const uint8_t *cones_type, const int32_t *offsets, uint8_t *out_nbr, uint8_t *out_deg, void *stream);

/* fsd_match_batch: ConeMatching on given sort indices; writes n_wv, left_wv, right_wv, l2r, r2l of `inter`
 * (all five must be non-NULL). */
int fsd_match_batch(const fsd_params *params, int n_frames, const float *cones_xy, const int32_t *offsets,
                    const float *pos, const float *dir, const int16_t *left_idx, const int16_t *right_idx,
                    const fsd_intermediate *inter, uint32_t *out_status, void *stream);

/* The two launches fsd_plan_batch is made of, callable one by one (e.g. to time each kernel).
 * coords_f64 != 0: cones_xy / pos / dir are fp64, else fp32.
 * fsd_sort_match_batch: ConeSorting + ConeMatching.  inter->n_wv, left_wv, right_wv, l2r, r2l must be non-NULL
 * (sort_dbg optional).  Writes out_status. */
int fsd_sort_match_batch(const fsd_params *params, int n_frames, int coords_f64, const void *cones_xy,
                         const uint8_t *cones_type, const int32_t *offsets, const void *pos, const void *dir,
                         int16_t *out_left_idx, int16_t *out_right_idx, const fsd_intermediate *inter,
                         uint32_t *out_status, void *stream);

/* fsd_path_batch: CalculatePath on given matching results.  inter->n_wv, left_wv, right_wv, l2r, r2l (inputs) and
 * path_f64 (output) must be non-NULL (grid optional); workspace as for fsd_plan_batch (fsd_workspace_bytes).
 * Status bits are OR-ed INTO out_status (zero it when the
 * stage is used on its own).  prev_path NULL = initial path of a fresh planner (default spline parameters only). */
int fsd_path_batch(const fsd_params *params, int n_frames, int coords_f64, const void *pos, const void *dir,
                   const fsd_intermediate *inter, const int16_t *force_P, const double *prev_path,
                   int prev_path_stride, float *out_path, uint32_t *out_status, void *workspace,
                   size_t workspace_bytes, void *stream);
/* fsd_path_batch with the fused all-gather of fsd_plan_batch_gather (gather nullable). */
int fsd_path_batch_gather(const fsd_params *params, int n_frames, int coords_f64, const void *pos, const void *dir,
                   const fsd_intermediate *inter, const int16_t *force_P, const double *prev_path,
                   int prev_path_stride, float *out_path, uint32_t *out_status, void *workspace,
                   size_t workspace_bytes, void *stream,
                          const fsd_gather *gather);

/*
 * fsd_global_path_batch: CalculatePath.run_path_calculation with a GLOBAL PATH
 * (fsd_path_planning/calculate_path/core_calculate_path.py:516-528): what PathPlanner does after set_global_path(path)
 * (full_pipeline.py:81) and what the acceleration / EBS missions do with their known map once relocalized
 * (acceleration_relocalization.py:168-169).  For each of n_poses poses: the points of global_path [n_points][2] (fp64,
 * shared by all poses) within 30 m of the position, starting n_points / 3 points before the closest one, are the centre
 * line; then fit, validity check, MPC tail as in fsd_path_batch.  More than 704 such points: FSD_ST_OVERFLOW, the previous
 * path is returned.  prev_path / force_P / outputs as for fsd_path_batch; poses are fp64.
 */
size_t fsd_global_path_workspace_bytes(int n_poses);
int fsd_global_path_batch(const fsd_params *params, int n_poses, const double *pos, const double *dir,
                          const double *global_path, int n_points, const int16_t *force_P, const double *prev_path,
                          int prev_path_stride, float *out_path, double *out_path_f64, int16_t *out_grid,
                          uint32_t *out_status, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Skidpad mission (MissionTypes.skidpad): relocalization + stateful tracking of the canonical skidpad path.
 *   SkidpadRelocalizer.do_relocalization_once   fsd_path_planning/relocalization/skidpad/skidpad_relocalizer.py:198-240
 *   SkidpadCalculatePath.fit_matches_as_spline  fsd_path_planning/calculate_path/skidpad_calculate_path.py:49-71
 *   pose into / path out of the map frame       fsd_path_planning/full_pipeline/full_pipeline.py:122-140, 178-194
 * Coordinates are fp64 (SLAM frames sit hundreds of metres from the origin).
 *
 * fsd_skidpad_relocalize_batch: one relocalization attempt for each of n_traj trajectories.
 *   cones_xy [total][2] / offsets [T+1]: the cones seen by each trajectory; pos [T][2] current position;
 *   orig_pos / orig_dir [T][2]: pose at the FIRST attempt of the trajectory; jitter [1140*6]: numpy
 *   RandomState(42).randn (skidpad_relocalizer.py:38, 53); ref_centers [4]: right (x, y), left (x, y) circle centres of
 *   the canonical path; reloc [T][8] out: translation(2), rotation, right reference centre(2), right calculated
 *   centre(2), success flag (1.0 / 0.0); n_accepted [T] (nullable): circles that passed the filters.
 */
int fsd_skidpad_relocalize_batch(const fsd_params *params, int n_traj, const double *cones_xy, const int32_t *offsets,
                                 const double *pos, const double *orig_pos, const double *orig_dir,
                                 const double *jitter, const double *ref_centers, double *reloc, int32_t *n_accepted,
                                 void *stream);

/*
 * fsd_skidpad_plan_batch: n_steps planner steps of n_traj trajectories (steps of trajectory t are
 * step_offsets[t] .. step_offsets[t+1]-1, in time order).  index_state [T] in/out: index_along_path of each
 * trajectory.  path_table [n_table][2]: the canonical path the planner tracks (every 2nd point of the track table).
 * prev_path: as for fsd_plan_batch, in the MAP frame (= out_internal_f64 of the previous step).  Outputs per step:
 * out_path fp32 (nullable) and out_path_f64 in the SLAM frame, out_internal_f64 in the map frame, out_index (path
 * index used, -1 before relocalization), out_grid (nullable), out_status.
 */
size_t fsd_skidpad_workspace_bytes(int n_steps);
int fsd_skidpad_plan_batch(const fsd_params *params, int n_traj, int n_steps, const int32_t *step_offsets,
                           const double *pos, const double *dir, const double *reloc, int32_t *index_state,
                           const double *path_table, int n_table, const int16_t *force_P, const double *prev_path,
                           int prev_path_stride, float *out_path, double *out_path_f64, double *out_internal_f64,
                           int32_t *out_index, int16_t *out_grid, uint32_t *out_status, void *workspace,
                           size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FSDPLAN_H */
