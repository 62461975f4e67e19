// HOST-CHECK build of the kernels' per-frame code (g++ -DFSD_HOSTCHECK, a warp of ONE lane).
//
// Test infrastructure for `pytest -m "not gpu"`: it lets the CPU-only container exercise the
// control flow and numerics of sort.cuh / match.cuh / spline.cuh / path.cuh against the golden
// vectors before any GPU time is spent.  It is NOT a product path: the package never loads this
// library (ft_fsd_path_planning_b200/_lib.py loads libfsdplan.so only and raises without CUDA).
#define FSD_HOSTCHECK 1
#include <cstdlib>
#include <cstring>
#include <new>

#include "frame.cuh"
#include "skidpad.cuh"

using namespace fsd;

// Capacities of the working memory handed to the per-frame code: NCAP knot records / PCAP path points as compiled, or
// more (fsd_hostcheck_set_caps) -- the records then continue BEHIND the PathSmem image, exactly as in path_kernel's
// second chance, where frame slot 0's arena is extended over the shared memory of the whole CTA.
// g_resume: the extra records are held in reserve -- fits start with NCAP records, suspend when they outgrow them and
// are resumed with all g_cap records (path_kernel's in-kernel second chance, spline.cuh FIT_SUSPENDED / fit_resume).
// start_cap: the records a fit starts with in that mode (NCAP, or fewer to make ordinary frames suspend in a test).
static int g_cap = NCAP, g_pcap = PCAP, g_resume = 0, g_start = NCAP;
extern "C" void fsd_hostcheck_set_caps(int cap, int pcap, int resume, int start_cap) {
  g_cap = cap > NCAP ? cap : NCAP;
  g_pcap = pcap > 0 ? pcap : PCAP;
  g_start = start_cap >= 8 && start_cap <= NCAP ? start_cap : NCAP;
  g_resume = resume && g_cap > g_start;
}
static PathSmem *new_path_smem() {
  void *raw = std::calloc(1, sizeof(PathSmem) + (size_t)(g_cap - NCAP) * sizeof(KnotRec));
  PathSmem *S = new (raw) PathSmem();
  S->pts = new d2[g_pcap];
  S->u = new double[g_pcap];
  S->pcap = g_pcap;
  S->W.cap = g_resume ? g_start : g_cap;
  S->W.suspendable = g_resume;
  return S;
}
static void free_path_smem(PathSmem *S) {
  delete[] S->pts;
  delete[] S->u;
  std::free(S);
}

extern "C" int fsd_hostcheck_initial_path(const fsd_params *params, double *out) {
  DevParams P = make_dev_params(*params);
  PathSmem *S = new_path_smem();
  unsigned st = initial_path_frame(*S, P, out);
  free_path_smem(S);
  return (int)st;
}

extern "C" int fsd_hostcheck_plan(const fsd_params *params, int n_frames, const double *xy, const uint8_t *type,
                                  const int32_t *offsets, const double *pos, const double *dir,
                                  const int16_t *force_P, const double *prev, double *out_path, int16_t *left_idx,
                                  int16_t *right_idx, int16_t *n_wv, double *left_wv, double *right_wv, int16_t *l2r,
                                  int16_t *r2l, int16_t *grid, int16_t *sort_dbg, uint32_t *status) {
  DevParams P = make_dev_params(*params);
  SortSmem *S = new SortSmem();
  MatchSmem *MS = new MatchSmem();
  PathSmem *Q = new_path_smem();
  double initial[FSD_HORIZON * 4];
  if (!prev) {
    initial_path_frame(*Q, P, initial);
    prev = initial;
  }
  StageOut O = {left_idx, right_idx, sort_dbg, n_wv, left_wv, right_wv, l2r, r2l, status};
  for (int b = 0; b < n_frames; ++b) {
    int lo = offsets[b], n = offsets[b + 1] - lo;
    unsigned st = 0;
    if (n > FSD_MAX_CONES) {
      n = FSD_MAX_CONES;
      st |= FSD_ST_OVERFLOW;
    }
    FramePose F = make_pose(pos[2 * b], pos[2 * b + 1], dir[2 * b], dir[2 * b + 1]);
    load_frame_plain(*S, xy + 2 * (size_t)lo, type + lo, n);
    st |= sort_frame(*S, n, F, P, sort_dbg + 8 * (size_t)b);
    store_sort(*S, b, O);
    st |= match_from_sort(*S, *MS, F, P);
    store_match(*MS, b, O);
    status[b] = st;
    path_from_tensors(*Q, b, O, F, force_P ? force_P[b] : 0, prev, P, out_path, nullptr, grid, g_resume ? g_cap : 0);
  }
  delete S;
  delete MS;
  free_path_smem(Q);
  return 0;
}

// path calculation along a global path (fsd_global_path_batch)
extern "C" int fsd_hostcheck_global_path(const fsd_params *params, int n_poses, const double *pos, const double *dir,
                                         const double *gpath, int n_points, const int16_t *force_P, const double *prev,
                                         int prev_stride, double *out, int16_t *grid, uint32_t *status) {
  DevParams P = make_dev_params(*params);
  PathSmem *Q = new_path_smem();
  double initial[FSD_HORIZON * 4];
  if (!prev) {
    initial_path_frame(*Q, P, initial);
    prev = initial;
    prev_stride = 0;
  }
  for (int b = 0; b < n_poses; ++b) {
    const FramePose F = make_pose(pos[2 * b], pos[2 * b + 1], dir[2 * b], dir[2 * b + 1]);
    int g[2] = {0, 0};
    status[b] = path_global(*Q, gpath, n_points, F, force_P ? force_P[b] : 0, prev + (size_t)b * prev_stride, P,
                            out + 160 * (size_t)b, g);
    grid[2 * b] = (int16_t)g[0];
    grid[2 * b + 1] = (int16_t)g[1];
  }
  free_path_smem(Q);
  return 0;
}

// the cost-matrix step alone (layout of fsd_knn_batch)
extern "C" int fsd_hostcheck_knn(const fsd_params *params, int n_frames, const double *xy, const uint8_t *type,
                                 const int32_t *offsets, uint8_t *out_nbr, uint8_t *out_deg) {
  DevParams P = make_dev_params(*params);
  SortSmem *S = new SortSmem();
  for (int b = 0; b < n_frames; ++b) {
    const int lo = offsets[b];
    int n = offsets[b + 1] - lo;
    if (n > FSD_MAX_CONES) n = FSD_MAX_CONES;
    load_frame_plain(*S, xy + 2 * (size_t)lo, type + lo, n);
    if (n >= 3) build_knn(*S, n, P);
    for (int i = 0; i < n; ++i)
      for (int s = 0; s < 2; ++s) {
        for (int q = 0; q < 5; ++q) out_nbr[((size_t)lo + i) * 10 + s * 5 + q] = n >= 3 ? S->nbr[s][i][q] : 0;
        out_deg[((size_t)lo + i) * 2 + s] = n >= 3 ? S->deg[s][i] : 0;
      }
  }
  delete S;
  return 0;
}

extern "C" int fsd_hostcheck_params_default(fsd_params *p) {
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
  return 0;
}

// the spline fit alone, for the comparison with scipy.interpolate.splprep
extern "C" int fsd_hostcheck_fit(const double *pts, int m, double s, double *t, int *n, double *c, int *k) {
  PathSmem *Q = new_path_smem();
  if (m > PCAP) {
    free_path_smem(Q);
    return 10;
  }
  for (int i = 0; i < m; ++i) {
    Q->pts[i].x = pts[2 * i];
    Q->pts[i].y = pts[2 * i + 1];
  }
  chord_params(Q->pts, m, Q->u);
  unsigned st = 0;
  int ier = fit_curve(Q->W, Q->pts, Q->u, m, s, &st);
  if (ier != 10) {
    *n = Q->W.n;
    *k = Q->W.k;
    for (int i = 0; i < Q->W.n; ++i) t[i] = Q->W.r[i].t;
    for (int i = 0; i < Q->W.n - Q->W.k - 1; ++i) {
      c[2 * i] = Q->W.r[i].c[0];
      c[2 * i + 1] = Q->W.r[i].c[1];
    }
  }
  free_path_smem(Q);
  return ier;
}

// ---- skidpad mission ------------------------------------------------------------------------------------------
extern "C" int fsd_hostcheck_skidpad_relocalize(const double *cones, int n, const double *pos, const double *orig_pos,
                                                const double *orig_dir, const double *jitter, const double *ref,
                                                double *out8, int *n_accepted) {
  SkidSmem *S = new SkidSmem();
  SkidReloc R;
  std::memset(&R, 0, sizeof(R));
  skidpad_relocalize(*S, cones, n, pos[0], pos[1], orig_pos[0], orig_pos[1], orig_dir[0], orig_dir[1], jitter, ref, &R,
                     n_accepted);
  std::memcpy(out8, &R, sizeof(R));
  delete S;
  return 0;
}

extern "C" int fsd_hostcheck_skidpad_steps(const fsd_params *params, const double *reloc8, const double *table,
                                           int n_table, const double *pos, const double *dir, int n_steps, int *state,
                                           const int16_t *force_P, const double *prev, int prev_stride, double *out,
                                           double *out_internal, int *index_out, uint32_t *status, int16_t *grid) {
  DevParams P = make_dev_params(*params);
  SkidReloc R;
  std::memcpy(&R, reloc8, sizeof(R));
  PathSmem *Q = new_path_smem();
  double *known = new double[4 * (size_t)n_steps];
  skidpad_track(R, table, n_table, pos, dir, n_steps, state, known, index_out);
  for (int s = 0; s < n_steps; ++s) {
    int g[2] = {0, 0};
    status[s] = skidpad_step(*Q, R, table, n_table, index_out[s], known + 4 * s, force_P ? force_P[s] : 0,
                             prev + (size_t)s * prev_stride, P, out + 160 * (size_t)s, out_internal + 160 * (size_t)s, g);
    grid[2 * s] = (int16_t)g[0];
    grid[2 * s + 1] = (int16_t)g[1];
  }
  delete[] known;
  free_path_smem(Q);
  return 0;
}
