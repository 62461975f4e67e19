/*
 * ORACLE (test infrastructure).  Entry points: whole-frame planner, packed-batch driver
 * (pthreads, for bench.py's cpu_baseline / --impl reference legs) and stage wrappers.
 * Restates PathPlanner.calculate_path_in_global_frame for trackdrive/autocross
 * (/root/reference/fsd_path_planning/full_pipeline/full_pipeline.py:84-207) with the
 * "fresh planner per frame" batch semantic of SURVEY.md Q12.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "fsd_oracle.h"
#include "oracle_internal.h"

static void result_init(fsd_oracle_result *out) {
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < FSD_O_MAX_SORTED; ++i) out->left_idx[i] = out->right_idx[i] = -1;
  for (int i = 0; i < FSD_O_MAX_WV; ++i) out->l2r[i] = out->r2l[i] = -2;
  out->first_k[0][0] = out->first_k[0][1] = out->first_k[1][0] = out->first_k[1][1] = -1;
}

int fsd_oracle_sort(const double *cones_xy, const unsigned char *cones_type, int n, const double *pos,
                    const double *dir, fsd_oracle_result *out) {
  result_init(out);
  fsd_o_frame f = {cones_xy, cones_type, n, {pos[0], pos[1]}, {dir[0], dir[1]}};
  return fsd_o_sort_frame(&f, out);
}

int fsd_oracle_match(const double *left, int nl, const double *right, int nr, const double *pos, const double *dir,
                     fsd_oracle_result *out) {
  result_init(out);
  return fsd_o_match(left, nl, right, nr, pos, dir, out);
}

int fsd_oracle_path(const double *left_wv, int nl, const double *right_wv, int nr, const int *l2r, const int *r2l,
                    const double *pos, const double *dir, int force_P, const double *prev_path,
                    fsd_oracle_result *out) {
  result_init(out);
  return fsd_o_path(left_wv, nl, right_wv, nr, l2r, r2l, pos, dir, force_P, prev_path, out);
}

/* run_path_calculation with a global path (set_global_path / the acceleration mission), core_calculate_path.py:516-528 */
int fsd_oracle_path_global(const double *global_path, int n_points, const double *pos, const double *dir, int force_P,
                           const double *prev_path, fsd_oracle_result *out) {
  result_init(out);
  return fsd_o_path_global(global_path, n_points, pos, dir, force_P, prev_path, out);
}

int fsd_oracle_plan_frame(const double *cones_xy, const unsigned char *cones_type, int n, const double *pos,
                          const double *dir, int force_P, const double *prev_path, fsd_oracle_result *out) {
  result_init(out);
  fsd_o_frame f = {cones_xy, cones_type, n, {pos[0], pos[1]}, {dir[0], dir[1]}};
  fsd_o_sort_frame(&f, out);
  double left[2 * FSD_O_MAX_SORTED], right[2 * FSD_O_MAX_SORTED];
  for (int i = 0; i < out->n_left; ++i) {
    left[2 * i] = cones_xy[2 * out->left_idx[i]];
    left[2 * i + 1] = cones_xy[2 * out->left_idx[i] + 1];
  }
  for (int i = 0; i < out->n_right; ++i) {
    right[2 * i] = cones_xy[2 * out->right_idx[i]];
    right[2 * i + 1] = cones_xy[2 * out->right_idx[i] + 1];
  }
  fsd_o_match(left, out->n_left, right, out->n_right, pos, dir, out);
  fsd_o_path(&out->left_wv[0][0], out->n_left_wv, &out->right_wv[0][0], out->n_right_wv, out->l2r, out->r2l, pos,
             dir, force_P, prev_path, out);
  return 0;
}

typedef struct {
  const double *xy;
  const unsigned char *type;
  const int *offsets;
  const double *pos, *dir;
  const short *force_P;
  fsd_oracle_result *results;
  int n_frames, tid, nthreads;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (int b = j->tid; b < j->n_frames; b += j->nthreads) {
    int lo = j->offsets[b], hi = j->offsets[b + 1];
    fsd_oracle_plan_frame(j->xy + 2 * (size_t)lo, j->type + lo, hi - lo, j->pos + 2 * b, j->dir + 2 * b,
                          j->force_P ? j->force_P[b] : 0, NULL, &j->results[b]);
  }
  return NULL;
}

int fsd_oracle_plan_batch(const double *cones_xy, const unsigned char *cones_type, const int *offsets, int n_frames,
                          const double *pos, const double *dir, const short *force_P, int threads,
                          fsd_oracle_result *results) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  double warm[FSD_O_HORIZON * 4];
  fsd_oracle_initial_path(warm);
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < threads; ++t) {
    job_t j = {cones_xy, cones_type, offsets, pos, dir, force_P, results, n_frames, t, threads};
    jobs[t] = j;
    if (threads > 1) pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  if (threads == 1)
    worker(&jobs[0]);
  else
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  return 0;
}
