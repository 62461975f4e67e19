/*
 * ORACLE (test infrastructure).  Skidpad mission: relocalization (K1) and the stateful global-path tracker (K2)
 * of SURVEY.md section 8(a).  Restates
 *   /root/reference/fsd_path_planning/relocalization/skidpad/skidpad_relocalizer.py:31-240
 *   /root/reference/fsd_path_planning/calculate_path/skidpad_calculate_path.py:49-71
 *   /root/reference/fsd_path_planning/full_pipeline/full_pipeline.py:122-140, 178-194
 * Third-party pieces: sklearn.cluster.DBSCAN(eps=3, min_samples=1) == connected components of the "distance <= 3"
 * graph, labelled in order of first appearance; numpy RandomState(42).randn jitter is passed in by the caller
 * (it is a data-independent constant sequence, generated with numpy itself).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fsd_oracle.h"
#include "oracle_internal.h"

#define MAX_NEAR 20

static int cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

static double median(double *v, int n) {
  qsort(v, n, sizeof(double), cmp_double);
  return n % 2 ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]);
}

int fsd_oracle_skidpad_relocalize(const double *cones_xy, int n, const double *pos, const double *orig_pos,
                                  const double *orig_dir, const double *jitter, const double *ref_centers,
                                  fsd_oracle_reloc *out) {
  memset(out, 0, sizeof(*out));
  /* the 20 cones nearest to the vehicle, nearest first (skidpad_relocalizer.py:207-212) */
  int m = n < MAX_NEAR ? n : MAX_NEAR;
  double near[2 * MAX_NEAR];
  {
    char *used = (char *)calloc(n > 0 ? n : 1, 1);
    for (int q = 0; q < m; ++q) {
      int best = -1;
      double bd = 0.0;
      for (int i = 0; i < n; ++i) {
        if (used[i]) continue;
        double dx = cones_xy[2 * i] - pos[0], dy = cones_xy[2 * i + 1] - pos[1];
        double d = sqrt(dx * dx + dy * dy);
        if (best < 0 || d < bd) {
          bd = d;
          best = i;
        }
      }
      used[best] = 1;
      near[2 * q] = cones_xy[2 * best];
      near[2 * q + 1] = cones_xy[2 * best + 1];
    }
    free(used);
  }
  /* circle_fit_powerset :31-64: only the 3-subsets are ever fitted */
  int cap = 1140, nacc = 0, t = 0;
  double *cen = (double *)malloc(sizeof(double) * 2 * cap);
  for (int a = 0; a < m; ++a)
    for (int b = a + 1; b < m; ++b)
      for (int c = b + 1; c < m; ++c, ++t) {
        int idx[3] = {a, b, c};
        double p[6];
        for (int q = 0; q < 3; ++q) {
          p[2 * q] = near[2 * idx[q]];
          p[2 * q + 1] = near[2 * idx[q] + 1];
        }
        /* mean distance to the closest other point of the subset */
        double mean_d = 0.0;
        for (int q = 0; q < 3; ++q) {
          double mn = INFINITY;
          for (int r = 0; r < 3; ++r) {
            if (r == q) continue;
            double d2 = fsd_o_cdist_sq(p[2 * r], p[2 * r + 1], p[2 * q], p[2 * q + 1]);
            double d = sqrt(d2);
            if (d < mn) mn = d;
          }
          mean_d += mn;
        }
        mean_d /= 3.0;
        for (int q = 0; q < 6; ++q) p[q] += jitter[6 * t + q] * 1e-3;
        double cx, cy, r;
        fsd_o_circle_fit(p, 3, &cx, &cy, &r);
        double resid = 0.0;
        for (int q = 0; q < 3; ++q) {
          double dx = cx - p[2 * q], dy = cy - p[2 * q + 1];
          resid += fabs(sqrt(dx * dx + dy * dy) - r);
        }
        resid /= 3.0;
        if (fabs(r - 7.625) < 1.0 && fabs(mean_d - 2.4) < 1.5 && resid < 0.4) {
          cen[2 * nacc] = cx;
          cen[2 * nacc + 1] = cy;
          nacc++;
        }
      }
  out->n_accepted = nacc;
  if (nacc < 3) {
    free(cen);
    return 0;
  }
  /* DBSCAN(eps=3, min_samples=1): connected components, labels in order of first appearance */
  int *label = (int *)malloc(sizeof(int) * nacc);
  for (int i = 0; i < nacc; ++i) label[i] = -1;
  int nlab = 0;
  int *stack = (int *)malloc(sizeof(int) * nacc);
  for (int i = 0; i < nacc; ++i) {
    if (label[i] >= 0) continue;
    int sp = 0;
    stack[sp++] = i;
    label[i] = nlab;
    while (sp > 0) {
      int a = stack[--sp];
      for (int b = 0; b < nacc; ++b) {
        if (label[b] >= 0) continue;
        double dx = cen[2 * a] - cen[2 * b], dy = cen[2 * a + 1] - cen[2 * b + 1];
        if (sqrt(dx * dx + dy * dy) <= 3.0) {
          label[b] = nlab;
          stack[sp++] = b;
        }
      }
    }
    nlab++;
  }
  free(stack);
  int ok = 0;
  if (nlab > 1) {
    /* calculate_circle_centers :67-98: pair of cluster medians closest to 18.25 m apart */
    double *med = (double *)malloc(sizeof(double) * 2 * nlab);
    double *tmp = (double *)malloc(sizeof(double) * nacc);
    for (int l = 0; l < nlab; ++l)
      for (int d = 0; d < 2; ++d) {
        int cnt = 0;
        for (int i = 0; i < nacc; ++i)
          if (label[i] == l) tmp[cnt++] = cen[2 * i + d];
        med[2 * l + d] = median(tmp, cnt);
      }
    double best = 1000.0;
    int b0 = -1, b1 = -1;
    for (int a = 0; a < nlab; ++a)
      for (int b = a + 1; b < nlab; ++b) {
        double dx = med[2 * a] - med[2 * b], dy = med[2 * a + 1] - med[2 * b + 1];
        double dist = fabs(18.25 - sqrt(dx * dx + dy * dy));
        if (dist < best) {
          best = dist;
          b0 = a;
          b1 = b;
        }
      }
    if (!(best > 0.5)) {
      /* calculate_transformation :101-169 */
      double c[2][2] = {{med[2 * b0], med[2 * b0 + 1]}, {med[2 * b1], med[2 * b1 + 1]}};
      double yaw = atan2(orig_dir[1], orig_dir[0]);
      int right = -1, left = -1;
      for (int q = 0; q < 2; ++q) {
        double rx, ry;
        fsd_o_rotate(c[q][0] - orig_pos[0], c[q][1] - orig_pos[1], -yaw, &rx, &ry);
        if (ry < 0.0) {
          if (right < 0) right = q;
        } else {
          if (left < 0) left = q;
        }
      }
      if (right >= 0 && left >= 0) {
        const double *rr = ref_centers, *lr = ref_centers + 2;
        out->translation[0] = rr[0] - c[right][0];
        out->translation[1] = rr[1] - c[right][1];
        double ref_angle = atan2(lr[1] - rr[1], lr[0] - rr[0]);
        double calc_angle = atan2(c[left][1] - c[right][1], c[left][0] - c[right][0]);
        out->rotation = ref_angle - calc_angle;
        out->right_ref[0] = rr[0];
        out->right_ref[1] = rr[1];
        out->right_calc[0] = c[right][0];
        out->right_calc[1] = c[right][1];
        out->relocalized = 1;
        ok = 1;
      }
    }
    free(med);
    free(tmp);
  }
  free(label);
  free(cen);
  return ok;
}

static void to_known(const fsd_oracle_reloc *r, const double *p, double *out) {
  /* transform_pose :133-147 */
  double rx, ry;
  fsd_o_rotate(p[0] + r->translation[0] - r->right_ref[0], p[1] + r->translation[1] - r->right_ref[1], r->rotation, &rx,
               &ry);
  out[0] = rx + r->right_ref[0];
  out[1] = ry + r->right_ref[1];
}

static void to_original(const fsd_oracle_reloc *r, const double *p, double *out) {
  /* transform_back_to_original :149-163 */
  double rx, ry;
  fsd_o_rotate(p[0] - r->translation[0] - r->right_calc[0], p[1] - r->translation[1] - r->right_calc[1], -r->rotation,
               &rx, &ry);
  out[0] = rx + r->right_calc[0];
  out[1] = ry + r->right_calc[1];
}

int fsd_oracle_skidpad_step(const double *path, int n_path, int *index_along_path, const fsd_oracle_reloc *reloc,
                            const double *pos, const double *dir, int force_P, const double *prev_path,
                            double *path_internal, fsd_oracle_result *out) {
  memset(out, 0, sizeof(*out));
  double prev[FSD_O_HORIZON * 4];
  if (prev_path)
    memcpy(prev, prev_path, sizeof(prev));
  else
    fsd_oracle_initial_path(prev);
  double update[2 * 4096];
  int nu = 0;
  double p[2] = {pos[0], pos[1]}, d[2] = {dir[0], dir[1]};
  int relocalized = reloc && reloc->relocalized;
  if (relocalized) {
    /* full_pipeline.py:127-134 */
    double yaw = atan2(dir[1], dir[0]) + reloc->rotation;
    to_known(reloc, pos, p);
    d[0] = cos(yaw);
    d[1] = sin(yaw);
    /* SkidpadCalculatePath.fit_matches_as_spline, skidpad_calculate_path.py:49-71 */
    double mean = 0.0;
    for (int i = 0; i < 9; ++i) {
      double dx = path[2 * i + 2] - path[2 * i], dy = path[2 * i + 3] - path[2 * i + 1];
      mean += sqrt(dx * dx + dy * dy);
    }
    mean /= 9.0;
    int mac = (int)(20.0 / mean);
    int lo = *index_along_path - mac < 0 ? 0 : *index_along_path - mac;
    int hi = *index_along_path + mac > n_path ? n_path : *index_along_path + mac;
    int best = lo;
    double bd = INFINITY;
    for (int i = lo; i < hi; ++i) {
      double dx = p[0] - path[2 * i], dy = p[1] - path[2 * i + 1];
      double dd = sqrt(dx * dx + dy * dy);
      if (dd < bd) {
        bd = dd;
        best = i;
      }
    }
    *index_along_path = best;
    int fin = best + (int)(25.0 / mean);
    if (fin > n_path) fin = n_path;
    nu = fin - best;
    memcpy(update, path + 2 * best, sizeof(double) * 2 * nu);
  } else {
    /* calculate_trivial_path, core_calculate_path.py:127-134 */
    double chord[2 * FSD_O_HORIZON];
    fsd_o_almost_straight_path(chord);
    double yaw = atan2(dir[1], dir[0]);
    nu = FSD_O_HORIZON - 1;
    for (int i = 0; i < nu; ++i) {
      double rx, ry;
      fsd_o_rotate(chord[2 * (i + 1)], chord[2 * (i + 1) + 1], yaw, &rx, &ry);
      update[2 * i] = rx + pos[0];
      update[2 * i + 1] = ry + pos[1];
    }
  }
  fsd_o_path_from_update(update, nu, p, d, force_P, prev, out);
  if (path_internal) memcpy(path_internal, out->path, sizeof(out->path));
  if (relocalized)
    for (int i = 0; i < FSD_O_HORIZON; ++i) {
      double q[2] = {out->path[i][1], out->path[i][2]}, o[2];
      to_original(reloc, q, o);
      out->path[i][1] = o[0];
      out->path[i][2] = o[1];
    }
  return 0;
}
