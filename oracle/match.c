/*
 * ORACLE (test infrastructure).  Cone matching: M1-M6 of SURVEY.md section 8(a).
 * Restates /root/reference/fsd_path_planning/cone_matching/functional_cone_matching.py and
 * match_directions.py with the parameters of core_cone_matching.py:101-117 / config.py:124-129, 162:
 * min_track_width 3, major radius 7.5, minor radius 3, max search angle 50 deg, non-monotonic.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fsd_oracle.h"
#include "oracle_internal.h"

#define MAXC 64

static void match_directions(const double *c, int n, int side, double *out) {
  /* calculate_match_search_direction, match_directions.py:23-44 */
  for (int i = 0; i < n; ++i) {
    int a = i == 0 ? 0 : (i == n - 1 ? n - 2 : i - 1);
    int b = i == 0 ? 1 : (i == n - 1 ? n - 1 : i + 1);
    double tx = c[2 * b] - c[2 * a], ty = c[2 * b + 1] - c[2 * a + 1], rx, ry;
    fsd_o_rotate(tx, ty, side == FSD_O_RIGHT ? M_PI / 2.0 : -M_PI / 2.0, &rx, &ry);
    double nrm = sqrt(rx * rx + ry * ry);
    out[2 * i] = rx / nrm;
    out[2 * i + 1] = ry / nrm;
  }
}

/* calculate_matches_for_side :340-384 (find_boolean_mask_of_all_potential_matches :73-144 and
 * select_best_match_candidate :147-175 inlined).  Returns 0, or 1 when the reference raises. */
static int matches_for_side(const double *cones, int n, int side, const double *other, int m, int *match,
                            double *dirs) {
  const double major = 5.0 * 1.5, minor = 3.0, max_angle = 50.0 * M_PI / 180.0;
  if (n <= 1) {
    for (int i = 0; i < n; ++i) match[i] = -1;
    return 0;
  }
  match_directions(cones, n, side, dirs);
  double odirs[2 * MAXC];
  int have_odirs = m > 1;
  if (have_odirs) match_directions(other, m, side == FSD_O_RIGHT ? FSD_O_LEFT : FSD_O_RIGHT, odirs);
  if (m == 0) {
    for (int i = 0; i < n; ++i) match[i] = -1;
    return 0;
  }
  if (!have_odirs) {
    /* m == 1: the direction array of the other side is (0, 2) and the boolean index at
     * functional_cone_matching.py:130 has the wrong length -> IndexError in the reference */
    for (int i = 0; i < n; ++i) match[i] = -1;
    return 1;
  }
  for (int i = 0; i < n; ++i) {
    double ang = atan2(dirs[2 * i + 1], dirs[2 * i]);
    int any = 0, best = 0;
    double best_d = 0.0;
    for (int j = 0; j < m; ++j) {
      double vx = other[2 * j] - cones[2 * i], vy = other[2 * j + 1] - cones[2 * i + 1], rx, ry;
      fsd_o_rotate(vx, vy, -ang, &rx, &ry);
      int ok = (rx * rx / (major * major) + ry * ry / (minor * minor)) < 1.0;
      double a = atan2(ry, rx);
      if (fabs(a / 2.0) > max_angle) ok = 0;
      double dd = fsd_o_angle_between(dirs[2 * i], dirs[2 * i + 1], odirs[2 * j], odirs[2 * j + 1]);
      if (dd < M_PI / 2.0) ok = 0;
      any |= ok;
      /* the match itself is the argmin over ALL cones of the other side (:162) */
      double d2 = fsd_o_cdist_sq(cones[2 * i], cones[2 * i + 1], other[2 * j], other[2 * j + 1]);
      if (j == 0 || d2 < best_d) {
        best_d = d2;
        best = j;
      }
    }
    match[i] = any ? best : -1;
  }
  return 0;
}

/* insert_virtual_cones_to_existing :195-261 */
static int insert_virtual(const double *other, int no, const double *virt, int nv, const double *car, double *out) {
  double ex[2 * MAXC], ins[2 * MAXC];
  int ne, ni;
  if (no > nv) {
    memcpy(ex, other, sizeof(double) * 2 * no);
    ne = no;
    memcpy(ins, virt, sizeof(double) * 2 * nv);
    ni = nv;
  } else {
    memcpy(ex, virt, sizeof(double) * 2 * nv);
    ne = nv;
    memcpy(ins, other, sizeof(double) * 2 * no);
    ni = no;
  }
  /* order of insertion: ascending distance to the nearest existing cone (:212) */
  double key[MAXC];
  int order[MAXC];
  for (int i = 0; i < ni; ++i) {
    double mn = 0.0;
    for (int j = 0; j < ne; ++j) {
      double d2 = fsd_o_cdist_sq(ins[2 * i], ins[2 * i + 1], ex[2 * j], ex[2 * j + 1]);
      if (j == 0 || d2 < mn) mn = d2;
    }
    key[i] = mn;
    order[i] = i;
  }
  for (int a = 1; a < ni; ++a) {
    int v = order[a], b = a - 1;
    while (b >= 0 && key[order[b]] > key[v]) {
      order[b + 1] = order[b];
      b--;
    }
    order[b + 1] = v;
  }
  for (int oi = 0; oi < ni; ++oi) {
    double cx = ins[2 * order[oi]], cy = ins[2 * order[oi] + 1];
    int index;
    if (ne == 1) {
      /* calculate_insert_index_for_one_cone :264-282 */
      double dv = sqrt((cx - car[0]) * (cx - car[0]) + (cy - car[1]) * (cy - car[1]));
      double de = sqrt((ex[0] - car[0]) * (ex[0] - car[0]) + (ex[1] - car[1]) * (ex[1] - car[1]));
      index = dv < de ? 0 : 1;
    } else {
      int c1 = -1, c2 = -1;
      double d1 = 0.0, d2 = 0.0;
      for (int j = 0; j < ne; ++j) {
        double dx = ex[2 * j] - cx, dy = ex[2 * j + 1] - cy;
        double d = sqrt(dx * dx + dy * dy);
        if (c1 < 0 || d < d1) {
          c2 = c1;
          d2 = d1;
          c1 = j;
          d1 = d;
        } else if (c2 < 0 || d < d2) {
          c2 = j;
          d2 = d;
        }
      }
      if (abs(c1 - c2) != 1) continue; /* :226-227 virtual cone skipped */
      double a = fsd_o_angle_between(ex[2 * c1] - cx, ex[2 * c1 + 1] - cy, ex[2 * c2] - cx, ex[2 * c2 + 1] - cy);
      if (a > M_PI / 2.0)
        index = (c1 < c2 ? c1 : c2) + 1; /* calculate_insert_index_of_new_cone :285-303 */
      else
        index = c1 < c2 ? c1 : c1 + 1;
    }
    if (ne >= MAXC) continue;
    for (int j = ne; j > index; --j) {
      ex[2 * j] = ex[2 * (j - 1)];
      ex[2 * j + 1] = ex[2 * (j - 1) + 1];
    }
    ex[2 * index] = cx;
    ex[2 * index + 1] = cy;
    ne++;
  }
  /* drop interior points whose polyline angle is below 85 deg (:252-259) */
  char drop[MAXC];
  memset(drop, 0, sizeof(drop));
  for (int i = 1; i + 1 < ne; ++i) {
    double a = fsd_o_angle_between(ex[2 * (i + 1)] - ex[2 * i], ex[2 * (i + 1) + 1] - ex[2 * i + 1],
                                   ex[2 * (i - 1)] - ex[2 * i], ex[2 * (i - 1) + 1] - ex[2 * i + 1]);
    drop[i] = (char)(a < 85.0 * M_PI / 180.0);
  }
  int w = 0;
  for (int i = 0; i < ne; ++i)
    if (!drop[i]) {
      out[2 * w] = ex[2 * i];
      out[2 * w + 1] = ex[2 * i + 1];
      w++;
    }
  return w;
}

/* calculate_cones_for_other_side :387-440: cones of `side` produce the other side with virtual cones */
static int cones_for_other_side(const double *cones, int n, int side, const double *other, int m, const double *car,
                                double *out, int *raises) {
  int match[MAXC];
  double dirs[2 * MAXC], virt[2 * MAXC];
  *raises |= matches_for_side(cones, n, side, other, m, match, dirs);
  int nv = 0;
  for (int i = 0; i < n; ++i)
    if (match[i] == -1) {
      /* calculate_positions_of_virtual_cones :178-192, min_track_width = 3 */
      virt[2 * nv] = cones[2 * i] + dirs[2 * i] * 3.0;
      virt[2 * nv + 1] = cones[2 * i + 1] + dirs[2 * i + 1] * 3.0;
      nv++;
    }
  int no;
  /* combine_and_sort_virtual_with_real :306-337 */
  if (m == 0) {
    memcpy(out, virt, sizeof(double) * 2 * nv);
    no = nv;
  } else if (nv == 0) {
    memcpy(out, other, sizeof(double) * 2 * m);
    no = m;
  } else {
    no = insert_virtual(other, m, virt, nv, car, out);
  }
  if (no < 2) {
    memcpy(out, other, sizeof(double) * 2 * m);
    no = m;
  }
  return no;
}

int fsd_o_match(const double *left_in, int nl, const double *right_in, int nr, const double *pos, const double *dir,
                fsd_oracle_result *out) {
  /* calculate_virtual_cones_for_both_sides :479-588 */
  (void)dir;
  double left[2 * MAXC], right[2 * MAXC];
  memcpy(left, left_in, sizeof(double) * 2 * nl);
  memcpy(right, right_in, sizeof(double) * 2 * nr);
  out->n_left_wv = out->n_right_wv = 0;
  if (nl < 2 && nr < 2) return 0;
  int mn = nl < nr ? nl : nr, mx = nl < nr ? nr : nl;
  int discard = mn == 0 || ((double)mx / (double)mn > 2.0);
  if (discard) {
    if (nl < nr)
      nl = 0;
    else
      nr = 0;
  }
  int raises = 0;
  double lwv[2 * MAXC], rwv[2 * MAXC];
  int nrw, nlw;
  if (nl >= 2)
    nrw = cones_for_other_side(left, nl, FSD_O_LEFT, right, nr, pos, rwv, &raises);
  else {
    memcpy(rwv, right, sizeof(double) * 2 * nr);
    nrw = nr;
  }
  if (nr >= 2)
    nlw = cones_for_other_side(right, nr, FSD_O_RIGHT, left, nl, pos, lwv, &raises);
  else {
    memcpy(lwv, left, sizeof(double) * 2 * nl);
    nlw = nl;
  }
  /* match_both_sides_with_virtual_cones :443-476 */
  int l2r[MAXC], r2l[MAXC];
  double dirs[2 * MAXC];
  raises |= matches_for_side(lwv, nlw, FSD_O_LEFT, rwv, nrw, l2r, dirs);
  raises |= matches_for_side(rwv, nrw, FSD_O_RIGHT, lwv, nlw, r2l, dirs);
  if (nlw > FSD_O_MAX_WV || nrw > FSD_O_MAX_WV) {
    out->status |= FSD_O_OVERFLOW;
    if (nlw > FSD_O_MAX_WV) nlw = FSD_O_MAX_WV;
    if (nrw > FSD_O_MAX_WV) nrw = FSD_O_MAX_WV;
  }
  out->n_left_wv = nlw;
  out->n_right_wv = nrw;
  for (int i = 0; i < nlw; ++i) {
    out->left_wv[i][0] = lwv[2 * i];
    out->left_wv[i][1] = lwv[2 * i + 1];
    out->l2r[i] = l2r[i];
  }
  for (int i = 0; i < nrw; ++i) {
    out->right_wv[i][0] = rwv[2 * i];
    out->right_wv[i][1] = rwv[2 * i + 1];
    out->r2l[i] = r2l[i];
  }
  if (raises) out->status |= FSD_O_REF_RAISES;
  return 0;
}
