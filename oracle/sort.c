/*
 * ORACLE (test infrastructure).  Cone sorting: S1-S8 of SURVEY.md section 8(a).
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/fsd_path_planning/).  Scalar fp64, literal control flow.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fsd_oracle.h"
#include "oracle_internal.h"

/* ---- shared math (utils/math_utils.py) ---------------------------------------------- */

double fsd_o_angle_between(double ax, double ay, double bx, double by) {
  /* vec_angle_between, math_utils.py:70-100 (cos clipped to [-1, 1]) */
  double c = (ax * bx + ay * by) / (sqrt(ax * ax + ay * ay) * sqrt(bx * bx + by * by));
  if (c < -1.0) c = -1.0;
  if (c > 1.0) c = 1.0;
  return acos(c);
}

double fsd_o_angle_difference(double a1, double a2) {
  /* angle_difference, math_utils.py:663-676: (a1 - a2 + 3pi) % 2pi - pi, Python modulo */
  double v = fmod(a1 - a2 + 3.0 * M_PI, 2.0 * M_PI);
  if (v < 0.0) v += 2.0 * M_PI;
  return v - M_PI;
}

double fsd_o_sign(double v) { return (double)((v > 0.0) - (v < 0.0)); }

void fsd_o_rotate(double px, double py, double theta, double *rx, double *ry) {
  /* rotate, math_utils.py:103-117: points @ [[c, -s], [s, c]].T */
  double c = cos(theta), s = sin(theta);
  *rx = px * c - py * s;
  *ry = px * s + py * c;
}

double fsd_o_cdist_sq(double xi, double yi, double xj, double yj) {
  /* my_cdist_sq_euclidean, math_utils.py:120-150: [1,1,x,y,x^2,y^2] . [x'^2,y'^2,-2x',-2y',1,1] */
  double acc = 1.0 * (xj * xj);
  acc += 1.0 * (yj * yj);
  acc += xi * (-2.0 * xj);
  acc += yi * (-2.0 * yj);
  acc += (xi * xi) * 1.0;
  acc += (yi * yi) * 1.0;
  return acc;
}

int fsd_o_inside_ellipse(double px, double py, double cx, double cy, double dx, double dy, double major,
                         double minor) {
  /* points_inside_ellipse, math_utils.py:493-530 */
  double ang = atan2(dy, dx), rx, ry;
  fsd_o_rotate(px - cx, py - cy, -ang, &rx, &ry);
  double crit = rx * rx / (major * major) + ry * ry / (minor * minor);
  return crit < 1.0;
}

/* lines_segments_intersect_indicator, sorting_cones/trace_sorter/line_segment_intersection.py:136-200 */
static int parallel_case(const double *a0, const double *a1, const double *b0, const double *b1, double eps) {
  /* :34-72 -- note `difference[0] < epsilon` without abs */
  double dx = a1[0] - a0[0], dy = a1[1] - a0[1];
  int maybe_overlap;
  double slope;
  if (dx < eps) {
    maybe_overlap = fabs(a0[0] - b0[0]) < eps;
    slope = INFINITY;
  } else {
    slope = dy / dx;
    double ia = a0[1] - slope * a0[0];
    double ib = b0[1] - slope * b0[0];
    maybe_overlap = fabs(ia - ib) < eps;
  }
  if (!maybe_overlap) return 0;
  int ax = slope > 1.0 ? 1 : 0;
  double left_end, right_start;
  if (a0[ax] < b0[ax]) {
    left_end = a1[ax];
    right_start = b0[ax] < b1[ax] ? b0[ax] : b1[ax];
  } else {
    left_end = b1[ax];
    right_start = a0[ax] < a1[ax] ? a0[ax] : a1[ax];
  }
  return left_end >= right_start;
}

int fsd_o_segments_intersect(const double *a0, const double *a1, const double *b0, const double *b1) {
  const double eps = 1e-6;
  /* homogeneous lines: cross((x0,y0,1),(x1,y1,1)) */
  double la[3] = {a0[1] - a1[1], a1[0] - a0[0], a0[0] * a1[1] - a0[1] * a1[0]};
  double lb[3] = {b0[1] - b1[1], b1[0] - b0[0], b0[0] * b1[1] - b0[1] * b1[0]};
  double ix = la[1] * lb[2] - la[2] * lb[1];
  double iy = la[2] * lb[0] - la[0] * lb[2];
  double iz = la[0] * lb[1] - la[1] * lb[0];
  if (fabs(iz) < eps) return parallel_case(a0, a1, b0, b1, eps);
  double x = ix / iz, y = iy / iz;
  double al = fmin(a0[0], a1[0]), ar = fmax(a0[0], a1[0]);
  double bl = fmin(b0[0], b1[0]), br = fmax(b0[0], b1[0]);
  double ab = fmin(a0[1], a1[1]), at = fmax(a0[1], a1[1]);
  double bb = fmin(b0[1], b1[1]), bt = fmax(b0[1], b1[1]);
  return (al - eps <= x && x <= ar + eps) && (bl - eps <= x && x <= br + eps) && (ab - eps <= y && y <= at + eps) &&
         (bb - eps <= y && y <= bt + eps);
}

/* ---- S2: starting cones (sorting_cones/trace_sorter/core_trace_sorter.py:344-465) ---- */

static void mask_first(const fsd_o_frame *f, int side, double *dist, char *valid) {
  /* mask_cone_can_be_first_in_config :379-407 */
  const double max_dist_to_first = 6.0;
  int opp = side == FSD_O_LEFT ? FSD_O_RIGHT : FSD_O_LEFT;
  double ang = atan2(f->dir[1], f->dir[0]);
  for (int i = 0; i < f->n; ++i) {
    double rx, ry;
    fsd_o_rotate(f->xy[2 * i] - f->pos[0], f->xy[2 * i + 1] - f->pos[1], -ang, &rx, &ry);
    double a = atan2(ry, rx);
    dist[i] = sqrt(rx * rx + ry * ry);
    double major = max_dist_to_first * 1.5, minor = max_dist_to_first / 1.5;
    int in_ellipse = (rx * rx / (major * major) + ry * ry / (minor * minor)) < 1.0;
    int valid_side = fsd_o_sign(a) == (side == FSD_O_LEFT ? 1.0 : -1.0);
    int ang_ok = fabs(a) < M_PI - M_PI / 5.0;
    int ang_min = fabs(a) > M_PI / 10.0;
    int right_color = f->type[i] == side;
    int mask_side = (valid_side && ang_ok && ang_min) || right_color;
    int not_opp = f->type[i] != opp;
    valid[i] = (char)(in_ellipse && mask_side && not_opp);
  }
}

static int select_start(const fsd_o_frame *f, const double *dist, const char *valid, const char *skip) {
  /* select_starting_cone :344-377 */
  int best = -1;
  for (int i = 0; i < f->n; ++i) {
    if (!valid[i] || (skip && skip[i])) continue;
    if (best < 0 || dist[i] < dist[best]) best = i;
  }
  if (best < 0) return -1;
  if (dist[best] > 6.0) return -1;
  return best;
}

static int select_first_k(const fsd_o_frame *f, int side, int *first_k) {
  /* select_first_k_starting_cones :409-465; returns count (0 = None) */
  int n = f->n;
  double *dist = (double *)malloc(sizeof(double) * n);
  char *valid = (char *)malloc(n), *skip = (char *)malloc(n);
  mask_first(f, side, dist, valid);
  int i1 = select_start(f, dist, valid, NULL);
  int count = 0;
  if (i1 >= 0) {
    for (int i = 0; i < n; ++i) {
      double a = fsd_o_angle_between(f->xy[2 * i] - f->pos[0], f->xy[2 * i + 1] - f->pos[1], f->dir[0], f->dir[1]);
      skip[i] = (char)(fabs(a) < M_PI / 2.0);
    }
    skip[i1] = 1;
    int i2 = select_start(f, dist, valid, skip);
    if (i2 < 0) {
      first_k[0] = i1;
      count = 1;
    } else {
      double d1x = f->xy[2 * i1] - f->xy[2 * i2], d1y = f->xy[2 * i1 + 1] - f->xy[2 * i2 + 1];
      double a1 = fsd_o_angle_between(d1x, d1y, f->dir[0], f->dir[1]);
      double a2 = fsd_o_angle_between(-d1x, -d1y, f->dir[0], f->dir[1]);
      if (a1 > a2) {
        int tmp = i1;
        i1 = i2;
        i2 = tmp;
      }
      double d = sqrt(d1x * d1x + d1y * d1y);
      if (d > 6.5 * 1.1 || d < 1.4) {
        first_k[0] = i1;
        count = 1;
      } else {
        first_k[0] = i2;
        first_k[1] = i1;
        count = 2;
      }
    }
  }
  free(dist);
  free(valid);
  free(skip);
  return count;
}

/* ---- S3: adjacency (sorting_cones/trace_sorter/adjacency_matrix.py:60-128) ------------- */

typedef struct {
  int *nbr;  /* n x 5 neighbour lists, ascending index */
  int *deg;  /* n */
  int reachable;
} fsd_o_graph;

static void build_graph(const fsd_o_frame *f, int side, int start, fsd_o_graph *g) {
  int n = f->n, k = n - 1 < 5 ? n - 1 : 5;
  int opp = side == FSD_O_LEFT ? FSD_O_RIGHT : FSD_O_LEFT;
  const double max_d2 = 6.5 * 6.5;
  int *knn = (int *)malloc(sizeof(int) * n * 5);
  int *kcnt = (int *)calloc(n, sizeof(int));
  double best[5];
  for (int i = 0; i < n; ++i) {
    int cnt = 0;
    if (f->type[i] != opp) {
      for (int j = 0; j < n; ++j) {
        if (j == i || f->type[j] == opp) continue;
        double d2 = fsd_o_cdist_sq(f->xy[2 * i], f->xy[2 * i + 1], f->xy[2 * j], f->xy[2 * j + 1]);
        /* the k smallest of the row; entries above max_dist^2 are cut afterwards (:102-107) */
        int pos = cnt;
        while (pos > 0 && d2 < best[pos - 1]) pos--;
        if (pos >= k) continue;
        int last = cnt < k ? cnt : k - 1;
        for (int q = last; q > pos; --q) {
          best[q] = best[q - 1];
          knn[i * 5 + q] = knn[i * 5 + q - 1];
        }
        best[pos] = d2;
        knn[i * 5 + pos] = j;
        if (cnt < k) cnt++;
      }
      int w = 0;
      for (int q = 0; q < cnt; ++q)
        if (!(best[q] > max_d2)) knn[i * 5 + w++] = knn[i * 5 + q];
      cnt = w;
    }
    kcnt[i] = cnt;
  }
  /* undirected: keep edges present in both directions (:110); CSR order = ascending (:49) */
  for (int i = 0; i < n; ++i) {
    int d = 0, tmp[5];
    for (int q = 0; q < kcnt[i]; ++q) {
      int j = knn[i * 5 + q], back = 0;
      for (int r = 0; r < kcnt[j]; ++r)
        if (knn[j * 5 + r] == i) back = 1;
      if (back) tmp[d++] = j;
    }
    for (int a = 1; a < d; ++a) {
      int v = tmp[a], b = a - 1;
      while (b >= 0 && tmp[b] > v) {
        tmp[b + 1] = tmp[b];
        b--;
      }
      tmp[b + 1] = v;
    }
    for (int q = 0; q < d; ++q) g->nbr[i * 5 + q] = tmp[q];
    g->deg[i] = d;
  }
  /* breadth_first_order, sorting_cones/trace_sorter/common.py:36-67: only the count is used */
  char *vis = (char *)calloc(n, 1);
  int *queue = (int *)malloc(sizeof(int) * n);
  int head = 0, tail = 0;
  queue[tail++] = start;
  vis[start] = 1;
  while (head < tail) {
    int node = queue[head++];
    for (int q = 0; q < g->deg[node]; ++q) {
      int j = g->nbr[node * 5 + q];
      if (!vis[j]) {
        vis[j] = 1;
        queue[tail++] = j;
      }
    }
  }
  g->reachable = tail;
  free(vis);
  free(queue);
  free(knn);
  free(kcnt);
}

/* stage wrapper for the adjacency parity test: neighbour lists (n x 5, ascending) and degrees of one side's graph */
int fsd_oracle_adjacency(const double *cones_xy, const unsigned char *cones_type, int n, int side, int *nbr, int *deg) {
  if (n < 1) return 0;
  fsd_o_frame f = {cones_xy, cones_type, n, {0.0, 0.0}, {1.0, 0.0}};
  fsd_o_graph g = {nbr, deg, 0};
  build_graph(&f, side, 0, &g);
  return g.reachable;
}

/* ---- S5: admissibility (sorting_cones/trace_sorter/end_configurations.py:108-278) ------ */

static void can_be_added(const fsd_o_frame *f, int side, const int *attempt, int pos, const int *nb, int nnb,
                         char *can) {
  const double thr_dir = 40.0 * M_PI / 180.0, thr_abs = 65.0 * M_PI / 180.0, car_size = 2.1;
  const double *xy = f->xy;
  double dn = sqrt(f->dir[0] * f->dir[0] + f->dir[1] * f->dir[1]);
  double dnx = f->dir[0] / dn, dny = f->dir[1] / dn;
  int last = attempt[pos];
  for (int i = 0; i < nnb; ++i) {
    can[i] = 1;
    for (int q = 0; q <= pos; ++q)
      if (attempt[q] == nb[i]) can[i] = 0; /* my_in1d :126 */
  }
  if (pos >= 1) {
    /* calculate_mask_within_ellipse :281-300 */
    int prev = attempt[pos - 1];
    double ddx = xy[2 * last] - xy[2 * prev], ddy = xy[2 * last + 1] - xy[2 * prev + 1];
    for (int i = 0; i < nnb; ++i)
      if (!fsd_o_inside_ellipse(xy[2 * nb[i]], xy[2 * nb[i] + 1], xy[2 * last], xy[2 * last + 1], ddx, ddy, 6.0, 3.0))
        can[i] = 0;
  }
  if (pos == 0) {
    /* mask_second_in_attempt_is_on_right_vehicle_side :260-278 */
    double ang_car = atan2(dny, dnx);
    for (int i = 0; i < nnb; ++i) {
      double a = atan2(xy[2 * nb[i] + 1] - f->pos[1], xy[2 * nb[i]] - f->pos[0]);
      double diff = fsd_o_angle_difference(a, ang_car);
      double expected = side == FSD_O_LEFT ? 1.0 : -1.0;
      int ok = (fsd_o_sign(diff) == expected) || (fabs(diff) < 5.0 * M_PI / 180.0);
      if (!ok) can[i] = 0;
    }
  }
  for (int i = 0; i < nnb; ++i) {
    if (!can[i]) continue;
    int cand = nb[i];
    /* check_if_neighbor_lies_between_last_in_attempt_and_candidate :226-257 */
    for (int q = 0; q < nnb; ++q) {
      int nbq = nb[q];
      if (nbq == cand) continue;
      double v1x = xy[2 * last] - xy[2 * nbq], v1y = xy[2 * last + 1] - xy[2 * nbq + 1];
      double v2x = xy[2 * cand] - xy[2 * nbq], v2y = xy[2 * cand + 1] - xy[2 * nbq + 1];
      double d_c = sqrt(v2x * v2x + v2y * v2y), d_l = sqrt(v1x * v1x + v1y * v1y);
      if (d_c < 6.0 && d_l < 6.0 && fsd_o_angle_between(v1x, v1y, v2x, v2y) > 150.0 * M_PI / 180.0) {
        can[i] = 0;
        break;
      }
    }
    double cx = xy[2 * cand], cy = xy[2 * cand + 1];
    if (can[i] && pos >= 1) {
      int prev = attempt[pos - 1];
      double ax = xy[2 * last] - xy[2 * prev], ay = xy[2 * last + 1] - xy[2 * prev + 1];
      double bx = cx - xy[2 * last], by = cy - xy[2 * last + 1];
      double angle_1 = atan2(ay, ax), angle_2 = atan2(by, bx);
      double difference = fsd_o_angle_difference(angle_2, angle_1);
      double len = sqrt(bx * bx + by * by);
      if (fabs(difference) > thr_abs)
        can[i] = 0;
      else if (side == FSD_O_LEFT)
        can[i] = (char)(difference < thr_dir || len < 4.0);
      else
        can[i] = (char)(difference > -thr_dir || len < 4.0);
      if (pos >= 2) {
        int pp = attempt[pos - 2];
        double zx = xy[2 * prev] - xy[2 * pp], zy = xy[2 * prev + 1] - xy[2 * pp + 1];
        double angle_3 = atan2(zy, zx);
        double difference_2 = fsd_o_angle_difference(angle_1, angle_3);
        if (fsd_o_sign(difference) != fsd_o_sign(difference_2) && fabs(difference - difference_2) > 1.3) can[i] = 0;
      }
    }
    if (can[i] && pos == 1) {
      int start = attempt[0];
      double off = fsd_o_angle_between(f->dir[0], f->dir[1], cx - xy[2 * start], cy - xy[2 * start + 1]);
      if (!(off < M_PI / 2.0)) can[i] = 0;
    }
    if (can[i]) {
      double cs[2] = {f->pos[0] - dnx * car_size / 2.0, f->pos[1] - dny * car_size / 2.0};
      double ce[2] = {f->pos[0] + dnx * car_size, f->pos[1] + dny * car_size};
      double a0[2] = {xy[2 * last], xy[2 * last + 1]}, a1[2] = {cx, cy};
      if (fsd_o_segments_intersect(a0, a1, cs, ce)) can[i] = 0;
    }
  }
}

/* ---- S4: exhaustive DFS + post-filter (end_configurations.py:320-520) ------------------- */

typedef struct {
  int *rows; /* count x L */
  int count, cap, L;
} cfg_list;

static void cfg_push(cfg_list *c, const int *row) {
  if (c->count == c->cap) {
    c->cap = c->cap ? c->cap * 2 : 16;
    c->rows = (int *)realloc(c->rows, sizeof(int) * c->cap * c->L);
  }
  memcpy(c->rows + (size_t)c->count * c->L, row, sizeof(int) * c->L);
  c->count++;
}

static int row_len(const int *row, int L) {
  int n = 0;
  for (int i = 0; i < L; ++i) n += row[i] != -1;
  return n;
}

static int row_cmp(const int *a, const int *b, int L) {
  for (int i = 0; i < L; ++i)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}

static long find_end_configurations(const fsd_o_frame *f, int side, const fsd_o_graph *g, const int *first_k, int nfk,
                                    int L, cfg_list *out, unsigned *status) {
  const long MAX_POPS = 1L << 22;
  int attempt[FSD_O_MAX_SORTED + 2];
  for (int i = 0; i < L + 2 && i < FSD_O_MAX_SORTED + 2; ++i) attempt[i] = -1;
  int cap = 64, sp = 0;
  int *stack = (int *)malloc(sizeof(int) * 2 * cap);
  if (nfk > 1) {
    attempt[0] = first_k[0];
    stack[0] = first_k[1];
    stack[1] = 1;
  } else {
    stack[0] = first_k[0];
    stack[1] = 0;
  }
  cfg_list raw = {0, 0, 0, L};
  long pops = 0;
  char can[5];
  while (sp >= 0) {
    if (++pops > MAX_POPS) {
      *status |= FSD_O_OVERFLOW;
      break;
    }
    int node = stack[2 * sp], pos = stack[2 * sp + 1];
    sp--;
    if (pos >= L) {
      /* target_length < 2 with two seeds: the reference writes out of bounds here (:369);
       * such a search cannot yield >= 3 cones, so the side ends with no configuration. */
      continue;
    }
    attempt[pos] = node;
    for (int q = pos + 1; q < L; ++q) attempt[q] = -1;
    const int *nb = g->nbr + node * 5;
    int nnb = g->deg[node];
    can_be_added(f, side, attempt, pos, nb, nnb, can);
    int any = 0;
    for (int i = 0; i < nnb; ++i) any |= can[i];
    if (pos < L - 1 && any) {
      for (int i = 0; i < nnb; ++i) {
        if (!can[i]) continue;
        sp++;
        if (sp >= cap) {
          cap *= 2;
          stack = (int *)realloc(stack, sizeof(int) * 2 * cap);
        }
        stack[2 * sp] = nb[i];
        stack[2 * sp + 1] = pos + 1;
      }
    } else {
      cfg_push(&raw, attempt);
    }
  }
  free(stack);
  /* post-filter :484-518 */
  cfg_list kept = {0, 0, 0, L};
  for (int r = 0; r < raw.count; ++r) {
    int *row = raw.rows + (size_t)r * L;
    if (row_len(row, L) <= 2) continue; /* :420 */
    int ok = 1;
    if (nfk > 1)
      for (int q = 0; q < nfk; ++q) ok &= row[q] == first_k[q];
    if (!ok) continue;
    int n = row_len(row, L);
    if (f->type[row[n - 1]] != side) row[n - 1] = -1; /* :492-500 */
    if (row_len(row, L) < 3) continue;
    cfg_push(&kept, row);
  }
  /* np.unique(axis=0): lexicographic sort + dedup */
  for (int a = 1; a < kept.count; ++a) {
    int tmp[FSD_O_MAX_SORTED];
    memcpy(tmp, kept.rows + (size_t)a * L, sizeof(int) * L);
    int b = a - 1;
    while (b >= 0 && row_cmp(kept.rows + (size_t)b * L, tmp, L) > 0) {
      memcpy(kept.rows + (size_t)(b + 1) * L, kept.rows + (size_t)b * L, sizeof(int) * L);
      b--;
    }
    memcpy(kept.rows + (size_t)(b + 1) * L, tmp, sizeof(int) * L);
  }
  cfg_list uniq = {0, 0, 0, L};
  for (int r = 0; r < kept.count; ++r)
    if (r == 0 || row_cmp(kept.rows + (size_t)r * L, kept.rows + (size_t)(r - 1) * L, L) != 0)
      cfg_push(&uniq, kept.rows + (size_t)r * L);
  /* drop rows that are a strict prefix of another row (:509-515) */
  for (int j = 0; j < uniq.count; ++j) {
    const int *rj = uniq.rows + (size_t)j * L;
    int covered = 0;
    for (int i = 0; i < uniq.count; ++i) {
      const int *ri = uniq.rows + (size_t)i * L;
      int all = 1;
      for (int q = 0; q < L; ++q) all &= (ri[q] == rj[q]) || (rj[q] == -1);
      covered += all;
    }
    if (!(covered > 1)) cfg_push(out, rj);
  }
  free(raw.rows);
  free(kept.rows);
  free(uniq.rows);
  return pops;
}

/* ---- S6/S7: cost (cost_function.py, cone_distance_cost.py, nearby_cone_search.py) -------- */

static void search_direction(const double *xy, int a, int b, int side, double *ox, double *oy) {
  /* calculate_search_direction_for_one, cone_matching/match_directions.py:7-20 */
  double tx = xy[2 * b] - xy[2 * a], ty = xy[2 * b + 1] - xy[2 * a + 1];
  double rx, ry;
  fsd_o_rotate(tx, ty, side == FSD_O_RIGHT ? M_PI / 2.0 : -M_PI / 2.0, &rx, &ry);
  double nrm = sqrt(rx * rx + ry * ry);
  *ox = rx / nrm;
  *oy = ry / nrm;
}

static int lower_bound(const int *a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

static void cones_on_either_side(const fsd_o_frame *f, int side, const cfg_list *cfgs, int *n_good, int *n_bad) {
  /* _impl_number_cones_on_each_side_for_each_config, nearby_cone_search.py:212-297 */
  int n = f->n, L = cfgs->L;
  const double *xy = f->xy;
  const double range2 = 6.0 * 6.0, half_angle = (M_PI / 1.5) / 2.0;
  char *in_cfg = (char *)calloc(n, 1);
  for (int r = 0; r < cfgs->count; ++r)
    for (int q = 0; q < L; ++q) {
      int v = cfgs->rows[(size_t)r * L + q];
      if (v != -1) in_cfg[v] = 1;
    }
  int *idxs = (int *)malloc(sizeof(int) * n), nidx = 0;
  for (int i = 0; i < n; ++i)
    if (in_cfg[i]) idxs[nidx++] = i;
  /* find_nearby_cones_for_idxs :97-103 */
  int *all = (int *)malloc(sizeof(int) * n), nall = 0;
  for (int j = 0; j < n; ++j) {
    int near = 0;
    for (int q = 0; q < nidx && !near; ++q) {
      int i = idxs[q];
      if (i == j) continue; /* diagonal = 1e7 */
      near = fsd_o_cdist_sq(xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]) < range2;
    }
    if (near) all[nall++] = j;
  }
  /* sorted_set_diff :88-94 -- mask[searchsorted(a, b)] = False with no membership test */
  char *keep = (char *)malloc(nall + 1);
  memset(keep, 1, nall + 1);
  for (int q = 0; q < nidx; ++q) {
    int p = lower_bound(all, nall, idxs[q]);
    if (p < nall) keep[p] = 0; /* p == nall: out-of-bounds write in the reference, no visible effect */
  }
  int *close = (int *)malloc(sizeof(int) * (nall + 1)), nclose = 0;
  for (int q = 0; q < nall; ++q)
    if (keep[q]) close[nclose++] = all[q];
  int *other = (int *)malloc(sizeof(int) * (nall + nidx + 1));
  for (int r = 0; r < cfgs->count; ++r) {
    const int *c = cfgs->rows + (size_t)r * L;
    int len = row_len(c, L);
    int nother = 0;
    for (int q = 0; q < nclose; ++q) other[nother++] = close[q];
    for (int q = 0; q < nidx; ++q) {
      int member = 0;
      for (int w = 0; w < len; ++w) member |= c[w] == idxs[q];
      if (!member) other[nother++] = idxs[q];
    }
    int good = 0, bad = 0;
    for (int j = 0; j < len; ++j) {
      double sx, sy;
      if (j == 0)
        search_direction(xy, c[0], c[1], side, &sx, &sy);
      else if (j == len - 1)
        search_direction(xy, c[j - 1], c[j], side, &sx, &sy);
      else
        search_direction(xy, c[j - 1], c[j + 1], side, &sx, &sy);
      int cj = c[j];
      for (int q = 0; q < nother; ++q) {
        int o = other[q];
        if (o == cj) continue; /* diagonal of the distance matrix is 1e7 */
        if (!(fsd_o_cdist_sq(xy[2 * cj], xy[2 * cj + 1], xy[2 * o], xy[2 * o + 1]) < range2)) continue;
        double vx = xy[2 * o] - xy[2 * cj], vy = xy[2 * o + 1] - xy[2 * cj + 1];
        good += fsd_o_angle_between(vx, vy, sx, sy) < half_angle;
        bad += fsd_o_angle_between(vx, vy, -sx, -sy) < half_angle;
      }
    }
    n_good[r] = good;
    n_bad[r] = bad;
  }
  free(in_cfg);
  free(idxs);
  free(all);
  free(keep);
  free(close);
  free(other);
}

static void cost_configurations(const fsd_o_frame *f, int side, const cfg_list *cfgs, double *costs) {
  /* cost_function.py:213-304 */
  int L = cfgs->L, C = cfgs->count, n = f->n;
  const double *xy = f->xy;
  int *good = (int *)malloc(sizeof(int) * C), *bad = (int *)malloc(sizeof(int) * C);
  cones_on_either_side(f, side, cfgs, good, bad);
  int mn = 0;
  for (int r = 0; r < C; ++r)
    if (r == 0 || good[r] - bad[r] < mn) mn = good[r] - bad[r];
  const double w[7] = {1000.0, 200.0, 5000.0, 1000.0, 0.0, 1000.0, 1000.0};
  double wsum = 0.0;
  for (int i = 0; i < 7; ++i) wsum += w[i];
  for (int r = 0; r < C; ++r) {
    const int *c = cfgs->rows + (size_t)r * L;
    int len = row_len(c, L);
    double px[FSD_O_MAX_SORTED], py[FSD_O_MAX_SORTED];
    for (int q = 0; q < L; ++q) {
      int idx = c[q] == -1 ? n - 1 : c[q]; /* -1 indexes the last cone (numpy wrap-around) */
      px[q] = xy[2 * idx];
      py[q] = xy[2 * idx + 1];
    }
    /* angle cost :41-79 */
    double asum = 0.0;
    int acnt = 0, under = 0;
    for (int q = 0; q + 2 < L; ++q) {
      if (c[q + 2] == -1) continue;
      double ax = px[q + 1] - px[q + 2], ay = py[q + 1] - py[q + 2]; /* all_to_next[q+1] */
      double bx = -(px[q] - px[q + 1]), by = -(py[q] - py[q + 1]);   /* -all_to_next[q] */
      double th = fsd_o_angle_between(ax, ay, bx, by);
      asum += (M_PI - th) / M_PI;
      acnt++;
      under += th < 40.0 * M_PI / 180.0;
    }
    double angle_cost = asum / acnt * (under + 1);
    /* residual distance :14-32 of cone_distance_cost.py */
    double resid = 0.0;
    for (int q = 0; q + 1 < L; ++q) {
      if (c[q + 1] == -1) continue;
      double dx = px[q + 1] - px[q], dy = py[q + 1] - py[q];
      double d = sqrt(dx * dx + dy * dy) - 3.0;
      resid += d > 0.0 ? d : 0.0;
    }
    double ncone_cost = 1.0 / len;
    double init_dir = fsd_o_angle_between(px[1] - px[0], py[1] - py[0], f->dir[0], f->dir[1]);
    /* cones on either side :191-210 */
    int diff = good[r] - bad[r] + abs(mn) + 1;
    double either = 1.0 / diff;
    /* wrong direction :149-188 */
    double wrong = 0.0;
    if (len != 3) {
      double ang[FSD_O_MAX_SORTED];
      for (int q = 0; q + 1 < len; ++q) ang[q] = atan2(py[q + 1] - py[q], px[q + 1] - px[q]);
      double sum = 0.0;
      double unwanted = side == FSD_O_LEFT ? 1.0 : -1.0;
      for (int q = 0; q + 2 < len; ++q) {
        double d = fsd_o_angle_difference(ang[q], ang[q + 1]);
        if (fsd_o_sign(d) == unwanted && fabs(d) > 40.0 * M_PI / 180.0) sum += d;
      }
      wrong = fabs(sum);
    }
    double terms[7] = {angle_cost, resid, ncone_cost, init_dir, 0.0, either, wrong};
    double total = 0.0;
    for (int i = 0; i < 7; ++i) total += terms[i] * (w[i] / wsum);
    costs[r] = total;
  }
  free(good);
  free(bad);
}

/* one side: calc_configurations_with_score_for_one_side, core_trace_sorter.py:252-327.
 * returns the length of the best configuration (0 = no result) */
static int sort_one_side(const fsd_o_frame *f, int side, int *best, fsd_oracle_result *dbg, int sidx,
                         unsigned *status) {
  dbg->first_k[sidx][0] = dbg->first_k[sidx][1] = -1;
  dbg->n_configs[sidx] = 0;
  dbg->n_pops[sidx] = 0;
  if (f->n < 3) return 0;
  int first_k[2] = {-1, -1};
  int nfk = select_first_k(f, side, first_k);
  if (nfk == 0) return 0;
  dbg->first_k[sidx][0] = first_k[0];
  dbg->first_k[sidx][1] = nfk > 1 ? first_k[1] : -1;
  fsd_o_graph g;
  g.nbr = (int *)malloc(sizeof(int) * f->n * 5);
  g.deg = (int *)calloc(f->n, sizeof(int));
  build_graph(f, side, first_k[0], &g);
  int L = g.reachable < FSD_O_MAX_SORTED ? g.reachable : FSD_O_MAX_SORTED; /* find_configs_and_scores.py:76 */
  int len = 0;
  if (L >= 3) {
    cfg_list cfgs = {0, 0, 0, L};
    dbg->n_pops[sidx] = (int)find_end_configurations(f, side, &g, first_k, nfk, L, &cfgs, status);
    dbg->n_configs[sidx] = cfgs.count;
    if (cfgs.count > 0) {
      int arg = 0;
      if (cfgs.count > 1) {
        double *costs = (double *)malloc(sizeof(double) * cfgs.count);
        cost_configurations(f, side, &cfgs, costs);
        for (int r = 1; r < cfgs.count; ++r)
          if (costs[r] < costs[arg]) arg = r;
        free(costs);
      }
      const int *row = cfgs.rows + (size_t)arg * L;
      for (int q = 0; q < L; ++q)
        if (row[q] != -1) best[len++] = row[q];
    }
    free(cfgs.rows);
  }
  free(g.nbr);
  free(g.deg);
  return len;
}

/* ---- S8: combine (sorting_cones/trace_sorter/combine_traces.py) --------------------------- */

static double angle_change_at(const double *xy, const int *cfg, int p) {
  /* calc_angle_change_at_position :260-275 */
  int a = cfg[p - 1], b = cfg[p], c = cfg[p + 1];
  double an = atan2(xy[2 * c + 1] - xy[2 * b + 1], xy[2 * c] - xy[2 * b]);
  double ap = atan2(xy[2 * a + 1] - xy[2 * b + 1], xy[2 * a] - xy[2 * b]);
  return fsd_o_angle_difference(an, ap);
}

static void combine(const fsd_o_frame *f, int *left, int *nl, int *right, int *nr) {
  /* handle_same_cone_in_both_configs :115-147 */
  const double *xy = f->xy;
  int li = -1, ri = -1;
  for (int a = 0; a < *nl && li < 0; ++a)
    for (int b = 0; b < *nr; ++b)
      if (left[a] == right[b]) {
        li = a;
        break;
      }
  if (li < 0) return;
  for (int b = 0; b < *nr && ri < 0; ++b)
    for (int a = 0; a < *nl; ++a)
      if (left[a] == right[b]) {
        ri = b;
        break;
      }
  /* calc_new_length_for_configs_for_same_cone_intersection :150-257 */
  int ls = -1, rs = -1, have = 0;
  if (li > 0 && ri > 0) {
    int pl = left[li - 1], pr = right[ri - 1], ic = left[li];
    double dl, dr;
    dl = sqrt((xy[2 * ic] - xy[2 * pl]) * (xy[2 * ic] - xy[2 * pl]) +
              (xy[2 * ic + 1] - xy[2 * pl + 1]) * (xy[2 * ic + 1] - xy[2 * pl + 1]));
    dr = sqrt((xy[2 * ic] - xy[2 * pr]) * (xy[2 * ic] - xy[2 * pr]) +
              (xy[2 * ic + 1] - xy[2 * pr + 1]) * (xy[2 * ic + 1] - xy[2 * pr + 1]));
    int l_low = dl < 3.0, r_low = dr < 3.0;
    if ((l_low || r_low) && !(l_low && r_low)) {
      if (l_low) {
        ls = *nl;
        rs = ri;
      } else {
        ls = li;
        rs = *nr;
      }
      have = 1;
    }
  }
  if (!have && left[li] == right[ri] && li >= 1 && li <= *nl - 2 && ri >= 1 && ri <= *nr - 2) {
    double al = angle_change_at(xy, left, li), ar = angle_change_at(xy, right, ri);
    double sl = fsd_o_sign(al), sr = fsd_o_sign(ar);
    double adiff = fabs(fabs(al) - fabs(ar));
    int ndiff = abs(*nl - *nr);
    if (sl == sr) {
      if (sl == 1.0) {
        ls = *nl;
        rs = ri;
      } else {
        ls = li;
        rs = *nr;
      }
    } else if (ndiff > 2) {
      if (*nl > *nr) {
        ls = *nl;
        rs = ri;
      } else {
        ls = li;
        rs = *nr;
      }
    } else if (adiff > 5.0 * M_PI / 180.0) {
      if (fabs(al) > fabs(ar)) {
        ls = *nl;
        rs = ri;
      } else {
        ls = li;
        rs = *nr;
      }
    } else {
      ls = li;
      rs = ri;
    }
  } else if (!have) {
    int l_end = li == *nl - 1, r_end = ri == *nr - 1;
    if (l_end && r_end) {
      ls = *nl - 1;
      rs = *nr - 1;
    } else if (l_end) {
      rs = *nr;
      ls = li;
    } else if (r_end) {
      ls = *nl;
      rs = ri;
    } else {
      ls = li;
      rs = ri;
    }
  }
  *nl = ls;
  *nr = rs;
}

int fsd_o_sort_frame(const fsd_o_frame *f, fsd_oracle_result *out) {
  /* TraceSorter.sort_left_right, core_trace_sorter.py:148-216 */
  unsigned status = 0;
  int left[FSD_O_MAX_SORTED], right[FSD_O_MAX_SORTED];
  int nl = sort_one_side(f, FSD_O_LEFT, left, out, 0, &status);
  int nr = sort_one_side(f, FSD_O_RIGHT, right, out, 1, &status);
  if (nl == 0) status |= FSD_O_NO_LEFT;
  if (nr == 0) status |= FSD_O_NO_RIGHT;
  if (nl > 0 && nr > 0) combine(f, left, &nl, right, &nr);
  out->n_left = nl;
  out->n_right = nr;
  for (int i = 0; i < FSD_O_MAX_SORTED; ++i) {
    out->left_idx[i] = i < nl ? left[i] : -1;
    out->right_idx[i] = i < nr ? right[i] : -1;
  }
  out->status |= status;
  return 0;
}
