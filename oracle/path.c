/*
 * ORACLE (test infrastructure).  Path calculation: P1-P4 of SURVEY.md section 8(a).
 * Restates /root/reference/fsd_path_planning/calculate_path/core_calculate_path.py,
 * path_parameterization.py, path_calculator_helpers.py and utils/spline_fit.py, with the
 * parameters of config.py:48, 55-59 (s=0.2, step 0.1, degree 3, 5 m validity, 20 m horizon,
 * 40 samples).  Spline arithmetic: oracle/fitpack.c.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "fsd_oracle.h"
#include "oracle_internal.h"

#define MAXP 4096
#define HORIZON FSD_O_HORIZON

enum { OK = 0, VALUE_ERROR = 1, RAISES = 2, UNSUPPORTED = 3 };

/* numpy's pairwise summation (np.sum / np.mean of a contiguous float64 vector) */
static double np_pairwise_sum(const double *a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

void fsd_o_circle_fit(const double *pts, int n, double *cx, double *cy, double *r) {
  /* circle_fit (hyper fit), utils/math_utils.py:579-646 */
  double mx = 0.0, my = 0.0;
  for (int i = 0; i < n; ++i) {
    mx += pts[2 * i];
    my += pts[2 * i + 1];
  }
  mx /= n;
  my /= n;
  double Mxy = 0, Mxx = 0, Myy = 0, Mxz = 0, Myz = 0, Mzz = 0;
  for (int i = 0; i < n; ++i) {
    double xi = pts[2 * i] - mx, yi = pts[2 * i + 1] - my, zi = xi * xi + yi * yi;
    Mxy += xi * yi;
    Mxx += xi * xi;
    Myy += yi * yi;
    Mxz += xi * zi;
    Myz += yi * zi;
    Mzz += zi * zi;
  }
  Mxy /= n;
  Mxx /= n;
  Myy /= n;
  Mxz /= n;
  Myz /= n;
  Mzz /= n;
  double Mz = Mxx + Myy, Cov_xy = Mxx * Myy - Mxy * Mxy, Var_z = Mzz - Mz * Mz;
  double A2 = 4.0 * Cov_xy - 3.0 * Mz * Mz - Mzz;
  double A1 = Var_z * Mz + 4.0 * Cov_xy * Mz - Mxz * Mxz - Myz * Myz;
  double A0 = Mxz * (Mxz * Myy - Myz * Mxy) + Myz * (Myz * Mxx - Mxz * Mxy) - Var_z * Cov_xy;
  double A22 = A2 + A2;
  double y = A0, x = 0.0;
  for (int it = 0; it < 99; ++it) {
    double Dy = A1 + x * (A22 + 16.0 * x * x);
    double xn = x - y / Dy;
    if (xn == x || !isfinite(xn)) break;
    double yn = A0 + xn * (A1 + xn * (A2 + 4.0 * xn * xn));
    if (fabs(yn) >= fabs(y)) break;
    x = xn;
    y = yn;
  }
  double det = x * x - x * Mz + Cov_xy;
  double Xc = (Mxz * (Myy - x) - Myz * Mxy) / det / 2.0;
  double Yc = (Myz * (Mxx - x) - Mxz * Mxy) / det / 2.0;
  *cx = Xc + mx;
  *cy = Yc + my;
  *r = sqrt(fabs(Xc * Xc + Yc * Yc + Mz));
}

/* ---- SplineFitterFactory.fit / SplineEvaluator.predict (utils/spline_fit.py:46-128) --------- */

typedef struct {
  double *t, *c; /* t[nest], c[2*nest] */
  int n, k, nest;
  double max_u;
} spline_t;

static void spline_free(spline_t *s) {
  free(s->t);
  free(s->c);
  s->t = s->c = NULL;
}

/* returns OK, VALUE_ERROR (scipy raises ValueError: ier == 10) or RAISES (m < 2: NullSplineEvaluator) */
static int spline_fit(const double *pts, int m, double smoothing, spline_t *sp) {
  sp->t = sp->c = NULL;
  if (m < 2) return RAISES;
  int k = m - 1 < 1 ? 1 : (m - 1 > 3 ? 3 : m - 1);
  double *u = (double *)malloc(sizeof(double) * m);
  u[0] = 0.0;
  for (int i = 1; i < m; ++i) {
    double dx = pts[2 * i] - pts[2 * i - 2], dy = pts[2 * i + 1] - pts[2 * i - 1];
    u[i] = u[i - 1] + sqrt(dx * dx + dy * dy); /* np.cumsum: sequential */
  }
  sp->nest = m + 2 * k;
  sp->t = (double *)calloc(sp->nest, sizeof(double));
  sp->c = (double *)calloc(2 * sp->nest, sizeof(double));
  sp->k = k;
  sp->max_u = u[m - 1];
  double fp;
  int ier = fsd_oracle_parcur(pts, u, m, k, smoothing, sp->t, &sp->n, sp->c, &fp);
  free(u);
  if (ier == 10) {
    spline_free(sp);
    return VALUE_ERROR;
  }
  return OK;
}

static void spline_eval(const spline_t *sp, const double *us, int n, double *out) {
  double *tmp = (double *)malloc(sizeof(double) * n);
  fsd_oracle_splev(sp->t, sp->n, sp->c, sp->k, us, n, tmp);
  for (int i = 0; i < n; ++i) out[2 * i] = tmp[i];
  fsd_oracle_splev(sp->t, sp->n, sp->c + sp->nest, sp->k, us, n, tmp);
  for (int i = 0; i < n; ++i) out[2 * i + 1] = tmp[i];
  free(tmp);
}

/* len(np.arange(0, max_u, step)) */
static int arange_len(double max_u, double step) {
  double q = ceil(max_u / step);
  if (!(q > 0.0)) return 0;
  return (int)q;
}

/* fit(points).predict(der=0) at every `step` up to max_u (fit's own max_u when max_u_override <= 0) */
static int fit_predict(const double *pts, int m, double smoothing, double step, double max_u_override, double *out,
                       int *n_out) {
  spline_t sp;
  int rc = spline_fit(pts, m, smoothing, &sp);
  if (rc != OK) return rc;
  double mu = max_u_override > 0.0 ? max_u_override : sp.max_u;
  int n = arange_len(mu, step);
  if (n > MAXP) {
    spline_free(&sp);
    return UNSUPPORTED;
  }
  double *us = (double *)malloc(sizeof(double) * (n + 1));
  for (int i = 0; i < n; ++i) us[i] = (double)i * step;
  spline_eval(&sp, us, n, out);
  *n_out = n;
  free(us);
  spline_free(&sp);
  return OK;
}

/* ---- PathParameterizer (calculate_path/path_parameterization.py) -------------------------- */

static double det3(const double *p0, const double *p1, const double *p2) {
  /* det([[1,x0,y0],[1,x1,y1],[1,x2,y2]]) (np.linalg.det in the reference) */
  return (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]);
}

static int parameterize_path(const double *path, int n, int force_P, double out[HORIZON][4], int *P_out,
                             unsigned *status) {
  /* _refit_spline :125-161 */
  if (n < 2) return RAISES;
  double *dist = (double *)malloc(sizeof(double) * n);
  for (int i = 0; i + 1 < n; ++i) {
    double dx = path[2 * i + 2] - path[2 * i], dy = path[2 * i + 3] - path[2 * i + 1];
    dist[i] = sqrt(dx * dx + dy * dy);
  }
  double path_length = np_pairwise_sum(dist, n - 1);
  int nm = n - 1 < 10 ? n - 1 : 10;
  double mean_dist = np_pairwise_sum(dist, nm) / nm;
  free(dist);
  double predict_every = path_length / HORIZON / 3;
  double ratio = predict_every / mean_dist;
  int skip = 1;
  if (isfinite(ratio) && (int)ratio > 1) skip = (int)ratio;
  int ms = (n + skip - 1) / skip;
  double *sk = (double *)malloc(sizeof(double) * 2 * ms);
  for (int i = 0; i < ms; ++i) {
    sk[2 * i] = path[2 * i * skip];
    sk[2 * i + 1] = path[2 * i * skip + 1];
  }
  spline_t sp;
  int rc = spline_fit(sk, ms, 0.01, &sp);
  free(sk);
  if (rc != OK) return rc;
  /* evaluation grid np.arange(0, max_u, predict_every): SURVEY Q13 */
  int P;
  double q = sp.max_u / predict_every;
  if (force_P > 0) {
    P = force_P;
  } else {
    double r = nearbyint(q);
    if (fabs(q - r) < 1e-9) {
      P = (int)r;
      *status |= FSD_O_TIE_P;
    } else {
      P = (int)ceil(q);
    }
  }
  *P_out = P;
  if (P < HORIZON || P > MAXP) {
    /* np.linspace(..., dtype=int) yields repeated indices -> ValueError (:284-285) */
    spline_free(&sp);
    return P > MAXP ? UNSUPPORTED : VALUE_ERROR;
  }
  double *us = (double *)malloc(sizeof(double) * P);
  double *pts = (double *)malloc(sizeof(double) * 2 * P);
  for (int i = 0; i < P; ++i) us[i] = (double)i * predict_every;
  spline_eval(&sp, us, P, pts);
  spline_free(&sp);
  /* _calculate_path_curvature :163-193, calculate_path_curvature :49-93 */
  int window = P / 5 < 30 ? P / 5 : 30;
  if (window % 2 == 0) window += 1;
  int hw = window / 2;
  double *curv = (double *)malloc(sizeof(double) * P);
  int *widx = (int *)malloc(sizeof(int) * window);
  double *wpts = (double *)malloc(sizeof(double) * 2 * window);
  for (int i = 0; i < P; ++i) {
    int wn = window;
    for (int j = 0; j < window; ++j) widx[j] = ((i - hw + j) % P + P) % P;
    int cut = -1;
    for (int j = 0; j + 1 < window; ++j)
      if (widx[j + 1] - widx[j] != 1) {
        cut = j + 1;
        break;
      }
    int lo = 0;
    if (cut >= 0) {
      if (i < window)
        lo = cut;
      else
        wn = cut;
    }
    int cnt = 0;
    for (int j = lo; j < wn; ++j) {
      wpts[2 * cnt] = pts[2 * widx[j]];
      wpts[2 * cnt + 1] = pts[2 * widx[j] + 1];
      cnt++;
    }
    double cx, cy, r;
    fsd_o_circle_fit(wpts, cnt, &cx, &cy, &r);
    r = fmin(fmax(r, 1.0), 3000.0);
    double sgn = fsd_o_sign(det3(&wpts[0], &wpts[2 * (cnt / 2)], &wpts[2 * (cnt - 1)]));
    curv[i] = (1.0 / r) * sgn;
  }
  /* scipy.ndimage.uniform_filter1d(size, mode="nearest"): window [i - size/2, i + size - size/2 - 1] */
  int fs = window / 2 > 2 ? window / 2 : 2;
  double *filt = (double *)malloc(sizeof(double) * P);
  for (int i = 0; i < P; ++i) {
    double acc = 0.0;
    for (int j = i - fs / 2; j <= i + fs - fs / 2 - 1; ++j) {
      int jj = j < 0 ? 0 : (j > P - 1 ? P - 1 : j);
      acc += curv[jj];
    }
    filt[i] = acc / fs;
  }
  /* _sample_path_parameters_for_prediction_horizon :252-295: np.linspace(0, P-1, 40, dtype=int) */
  double step = (double)(P - 1) / (double)(HORIZON - 1);
  for (int i = 0; i < HORIZON; ++i) {
    int idx = i == HORIZON - 1 ? P - 1 : (int)floor((double)i * step);
    out[i][0] = us[idx];
    out[i][1] = pts[2 * idx];
    out[i][2] = pts[2 * idx + 1];
    out[i][3] = filt[idx];
  }
  free(us);
  free(pts);
  free(curv);
  free(widx);
  free(wpts);
  free(filt);
  return OK;
}

/* ---- CalculatePath tail (core_calculate_path.py:239-499) ----------------------------------- */

static int mpc_tail(const double *update, int n_in, const double *pos, const double *dir, int force_P,
                    double out[HORIZON][4], int *P_out, int *n_trim, unsigned *status) {
  double *path = (double *)malloc(sizeof(double) * 2 * (MAXP + 64));
  int n = 0;
  if (n_in < 1 || n_in > MAXP) {
    free(path);
    return n_in < 1 ? RAISES : UNSUPPORTED;
  }
  /* connect_path_to_car :430-457 */
  {
    double fx = update[0] - pos[0], fy = update[1] - pos[1];
    double d = sqrt(fx * fx + fy * fy);
    double a = fsd_o_angle_between(fx, fy, dir[0], dir[1]);
    if (!(d < 0.5 || a > M_PI / 2.0)) {
      path[0] = pos[0] + fx / d * 0.2;
      path[1] = pos[1] + fy / d * 0.2;
      n = 1;
    }
    memcpy(path + 2 * n, update, sizeof(double) * 2 * n_in);
    n += n_in;
  }
  /* extend_path :261-334 */
  {
    int first = n;
    for (int i = 0; i < n; ++i)
      if ((path[2 * i] - pos[0]) * dir[0] + (path[2 * i + 1] - pos[1]) * dir[1] > 0.0) {
        first = i;
        break;
      }
    int start = n - 20 < 0 ? 0 : n - 20;
    if (first < start) start = first;
    int nf = n - start;
    const double *front = path + 2 * start;
    if (nf < 2) {
      free(path);
      return RAISES; /* cumsum of an empty array indexed with [-1] */
    }
    double plen = 0.0;
    for (int i = 0; i + 1 < nf; ++i) {
      double dx = front[2 * i + 2] - front[2 * i], dy = front[2 * i + 3] - front[2 * i + 1];
      plen += sqrt(dx * dx + dy * dy);
    }
    if (!(plen > 20.0)) {
      int nr = nf < 20 ? nf : 20;
      const double *rel = front + 2 * (nf - nr);
      double cx, cy, radius;
      fsd_o_circle_fit(rel, nr, &cx, &cy, &radius);
      double r_use = fmin(fmax(radius, 10.0), 100.0);
      double lastx = path[2 * (n - 1)], lasty = path[2 * (n - 1) + 1];
      if (r_use < 80.0) {
        double p0[2] = {rel[0] - cx, rel[1] - cy};
        double p1[2] = {rel[2 * (nr / 2)] - cx, rel[2 * (nr / 2) + 1] - cy};
        double p2[2] = {rel[2 * (nr - 1)] - cx, rel[2 * (nr - 1) + 1] - cy};
        double sgn = fsd_o_sign(det3(p0, p1, p2));
        double a0 = atan2(p0[1], p0[0]);
        double a1 = a0 + sgn * M_PI;
        /* np.linspace(a0, a1) -> 50 angles; the first point is dropped (:333) */
        double stepa = (a1 - a0) / 49.0;
        double r0x = cos(a0) * r_use, r0y = sin(a0) * r_use;
        for (int i = 1; i < 50; ++i) {
          double ang = i == 49 ? a1 : (double)i * stepa + a0;
          path[2 * n] = cos(ang) * r_use - r0x + lastx;
          path[2 * n + 1] = sin(ang) * r_use - r0y + lasty;
          n++;
        }
      } else {
        double dx = lastx - path[2 * (n - 2)], dy = lasty - path[2 * (n - 2) + 1];
        double nrm = sqrt(dx * dx + dy * dy);
        dx /= nrm;
        dy /= nrm;
        for (int i = 1; i < 30; ++i) {
          path[2 * n] = lastx + dx * (double)i;
          path[2 * n + 1] = lasty + dy * (double)i;
          n++;
        }
      }
    }
  }
  /* remove_path_behind_car :459-465 */
  int i0 = 0;
  {
    double best = 0.0;
    for (int i = 0; i < n; ++i) {
      double dx = pos[0] - path[2 * i], dy = pos[1] - path[2 * i + 1];
      double d = sqrt(dx * dx + dy * dy);
      if (i == 0 || d < best) {
        best = d;
        i0 = i;
      }
    }
  }
  /* refit_path_for_mpc_with_safety_factor :239-259: predict up to u = 20 * 1.5 */
  double *fixed = (double *)malloc(sizeof(double) * 2 * 512);
  int nfix = 0;
  int rc = fit_predict(path + 2 * i0, n - i0, 0.2, 0.1, 20.0 * 1.5, fixed, &nfix);
  free(path);
  if (rc == RAISES) rc = UNSUPPORTED; /* NullSplineEvaluator -> (40,4) previous path re-parameterised (latent bug) */
  if (rc != OK) {
    free(fixed);
    return rc;
  }
  /* remove_path_not_in_prediction_horizon :467-499 */
  int keep;
  {
    double cum = 0.0;
    int first_over = -1;
    for (int i = 0; i + 1 < nfix; ++i) {
      double dx = fixed[2 * i + 2] - fixed[2 * i], dy = fixed[2 * i + 3] - fixed[2 * i + 1];
      cum += sqrt(dx * dx + dy * dy);
      if (cum > 20.0) {
        first_over = i;
        break;
      }
    }
    keep = first_over < 0 ? nfix - 1 : first_over;
  }
  *n_trim = keep;
  rc = parameterize_path(fixed, keep, force_P, out, P_out, status);
  free(fixed);
  return rc;
}

/* ---- initial path of a fresh planner (core_calculate_path.py:103-121) ----------------------- */

static double g_initial[HORIZON][4];
static pthread_once_t g_initial_once = PTHREAD_ONCE_INIT;

void fsd_o_almost_straight_path(double *chord) {
  /* calculate_almost_straight_path, path_calculator_helpers.py:26-68 */
  double max_angle = M_PI / 50.0, radius = 1000.0;
  double step = max_angle / (HORIZON - 1);
  for (int i = 0; i < HORIZON; ++i) {
    double a = i == HORIZON - 1 ? max_angle : (double)i * step;
    double px = (cos(a) - 1.0) * radius, py = (sin(a) - 0.0) * radius, rx, ry;
    fsd_o_rotate(px, py, -(M_PI / 2.0), &rx, &ry);
    chord[2 * i] = rx;
    chord[2 * i + 1] = ry * 1.0;
  }
}

static void compute_initial(void) {
  double chord[2 * HORIZON];
  fsd_o_almost_straight_path(chord);
  double *dense = (double *)malloc(sizeof(double) * 2 * MAXP);
  int nd = 0, P = 0;
  unsigned st = 0;
  fit_predict(chord, HORIZON, 0.2, 0.1, -1.0, dense, &nd);
  parameterize_path(dense, nd, 0, g_initial, &P, &st);
  free(dense);
}

void fsd_oracle_initial_path(double *out40x4) {
  pthread_once(&g_initial_once, compute_initial);
  memcpy(out40x4, g_initial, sizeof(g_initial));
}

/* second half of run_path_calculation (core_calculate_path.py:555-575): validity check, MPC tail, fallback.
 * update: path update (capacity >= 40 points), modified in place. */
int fsd_o_path_from_update(double *update, int nu, const double *pos, const double *dir, int force_P,
                           const double *prev_path, fsd_oracle_result *out) {
  double prev[HORIZON][4];
  memcpy(prev, prev_path, sizeof(prev));
  double prev_xy[2 * HORIZON];
  for (int i = 0; i < HORIZON; ++i) {
    prev_xy[2 * i] = prev[i][1];
    prev_xy[2 * i + 1] = prev[i][2];
  }
  int rc;
  /* overwrite_path_if_it_is_too_far_away :225-237 */
  {
    double best = INFINITY;
    for (int i = 0; i < nu; ++i) {
      double dx = pos[0] - update[2 * i], dy = pos[1] - update[2 * i + 1];
      double d = sqrt(dx * dx + dy * dy);
      if (d < best) best = d;
    }
    if (best > 5.0) {
      out->status |= FSD_O_PATH_TOO_FAR;
      memcpy(update, prev_xy, sizeof(prev_xy));
      nu = HORIZON;
    }
  }
  /* do_all_mpc_parameter_calculations with the ValueError fallback :561-570 */
  unsigned st = 0;
  rc = mpc_tail(update, nu, pos, dir, force_P, out->path, &out->P, &out->n_trim, &st);
  if (rc == VALUE_ERROR) {
    out->status |= FSD_O_MPC_FAILED;
    st = 0;
    rc = mpc_tail(prev_xy, HORIZON, pos, dir, force_P, out->path, &out->P, &out->n_trim, &st);
  }
  out->status |= st;
  if (rc != OK) {
    out->status |= rc == UNSUPPORTED ? FSD_O_UNSUPPORTED : FSD_O_REF_RAISES;
    memcpy(out->path, prev, sizeof(prev));
  }
  return 0;
}

/* fit_matches_as_spline :207-223 and everything after it, on a given centre line */
static int path_from_centerline(const double *cl, int ncl, double prev[HORIZON][4], const double *prev_xy,
                                const double *pos, const double *dir, int force_P, fsd_oracle_result *out) {
  double *update = (double *)malloc(sizeof(double) * 2 * MAXP);
  int nu = 0;
  int rc = fit_predict(cl, ncl, 0.2, 0.1, -1.0, update, &nu);
  if (rc == VALUE_ERROR) {
    out->status |= FSD_O_FIT1_FAILED;
    rc = fit_predict(prev_xy, HORIZON, 0.2, 0.1, -1.0, update, &nu);
  }
  if (rc != OK || nu < 1) {
    out->status |= rc == UNSUPPORTED ? (FSD_O_UNSUPPORTED | FSD_O_OVERFLOW) : FSD_O_REF_RAISES;
    memcpy(out->path, prev, sizeof(double) * HORIZON * 4);
    free(update);
    return 0;
  }
  fsd_o_path_from_update(update, nu, pos, dir, force_P, &prev[0][0], out);
  free(update);
  return 0;
}

/* ---- run_path_calculation with a global path (core_calculate_path.py:516-528): the centre line is the part of the
 * global path within 30 m of the car, rolled so that the closest point sits at index M / 3 ------------------------ */
int fsd_o_path_global(const double *gpath, int M, const double *pos, const double *dir, int force_P,
                      const double *prev_path, fsd_oracle_result *out) {
  double prev[HORIZON][4];
  if (prev_path)
    memcpy(prev, prev_path, sizeof(prev));
  else
    fsd_oracle_initial_path(&prev[0][0]);
  double prev_xy[2 * HORIZON];
  for (int i = 0; i < HORIZON; ++i) {
    prev_xy[2 * i] = prev[i][1];
    prev_xy[2 * i + 1] = prev[i][2];
  }
  double *dist = (double *)malloc(sizeof(double) * (M > 0 ? M : 1));
  double *cl = (double *)malloc(sizeof(double) * 2 * (M > 0 ? M : 1));
  int best = 0;
  for (int i = 0; i < M; ++i) {
    double dx = pos[0] - gpath[2 * i], dy = pos[1] - gpath[2 * i + 1];
    dist[i] = sqrt(dx * dx + dy * dy);
    if (dist[i] < dist[best]) best = i; /* np.argmin: first minimum */
  }
  /* np.roll(a, r)[i] = a[(i - r) mod M] with r = -best + M / 3 */
  int ncl = 0;
  for (int i = 0; i < M; ++i) {
    int src = ((i + best - M / 3) % M + M) % M;
    if (dist[src] < 30.0) {
      cl[2 * ncl] = gpath[2 * src];
      cl[2 * ncl + 1] = gpath[2 * src + 1];
      ncl++;
    }
  }
  int rc = path_from_centerline(cl, ncl, prev, prev_xy, pos, dir, force_P, out);
  free(dist);
  free(cl);
  return rc;
}

/* ---- CalculatePath.run_path_calculation (core_calculate_path.py:514-575), global_path None ---- */

int fsd_o_path(const double *left, int nl, const double *right, int nr, const int *l2r, const int *r2l,
               const double *pos, const double *dir, int force_P, const double *prev_path, fsd_oracle_result *out) {
  double prev[HORIZON][4];
  if (prev_path)
    memcpy(prev, prev_path, sizeof(prev));
  else
    fsd_oracle_initial_path(&prev[0][0]);
  double prev_xy[2 * HORIZON];
  for (int i = 0; i < HORIZON; ++i) {
    prev_xy[2 * i] = prev[i][1];
    prev_xy[2 * i + 1] = prev[i][2];
  }
  double centre[2 * 64];
  int nc = 0;
  const double *cl = prev_xy;
  int ncl = HORIZON;
  if (nl < 3 && nr < 3) {
    out->status |= FSD_O_FEW_CONES;
  } else {
    /* select_side_to_use :165-183 / side_score :151-163: max((n_matches, sum idx)), ties -> left */
    int nml = 0, nmr = 0;
    long sl = 0, sr = 0;
    for (int i = 0; i < nl; ++i)
      if (l2r[i] != -1) {
        nml++;
        sl += l2r[i];
      }
    for (int i = 0; i < nr; ++i)
      if (r2l[i] != -1) {
        nmr++;
        sr += r2l[i];
      }
    int use_left = !(nmr > nml || (nmr == nml && sr > sl));
    const double *side = use_left ? left : right, *other = use_left ? right : left;
    const int *mt = use_left ? l2r : r2l;
    int ns = use_left ? nl : nr;
    /* calculate_centerline_points_of_matches :185-205 */
    for (int i = 0; i < ns; ++i)
      if (mt[i] != -1) {
        centre[2 * nc] = (side[2 * i] + other[2 * mt[i]]) / 2.0;
        centre[2 * nc + 1] = (side[2 * i + 1] + other[2 * mt[i] + 1]) / 2.0;
        nc++;
      }
    if (nc < 2) {
      out->status |= FSD_O_FEW_MATCHES;
    } else {
      cl = centre;
      ncl = nc;
    }
  }
  return path_from_centerline(cl, ncl, prev, prev_xy, pos, dir, force_P, out);
}
