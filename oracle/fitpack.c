/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU restatement of the smoothing-spline curve fit the reference reaches through
 *   scipy.interpolate.splprep(x, s=s, k=k, u=u, per=False)   and   scipy.interpolate.splev
 * (reference call sites: fsd_path_planning/utils/spline_fit.py:61, :117-119).
 * The arithmetic lives in a third-party dependency that is NOT under /root/reference:
 * SciPy (pyproject.toml:9 lists `scipy`, unpinned; validated here against scipy 1.18.1),
 * i.e. P. Dierckx' FITPACK routines parcur -> fppara, fpknot, fpbspl, fpgivs, fprota,
 * fpback, fpdisc, fprati and splev.  This file restates the published algorithm
 *  for exactly the reference's usage: iopt=0, ipar=1 (u given),
 * ub=u[0], ue=u[m-1], unit weights, idim=2, nest=m+2k, tol=1e-3, maxit=20.
 *
 * Pinned by tests/test_oracle_fitpack.py against scipy.interpolate.splprep(full_output=1)
 * (knots, coefficients, fp, ier) and splev, and by the golden vectors in tests/golden/.
 *
 * All index arithmetic below is written 1-based like the algorithm description and shifted
 * at the point of access (macros), which keeps the control flow auditable against it.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fsd_oracle.h"

/* ---- primitives ------------------------------------------------------------------- */

static void fpgivs(double piv, double *ww, double *cs, double *sn) {
  double store = fabs(piv), dd;
  if (store >= *ww) {
    double r = *ww / piv;
    dd = store * sqrt(1.0 + r * r);
  } else {
    double r = piv / *ww;
    dd = *ww * sqrt(1.0 + r * r);
  }
  *cs = *ww / dd;
  *sn = piv / dd;
  *ww = dd;
}

static void fprota(double cs, double sn, double *a, double *b) {
  double s1 = *a, s2 = *b;
  *b = cs * s2 + sn * s1;
  *a = cs * s1 - sn * s2;
}

/* B-spline basis values h[0..k] of degree k at x, knot interval t(l) <= x < t(l+1), l 1-based */
static void fpbspl(const double *t, int k, double x, int l, double *h) {
#define T(i) t[(i)-1]
  double hh[6];
  h[0] = 1.0;
  for (int j = 1; j <= k; ++j) {
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
    for (int i = 1; i <= j; ++i) {
      int li = l + i, lj = li - j;
      if (T(li) == T(lj)) {
        h[i] = 0.0;
      } else {
        double f = hh[i - 1] / (T(li) - T(lj));
        h[i - 1] += f * (T(li) - x);
        h[i] = f * (x - T(lj));
      }
    }
  }
#undef T
}

/* a is (nest x ncol) row-major: a(i,j) = a[(i-1)*ncol + (j-1)] */
static void fpback(const double *a, int ncol, const double *z, int n, int kb, double *c) {
#define A(i, j) a[((i)-1) * ncol + ((j)-1)]
  c[n - 1] = z[n - 1] / A(n, 1);
  for (int i = n - 1; i >= 1; --i) {
    double store = z[i - 1];
    int i1 = kb - 1;
    if (n - i < i1) i1 = n - i;
    for (int l = 1; l <= i1; ++l) store -= c[i + l - 1] * A(i, l + 1);
    c[i - 1] = store / A(i, 1);
  }
#undef A
}

/* discontinuity jumps of the k-th derivative: b is ((n-2*k1) x k2) row-major */
static void fpdisc(const double *t, int n, int k2, double *b) {
#define T(i) t[(i)-1]
#define B(i, j) b[((i)-1) * k2 + ((j)-1)]
  double h[12];
  int k1 = k2 - 1, k = k1 - 1, nk1 = n - k1, nrint = nk1 - k;
  double an = nrint;
  double fac = an / (T(nk1 + 1) - T(k1));
  for (int l = k2; l <= nk1; ++l) {
    int lmk = l - k1;
    for (int j = 1; j <= k1; ++j) {
      int ik = j + k1, lj = l + j, lk = lj - k2;
      h[j - 1] = T(l) - T(lk);
      h[ik - 1] = T(l) - T(lj);
    }
    int lp = lmk;
    for (int j = 1; j <= k2; ++j) {
      int jk = j;
      double prod = h[j - 1];
      for (int i = 1; i <= k; ++i) {
        jk++;
        prod = prod * h[jk - 1] * fac;
      }
      int lk = lp + k1;
      B(lmk, j) = (T(lk) - T(lp)) / prod;
      lp++;
    }
  }
#undef T
#undef B
}

static void fpknot(const double *x, double *t, int *n, double *fpint, int *nrdata, int *nrint) {
#define T(i) t[(i)-1]
#define FPINT(i) fpint[(i)-1]
#define NRDATA(i) nrdata[(i)-1]
  int k = (*n - *nrint - 1) / 2;
  double fpmax = 0.0;
  int jbegin = 1, number = 1, maxpt = 0, maxbeg = 1;
  for (int j = 1; j <= *nrint; ++j) {
    int jpoint = NRDATA(j);
    if (!(fpmax >= FPINT(j) || jpoint == 0)) {
      fpmax = FPINT(j);
      number = j;
      maxpt = jpoint;
      maxbeg = jbegin;
    }
    jbegin += jpoint + 1;
  }
  int ihalf = maxpt / 2 + 1;
  int nrx = maxbeg + ihalf;
  int next = number + 1;
  if (next <= *nrint) {
    for (int j = next; j <= *nrint; ++j) {
      int jj = next + *nrint - j;
      FPINT(jj + 1) = FPINT(jj);
      NRDATA(jj + 1) = NRDATA(jj);
      int jk = jj + k;
      T(jk + 1) = T(jk);
    }
  }
  NRDATA(number) = ihalf - 1;
  NRDATA(next) = maxpt - ihalf;
  double am = maxpt ? maxpt : 1;
  double an = NRDATA(number);
  FPINT(number) = fpmax * an / am;
  an = NRDATA(next);
  FPINT(next) = fpmax * an / am;
  int jk = next + k;
  T(jk) = x[nrx - 1];
  *n += 1;
  *nrint += 1;
#undef T
#undef FPINT
#undef NRDATA
}

static double fprati(double *p1, double *f1, double p2, double f2, double *p3, double *f3) {
  double p;
  if (*p3 > 0.0) {
    double h1 = *f1 * (f2 - *f3);
    double h2 = f2 * (*f3 - *f1);
    double h3 = *f3 * (*f1 - f2);
    p = -(*p1 * p2 * h3 + p2 * *p3 * h1 + *p3 * *p1 * h2) / (*p1 * h1 + p2 * h2 + *p3 * h3);
  } else {
    p = (*p1 * (*f1 - *f3) * f2 - p2 * (f2 - *f3) * *f1) / ((*f1 - f2) * *f3);
  }
  if (f2 < 0.0) {
    *p3 = p2;
    *f3 = f2;
  } else {
    *p1 = p2;
    *f1 = f2;
  }
  return p;
}

/* ---- parcur / fppara --------------------------------------------------------------- */

/*
 * x: m points, interleaved (x0,y0,x1,y1,...). u: m strictly increasing parameters.
 * Outputs: t[nest], *n, c[2*nest] laid out as cx at c[0..], cy at c[nest..] (only the
 * first n-k-1 of each are meaningful), *fp.  Returns ier (FITPACK convention; 10 = invalid
 * input, the case scipy turns into ValueError).
 */
int fsd_oracle_parcur(const double *x, const double *u, int m, int k, double s, double *t, int *n_out,
                      double *c, double *fp_out) {
  const int idim = 2;
  const double tol = 1e-3;
  const int maxit = 20;
  const double con1 = 0.1, con9 = 0.9, con4 = 0.04, half = 0.5;
  int k1 = k + 1, k2 = k + 2, nmin = 2 * k1;
  int nest = m + 2 * k;
  int ier = 10;
  *n_out = 0;
  *fp_out = 0.0;
  if (k < 1 || k > 5 || m < k1 || nest < nmin) return 10;
  for (int i = 1; i < m; ++i)
    if (!(u[i - 1] < u[i])) return 10;
  if (s < 0.0) return 10;
  double ub = u[0], ue = u[m - 1];

  int nmax = m + k1;
  double acc = tol * s;

  /* work arrays */
  double *a = (double *)calloc((size_t)nest * k1, sizeof(double));
  double *g = (double *)calloc((size_t)nest * k2, sizeof(double));
  double *b = (double *)calloc((size_t)nest * k2, sizeof(double));
  double *q = (double *)calloc((size_t)m * k1, sizeof(double));
  double *z = (double *)calloc((size_t)nest * idim, sizeof(double));
  double *fpint = (double *)calloc((size_t)nest + 2, sizeof(double));
  int *nrdata = (int *)calloc((size_t)nest + 2, sizeof(int));
  double h[8], xi[2];
  memset(t, 0, sizeof(double) * nest);
  memset(c, 0, sizeof(double) * nest * idim);

#define A(i, j) a[((i)-1) * k1 + ((j)-1)]
#define G(i, j) g[((i)-1) * k2 + ((j)-1)]
#define BB(i, j) b[((i)-1) * k2 + ((j)-1)]
#define Q(i, j) q[((i)-1) * k1 + ((j)-1)]
#define Z(i, d) z[(d)*nest + ((i)-1)]
#define C(i, d) c[(d)*nest + ((i)-1)]
#define T(i) t[(i)-1]
#define U(i) u[(i)-1]
#define X(i, d) x[((i)-1) * idim + (d)]

  int n, nplus = 0, nrint, nk1 = 0;
  double fp = 0.0, fpold = 0.0, fp0 = 0.0, fpms = 0.0;
  int done = 0;

  if (s == 0.0) {
    /* interpolating curve: knots at the data abscissae */
    n = nmax;
    if (nmax > nest) {
      ier = 10;
      goto cleanup;
    }
    int mk1 = m - k1;
    if (mk1 > 0) {
      int k3 = k / 2, i = k2, j = k3 + 2;
      if (k3 * 2 != k) {
        for (int l = 1; l <= mk1; ++l) {
          T(i) = U(j);
          i++;
          j++;
        }
      } else {
        for (int l = 1; l <= mk1; ++l) {
          T(i) = (U(j) + U(j - 1)) * half;
          i++;
          j++;
        }
      }
    }
    ier = 0;
  } else {
    n = nmin;
    fpold = 0.0;
    nplus = 0;
    nrdata[0] = m - 2;
    ier = 0;
  }

  for (int iter = 1; iter <= m && !done; ++iter) {
    if (n == nmin) ier = -2;
    nrint = n - nmin + 1;
    nk1 = n - k1;
    {
      int i = n;
      for (int j = 1; j <= k1; ++j) {
        T(j) = ub;
        T(i) = ue;
        i--;
      }
    }
    fp = 0.0;
    for (int i = 0; i < nest * idim; ++i) z[i] = 0.0;
    for (int i = 1; i <= nk1; ++i)
      for (int j = 1; j <= k1; ++j) A(i, j) = 0.0;
    int l = k1;
    for (int it = 1; it <= m; ++it) {
      double ui = U(it);
      xi[0] = X(it, 0);
      xi[1] = X(it, 1);
      while (ui >= T(l + 1) && l != nk1) l++;
      fpbspl(t, k, ui, l, h);
      for (int i = 1; i <= k1; ++i) Q(it, i) = h[i - 1];
      int j = l - k1;
      for (int i = 1; i <= k1; ++i) {
        j++;
        double piv = h[i - 1];
        if (piv == 0.0) continue;
        double cs, sn;
        fpgivs(piv, &A(j, 1), &cs, &sn);
        for (int d = 0; d < idim; ++d) fprota(cs, sn, &xi[d], &Z(j, d));
        if (i == k1) break;
        int i2 = 1;
        for (int i1 = i + 1; i1 <= k1; ++i1) {
          i2++;
          fprota(cs, sn, &h[i1 - 1], &A(j, i2));
        }
      }
      for (int d = 0; d < idim; ++d) fp += xi[d] * xi[d];
    }
    if (ier == -2) fp0 = fp;
    fpint[n - 1] = fp0;
    fpint[n - 2] = fpold;
    nrdata[n - 1] = nplus;
    for (int d = 0; d < idim; ++d) fpback(a, k1, &Z(1, d), nk1, k1, &C(1, d));
    fpms = fp - s;
    if (fabs(fpms) < acc) {
      done = 1;
      break;
    }
    if (fpms < 0.0) break; /* -> part 2 */
    if (n == nmax) {
      ier = -1;
      done = 1;
      break;
    }
    if (n == nest) {
      ier = 1;
      done = 1;
      break;
    }
    if (ier == 0) {
      int npl1 = nplus * 2;
      double rn = nplus;
      if (fpold - fp > acc) npl1 = (int)(rn * fpms / (fpold - fp));
      int mx = npl1;
      if (nplus / 2 > mx) mx = nplus / 2;
      if (1 > mx) mx = 1;
      nplus = nplus * 2 < mx ? nplus * 2 : mx;
    } else {
      nplus = 1;
      ier = 0;
    }
    fpold = fp;
    /* residual sum per knot interval */
    {
      double fpart = 0.0;
      int i = 1, newk = 0;
      l = k2;
      for (int it = 1; it <= m; ++it) {
        if (!(U(it) < T(l) || l > nk1)) {
          newk = 1;
          l++;
        }
        double term = 0.0;
        int l0 = l - k2;
        for (int d = 0; d < idim; ++d) {
          double fac = 0.0;
          for (int j = 1; j <= k1; ++j) fac += C(l0 + j, d) * Q(it, j);
          double r = fac - X(it, d);
          term += r * r;
        }
        fpart += term;
        if (newk) {
          double store = term * half;
          fpint[i - 1] = fpart - store;
          i++;
          fpart = store;
          newk = 0;
        }
      }
      fpint[nrint - 1] = fpart;
    }
    for (int lk = 1; lk <= nplus; ++lk) {
      fpknot(u, t, &n, fpint, nrdata, &nrint);
      if (n == nmax) {
        /* all data abscissae become knots */
        int mk1 = m - k1;
        int k3 = k / 2, i = k2, j = k3 + 2;
        if (k3 * 2 != k) {
          for (int l2 = 1; l2 <= mk1; ++l2) {
            T(i) = U(j);
            i++;
            j++;
          }
        } else {
          for (int l2 = 1; l2 <= mk1; ++l2) {
            T(i) = (U(j) + U(j - 1)) * half;
            i++;
            j++;
          }
        }
        break;
      }
      if (n == nest) break;
    }
  }

  if (!done && ier != -2) {
    /* part 2: find the smoothing parameter p with F(p) = s */
    fpdisc(t, n, k2, b);
    double p1 = 0.0, f1 = fp0 - s, p3 = -1.0, f3 = fpms, p = 0.0;
    for (int i = 1; i <= nk1; ++i) p += A(i, 1);
    p = (double)nk1 / p;
    int ich1 = 0, ich3 = 0, n8 = n - nmin;
    double *cz = (double *)calloc((size_t)nest * idim, sizeof(double));
#define CZ(i, d) cz[(d)*nest + ((i)-1)]
    for (int iter = 1; iter <= maxit; ++iter) {
      double pinv = 1.0 / p;
      memcpy(cz, z, sizeof(double) * nest * idim);
      for (int i = 1; i <= nk1; ++i) {
        G(i, k2) = 0.0;
        for (int j = 1; j <= k1; ++j) G(i, j) = A(i, j);
      }
      for (int it = 1; it <= n8; ++it) {
        for (int i = 1; i <= k2; ++i) h[i - 1] = BB(it, i) * pinv;
        xi[0] = xi[1] = 0.0;
        for (int j = it; j <= nk1; ++j) {
          double piv = h[0], cs, sn;
          fpgivs(piv, &G(j, 1), &cs, &sn);
          for (int d = 0; d < idim; ++d) fprota(cs, sn, &xi[d], &CZ(j, d));
          if (j == nk1) break;
          int i2 = k1;
          if (j > n8) i2 = nk1 - j;
          for (int i = 1; i <= i2; ++i) {
            fprota(cs, sn, &h[i], &G(j, i + 1));
            h[i - 1] = h[i];
          }
          h[i2] = 0.0;
        }
      }
      for (int d = 0; d < idim; ++d) fpback(g, k2, &CZ(1, d), nk1, k2, &C(1, d));
      fp = 0.0;
      int l = k2;
      for (int it = 1; it <= m; ++it) {
        if (!(U(it) < T(l) || l > nk1)) l++;
        int l0 = l - k2;
        double term = 0.0;
        for (int d = 0; d < idim; ++d) {
          double fac = 0.0;
          for (int j = 1; j <= k1; ++j) fac += C(l0 + j, d) * Q(it, j);
          double r = fac - X(it, d);
          term += r * r;
        }
        fp += term;
      }
      fpms = fp - s;
      if (fabs(fpms) < acc) {
        ier = 0;
        break;
      }
      if (iter == maxit) {
        ier = 3;
        break;
      }
      double p2 = p, f2 = fpms;
      if (ich3 == 0) {
        if (!((f2 - f3) > acc)) {
          p3 = p2;
          f3 = f2;
          p = p * con4;
          if (p <= p1) p = p1 * con9 + p2 * con1;
          continue;
        }
        if (f2 < 0.0) ich3 = 1;
      }
      if (ich1 == 0) {
        if (!((f1 - f2) > acc)) {
          p1 = p2;
          f1 = f2;
          p = p / con4;
          if (p3 < 0.0) continue;
          if (p >= p3) p = p2 * con1 + p3 * con9;
          continue;
        }
        if (f2 > 0.0) ich1 = 1;
      }
      if (f2 >= f1 || f2 <= f3) {
        ier = 2;
        break;
      }
      p = fprati(&p1, &f1, p2, f2, &p3, &f3);
    }
    free(cz);
#undef CZ
  }

cleanup:
  *n_out = n;
  *fp_out = fp;
  free(a);
  free(g);
  free(b);
  free(q);
  free(z);
  free(fpint);
  free(nrdata);
  return ier;
#undef A
#undef G
#undef BB
#undef Q
#undef Z
#undef C
#undef T
#undef U
#undef X
}

/* splev with ext=0 (extrapolate with the end polynomial pieces); xs must be ascending for
 * speed but any order is handled.  c: nk1 coefficients of ONE coordinate. */
void fsd_oracle_splev(const double *t, int n, const double *c, int k, const double *xs, int mx, double *ys) {
#define T(i) t[(i)-1]
  int k1 = k + 1, k2 = k1 + 1, nk1 = n - k1;
  int l = k1, l1 = l + 1;
  double h[8];
  for (int i = 0; i < mx; ++i) {
    double arg = xs[i];
    while (!(arg >= T(l) || l1 == k2)) {
      l1 = l;
      l = l - 1;
    }
    while (!(arg < T(l1) || l == nk1)) {
      l = l1;
      l1 = l + 1;
    }
    fpbspl(t, k, arg, l, h);
    double sp = 0.0;
    int ll = l - k1;
    for (int j = 1; j <= k1; ++j) {
      ll++;
      sp += c[ll - 1] * h[j - 1];
    }
    ys[i] = sp;
  }
#undef T
}
