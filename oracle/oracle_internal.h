/* ORACLE (test infrastructure): declarations shared between the oracle's translation units. */
#ifndef FSD_ORACLE_INTERNAL_H
#define FSD_ORACLE_INTERNAL_H

#include "fsd_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
int fsd_o_path_from_update(double *update, int nu, const double *pos, const double *dir, int force_P,
                           const double *prev_path, fsd_oracle_result *out);
void fsd_o_almost_straight_path(double *chord40x2);

#endif

#define FSD_O_UNKNOWN 0
#define FSD_O_RIGHT 1 /* yellow */
#define FSD_O_LEFT 2  /* blue */

#define FSD_O_UNSUPPORTED (1u << 10) /* reference takes a latent-bug path the oracle does not restate */

typedef struct {
  const double *xy;          /* n x 2 */
  const unsigned char *type; /* n */
  int n;
  double pos[2], dir[2];
} fsd_o_frame;

double fsd_o_angle_between(double ax, double ay, double bx, double by);
double fsd_o_angle_difference(double a1, double a2);
double fsd_o_sign(double v);
void fsd_o_rotate(double px, double py, double theta, double *rx, double *ry);
double fsd_o_cdist_sq(double xi, double yi, double xj, double yj);
int fsd_o_inside_ellipse(double px, double py, double cx, double cy, double dx, double dy, double major,
                         double minor);
int fsd_o_segments_intersect(const double *a0, const double *a1, const double *b0, const double *b1);
void fsd_o_circle_fit(const double *pts, int n, double *cx, double *cy, double *r);

int fsd_o_sort_frame(const fsd_o_frame *f, fsd_oracle_result *out);
int fsd_o_match(const double *left, int nl, const double *right, int nr, const double *pos, const double *dir,
                fsd_oracle_result *out);
int fsd_o_path(const double *left_wv, int nl, const double *right_wv, int nr, const int *l2r, const int *r2l,
               const double *pos, const double *dir, int force_P, const double *prev_path, fsd_oracle_result *out);
int fsd_o_path_global(const double *gpath, int M, const double *pos, const double *dir, int force_P,
                      const double *prev_path, fsd_oracle_result *out);

int fsd_o_path_from_update(double *update, int nu, const double *pos, const double *dir, int force_P,
                           const double *prev_path, fsd_oracle_result *out);
void fsd_o_almost_straight_path(double *chord40x2);

#endif
