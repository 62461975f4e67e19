/*
 * ORACLE (test infrastructure).  CPU restatement, in plain C / fp64, of the reference's
 * sorting_cones -> cone_matching -> calculate_path hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the reported CPU baseline.  The product
 * (ft_fsd_path_planning_b200 + libfsdplan.so) never links, imports or calls it.
 *
 * Parity pin: there are no tests or golden vectors in the reference, so
 * this restatement is pinned against outputs of the UNMODIFIED reference run in the build
 * container: the .npz files under tests/golden/, produced by tests/golden/make_goldens.py.
 */
#ifndef FSD_ORACLE_H
#define FSD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define FSD_O_MAX_SORTED 12
#define FSD_O_MAX_WV 32
#define FSD_O_HORIZON 40

/* status bits (same numbering as include/fsdplan.h) */
#define FSD_O_NO_LEFT (1u << 0)
#define FSD_O_NO_RIGHT (1u << 1)
#define FSD_O_FEW_CONES (1u << 2)     /* both sides < 3 cones -> previous path as centre line */
#define FSD_O_FEW_MATCHES (1u << 3)   /* < 2 centre points -> previous path */
#define FSD_O_FIT1_FAILED (1u << 4)   /* first spline fit invalid -> previous path re-fitted */
#define FSD_O_PATH_TOO_FAR (1u << 5)  /* path > 5 m from the car -> previous path */
#define FSD_O_MPC_FAILED (1u << 6)    /* tail failed (ValueError in the reference) -> redone with previous path */
#define FSD_O_TIE_P (1u << 7)         /* evaluation-grid size decided by the tie rule */
#define FSD_O_OVERFLOW (1u << 8)      /* a static bound was exceeded */
#define FSD_O_REF_RAISES (1u << 9)    /* the reference raises an exception on this input */

typedef struct {
  int n_left, n_right;
  int left_idx[FSD_O_MAX_SORTED], right_idx[FSD_O_MAX_SORTED];
  int n_left_wv, n_right_wv;
  double left_wv[FSD_O_MAX_WV][2], right_wv[FSD_O_MAX_WV][2];
  int l2r[FSD_O_MAX_WV], r2l[FSD_O_MAX_WV];
  double path[FSD_O_HORIZON][4]; /* u, x, y, curvature */
  int P;                         /* size of the last evaluation grid */
  int n_trim;                    /* points entering the last re-fit */
  unsigned status;
  /* debugging aids */
  int first_k[2][2];  /* [side 0=left,1=right][..], -1 padded */
  int n_configs[2];   /* configurations after the post-filter */
  int n_pops[2];      /* DFS node pops */
} fsd_oracle_result;

/* force_P <= 0: default rule (round when within 1e-9 of an integer, else ceil).
 * prev_path: 40x4 previous path (NULL -> the planner's constant initial path). */
int fsd_oracle_plan_frame(const double *cones_xy, const unsigned char *cones_type, int n, const double *pos,
                          const double *dir, int force_P, const double *prev_path, fsd_oracle_result *out);

/* Packed batch, `threads` host threads (pthreads); results[b] for every frame. */
int fsd_oracle_plan_batch(const double *cones_xy, const unsigned char *cones_type, const int *offsets, int n_frames,
                          const double *pos, const double *dir, const short *force_P, int threads,
                          fsd_oracle_result *results);

/* constant initial path of a fresh planner (core_calculate_path.py:103-107) */
void fsd_oracle_initial_path(double *out40x4);

/* FITPACK restatement (oracle/fitpack.c) */
int fsd_oracle_parcur(const double *x, const double *u, int m, int k, double s, double *t, int *n_out, double *c,
                      double *fp_out);
void fsd_oracle_splev(const double *t, int n, const double *c, int k, const double *xs, int mx, double *ys);

/* CalculatePath.run_path_calculation with a global path (core_calculate_path.py:516-528) */
int fsd_oracle_path_global(const double *global_path, int n_points, const double *pos, const double *dir, int force_P,
                           const double *prev_path, fsd_oracle_result *out);

/* stage entry points for stage-level checks */
/* create_adjacency_matrix (sorting_cones/trace_sorter/adjacency_matrix.py:60-128) for one side (1 = right / yellow,
 * 2 = left / blue): neighbour lists nbr [n][5] in ascending index order, degrees deg [n] */
int fsd_oracle_adjacency(const double *cones_xy, const unsigned char *cones_type, int n, int side, int *nbr, int *deg);
int fsd_oracle_sort(const double *cones_xy, const unsigned char *cones_type, int n, const double *pos,
                    const double *dir, fsd_oracle_result *out);
int fsd_oracle_match(const double *left, int nl, const double *right, int nr, const double *pos, const double *dir,
                     fsd_oracle_result *out);
int fsd_oracle_path(const double *left_wv, int nl, const double *right_wv, int nr, const int *l2r, const int *r2l,
                    const double *pos, const double *dir, int force_P, const double *prev_path,
                    fsd_oracle_result *out);

/* ---- skidpad mission (oracle/skidpad.c) ---- */
typedef struct {
  int relocalized, n_accepted;
  double translation[2], rotation, right_ref[2], right_calc[2];
} fsd_oracle_reloc;

/* jitter: RandomState(42).randn(1140 * 6); ref_centers: [right(x,y), left(x,y)] of the canonical path */
int fsd_oracle_skidpad_relocalize(const double *cones_xy, int n, const double *pos, const double *orig_pos,
                                  const double *orig_dir, const double *jitter, const double *ref_centers,
                                  fsd_oracle_reloc *out);
/* one planner step; path = canonical path [::2]; index_along_path in/out; reloc may be NULL (not relocalised);
 * path_internal: the (40,4) path before the transform back to the SLAM frame (next step's prev_path) */
int fsd_oracle_skidpad_step(const double *path, int n_path, int *index_along_path, const fsd_oracle_reloc *reloc,
                            const double *pos, const double *dir, int force_P, const double *prev_path,
                            double *path_internal, fsd_oracle_result *out);

#ifdef __cplusplus
}
#endif
#endif
