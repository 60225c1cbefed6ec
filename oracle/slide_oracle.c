/*
 * slide_oracle.c -- CPU ORACLE (test infrastructure, see slide_oracle.h).
 *
 * Restates the reference's arithmetic in fp64 with the reference's loop order.  Build with
 * -ffp-contract=off (place_recognition.cpp is compiled without FMA).
 * "PR.cpp" = backend/sloam/src/core/place_recognition.cpp,
 * "SC.cpp" = backend/sloam/clipper_semantic_object/src/semantic_clipper.cpp.
 */
#include "slide_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ params */

double slide_oracle_deg2rad(double deg) { return deg * M_PI / 180.; } /* PR.cpp:35 */

void slide_oracle_default_params(slide_oracle_params *p) { /* PR.cpp:24-75 */
  p->compute_budget_sec = -1.0; /* the oracle disables the anytime budget by default */
  p->dilation_factor = 1.2;
  p->match_xy_step_size = 0.5;
  p->match_yaw_half_range = slide_oracle_deg2rad(180.);
  p->disable_yaw_search = 0;
  p->match_yaw_angle_step_size = slide_oracle_deg2rad(2.0);
  p->match_threshold = 0.5;
  p->match_threshold_dimension = 1.0;
  p->ignore_dimension = 0;
  p->min_num_inliers = 5;
  p->use_lsq = 1;
  p->min_num_map_objects_to_start = 1;
  p->match_x_half_range_intra = 5.0;
  p->match_y_half_range_intra = 5.0;
  p->match_yaw_half_range_intra = slide_oracle_deg2rad(10.);
  p->inter_loop_closure = 1;
}

/* ------------------------------------------------------------------ scoring */

/* PR.cpp:299-355: first reference object (ascending index) that has the query's label,
 * lies strictly within match_threshold in xy and passes the dimension rule. */
static int first_match(const slide_oracle_params *p, const double *ref7, int n_ref,
                       double label, double xq, double yq, const double *qdims) {
  for (int i = 0; i < n_ref; i++) {
    const double *r = ref7 + 7 * (size_t)i;
    if (r[0] != label) continue;                       /* PR.cpp:306 */
    double x_diff = r[1] - xq;                          /* PR.cpp:310-313 */
    double y_diff = r[2] - yq;
    double avg_dim_diff = 0;
    if (r[5] == 0 && r[6] == 0) {                       /* PR.cpp:318-322 (keyed on REF) */
      avg_dim_diff = fabs(r[4] - qdims[0]);
    } else {
      for (int d = 0; d < 3; d++) avg_dim_diff += fabs(r[4 + d] - qdims[d]);
      avg_dim_diff /= 3;                                /* PR.cpp:329 */
    }
    int distance_match = sqrt(x_diff * x_diff + y_diff * y_diff) < p->match_threshold;
    int dimension_match = p->ignore_dimension ? 1 : (avg_dim_diff < p->match_threshold_dimension);
    if (distance_match && dimension_match) return i;   /* PR.cpp:341-353 (break) */
  }
  return -1;
}

int slide_oracle_score_one(const slide_oracle_params *p, const double *ref7, int n_ref,
                           const double *qry7, int n_qry, double c, double s, double x,
                           double y, int *ref_idx_out, int *qry_idx_out) {
  int inliers = 0;
  const double ms = -s; /* cur_R_t(0,1) = -sin(yaw)  PR.cpp:247 */
  for (int j = 0; j < n_qry; j++) {
    const double *q = qry7 + 7 * (size_t)j;
    /* PR.cpp:257-261: R_t * [qx qy 1]^T, then / 1.0.  Eigen evaluates this fixed 3x3 times
     * (dynamic) 3x1 product coefficient-based (product_type_selector<Small,1,Small>); with its
     * default packet path (SSE2 / NEON, EIGEN_UNALIGNED_VECTORIZE) rows 0..1 come out of
     * etor_product_packet_impl = pmul, pmadd, pmadd: LEFT TO RIGHT (the contract adopted here).
     * Built with EIGEN_DONT_VECTORIZE the scalar path is a 3-term redux e0 + (e1 + e2): that
     * variant is compiled with -DSLIDE_ORACLE_ALT_SUM_ORDER into libslide_oracle_alt.so and
     * tests/test_oracle_pins.py checks that no golden result depends on the choice. */
#ifdef SLIDE_ORACLE_ALT_SUM_ORDER
    double xt = c * q[1] + (ms * q[2] + x);
    double yt = s * q[1] + (c * q[2] + y);
#else
    double xt = (c * q[1] + ms * q[2]) + x;
    double yt = (s * q[1] + c * q[2]) + y;
#endif
    int i = first_match(p, ref7, n_ref, q[0], xt, yt, q + 4);
    if (i >= 0) {
      if (ref_idx_out) ref_idx_out[inliers] = i;
      if (qry_idx_out) qry_idx_out[inliers] = j;
      inliers++;
    }
  }
  return inliers;
}

/* ------------------------------------------------------------------ lattice */

typedef struct {
  int n_yaw;
  double *yaw;      /* PR.cpp:136-146 */
  int rings;        /* outer_loop_steps */
  double ox, oy;    /* outer_loop_step_size_{x,y} */
  int sanity_fail;
} lattice_hdr;

static int build_yaw(const slide_oracle_params *p, double yaw_half, lattice_hdr *L) {
  int cap = 16, n = 0;
  double *v = (double *)malloc(sizeof(double) * cap);
  if (p->disable_yaw_search) {
    v[n++] = 0.0;
  } else {
    for (double yaw_raw = -yaw_half; yaw_raw < yaw_half; yaw_raw += p->match_yaw_angle_step_size) {
      if (n == cap) { cap *= 2; v = (double *)realloc(v, sizeof(double) * cap); }
      v[n++] = yaw_raw;
      if (n > 100000000) break; /* guard against a zero step */
    }
  }
  L->yaw = v;
  L->n_yaw = n;
  return n;
}

static void build_hdr(const slide_oracle_params *p, double half_x, double half_y, lattice_hdr *L) {
  double outer = 10 * p->match_xy_step_size;                       /* PR.cpp:154 */
  double steps_d = fmin(half_x, half_y) / outer;                   /* PR.cpp:155-156 */
  int steps = (int)ceil(steps_d);                                  /* PR.cpp:158 */
  L->rings = steps;
  L->ox = half_x / (double)steps;                                  /* PR.cpp:163-166 */
  L->oy = half_y / (double)steps;
  L->sanity_fail = (L->ox < p->match_xy_step_size || L->oy < p->match_xy_step_size); /* :169 */
}

long long slide_oracle_enumerate_lattice(const slide_oracle_params *p, double half_x,
                                         double half_y, double *tx_out, double *ty_out,
                                         int *ring_out, long long cap, double *yaw_out,
                                         int yaw_cap, int *n_yaw_out) {
  lattice_hdr L;
  build_yaw(p, p->inter_loop_closure ? p->match_yaw_half_range : p->match_yaw_half_range_intra, &L);
  if (n_yaw_out) *n_yaw_out = L.n_yaw;
  if (yaw_out)
    for (int i = 0; i < L.n_yaw && i < yaw_cap; i++) yaw_out[i] = L.yaw[i];
  free(L.yaw);
  build_hdr(p, half_x, half_y, &L);
  if (L.sanity_fail) return -1;
  const double step = p->match_xy_step_size;
  long long n = 0;
  for (int k = 0; k < L.rings; k++) {
    double kd = (double)k;
    double x_right_prev = kd * L.ox, x_left_prev = -kd * L.ox;     /* PR.cpp:204,210 */
    double x_pos_end = (kd + 1) * L.ox, x_neg_start = -(kd + 1) * L.ox;
    double y_right_prev = kd * L.oy, y_left_prev = -kd * L.oy;
    double y_pos_end = (kd + 1) * L.oy, y_neg_start = -(kd + 1) * L.oy;
    for (double x = x_neg_start; x <= x_pos_end; x += step) {      /* PR.cpp:230 */
      for (double y = y_neg_start; y <= y_pos_end; y += step) {    /* PR.cpp:232 */
        if ((x >= x_left_prev && x <= x_right_prev) && (y >= y_left_prev && y <= y_right_prev))
          continue;                                                /* PR.cpp:238-241 */
        if (n < cap) {
          if (tx_out) tx_out[n] = x;
          if (ty_out) ty_out[n] = y;
          if (ring_out) ring_out[n] = k;
        }
        n++;
      }
    }
  }
  return n;
}

static void result_init(slide_oracle_match_result *res) {
  memset(res, 0, sizeof(*res));
  res->best_num_inliers = -10000;                                  /* PR.cpp:125 */
  res->R_t[0] = res->R_t[4] = res->R_t[8] = 1.0;                   /* PR.cpp:127 */
  res->best_hyp_index = -1;
}

static void set_rt(double *R_t, double c, double s, double x, double y) {
  R_t[0] = c;  R_t[1] = -s; R_t[2] = x;                            /* PR.cpp:246-251 */
  R_t[3] = s;  R_t[4] = c;  R_t[5] = y;
  R_t[6] = 0;  R_t[7] = 0;  R_t[8] = 1;
}

int slide_oracle_match_maps(const slide_oracle_params *p, const double *ref7, int n_ref,
                            const double *qry7, int n_qry, double half_x, double half_y,
                            long long hyp_begin, long long hyp_end, int *ref_idx_out,
                            int *qry_idx_out, int *counts_out, long long counts_cap,
                            slide_oracle_match_result *res) {
  result_init(res);
  lattice_hdr L;
  build_yaw(p, p->inter_loop_closure ? p->match_yaw_half_range : p->match_yaw_half_range_intra, &L);
  build_hdr(p, half_x, half_y, &L);
  res->n_yaw = L.n_yaw;
  res->n_rings = L.rings;
  if (L.sanity_fail) { /* PR.cpp:169-175: return with outputs untouched */
    free(L.yaw);
    res->status = 1;
    return 1;
  }
  /* cos/sin through libm once per yaw candidate -- the reference calls cos(yaw)/sin(yaw)
   * per hypothesis with the same argument (PR.cpp:246-250), so the values are identical */
  double *cs = (double *)malloc(sizeof(double) * 2 * (size_t)(L.n_yaw > 0 ? L.n_yaw : 1));
  for (int a = 0; a < L.n_yaw; a++) { cs[2 * a] = cos(L.yaw[a]); cs[2 * a + 1] = sin(L.yaw[a]); }
  int *cur_ref = (int *)malloc(sizeof(int) * (size_t)(n_qry > 0 ? n_qry : 1));
  int *cur_qry = (int *)malloc(sizeof(int) * (size_t)(n_qry > 0 ? n_qry : 1));

  const double step = p->match_xy_step_size;
  int best = -10000;
  long long h = 0, scored = 0;
  time_t t0 = time(NULL);
  for (int k = 0; k < L.rings; k++) {
    if (p->compute_budget_sec > 0) { /* PR.cpp:181-191, whole seconds, strict '>' */
      double duration = (double)(long long)difftime(time(NULL), t0);
      if (duration > p->compute_budget_sec) break;
    }
    double kd = (double)k;
    double x_right_prev = kd * L.ox, x_left_prev = -kd * L.ox;
    double x_pos_end = (kd + 1) * L.ox, x_neg_start = -(kd + 1) * L.ox;
    double y_right_prev = kd * L.oy, y_left_prev = -kd * L.oy;
    double y_pos_end = (kd + 1) * L.oy, y_neg_start = -(kd + 1) * L.oy;
    for (double x = x_neg_start; x <= x_pos_end; x += step) {
      for (double y = y_neg_start; y <= y_pos_end; y += step) {
        if ((x >= x_left_prev && x <= x_right_prev) && (y >= y_left_prev && y <= y_right_prev))
          continue;
        if (h + L.n_yaw <= hyp_begin || (hyp_end >= 0 && h >= hyp_end)) { h += L.n_yaw; continue; }
        for (int a = 0; a < L.n_yaw; a++, h++) {
          if (h < hyp_begin || (hyp_end >= 0 && h >= hyp_end)) continue;
          int cur = slide_oracle_score_one(p, ref7, n_ref, qry7, n_qry, cs[2 * a], cs[2 * a + 1],
                                           x, y, cur_ref, cur_qry);
          scored++;
          if (counts_out && (h - hyp_begin) < counts_cap) counts_out[h - hyp_begin] = cur;
          if (cur > best) { /* PR.cpp:361 strict */
            best = cur;
            set_rt(res->R_t, cs[2 * a], cs[2 * a + 1], x, y);
            res->best_hyp_index = h;
            res->n_matched = cur;
            if (ref_idx_out) memcpy(ref_idx_out, cur_ref, sizeof(int) * (size_t)cur);
            if (qry_idx_out) memcpy(qry_idx_out, cur_qry, sizeof(int) * (size_t)cur);
          }
        }
      }
    }
  }
  res->best_num_inliers = best;
  res->hypotheses_scored = scored;
  free(cs); free(cur_ref); free(cur_qry); free(L.yaw);
  return 0;
}

int slide_oracle_match_maps_mt(const slide_oracle_params *p, const double *ref7, int n_ref,
                               const double *qry7, int n_qry, double half_x, double half_y,
                               long long hyp_begin, long long hyp_end, int *ref_idx_out,
                               int *qry_idx_out, int n_threads,
                               slide_oracle_match_result *res) {
  return slide_oracle_match_maps_mt_counts(p, ref7, n_ref, qry7, n_qry, half_x, half_y, hyp_begin, hyp_end,
                                           ref_idx_out, qry_idx_out, n_threads, NULL, 0, res);
}

int slide_oracle_match_maps_mt_counts(const slide_oracle_params *p, const double *ref7, int n_ref,
                                      const double *qry7, int n_qry, double half_x, double half_y,
                                      long long hyp_begin, long long hyp_end, int *ref_idx_out,
                                      int *qry_idx_out, int n_threads, int *counts_out,
                                      long long counts_cap, slide_oracle_match_result *res) {
  result_init(res);
  int n_yaw = 0;
  long long nt = slide_oracle_enumerate_lattice(p, half_x, half_y, NULL, NULL, NULL, 0, NULL, 0, &n_yaw);
  lattice_hdr L;
  build_hdr(p, half_x, half_y, &L);
  res->n_rings = L.rings;
  res->n_yaw = n_yaw;
  if (nt < 0) { res->status = 1; return 1; }
  double *tx = (double *)malloc(sizeof(double) * (size_t)(nt > 0 ? nt : 1));
  double *ty = (double *)malloc(sizeof(double) * (size_t)(nt > 0 ? nt : 1));
  double *yaw = (double *)malloc(sizeof(double) * (size_t)(n_yaw > 0 ? n_yaw : 1));
  slide_oracle_enumerate_lattice(p, half_x, half_y, tx, ty, NULL, nt, yaw, n_yaw, &n_yaw);
  double *cs = (double *)malloc(sizeof(double) * 2 * (size_t)(n_yaw > 0 ? n_yaw : 1));
  for (int a = 0; a < n_yaw; a++) { cs[2 * a] = cos(yaw[a]); cs[2 * a + 1] = sin(yaw[a]); }
  long long total = nt * (long long)n_yaw;
  long long hb = hyp_begin < 0 ? 0 : hyp_begin;
  long long he = (hyp_end < 0 || hyp_end > total) ? total : hyp_end;
  if (n_yaw <= 0) { hb = 0; he = 0; }
  int best = -10000;
  long long best_h = -1, scored = 0;
#ifdef _OPENMP
  if (n_threads < 1) n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads)
  {
    int lbest = -10000;
    long long lbest_h = -1, lscored = 0;
#pragma omp for schedule(dynamic, 8) nowait
    for (long long h = hb; h < he; h++) {
      const long long t = h / n_yaw;
      const int a = (int)(h % n_yaw);
      int cur = slide_oracle_score_one(p, ref7, n_ref, qry7, n_qry, cs[2 * a], cs[2 * a + 1],
                                       tx[t], ty[t], NULL, NULL);
      lscored++;
      if (counts_out && (h - hb) < counts_cap) counts_out[h - hb] = cur;
      if (cur > lbest || (cur == lbest && h < lbest_h)) { lbest = cur; lbest_h = h; }
    }
#pragma omp critical
    {
      scored += lscored;
      if (lbest_h >= 0 && (lbest > best || (lbest == best && lbest_h < best_h))) {
        best = lbest; best_h = lbest_h;
      }
    }
  }
  res->hypotheses_scored = scored;
  res->best_num_inliers = best;
  res->best_hyp_index = best_h;
  if (best_h >= 0) {
    long long t = best_h / n_yaw;
    int a = (int)(best_h % n_yaw);
    set_rt(res->R_t, cs[2 * a], cs[2 * a + 1], tx[t], ty[t]);
    res->n_matched = slide_oracle_score_one(p, ref7, n_ref, qry7, n_qry, cs[2 * a], cs[2 * a + 1],
                                            tx[t], ty[t], ref_idx_out, qry_idx_out);
  }
  free(tx); free(ty); free(yaw); free(cs);
  return 0;
}

/* ------------------------------------------------------------------ small linear algebra */

typedef struct { double c, s; } jrot; /* J = [[c, s], [-s, c]] */

static void rot_rows(double *M, int n, int p, int q, jrot j) { /* rows p,q <- J * rows */
  for (int i = 0; i < n; i++) {
    double x = M[p * n + i], y = M[q * n + i];
    M[p * n + i] = j.c * x + j.s * y;
    M[q * n + i] = -j.s * x + j.c * y;
  }
}
static void rot_cols(double *M, int n, int p, int q, jrot j) { /* cols p,q <- cols * J */
  for (int i = 0; i < n; i++) {
    double x = M[i * n + p], y = M[i * n + q];
    M[i * n + p] = j.c * x - j.s * y;
    M[i * n + q] = j.s * x + j.c * y;
  }
}
static jrot rot_mul(jrot a, jrot b) { jrot r = {a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c}; return r; }
static jrot rot_T(jrot a) { jrot r = {a.c, -a.s}; return r; }

/* symmetric 2x2 [[x,y],[y,z]] -> rotation that diagonalises it */
static jrot sym_jacobi(double x, double y, double z) {
  jrot r = {1.0, 0.0};
  double deno = 2.0 * fabs(y);
  if (deno < DBL_MIN) return r;
  double tau = (x - z) / deno;
  double w = sqrt(tau * tau + 1.0);
  double t = tau > 0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
  double sign_t = t > 0 ? 1.0 : -1.0;
  double n = 1.0 / sqrt(t * t + 1.0);
  r.s = -sign_t * (y / fabs(y)) * fabs(t) * n;
  r.c = n;
  return r;
}

/* two-sided Jacobi SVD of an n x n (n <= 3) row-major matrix: A = U diag(S) V^T */
static void jacobi_svd(const double *A, int n, double *U, double *S, double *V) {
  double W[9];
  double scale = 0;
  for (int i = 0; i < n * n; i++) scale = fmax(scale, fabs(A[i]));
  if (scale == 0) scale = 1;
  for (int i = 0; i < n * n; i++) { W[i] = A[i] / scale; U[i] = V[i] = 0; }
  for (int i = 0; i < n; i++) U[i * n + i] = V[i * n + i] = 1;
  const double precision = 2.0 * DBL_EPSILON, tiny = DBL_MIN;
  double max_diag = 0;
  for (int i = 0; i < n; i++) max_diag = fmax(max_diag, fabs(W[i * n + i]));
  for (int sweep = 0; sweep < 200; sweep++) {
    int finished = 1;
    for (int p = 1; p < n; p++)
      for (int q = 0; q < p; q++) {
        double thr = fmax(tiny, precision * max_diag);
        if (fabs(W[p * n + q]) > thr || fabs(W[q * n + p]) > thr) {
          finished = 0;
          /* 2x2 block [[Wpp,Wpq],[Wqp,Wqq]]: make it symmetric, then diagonalise */
          double m00 = W[p * n + p], m01 = W[p * n + q], m10 = W[q * n + p], m11 = W[q * n + q];
          jrot rot1 = {1.0, 0.0};
          double t = m00 + m11, d = m10 - m01;
          if (fabs(d) >= DBL_MIN) {
            double u = t / d, tmp = sqrt(1.0 + u * u);
            rot1.s = 1.0 / tmp;
            rot1.c = u / tmp;
          }
          /* rows of the 2x2 <- rot1 * rows */
          double a00 = rot1.c * m00 + rot1.s * m10, a01 = rot1.c * m01 + rot1.s * m11;
          double a11 = -rot1.s * m01 + rot1.c * m11;
          jrot jr = sym_jacobi(a00, a01, a11);
          jrot jl = rot_mul(rot1, rot_T(jr));
          rot_rows(W, n, p, q, jl);
          rot_cols(U, n, p, q, rot_T(jl));
          rot_cols(W, n, p, q, jr);
          rot_cols(V, n, p, q, jr);
          max_diag = fmax(max_diag, fmax(fabs(W[p * n + p]), fabs(W[q * n + q])));
        }
      }
    if (finished) break;
  }
  for (int i = 0; i < n; i++) {
    double a = W[i * n + i];
    S[i] = fabs(a) * scale;
    if (a < 0) for (int r = 0; r < n; r++) U[r * n + i] = -U[r * n + i];
  }
  for (int i = 0; i < n; i++) { /* descending order */
    int pos = i;
    for (int k = i + 1; k < n; k++) if (S[k] > S[pos]) pos = k;
    if (S[pos] == 0) break;
    if (pos != i) {
      double ts = S[i]; S[i] = S[pos]; S[pos] = ts;
      for (int r = 0; r < n; r++) {
        double tu = U[r * n + i]; U[r * n + i] = U[r * n + pos]; U[r * n + pos] = tu;
        double tv = V[r * n + i]; V[r * n + i] = V[r * n + pos]; V[r * n + pos] = tv;
      }
    }
  }
}

void slide_oracle_svd3(const double *A, double *U, double *S, double *V) { jacobi_svd(A, 3, U, S, V); }

static double det3(const double *R) {
  return R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
         R[2] * (R[3] * R[7] - R[4] * R[6]);
}

static void mat_mul_nt(const double *A, const double *B, int n, double *C) { /* C = A * B^T */
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double acc = 0;
      for (int k = 0; k < n; k++) acc += A[i * n + k] * B[j * n + k];
      C[i * n + j] = acc;
    }
}

static void get_xyz_yaw(const double *tf16, double *xyz_yaw4) { /* PR.cpp:697-711 */
  xyz_yaw4[0] = tf16[3];
  xyz_yaw4[1] = tf16[7];
  xyz_yaw4[2] = tf16[11];
  xyz_yaw4[3] = atan2(tf16[4], tf16[0]);
}

void slide_oracle_solve_lsq(const double *tgt3, const double *src3, int k, double *xyz_yaw4,
                            double *transform16) { /* PR.cpp:632-695 */
  double cs[3] = {0, 0, 0}, ct[3] = {0, 0, 0};
  for (int i = 0; i < k; i++)
    for (int d = 0; d < 3; d++) { cs[d] += src3[3 * i + d]; ct[d] += tgt3[3 * i + d]; }
  for (int d = 0; d < 3; d++) { cs[d] /= (double)k; ct[d] /= (double)k; }  /* :655-658 */
  double H[9] = {0};
  for (int i = 0; i < k; i++)                                             /* :665-671 */
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
        H[a * 3 + b] += (src3[3 * i + a] - cs[a]) * (tgt3[3 * i + b] - ct[b]);
  double U[9], S[3], V[9], R[9];
  jacobi_svd(H, 3, U, S, V);                                              /* :674-675 */
  mat_mul_nt(V, U, 3, R);                                                 /* :678 R = V U^T */
  if (det3(R) < 0) {                                                      /* :680-686 */
    double U2[9], S2[3], V2[9];
    jacobi_svd(R, 3, U2, S2, V2);
    for (int r = 0; r < 3; r++) V2[r * 3 + 2] = -V2[r * 3 + 2];
    mat_mul_nt(V2, U2, 3, R);
  }
  double t[3];
  for (int a = 0; a < 3; a++)                                             /* :689 */
    t[a] = ct[a] - (R[a * 3 + 0] * cs[0] + R[a * 3 + 1] * cs[1] + R[a * 3 + 2] * cs[2]);
  for (int i = 0; i < 16; i++) transform16[i] = 0;
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) transform16[a * 4 + b] = R[a * 3 + b];
    transform16[a * 4 + 3] = t[a];
  }
  transform16[15] = 1;
  get_xyz_yaw(transform16, xyz_yaw4);
}

/* ------------------------------------------------------------------ findTransformation */

static void mat4_mul(const double *A, const double *B, double *C) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double acc = 0;
      for (int k = 0; k < 4; k++) acc += A[i * 4 + k] * B[k * 4 + j];
      C[i * 4 + j] = acc;
    }
}

int slide_oracle_find_transformation(const slide_oracle_params *p, const double *ref7_in,
                                     int n_ref, const double *qry7_in, int n_qry,
                                     int *ref_idx_out, int *qry_idx_out, int n_threads,
                                     slide_oracle_tf_result *res) {
  memset(res, 0, sizeof(*res));
  size_t rb = sizeof(double) * 7 * (size_t)(n_ref > 0 ? n_ref : 1);
  size_t qb = sizeof(double) * 7 * (size_t)(n_qry > 0 ? n_qry : 1);
  double *ref7 = (double *)malloc(rb), *qry7 = (double *)malloc(qb);
  memcpy(ref7, ref7_in, sizeof(double) * 7 * (size_t)n_ref);
  memcpy(qry7, qry7_in, sizeof(double) * 7 * (size_t)n_qry);
  double cref[2] = {0, 0}, cqry[2] = {0, 0};
  double half_x, half_y;
  slide_oracle_params pp = *p;
  if (p->inter_loop_closure) {
    for (int i = 0; i < n_ref; i++) { cref[0] += ref7[7 * i + 1]; cref[1] += ref7[7 * i + 2]; } /* :713-722 */
    cref[0] /= (double)n_ref; cref[1] /= (double)n_ref;
    for (int i = 0; i < n_qry; i++) { cqry[0] += qry7[7 * i + 1]; cqry[1] += qry7[7 * i + 2]; }
    cqry[0] /= (double)n_qry; cqry[1] /= (double)n_qry;
    for (int i = 0; i < n_ref; i++) { ref7[7 * i + 1] -= cref[0]; ref7[7 * i + 2] -= cref[1]; } /* :755-765 */
    for (int i = 0; i < n_qry; i++) { qry7[7 * i + 1] -= cqry[0]; qry7[7 * i + 2] -= cqry[1]; }
    double bx_r = 0, by_r = 0, bx_q = 0, by_q = 0;                 /* :724-734 */
    for (int i = 0; i < n_ref; i++) { bx_r = fmax(bx_r, fabs(ref7[7 * i + 1])); by_r = fmax(by_r, fabs(ref7[7 * i + 2])); }
    for (int i = 0; i < n_qry; i++) { bx_q = fmax(bx_q, fabs(qry7[7 * i + 1])); by_q = fmax(by_q, fabs(qry7[7 * i + 2])); }
    double max_x = fmax(bx_r, bx_q), max_y = fmax(by_r, by_q);     /* :771-774 */
    if (!p->disable_yaw_search) { double m = fmax(max_x, max_y); max_x = m; max_y = m; } /* :777-782 */
    half_x = max_x * p->dilation_factor;                           /* :786-787 */
    half_y = max_y * p->dilation_factor;
    res->yaw_half = p->match_yaw_half_range;
  } else {
    half_x = p->match_x_half_range_intra;                          /* :808-810 */
    half_y = p->match_y_half_range_intra;
    res->yaw_half = p->match_yaw_half_range_intra;
  }
  res->half_x = half_x; res->half_y = half_y;
  res->centroid_ref[0] = cref[0]; res->centroid_ref[1] = cref[1];
  res->centroid_qry[0] = cqry[0]; res->centroid_qry[1] = cqry[1];

  slide_oracle_match_result mr;
  int *ri = ref_idx_out, *qi = qry_idx_out, own = 0;
  if (!ri || !qi) {
    own = 1;
    ri = (int *)malloc(sizeof(int) * (size_t)(n_qry > 0 ? n_qry : 1));
    qi = (int *)malloc(sizeof(int) * (size_t)(n_qry > 0 ? n_qry : 1));
  }
  if (n_threads > 1 || n_threads < 0)
    slide_oracle_match_maps_mt(&pp, ref7, n_ref, qry7, n_qry, half_x, half_y, 0, -1, ri, qi, n_threads, &mr);
  else
    slide_oracle_match_maps(&pp, ref7, n_ref, qry7, n_qry, half_x, half_y, 0, -1, ri, qi, NULL, 0, &mr);
  res->match_status = mr.status;
  res->hypotheses_scored = mr.hypotheses_scored;
  res->best_hyp_index = mr.best_hyp_index;
  memcpy(res->R_t, mr.R_t, sizeof(mr.R_t));
  /* findTransformation initialises best_num_inliers_out = 0 and MatchMaps leaves it alone on
   * the sanity-check return (PR.cpp:819, :169-175) */
  res->best_num_inliers = mr.status == 1 ? 0 : mr.best_num_inliers;
  res->n_matched = mr.status == 1 ? 0 : mr.n_matched;
  int found = 0;
  if (!(res->best_num_inliers < p->min_num_inliers)) {             /* :849 */
    found = 1;
    if (!p->use_lsq) {                                             /* :882-905 */
      double raw[16] = {0};
      raw[0] = mr.R_t[0]; raw[1] = mr.R_t[1]; raw[4] = mr.R_t[3]; raw[5] = mr.R_t[4];
      raw[10] = 1; raw[15] = 1;
      raw[3] = mr.R_t[2]; raw[7] = mr.R_t[5]; raw[11] = 0;
      if (p->inter_loop_closure) {                                 /* :947-967 */
        double H1[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1}, H2[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};
        H1[3] = cref[0]; H1[7] = cref[1];
        H2[3] = -cqry[0]; H2[7] = -cqry[1];
        double T1[16];
        mat4_mul(H1, raw, T1);
        mat4_mul(T1, H2, res->transform);
      } else {
        memcpy(res->transform, raw, sizeof(raw));
      }
      get_xyz_yaw(res->transform, res->xyz_yaw);
    } else {                                                       /* :906-944 */
      int k = res->n_matched;
      double *tgt = (double *)malloc(sizeof(double) * 3 * (size_t)(k > 0 ? k : 1));
      double *src = (double *)malloc(sizeof(double) * 3 * (size_t)(k > 0 ? k : 1));
      for (int m = 0; m < k; m++) {
        const double *r = ref7 + 7 * (size_t)ri[m];
        const double *q = qry7 + 7 * (size_t)qi[m];
        tgt[3 * m] = r[1]; tgt[3 * m + 1] = r[2]; tgt[3 * m + 2] = r[3];
        src[3 * m] = q[1]; src[3 * m + 1] = q[2]; src[3 * m + 2] = q[3];
        if (p->inter_loop_closure) {                               /* :925-937 */
          tgt[3 * m] += cref[0]; tgt[3 * m + 1] += cref[1];
          src[3 * m] += cqry[0]; src[3 * m + 1] += cqry[1];
        }
      }
      slide_oracle_solve_lsq(tgt, src, k, res->xyz_yaw, res->transform);
      free(tgt); free(src);
    }
  }
  res->found = found;
  if (own) { free(ri); free(qi); }
  free(ref7); free(qry7);
  return found;
}

int slide_oracle_find_inter_loop_closure(const slide_oracle_params *p, const double *ref7,
                                         int n_ref, const double *qry7, int n_qry,
                                         int n_threads, double *tf16,
                                         slide_oracle_tf_result *res_out) { /* PR.cpp:498-538 */
  slide_oracle_tf_result local, *res = res_out ? res_out : &local;
  memset(res, 0, sizeof(*res));
  int found = 0;
  if (!(n_ref < p->min_num_map_objects_to_start || n_qry < p->min_num_map_objects_to_start))
    found = slide_oracle_find_transformation(p, ref7, n_ref, qry7, n_qry, NULL, NULL, n_threads, res);
  if (!found) return 0;
  double x = res->xyz_yaw[0], y = res->xyz_yaw[1], z = res->xyz_yaw[2], yaw = res->xyz_yaw[3];
  for (int i = 0; i < 16; i++) tf16[i] = 0;
  tf16[0] = cos(yaw); tf16[1] = -sin(yaw); tf16[4] = sin(yaw); tf16[5] = cos(yaw);
  tf16[10] = 1; tf16[15] = 1;
  tf16[3] = x; tf16[7] = y; tf16[11] = z;
  return 1;
}

static int mat4_rigid_inverse(const double *A, double *Ainv) { /* Sophus SE3::inverse(): [R^T, -R^T t] */
  double T[16] = {0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[i * 4 + j] = A[j * 4 + i];
  for (int i = 0; i < 3; i++) T[i * 4 + 3] = -(T[i * 4] * A[3] + T[i * 4 + 1] * A[7] + T[i * 4 + 2] * A[11]);
  T[15] = 1;
  memcpy(Ainv, T, sizeof(T));
  return 1;
}

int slide_oracle_find_intra_loop_closure(const slide_oracle_params *p, const double *meas7, int n_meas,
                                         const double *submap7, int n_sub, const double *query_pose16,
                                         const double *candidate_pose16, int n_threads, double *tf16,
                                         slide_oracle_tf_result *res_out) { /* PR.cpp:389-496 */
  slide_oracle_tf_result local, *res = res_out ? res_out : &local;
  memset(res, 0, sizeof(*res));
  if (n_meas == 0 || n_sub == 0) return 0;                         /* :395-398 */
  if (n_meas < 4) return 0;                                        /* :400-403 */
  double *moved = (double *)malloc(sizeof(double) * 7 * (size_t)n_meas);
  const double *P = query_pose16;
  for (int i = 0; i < n_meas; i++) {                               /* :421-439 */
    const double *m = meas7 + 7 * (size_t)i;
    double v[4];
    /* Matrix4d * Vector4d, all sizes fixed: Eigen's packet product accumulates column by column */
    for (int r = 0; r < 4; r++) v[r] = ((P[r * 4] * m[1] + P[r * 4 + 1] * m[2]) + P[r * 4 + 2] * m[3]) + P[r * 4 + 3] * 1.0;
    double *o = moved + 7 * (size_t)i;
    o[0] = m[0]; o[1] = v[0] / v[3]; o[2] = v[1] / v[3]; o[3] = v[2] / v[3];
    o[4] = m[4]; o[5] = m[5]; o[6] = m[6];
  }
  slide_oracle_params pp = *p;
  pp.inter_loop_closure = 0;  /* the intra instance is constructed with inter_loop_closure = false (sloamNode.cpp:23) */
  int found = slide_oracle_find_transformation(&pp, submap7, n_sub, moved, n_meas, NULL, NULL, n_threads, res); /* :446 */
  free(moved);
  if (!found) return 0;                                            /* :449-453 */
  double yaw = res->xyz_yaw[3];
  double lc[16] = {0};                                             /* :455-470, z forced to 0 (:466) */
  lc[0] = cos(yaw); lc[1] = -sin(yaw); lc[4] = sin(yaw); lc[5] = cos(yaw);
  lc[10] = 1; lc[15] = 1;
  lc[3] = res->xyz_yaw[0]; lc[7] = res->xyz_yaw[1]; lc[11] = 0.0;
  double cinv[16], drift[16];
  mat4_rigid_inverse(candidate_pose16, cinv);
  mat4_mul(cinv, query_pose16, drift);                             /* :478 candidate^-1 * query */
  mat4_mul(drift, lc, tf16);                                       /* :483-494 */
  return 1;
}

/* ------------------------------------------------------------------ SlideGraph descriptor half */

void slide_oracle_triangle_descriptor(const double *t, double *desc3, int *perm3) { /* SC.cpp:66-90 */
  /* rowwise().mean() of a 2x3: (a+b+c)/3 with Eigen's 3-term redux a + (b + c)?  The fixed-size
   * redux of 3 splits as (a) + (b + c) in the non-vectorised unroller; the contract adopted
   * here is left-to-right (a + b) + c, FMA-free (the target is built with -mfma, so the
   * reference binary itself is not pinned at the ulp level -- SURVEY.md section 8c). */
  double cx = ((t[0] + t[2]) + t[4]) / 3.0, cy = ((t[1] + t[3]) + t[5]) / 3.0;
  double d[3];
  for (int i = 0; i < 3; i++) {
    double dx = t[2 * i] - cx, dy = t[2 * i + 1] - cy;
    d[i] = sqrt(dx * dx + dy * dy);
  }
  int idx[3] = {0, 1, 2}; /* argsort ascending; ties keep the lower index first */
  for (int a = 1; a < 3; a++)
    for (int b = a; b > 0 && d[idx[b]] < d[idx[b - 1]]; b--) { int tmp = idx[b]; idx[b] = idx[b - 1]; idx[b - 1] = tmp; }
  for (int i = 0; i < 3; i++) { desc3[i] = d[idx[i]]; perm3[i] = idx[i]; }
}

long long slide_oracle_match_triangles(const double *tm, int t_model, const double *td, int t_data,
                                       double threshold, int *mi, int *di, double *diff_out,
                                       long long cap) { /* SC.cpp:111-118 over SC.cpp:49-108 */
  double *dm = (double *)malloc(sizeof(double) * 3 * (size_t)(t_model > 0 ? t_model : 1));
  double *dd = (double *)malloc(sizeof(double) * 3 * (size_t)(t_data > 0 ? t_data : 1));
  int perm[3];
  for (int i = 0; i < t_model; i++) slide_oracle_triangle_descriptor(tm + 6 * (size_t)i, dm + 3 * (size_t)i, perm);
  for (int j = 0; j < t_data; j++) slide_oracle_triangle_descriptor(td + 6 * (size_t)j, dd + 3 * (size_t)j, perm);
  long long n = 0;
  for (int i = 0; i < t_model; i++)
    for (int j = 0; j < t_data; j++) {
      double diff = 0;
      for (int k = 0; k < 3; k++) { double e = dm[3 * i + k] - dd[3 * j + k]; diff += e * e; } /* pow(.,2) */
      diff = sqrt(diff);
      if (diff < threshold) {
        if (n < cap) { if (mi) mi[n] = i; if (di) di[n] = j; if (diff_out) diff_out[n] = diff; }
        n++;
      }
    }
  free(dm); free(dd);
  return n;
}

void slide_oracle_estimate_tf(const double *a, const double *b, int k, double *tf9) { /* SC.cpp:122-138 */
  double ca[2] = {0, 0}, cb[2] = {0, 0};
  for (int i = 0; i < k; i++) { ca[0] += a[2 * i]; ca[1] += a[2 * i + 1]; cb[0] += b[2 * i]; cb[1] += b[2 * i + 1]; }
  ca[0] /= (double)k; ca[1] /= (double)k; cb[0] /= (double)k; cb[1] /= (double)k;
  double H[4] = {0, 0, 0, 0};
  for (int i = 0; i < k; i++) {
    double ax = a[2 * i] - ca[0], ay = a[2 * i + 1] - ca[1];
    double bx = b[2 * i] - cb[0], by = b[2 * i + 1] - cb[1];
    H[0] += ax * bx; H[1] += ax * by; H[2] += ay * bx; H[3] += ay * by;
  }
  double U[4], S[2], V[4], R[4];
  jacobi_svd(H, 2, U, S, V);
  mat_mul_nt(V, U, 2, R);
  if (R[0] * R[3] - R[1] * R[2] < 0) { R[1] = -R[1]; R[3] = -R[3]; } /* R.col(1) *= -1 */
  double tx = cb[0] - (R[0] * ca[0] + R[1] * ca[1]);
  double ty = cb[1] - (R[2] * ca[0] + R[3] * ca[1]);
  tf9[0] = R[0]; tf9[1] = R[1]; tf9[2] = tx;
  tf9[3] = R[2]; tf9[4] = R[3]; tf9[5] = ty;
  tf9[6] = 0; tf9[7] = 0; tf9[8] = 1;
}
