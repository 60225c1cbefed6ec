/*
 * clipper_oracle.c -- CPU ORACLE for the SlideGraph half: CLIPPER affinity scoring and dense-clique
 * solver, and the run_semantic_clipper pipeline.  TEST INFRASTRUCTURE (see slide_oracle.h): only
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs may call it.
 *
 * Plain-C restatement (fp64, dense matrices, the reference's loop structure) of
 *   clipper::invariants::EuclideanDistance::operator()   CSO/src/invariants/euclidean_distance.cpp:13-30
 *   clipper::utils::createAllToAll / k2ij / findIndicesOfkLargest / findIndicesWhereAboveThreshold /
 *                   selectFromIndicator                   CSO/include/clipper/utils.h:60-70, CSO/src/utils.cpp:34-104
 *   clipper::CLIPPER::scorePairwiseConsistency            CSO/src/clipper.cpp:21-65
 *   clipper::CLIPPER::findDenseClique                     CSO/src/clipper.cpp:172-323
 *   clipper::CLIPPER::getAffinityMatrix / getSelectedAssociations   clipper.cpp:121-135
 *   clipper::dsd::solve (Goldberg's densest subgraph)     CSO/src/dsd.cpp:21-327
 *   semantic_clipper::run_semantic_clipper (after the triangulation)   CSO/src/semantic_clipper.cpp:140-275
 * with CSO = backend/sloam/clipper_semantic_object.
 *
 * PARITY PIN: this half is pinned by the reference's OWN known-answer tests --
 * CSO/test/affinity_test.cpp:93-107 (the 12 x 12 affinity matrix `Mtrue`),
 * CSO/test/clipper_test.cpp:15-68 (the selected clique), CSO/test/dsd_test.cpp:15-80 (the densest
 * subgraph of a 20-node graph) -- all reproduced in tests/test_clipper_oracle.py.  Unpinned: the
 * order in which Eigen accumulates its sparse selfadjoint products (the solver's iterates agree to
 * rounding, its selected nodes exactly), and clipper.solve()'s random u0 (std::random_device,
 * CSO/src/utils.cpp:22-29): the oracle takes u0 as an argument.
 */
#include "clipper_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

void clipper_oracle_default_params(clipper_oracle_params *p) { /* clipper.h:28-60, euclidean_distance.h:27-30 */
  p->sigma = 0.01; p->epsilon = 0.06; p->mindist = 0;
  p->tol_u = 1e-8; p->tol_F = 1e-9; p->tol_Fop = 1e-10;
  p->maxiniters = 200; p->maxoliters = 1000;
  p->beta = 0.25; p->maxlsiters = 99;
  p->eps = 1e-9; p->affinityeps = 1e-4;
  p->rescale_u0 = 1;
  p->rounding = CLIPPER_ROUND_DSD_HEU;
}

/* euclidean_distance.cpp:13-30.  Points are columns of a dim x n column-major matrix (Eigen's
 * invariants::Data).  (ai - aj).norm(): squares summed left to right, then sqrt. */
static double point_dist(const double *D, int dim, int i, int j) {
  double s = 0;
  for (int k = 0; k < dim; k++) { double e = D[(size_t)dim * i + k] - D[(size_t)dim * j + k]; s += e * e; }
  return sqrt(s);
}

double clipper_oracle_invariant(const clipper_oracle_params *p, const double *D1, const double *D2, int dim,
                                int a_i, int a_j, int b_i, int b_j) {
  const double l1 = point_dist(D1, dim, a_i, a_j);
  const double l2 = point_dist(D2, dim, b_i, b_j);
  if (p->mindist > 0 && (l1 < p->mindist || l2 < p->mindist)) return 0.0;      /* :22-24 */
  const double c = fabs(l1 - l2);                                                /* :27 */
  return (c < p->epsilon) ? exp(-0.5 * c * c / (p->sigma * p->sigma)) : 0;       /* :29 */
}

void clipper_oracle_all_to_all(int n1, int n2, int *A) { /* utils.h:60-70 */
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < n2; j++) { A[2 * (j + i * n2)] = i; A[2 * (j + i * n2) + 1] = j; }
}

void clipper_oracle_k2ij(long long k, long long n, long long *i_out, long long *j_out) { /* utils.cpp:81-92 */
  k += 1;
  const long long l = n * (n - 1) / 2 - k;
  const long long o = (long long)floor((sqrt(1 + 8 * (double)l) - 1) / 2.);
  const long long p = l - o * (o + 1) / 2;
  const long long i = n - (o + 1);
  const long long j = n - p;
  *i_out = i - 1; *j_out = j - 1;
}

/* clipper.cpp:21-65.  M: m x m row-major, only the strict upper triangle is written (the
 * reference's M_ = M.sparseView() of exactly that).  Returns the number of non-zeros. */
long long clipper_oracle_score_pairwise(const clipper_oracle_params *p, const double *D1, int n1, const double *D2, int n2,
                                        int dim, const int *A, int m, double *M) {
  (void)n1; (void)n2;
  memset(M, 0, sizeof(double) * (size_t)m * (size_t)m);
  long long nnz = 0;
  const long long n_pairs = (long long)m * (m - 1) / 2;
  for (long long k = 0; k < n_pairs; k++) {                                      /* :31-32 */
    long long i, j;
    clipper_oracle_k2ij(k, m, &i, &j);
    if (A[2 * i] == A[2 * j] || A[2 * i + 1] == A[2 * j + 1]) continue;          /* :34-37 distinctness */
    const double scr = clipper_oracle_invariant(p, D1, D2, dim, A[2 * i], A[2 * j], A[2 * i + 1], A[2 * j + 1]);
    if (scr > p->affinityeps) { M[(size_t)i * m + j] = scr; nnz++; }              /* :52-54 */
  }
  return nnz;
}

void clipper_oracle_affinity_matrix(const double *Mupper, int m, double *Mfull) { /* clipper.cpp:121-126 */
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++)
      Mfull[(size_t)i * m + j] = i == j ? 1.0 : (i < j ? Mupper[(size_t)i * m + j] : Mupper[(size_t)j * m + i]);
}

/* y = selfadjointView<Upper>(M) * u  (no diagonal stored); `ones`: the constraint matrix C_ = pattern of M_ */
static void sym_mv(const double *M, int n, const double *u, double *y, int ones) {
  for (int i = 0; i < n; i++) y[i] = 0;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      const double v = M[(size_t)i * n + j];
      if (v == 0) continue;
      const double w = ones ? 1.0 : v;
      y[i] += w * u[j];
      y[j] += w * u[i];
    }
}

static double vsum(const double *x, int n) { double s = 0; for (int i = 0; i < n; i++) s += x[i]; return s; }
static double vdot(const double *x, const double *y, int n) { double s = 0; for (int i = 0; i < n; i++) s += x[i] * y[i]; return s; }

/* idxD / num / den / mean of clipper.cpp:205-212 and :278-286 */
static int homotopy_step(const clipper_oracle_params *p, const double *M, int n, const double *u, double *Cbu, double *Mu,
                         double *tmp, int use_abs, double *out) {
  const double su = vsum(u, n);
  sym_mv(M, n, u, tmp, 1);
  for (int i = 0; i < n; i++) Cbu[i] = su - tmp[i] - u[i];                       /* ones*u.sum() - C*u - u */
  int cnt = 0;
  for (int i = 0; i < n; i++) cnt += (Cbu[i] > p->eps && u[i] > p->eps);
  if (cnt == 0) return 0;
  sym_mv(M, n, u, Mu, 0);
  double acc = 0;
  for (int i = 0; i < n; i++)
    if (Cbu[i] > p->eps && u[i] > p->eps) {
      const double q = (Mu[i] + u[i]) / Cbu[i];
      acc += use_abs ? fabs(q) : q;
    }
  *out = acc / (double)cnt;
  return 1;
}

static void grad(const double *M, int n, const double *u, double d, double *g, double *t1, double *t2) {
  /* (1 + d) * u - d * ones * u.sum() + M * u + C * u * d   (clipper.cpp:222, 244-247) */
  const double su = vsum(u, n);
  sym_mv(M, n, u, t1, 0);
  sym_mv(M, n, u, t2, 1);
  for (int i = 0; i < n; i++) g[i] = (1 + d) * u[i] - d * su + t1[i] + t2[i] * d;
}

/* utils.cpp:34-55: indices of the k largest entries, via a min-heap keyed (value, index); the output
 * is filled from the back, so it lists the entries in DESCENDING (value, index) order. */
typedef struct { double v; int i; } heap_item;
static int item_gt(heap_item a, heap_item b) { return a.v > b.v || (a.v == b.v && a.i > b.i); } /* std::greater<pair> */
static void heap_push(heap_item *h, int *n, heap_item x) {
  int k = (*n)++;
  h[k] = x;
  while (k > 0) { int par = (k - 1) / 2; if (item_gt(h[par], h[k])) { heap_item t = h[par]; h[par] = h[k]; h[k] = t; k = par; } else break; }
}
static heap_item heap_pop(heap_item *h, int *n) {
  heap_item top = h[0];
  h[0] = h[--(*n)];
  int k = 0;
  for (;;) {
    int l = 2 * k + 1, r = l + 1, s = k;
    if (l < *n && item_gt(h[s], h[l])) s = l;
    if (r < *n && item_gt(h[s], h[r])) s = r;
    if (s == k) break;
    heap_item t = h[s]; h[s] = h[k]; h[k] = t; k = s;
  }
  return top;
}

int clipper_oracle_k_largest(const double *x, int n, int k, int *idx_out) {
  if (k < 1) return 0;
  heap_item *h = (heap_item *)malloc(sizeof(heap_item) * (size_t)(k + 1));
  int hn = 0;
  for (int i = 0; i < n; i++) {
    heap_item it = {x[i], i};
    if (hn < k) heap_push(h, &hn, it);
    else if (h[0].v < x[i]) { heap_pop(h, &hn); heap_push(h, &hn, it); }
  }
  /* the reference pops k times even when fewer than k entries exist (undefined behaviour there):
   * the oracle returns what exists */
  const int got = hn;
  for (int i = 0; i < got; i++) idx_out[got - i - 1] = heap_pop(h, &hn).i;
  free(h);
  return got;
}

/* clipper.cpp:172-323 on the upper-triangular affinity M (n x n row-major, zero diagonal). */
int clipper_oracle_find_dense_clique(const clipper_oracle_params *p, const double *M, int n, const double *u0,
                                     clipper_oracle_solution *sol, int *nodes_out, double *u_out) {
  memset(sol, 0, sizeof(*sol));
  if (n <= 0) return 0;
  double *buf = (double *)malloc(sizeof(double) * 9 * (size_t)n);
  double *gradF = buf, *gradFnew = buf + n, *u = buf + 2 * n, *unew = buf + 3 * n, *Mu = buf + 4 * n, *Cbu = buf + 5 * n,
         *t1 = buf + 6 * n, *t2 = buf + 7 * n, *t3 = buf + 8 * n;
  if (p->rescale_u0) {                                                           /* :195-199 */
    sym_mv(M, n, u0, t1, 0);
    for (int i = 0; i < n; i++) u[i] = t1[i] + u0[i];
  } else {
    memcpy(u, u0, sizeof(double) * (size_t)n);
  }
  { const double nu = sqrt(vdot(u, u, n)); for (int i = 0; i < n; i++) u[i] /= nu; }  /* :200 */
  double d = 0;                                                                  /* :203-212 */
  { double v; if (homotopy_step(p, M, n, u, Cbu, Mu, t3, 0, &v)) d = v; }
  double F = 0;
  int i_out;
  for (i_out = 0; i_out < p->maxoliters; ++i_out) {                              /* :221 */
    grad(M, n, u, d, gradF, t1, t2);
    F = vdot(u, gradF, n);                                                       /* :223 */
    for (int j = 0; j < p->maxiniters; ++j) {                                    /* :229 */
      double alpha = 1, Fnew = 0, deltaF = 0;
      for (int k = 0; k < p->maxlsiters; ++k) {                                  /* :237 */
        for (int q = 0; q < n; q++) { const double v = u[q] + alpha * gradF[q]; unew[q] = v > 0 ? v : 0; }  /* :238-239 */
        { const double nu = sqrt(vdot(unew, unew, n)); if (nu > 0) for (int q = 0; q < n; q++) unew[q] /= nu; }  /* :240 normalize() */
        grad(M, n, unew, d, gradFnew, t1, t2);                                   /* :241-244 */
        Fnew = vdot(unew, gradFnew, n);                                          /* :245 */
        deltaF = Fnew - F;                                                       /* :247 */
        if (deltaF < -p->eps) alpha = alpha * p->beta;                           /* :249-251 */
        else break;
      }
      double du = 0;
      for (int q = 0; q < n; q++) { const double e = unew[q] - u[q]; du += e * e; }
      const double deltau = sqrt(du);                                            /* :256 */
      F = Fnew;                                                                  /* :259-261 */
      memcpy(u, unew, sizeof(double) * (size_t)n);
      memcpy(gradF, gradFnew, sizeof(double) * (size_t)n);
      if (deltau < p->tol_u || fabs(deltaF) < p->tol_F) break;                   /* :264 */
    }
    double deltad;                                                               /* :271-287 */
    if (homotopy_step(p, M, n, u, Cbu, Mu, t3, 1, &deltad)) d += deltad;
    else break;
  }
  int n_nodes = 0;
  if (p->rounding == CLIPPER_ROUND_NONZERO) {                                    /* :297-299, utils.cpp:59-69 */
    for (int q = 0; q < n; q++) if (u[q] > 0.0) nodes_out[n_nodes++] = q;
  } else if (p->rounding == CLIPPER_ROUND_DSD) {                                 /* :301-307 */
    int *S = (int *)malloc(sizeof(int) * (size_t)n), ns = 0;
    for (int q = 0; q < n; q++) if (u[q] > 0.0) S[ns++] = q;
    n_nodes = clipper_oracle_dsd(M, n, S, ns, nodes_out);
    free(S);
  } else {                                                                       /* :309-316 DSD_HEU */
    const int omega = (int)round(F);
    n_nodes = clipper_oracle_k_largest(u, n, omega, nodes_out);
  }
  sol->ifinal = i_out;
  sol->score = F;
  sol->n_nodes = n_nodes;
  if (u_out) memcpy(u_out, u, sizeof(double) * (size_t)n);
  free(buf);
  return n_nodes;
}

/* ------------------------------------------------------------------ dsd.cpp (Goldberg / Dinic) */
typedef struct {
  long long nverts, nedges;
  long long *Q, *fin, *pro, *another_pro, *pro3, *dist, *next, *to, *cut;
  double *flow, *cap;
} flow_net;

static void new_edge(flow_net *N, long long u, long long v, double w, long *nEdge) { /* dsd.cpp:41-53 */
  N->to[*nEdge] = v; N->cap[*nEdge] = w; N->flow[*nEdge] = 0; N->next[*nEdge] = N->fin[u]; N->fin[u] = (*nEdge)++;
  N->to[*nEdge] = u; N->cap[*nEdge] = w; N->flow[*nEdge] = w; N->next[*nEdge] = N->fin[v]; N->fin[v] = (*nEdge)++;
}
static int dinic_bfs(flow_net *N, long long src, long long dest) { /* :57-78 */
  long long st, en;
  for (long long i = 0; i < N->nverts; i++) N->dist[i] = -1;
  N->dist[src] = st = en = 0;
  N->Q[en++] = src;
  while (st < en) {
    long long u = N->Q[st++];
    for (long long i = N->fin[u]; i >= 0; i = N->next[i]) {
      long long v = N->to[i];
      if (N->flow[i] < N->cap[i] && N->dist[v] == -1) { N->dist[v] = N->dist[u] + 1; N->Q[en++] = v; }
    }
  }
  return N->dist[dest] != -1;
}
static double dinic_dfs(flow_net *N, long long u, double fl, long long src, long long dest) { /* :82-104 */
  if (u == dest) return fl;
  for (long long *e = &N->pro[u]; *e >= 0; *e = N->next[*e]) {
    long long v = N->to[*e];
    if (N->flow[*e] < N->cap[*e] && N->dist[v] == N->dist[u] + 1) {
      if (u == src || (N->cap[*e] - N->flow[*e]) <= fl) fl = N->cap[*e] - N->flow[*e];
      double df = dinic_dfs(N, v, fl, src, dest);
      if (df > 0) { N->flow[*e] += df; N->flow[*e ^ 1] -= df; return df; }
    }
  }
  return 0;
}
static void find_cut(flow_net *N, long long u) { /* :108-118 */
  N->cut[u] = 1;
  for (long long *e = &N->another_pro[u]; *e >= 0; *e = N->next[*e]) {
    long long v = N->to[*e];
    if (N->flow[*e] < N->cap[*e] && N->cut[v] == 0) find_cut(N, v);
  }
}

/* dsd::solve(A, S) -> densest_subgraph (dsd.cpp:167-326).  M: upper-triangular n x n row-major. */
int clipper_oracle_dsd(const double *M, int n, const int *S_in, int ns_in, int *nodes_out) {
  int *S = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1)), ns = ns_in;
  if (ns_in > 0) memcpy(S, S_in, sizeof(int) * (size_t)ns_in);
  else { ns = n; for (int i = 0; i < n; i++) S[i] = i; }                         /* :283-289 */
  const long long m = (long long)ns * ns - ns;                                   /* :292 */
  const long long nverts = n + 2, nedges = m + 2 * (long long)n;
  double (*E)[3] = (double (*)[3])malloc(sizeof(double) * 3 * (size_t)(nedges > 0 ? nedges : 1));
  double *degree = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  long long ne = 0;
  for (int a = 0; a < ns; a++)                                                   /* :300-313 */
    for (int b = 0; b < ns; b++) {
      const int i = S[a], j = S[b];
      if (i == j) continue;
      const double ew = (i < j) ? M[(size_t)i * n + j] : M[(size_t)j * n + i];
      E[ne][0] = i; E[ne][1] = j; E[ne][2] = ew;
      degree[i] += ew;                                                           /* :183-186 */
      E[ne][0] += 1; E[ne][1] += 1;
      ne++;
    }
  flow_net N;
  N.nverts = nverts; N.nedges = nedges;
  N.Q = (long long *)malloc(sizeof(long long) * (size_t)nverts); N.fin = (long long *)malloc(sizeof(long long) * (size_t)nverts);
  N.pro = (long long *)malloc(sizeof(long long) * (size_t)nverts); N.another_pro = (long long *)malloc(sizeof(long long) * (size_t)nverts);
  N.pro3 = (long long *)malloc(sizeof(long long) * (size_t)nverts); N.dist = (long long *)malloc(sizeof(long long) * (size_t)nverts);
  N.cut = (long long *)malloc(sizeof(long long) * (size_t)nverts);
  N.flow = (double *)malloc(sizeof(double) * 2 * (size_t)nedges); N.cap = (double *)malloc(sizeof(double) * 2 * (size_t)nedges);
  N.next = (long long *)malloc(sizeof(long long) * 2 * (size_t)nedges); N.to = (long long *)malloc(sizeof(long long) * 2 * (size_t)nedges);
  long long *final_cut = (long long *)calloc((size_t)nverts, sizeof(long long));
  double L = 0, U = (double)(m / 2);                                             /* :191-192 (integer division) */
  while ((double)n * (n - 1) * (U - L) >= 1) {                                   /* :213 */
    const double g = (U + L) / 2;
    const long long src = 0, dest = nverts - 1;
    for (long long i = m; i < m + n; i++) { E[i][0] = (double)src; E[i][1] = (double)(i - m + 1); E[i][2] = (double)(m / 2); }       /* :21-37 */
    for (long long i = n + m; i < m + 2 * (long long)n; i++) {
      E[i][0] = (double)(i - m - n + 1); E[i][1] = (double)dest; E[i][2] = (double)(m / 2) + 2 * g - degree[i - m - n];
    }
    for (long long i = 0; i < nverts; i++) { N.fin[i] = -1; N.cut[i] = 0; }      /* :122-131 */
    long nEdge = 0;
    for (long long i = 0; i < nedges; i++) new_edge(&N, (long long)E[i][0], (long long)E[i][1], E[i][2], &nEdge);
    while (dinic_bfs(&N, src, dest)) {                                           /* :135-150 */
      for (long long i = 0; i < nverts; i++) { N.pro[i] = N.fin[i]; N.another_pro[i] = N.fin[i]; N.pro3[i] = N.fin[i]; }
      for (;;) { double df = dinic_dfs(&N, src, 0, src, dest); if (!df) break; }
    }
    find_cut(&N, src);
    long long cs = 0;
    for (long long i = 0; i < nverts; i++) cs += N.cut[i];
    if (cs == 1) U = g;                                                          /* :222-229 */
    else { L = g; memcpy(final_cut, N.cut, sizeof(long long) * (size_t)nverts); }
  }
  final_cut[0] = 0; final_cut[nverts - 1] = 0;
  int num = 0;
  for (long long i = 1; i < nverts - 1; i++) if (final_cut[i] != 0) nodes_out[num++] = (int)(i - 1);  /* :237-247 */
  free(S); free(E); free(degree); free(N.Q); free(N.fin); free(N.pro); free(N.another_pro); free(N.pro3); free(N.dist);
  free(N.cut); free(N.flow); free(N.cap); free(N.next); free(N.to); free(final_cut);
  return num;
}

/* ------------------------------------------------------------------ run_semantic_clipper (SC.cpp:140-275), triangles given */
int clipper_oracle_run_semantic_clipper(const double *tris_model6, int t_model, const double *tris_data6, int t_data,
                                        double sigma, double epsilon, int min_num_pairs, double matching_threshold,
                                        const double *u0, int u0_len, double *tf16, clipper_oracle_sc_info *info) {
  memset(info, 0, sizeof(*info));
  /* match_triangles (SC.cpp:111-118): model-major, data-minor; 3 points per match in sorted order (SC.cpp:102-105) */
  long long n_match = slide_oracle_match_triangles(tris_model6, t_model, tris_data6, t_data, matching_threshold, NULL, NULL, NULL, 0);
  int *mi = (int *)malloc(sizeof(int) * (size_t)(n_match > 0 ? n_match : 1)), *di = (int *)malloc(sizeof(int) * (size_t)(n_match > 0 ? n_match : 1));
  slide_oracle_match_triangles(tris_model6, t_model, tris_data6, t_data, matching_threshold, mi, di, NULL, n_match);
  const int m = (int)(3 * n_match);
  info->n_triangle_matches = n_match;
  info->n_associations = m;
  double *Pm = (double *)malloc(sizeof(double) * 2 * (size_t)(m > 0 ? m : 1)), *Pd = (double *)malloc(sizeof(double) * 2 * (size_t)(m > 0 ? m : 1));
  for (long long k = 0; k < n_match; k++) {
    double desc[3]; int pm[3], pd[3];
    slide_oracle_triangle_descriptor(tris_model6 + 6 * (size_t)mi[k], desc, pm);
    slide_oracle_triangle_descriptor(tris_data6 + 6 * (size_t)di[k], desc, pd);
    for (int v = 0; v < 3; v++) {
      Pm[2 * (3 * k + v)] = tris_model6[6 * (size_t)mi[k] + 2 * pm[v]]; Pm[2 * (3 * k + v) + 1] = tris_model6[6 * (size_t)mi[k] + 2 * pm[v] + 1];
      Pd[2 * (3 * k + v)] = tris_data6[6 * (size_t)di[k] + 2 * pd[v]];  Pd[2 * (3 * k + v) + 1] = tris_data6[6 * (size_t)di[k] + 2 * pd[v] + 1];
    }
  }
  int found = 0;
  for (int i = 0; i < 16; i++) tf16[i] = (i % 5 == 0) ? 1.0 : 0.0;
  if (m > 0 && u0_len >= m) {
    int *A = (int *)malloc(sizeof(int) * 2 * (size_t)m);
    for (int i = 0; i < m; i++) { A[2 * i] = i; A[2 * i + 1] = i; }              /* SC.cpp:203-207 */
    clipper_oracle_params p;
    clipper_oracle_default_params(&p);
    p.sigma = sigma; p.epsilon = epsilon;                                        /* SC.cpp:210-212 */
    double *M = (double *)malloc(sizeof(double) * (size_t)m * (size_t)m);
    info->nnz = clipper_oracle_score_pairwise(&p, Pm, m, Pd, m, 2, A, m, M);     /* SC.cpp:224 */
    int *nodes = (int *)malloc(sizeof(int) * (size_t)m);
    clipper_oracle_solution sol;
    int nn = clipper_oracle_find_dense_clique(&p, M, m, u0, &sol, nodes, NULL);  /* SC.cpp:227 */
    info->n_inliers = nn;
    info->score = sol.score;
    if (!(nn < min_num_pairs)) {                                                 /* SC.cpp:249-255 */
      double *a = (double *)malloc(sizeof(double) * 2 * (size_t)(nn > 0 ? nn : 1)), *b = (double *)malloc(sizeof(double) * 2 * (size_t)(nn > 0 ? nn : 1));
      for (int k = 0; k < nn; k++) {                                             /* SC.cpp:238-246: Ainliers rows are (node, node) */
        a[2 * k] = Pm[2 * nodes[k]]; a[2 * k + 1] = Pm[2 * nodes[k] + 1];
        b[2 * k] = Pd[2 * nodes[k]]; b[2 * k + 1] = Pd[2 * nodes[k] + 1];
      }
      double tf9[9];
      slide_oracle_estimate_tf(a, b, nn, tf9);                                   /* SC.cpp:258: model -> data */
      const double yaw = atan2(tf9[3], tf9[0]);                                  /* SC.cpp:261-268 */
      tf16[3] = tf9[2]; tf16[7] = tf9[5];
      tf16[0] = cos(yaw); tf16[1] = -sin(yaw); tf16[4] = sin(yaw); tf16[5] = cos(yaw);
      found = 1;
      free(a); free(b);
    }
    free(A); free(M); free(nodes);
  }
  free(mi); free(di); free(Pm); free(Pd);
  info->found = found;
  return found;
}
