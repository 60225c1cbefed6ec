/*
 * clipper_oracle.h -- CPU ORACLE of the SlideGraph half (CLIPPER affinity + dense clique,
 * run_semantic_clipper).  TEST INFRASTRUCTURE, see clipper_oracle.c / slide_oracle.h.
 */
#ifndef CLIPPER_ORACLE_H
#define CLIPPER_ORACLE_H

#include "slide_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { CLIPPER_ROUND_NONZERO = 0, CLIPPER_ROUND_DSD = 1, CLIPPER_ROUND_DSD_HEU = 2 }; /* clipper.h:50 */

/* clipper::Params (clipper.h:28-60) + invariants::EuclideanDistance::Params (euclidean_distance.h:27-30) */
typedef struct clipper_oracle_params {
  double sigma, epsilon, mindist;
  double tol_u, tol_F, tol_Fop;
  int maxiniters, maxoliters;
  double beta;
  int maxlsiters;
  double eps, affinityeps;
  int rescale_u0;
  int rounding;
} clipper_oracle_params;

typedef struct clipper_oracle_solution { int ifinal; int n_nodes; double score; } clipper_oracle_solution;

void clipper_oracle_default_params(clipper_oracle_params *p);
/* EuclideanDistance::operator() on columns (a_i, a_j) of D1 and (b_i, b_j) of D2 (dim x n, column-major) */
double clipper_oracle_invariant(const clipper_oracle_params *p, const double *D1, const double *D2, int dim,
                                int a_i, int a_j, int b_i, int b_j);
void clipper_oracle_all_to_all(int n1, int n2, int *A /* (n1*n2) x 2 */);
void clipper_oracle_k2ij(long long k, long long n, long long *i_out, long long *j_out);
/* scorePairwiseConsistency: M (m x m row-major) receives the strict upper triangle; returns nnz */
long long clipper_oracle_score_pairwise(const clipper_oracle_params *p, const double *D1, int n1, const double *D2, int n2,
                                        int dim, const int *A, int m, double *M);
/* getAffinityMatrix: symmetric view + identity */
void clipper_oracle_affinity_matrix(const double *Mupper, int m, double *Mfull);
int clipper_oracle_k_largest(const double *x, int n, int k, int *idx_out);
/* findDenseClique with the given u0; nodes_out capacity n; u_out (optional) n.  Returns the node count. */
int clipper_oracle_find_dense_clique(const clipper_oracle_params *p, const double *M, int n, const double *u0,
                                     clipper_oracle_solution *sol, int *nodes_out, double *u_out);
/* dsd::solve(A, S): S_in may be NULL / ns_in = 0 (all nodes); returns the node count */
int clipper_oracle_dsd(const double *M, int n, const int *S_in, int ns_in, int *nodes_out);

typedef struct clipper_oracle_sc_info {
  int found;
  long long n_triangle_matches;
  int n_associations;
  long long nnz;
  int n_inliers;
  double score;
} clipper_oracle_sc_info;

/* run_semantic_clipper (SC.cpp:140-275) from the triangle lists on (the triangulation itself is qhull's);
 * u0: initial vector of clipper.solve (the reference draws it from std::random_device), at least
 * 3 * matches long.  tf16: the reference's tfFromQuery2Ref output (row-major).  Returns found. */
int clipper_oracle_run_semantic_clipper(const double *tris_model6, int t_model, const double *tris_data6, int t_data,
                                        double sigma, double epsilon, int min_num_pairs, double matching_threshold,
                                        const double *u0, int u0_len, double *tf16, clipper_oracle_sc_info *info);

#ifdef __cplusplus
}
#endif
#endif
