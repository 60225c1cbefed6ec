/*
 * slide_pr.h -- C-ABI of the B200-native SlideSLAM place-recognition (SlideMatch) search.
 *
 * This is the drop-in boundary for the hot path of lunarlab-gatech/SLIDE_SLAM's
 * backend/sloam place recognition.  The reference has no FFI layer: the boundary is the C++
 * class PlaceRecognition (backend/sloam/include/core/place_recognition.h:31-237).  Every entry
 * point below names the member function it replaces; include/slide_pr/place_recognition.hpp
 * is the header-only C++ adapter with the reference's own signatures.
 *
 * Conventions
 *   - landmark rows are the reference's Eigen::Vector7d record [label,x,y,z,d1,d2,d3]
 *     (place_recognition.h:51-57), contiguous doubles, stride 7, so
 *     reinterpret_cast<const double*>(vec.data()) of a std::vector<Eigen::Vector7d> works.
 *   - matrices are row-major.
 *   - all functions return SLIDE_PR_OK (0) or a negative error code; "closure not found" is a
 *     positive status, never an error (the reference cannot tell them apart: PR.cpp:515-519).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SLIDE_PR_ERR_CUDA and slide_pr_last_error() says why.
 *   - a handle is bound to one CUDA device and is not re-entrant (one call at a time per
 *     handle, exactly like one PlaceRecognition instance: PR.cpp:786-787 mutates members).
 */
#ifndef SLIDE_PR_H
#define SLIDE_PR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIDE_PR_ABI_VERSION 2

enum {
  SLIDE_PR_OK = 0,
  SLIDE_PR_NOT_FOUND = 1,          /* fewer than min_num_inliers (PR.cpp:849) or size gate (PR.cpp:508) */
  SLIDE_PR_SANITY_RETURN = 2,      /* MatchMaps' early return, outputs untouched (PR.cpp:169-175) */
  SLIDE_PR_ERR_INVALID = -1,       /* bad argument */
  SLIDE_PR_ERR_CUDA = -2,          /* CUDA runtime error / no device */
  SLIDE_PR_ERR_UNSUPPORTED = -3,   /* problem too large for the index structures */
  SLIDE_PR_ERR_NONFINITE = -4,     /* NaN/Inf coordinate in a map */
  SLIDE_PR_ERR_INTERNAL = -5       /* self-check failed (hash count != brute-force count) */
};

/* rosparams sloam/place_recognition/ * as stored in the members (PR.cpp:24-75); angles in radians,
 * computed by the caller as deg * M_PI / 180. like PR.cpp:35,41,60 (slide_pr_deg2rad does it). */
typedef struct slide_pr_params {
  double compute_budget_sec;          /* <= 0: unlimited.  > 0: checked between rings (PR.cpp:181-191) */
  double dilation_factor;             /* 1.2 */
  double match_xy_step_size;          /* 0.5 */
  double match_yaw_half_range;        /* pi */
  double match_yaw_angle_step_size;   /* 2 deg */
  double match_threshold;             /* 0.5 */
  double match_threshold_dimension;   /* 1.0 */
  double match_x_half_range_intra;    /* 5 */
  double match_y_half_range_intra;    /* 5 */
  double match_yaw_half_range_intra;  /* 10 deg */
  int32_t disable_yaw_search;         /* 0 */
  int32_t ignore_dimension;           /* 0 */
  int32_t min_num_inliers;            /* 5 */
  int32_t use_lsq;                    /* 1 (rosparam use_nonlinear_least_squares) */
  int32_t min_num_map_objects_to_start; /* 1 */
  int32_t inter_loop_closure;         /* 1 (public member PR.h:43) */
  int32_t device;                     /* CUDA device ordinal; -1 = current device */
  int32_t exhaustive_search;          /* not a rosparam: which kernels search the lattice.  0 (default): the pair-join scorer
                                         (exact inlier count of every hypothesis); 1: the lattice kernels, every hypothesis
                                         verified; 2: the lattice kernels, bound-and-verify.  Same winner, inlier count and
                                         correspondences in every case */
} slide_pr_params;

typedef struct slide_pr_handle slide_pr_handle;

/* Result of one MatchMaps-level search (PR.h:70-74 outputs + bookkeeping). */
typedef struct slide_pr_match_result {
  int32_t status;              /* SLIDE_PR_OK or SLIDE_PR_SANITY_RETURN */
  int32_t best_num_inliers;    /* -10000 when no hypothesis was scored (PR.cpp:125) */
  double  R_t[9];              /* [[c,-s,x],[s,c,y],[0,0,1]] of the winner (PR.cpp:244-251) */
  int32_t n_matched;           /* entries written to ref_idx_out / qry_idx_out */
  int32_t n_rings;
  int32_t n_yaw;
  int32_t rings_scored;        /* < n_rings only when compute_budget_sec cut the search */
  int64_t hypotheses_scored;
  int64_t best_hyp_index;      /* canonical index (ring -> x -> y -> yaw enumeration order); -1 if none */
  int64_t n_translations;      /* lattice translations (all rings) */
  float   kernel_ms;           /* device time of the search kernels (CUDA events on the call's stream) */
  float   prepare_ms;          /* host time spent building the index structures + H2D enqueue */
  int64_t gpu_launches;        /* kernels launched by this call */
  int64_t filter_hits;         /* bitmap hits verified in fp64 (0 unless stats were enabled) */
  int64_t h2d_bytes;           /* bytes copied host->device by prepare (+ re-chunking in search) */
  int64_t d2h_bytes;           /* bytes copied device->host by search/extract */
  int64_t groups_probed;       /* (warp, query group) pairs probed / skipped by the bounding-box test */
  int64_t groups_skipped;      /*   (0 unless stats were enabled) */
  int32_t reuse;               /* bit 0: lattice reused from the previous prepare, bit 1: reference-map index reused */
  int32_t search_mode;         /* 2: pair-join scorer (every hypothesis counted exactly); 0: lattice kernels, every hypothesis
                                  verified; 1: lattice kernels, bound-and-verify (same winner in every mode) */
} slide_pr_match_result;

/* Options for the sharded / sliced search (multi-GPU and tests). */
typedef struct slide_pr_search_opts {
  int64_t trans_begin;   /* score translations with canonical ordinal in [trans_begin, trans_end) */
  int64_t trans_end;     /* < 0: to the end */
  int32_t shard_index;   /* this rank's shard of the work (round-robin over chunk groups) */
  int32_t shard_count;   /* <= 1: no sharding */
  int32_t *counts_out;   /* optional HOST buffer: inlier count of every hypothesis of the slice,
                            counts_out[(t - trans_begin) * n_yaw + iyaw]; needs trans_end >= 0 */
  int64_t counts_cap;    /* capacity of counts_out in entries */
  void   *stream;        /* cudaStream_t to run on; NULL = the handle's own stream */
  int32_t collect_stats; /* 1: count filter hits (slower; implies exhaustive) */
  int32_t exhaustive;    /* engine of this search.  0 (default): the handle's default (the pair-join scorer unless
                            slide_pr_params.exhaustive_search / SLIDE_PR_ENGINE=lattice say otherwise).
                            4: the pair-join scorer -- the exact inlier count of every hypothesis of the slice / shard.
                            Lattice kernels: 1: every hypothesis verified exactly; 3: bound-and-verify -- a cheap upper
                            bound (bitmap filter hits) for every hypothesis, exact verification only where the bound
                            reaches the running best; 2: bound phase only (first half of a sharded bound-and-verify
                            search, and a test hook): the result holds the best exactly scored seed hypothesis;
                            counts_out, if given, receives the upper bounds.  collect_stats, reuse_bounds, more than
                            65535 query landmarks and a compute budget that can bind are served by the lattice kernels. */
  int32_t incumbent_inliers; /* > 0: an inlier count already reached elsewhere (another shard of the same pair, after an
                            all-reduce(max) of the shards' bound-phase results): hypotheses whose bound is below it are
                            not verified.  A shard that finds nothing >= the incumbent reports best_hyp_index = -1. */
  int32_t reuse_bounds;  /* lattice kernels: 1: the bounds of the preceding exhaustive = 2 call on the same prepared problem and
                            shard are still on the device: skip the bound phase and go straight to the verification */
} slide_pr_search_opts;

typedef struct slide_pr_tf_result {
  int32_t found;               /* return value of findTransformation */
  int32_t best_num_inliers;
  int32_t n_matched;
  int32_t reserved;
  double  R_t[9];              /* lattice winner in the (centroid-shifted) search frames */
  double  xyz_yaw[4];          /* PR.cpp:902 / :941 */
  double  transform[16];       /* transform_out */
  double  centroid_ref[2], centroid_qry[2];
  double  half_x, half_y, yaw_half;
  slide_pr_match_result match; /* the underlying MatchMaps call */
} slide_pr_tf_result;

/* ---- lifetime --------------------------------------------------------------------------- */
int  slide_pr_abi_version(void);
double slide_pr_deg2rad(double deg);                       /* deg * M_PI / 180.  (PR.cpp:35) */
void slide_pr_default_params(slide_pr_params *p);          /* PlaceRecognition::ParamInit defaults, PR.cpp:24-75; budget disabled */
int  slide_pr_create(const slide_pr_params *p, slide_pr_handle **out);   /* PlaceRecognition ctor PR.cpp:15-21 */
void slide_pr_destroy(slide_pr_handle *h);
int  slide_pr_set_params(slide_pr_handle *h, const slide_pr_params *p);  /* public members use_lsq / inter_loop_closure, PR.h:34-43 */
const char *slide_pr_last_error(const slide_pr_handle *h); /* h may be NULL: error of the last failed create */

/* ---- the hot path ------------------------------------------------------------------------ */
/* PlaceRecognition::MatchMaps (PR.h:70-74, PR.cpp:98-387).  half_x/half_y are the members
 * match_{x,y}_half_range_ that findTransformation sets before the call (PR.cpp:786-787, 808-809);
 * the yaw range comes from the params (inter or intra).  ref_idx_out/qry_idx_out: caller arrays
 * of capacity n_qry receiving, in query order, the indices of the matched reference / query
 * objects (the reference returns the rows themselves: PR.cpp:344-350). */
int slide_pr_match_maps(slide_pr_handle *h, const double *ref7, int32_t n_ref,
                        const double *qry7, int32_t n_qry, double half_x, double half_y,
                        int32_t *ref_idx_out, int32_t *qry_idx_out, slide_pr_match_result *out);

/* The same search split in its stages, for sharded multi-GPU use and for timing:
 *   prepare : host index build + H2D of both maps           (inputs become HBM-resident; the structures of the
 *             engine that is not the handle's default are built by the first search that asks for it)
 *   search  : the scoring kernels over this shard/slice     -> local best (count, canonical index)
 *   extract : correspondences + R_t of ONE hypothesis (the global winner after the all-gather) */
int slide_pr_prepare(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7,
                     int32_t n_qry, double half_x, double half_y);
int slide_pr_search(slide_pr_handle *h, const slide_pr_search_opts *opts, slide_pr_match_result *out);
/* size of the prepared search lattice (PR.cpp:136-241): translations over all rings, yaw candidates, rings */
int slide_pr_lattice_info(const slide_pr_handle *h, int64_t *n_translations, int32_t *n_yaw, int32_t *n_rings);
int slide_pr_extract(slide_pr_handle *h, int64_t hyp_index, int32_t *ref_idx_out,
                     int32_t *qry_idx_out, slide_pr_match_result *inout);

/* PlaceRecognition::findTransformation (PR.h:150-153, PR.cpp:736-945): centroid shift, search
 * range, MatchMaps, inlier gate, LSQ refinement or centroid un-shift.  Returns SLIDE_PR_OK when
 * found, SLIDE_PR_NOT_FOUND otherwise. */
int slide_pr_find_transformation(slide_pr_handle *h, const double *ref7, int32_t n_ref,
                                 const double *qry7, int32_t n_qry, int32_t *ref_idx_out,
                                 int32_t *qry_idx_out, slide_pr_tf_result *out);

/* PlaceRecognition::findInterLoopClosure (PR.h:103-106, PR.cpp:498-538): size gate +
 * findTransformation + yaw/xyz 4x4.  tf16: row-major tfFromQueryToRef. */
int slide_pr_find_inter_loop_closure(slide_pr_handle *h, const double *ref7, int32_t n_ref,
                                     const double *qry7, int32_t n_qry, double *tf16,
                                     slide_pr_tf_result *out /* may be NULL */);

/* PlaceRecognition::findIntraLoopClosure (PR.h:88-91, PR.cpp:389-496).  Poses are row-major 4x4
 * (Sophus SE3::matrix()).  measurements are in the query's local frame. */
int slide_pr_find_intra_loop_closure(slide_pr_handle *h, const double *measurements7, int32_t n_meas,
                                     const double *submap7, int32_t n_sub, const double *query_pose16,
                                     const double *candidate_pose16, double *tf16,
                                     slide_pr_tf_result *out /* may be NULL */);

/* Several candidate key poses for ONE set of measurements (the reference tries one candidate per attempt,
 * sloamNode.cpp:355-486).  submaps7[k] (n_subs[k] rows) and candidate_poses16 + 16 k describe
 * candidate k; out[k] and tf16_out + 16 k receive what slide_pr_find_intra_loop_closure returns for it
 * (out[k].found tells whether tf16_out + 16 k was written).  The candidates' searches are enqueued back to back --
 * the host prepares candidate k + 1 while the GPU scores candidate k -- and the host waits once for all of them. */
int slide_pr_find_intra_loop_closure_batch(slide_pr_handle *h, const double *measurements7, int32_t n_meas,
                                           const double *const *submaps7, const int32_t *n_subs, const double *query_pose16,
                                           const double *candidate_poses16, int32_t n_cand, double *tf16_out,
                                           slide_pr_tf_result *out /* n_cand */);

/* PlaceRecognition::solveLSQ (PR.h:127-130, PR.cpp:632-695): Kabsch on k matched pairs.
 * tgt3 = map objects, src3 = detection objects (k x 3). */
int slide_pr_solve_lsq(const double *tgt3, const double *src3, int32_t k, double *xyz_yaw4,
                       double *transform16);

/* PlaceRecognition::getxyzYawfromTF (PR.h:138, PR.cpp:697-711). */
void slide_pr_get_xyz_yaw_from_tf(const double *tf16, double *xyz_yaw4);

/* ---- batches (BASELINE configs 4 and 5) --------------------------------------------------- */
/* n_pairs independent findTransformation calls (all-pairs multi-robot matching / streaming
 * submap queries).  maps: array of n_maps pointers to n x 7 rows; pair p matches reference
 * maps[ref_of[p]] against query maps[qry_of[p]].  Every map goes through the map cache once: the
 * index structures of a reference map are built once, however many pairs use it and in whatever order. */
int slide_pr_find_transformation_batch(slide_pr_handle *h, const double *const *maps,
                                       const int32_t *map_sizes, int32_t n_maps,
                                       const int32_t *ref_of, const int32_t *qry_of, int32_t n_pairs,
                                       slide_pr_tf_result *out /* n_pairs */);

/* ---- device-resident map cache (one slot per robot) ---------------------------------------------- */
/* The reference keeps one object map per robot (databaseManager::robotMapDict_, databaseManager.h:99-102)
 * and deep-copies both maps for every attempt (sloamNode.cpp:603-614).  slide_pr_map_cache_put hands a map
 * over once per version; its reference-side index is built the first time the map is searched as a
 * reference and reused by every later pair.  Least recently used slots are dropped beyond 64 maps. */
int slide_pr_map_cache_put(slide_pr_handle *h, int64_t robot_id, uint64_t version, const double *rows7, int32_t n);
int slide_pr_map_cache_drop(slide_pr_handle *h, int64_t robot_id);   /* SLIDE_PR_NOT_FOUND if absent */
int32_t slide_pr_map_cache_size(const slide_pr_handle *h);
/* PlaceRecognition::findTransformation on two cached maps (inter-robot mode); same outputs as
 * slide_pr_find_transformation.  out->match.reuse bit 1 tells whether the reference index was reused. */
int slide_pr_find_transformation_cached(slide_pr_handle *h, int64_t ref_robot_id, int64_t qry_robot_id, int32_t *ref_idx_out,
                                        int32_t *qry_idx_out, slide_pr_tf_result *out);

/* ---- multi-GPU merge ---------------------------------------------------------------------- */
/* Packs a local result into the 16-byte record exchanged by the all-gather, and merges
 * n records deterministically: max inliers, ties to the smallest canonical index (the
 * reference's strict '>' first-wins rule, PR.cpp:361).  Returns the index of the winner. */
typedef struct slide_pr_topk_record { int64_t hyp_index; int32_t inliers; int32_t rank; } slide_pr_topk_record;
void slide_pr_pack_record(const slide_pr_match_result *r, int32_t rank, slide_pr_topk_record *rec);
int  slide_pr_merge_records(const slide_pr_topk_record *recs, int32_t n);

/* ---- SlideGraph descriptor half (semantic_clipper.cpp) ------------------------------------- */
/* compute_triangle_diff / match_triangles (SC.cpp:49-118) on the GPU: descriptors of all
 * triangles are built once, binned by their first component, and matched; output order is the
 * reference's (model-major, data-minor; restored by a radix sort of the match keys).  tris: t x 6
 * [x0,y0,x1,y1,x2,y2].  model_idx_out /
 * data_idx_out / perm_out (3 ints per match per side: the sorted vertex order) have capacity
 * cap matches; returns the total match count in *n_matches (may exceed cap). */
int slide_pr_match_triangles(slide_pr_handle *h, const double *tris_model6, int32_t t_model,
                             const double *tris_data6, int32_t t_data, double threshold,
                             int32_t *model_idx_out, int32_t *data_idx_out,
                             int32_t *perm_model_out, int32_t *perm_data_out,
                             int64_t cap, int64_t *n_matches);

/* The same with a class signature: labels_*3 hold the semantic label of every triangle vertex
 * (t x 3).  A pair is kept only if, in addition to the descriptor test, the vertices paired by
 * the sorted order carry equal labels -- the "semantic label check" the reference leaves as a TODO
 * (SC.cpp:114,186).  NULL label arrays give slide_pr_match_triangles. */
int slide_pr_match_triangles_labeled(slide_pr_handle *h, const double *tris_model6, const double *labels_model3,
                                     int32_t t_model, const double *tris_data6, const double *labels_data3,
                                     int32_t t_data, double threshold, int32_t *model_idx_out,
                                     int32_t *data_idx_out, int32_t *perm_model_out, int32_t *perm_data_out,
                                     int64_t cap, int64_t *n_matches);

/* semantic_clipper::estimate_tf (SC.cpp:122-138): 2-D Kabsch a -> b on k point pairs (k x 2 rows);
 * tf9 row-major [[R, t], [0, 0, 1]]. */
int slide_pr_estimate_tf(const double *pts_a2, const double *pts_b2, int32_t k, double *tf9);

/* One rigid-transform hypothesis (c, s, x, y) per triangle match: the 2-D Kabsch fit that maps the
 * data (query-map) triangle onto the model (reference-map) triangle with the vertices paired in
 * sorted-descriptor order (perm_*: 3 ints per match, as returned by slide_pr_match_triangles).
 * hyps4_out: n x 4, ready for slide_pr_score_hypotheses. */
int slide_pr_triangle_hypotheses(const double *tris_model6, const double *tris_data6, const int32_t *model_idx,
                                 const int32_t *data_idx, const int32_t *perm_model, const int32_t *perm_data,
                                 int64_t n, double *hyps4_out);

/* Scores an explicit list of rigid-transform hypotheses (c, s, x, y) -- e.g. the 2-D Kabsch
 * fits of matched triangles (SC.cpp:122-138) -- with the MatchMaps inlier predicate
 * (PR.cpp:272-357) against the maps given to slide_pr_prepare.  hyps: n x 4 doubles.
 * counts_out (optional, n ints).  The winner (max count, lowest index) goes to *out. */
int slide_pr_score_hypotheses(slide_pr_handle *h, const double *hyps4, int64_t n,
                              int32_t *counts_out, slide_pr_match_result *out);

/* The generator half in one call, entirely on the device: descriptors -> binning of the data
 * triangles by their first descriptor component -> windowed matching -> radix sort into the reference's
 * order (SC.cpp:111-118) -> one 2-D Kabsch hypothesis per match (SC.cpp:122-138) -> the MatchMaps
 * predicate on every hypothesis (PR.cpp:272-357) -> best hypothesis (max count, lowest match index).
 * Needs the map pair given to slide_pr_prepare.  Optional outputs (capacity cap matches each):
 * the match list, the hypotheses (4 doubles each: c, s, x, y) and their inlier counts. */
typedef struct slide_pr_generate_info {
  int64_t n_matches;
  int32_t n_triangles_model, n_triangles_data;
  float   match_ms;    /* descriptors + binning + matching + sort (device time) */
  float   kabsch_ms;
  float   score_ms;
  int32_t reserved;
} slide_pr_generate_info;
int slide_pr_generate_and_score(slide_pr_handle *h, const double *tris_model6, const double *labels_model3, int32_t t_model,
                                const double *tris_data6, const double *labels_data3, int32_t t_data, double threshold,
                                slide_pr_match_result *out, slide_pr_generate_info *info /* may be NULL */,
                                int32_t *model_idx_out, int32_t *data_idx_out, double *hyps4_out, int32_t *counts_out,
                                int64_t cap);

/* ---- SlideGraph: CLIPPER affinity scoring + dense-clique solver (clipper.cpp) ---------------- */
enum { SLIDE_CLIPPER_ROUND_NONZERO = 0, SLIDE_CLIPPER_ROUND_DSD = 1, SLIDE_CLIPPER_ROUND_DSD_HEU = 2 };  /* clipper.h:50 */

/* clipper::Params (clipper.h:28-60) + invariants::EuclideanDistance::Params (euclidean_distance.h:27-30) */
typedef struct slide_clipper_params {
  double sigma;        /* 0.01  spread of the exponential kernel */
  double epsilon;      /* 0.06  bound on the consistency score */
  double mindist;      /* 0     minimum distance between points of one dataset */
  double tol_u, tol_F, tol_Fop;   /* 1e-8, 1e-9, 1e-10 */
  int32_t maxiniters, maxoliters; /* 200, 1000 */
  double beta;         /* 0.25 */
  int32_t maxlsiters;  /* 99 */
  double eps;          /* 1e-9 */
  double affinityeps;  /* 1e-4 */
  int32_t rescale_u0;  /* 1 */
  int32_t rounding;    /* SLIDE_CLIPPER_ROUND_DSD_HEU */
} slide_clipper_params;

typedef struct slide_clipper_solution {   /* clipper::Solution (clipper.h:65-73) */
  int32_t n_nodes;     /* entries written to nodes_out */
  int32_t ifinal;      /* outer iterations */
  double  score;       /* objective value F */
  double  d;           /* final homotopy parameter */
  int64_t line_search_steps;
  float   kernel_ms;   /* device time of the persistent solver kernel */
  int32_t reserved;
} slide_clipper_solution;

void slide_clipper_default_params(slide_clipper_params *p);

/* CLIPPER::scorePairwiseConsistency (clipper.h:93-95, clipper.cpp:21-65) with the EuclideanDistance
 * invariant (euclidean_distance.cpp:13-30).  D1 / D2: dim x n1 / dim x n2, column-major -- the memory
 * of the reference's Eigen::MatrixXd arguments; dim <= 3.  A: m x 2 row-major associations, or NULL
 * (m ignored) for the all-to-all hypothesis.  The affinity matrix stays on the device in CSR form;
 * *nnz_upper receives the non-zeros of the reference's upper-triangular M_. */
int slide_pr_clipper_score_pairwise_consistency(slide_pr_handle *h, const slide_clipper_params *p, const double *D1, int32_t n1,
                                                const double *D2, int32_t n2, int32_t dim, const int32_t *A, int32_t m,
                                                int64_t *nnz_upper);
/* CLIPPER::getInitialAssociations (clipper.cpp:107-110): returns m; writes min(m, cap) rows */
int32_t slide_pr_clipper_get_initial_associations(slide_pr_handle *h, int32_t *A_out, int32_t cap);
/* CLIPPER::getAffinityMatrix (clipper.cpp:121-126): dense m x m row-major, symmetric, identity diagonal */
int slide_pr_clipper_get_affinity_matrix(slide_pr_handle *h, double *M_out, int64_t cap);
/* the same matrix as the device holds it: symmetric CSR without the diagonal; row_ptr has m + 1 entries.
 * col / val may be NULL to query sizes (row_ptr[m] = entries). */
int slide_pr_clipper_get_affinity_csr(slide_pr_handle *h, int64_t *row_ptr, int32_t *col, double *val, int64_t cap);
/* CLIPPER::solve + findDenseClique (clipper.cpp:69-78, 172-323).  u0: m doubles, or NULL for a
 * deterministic stand-in of the reference's std::random_device vector (utils.cpp:22-29): U[0,1)
 * from a splitmix64 stream seeded with `seed`.  nodes_out: capacity cap; u_out (optional): m. */
int slide_pr_clipper_solve(slide_pr_handle *h, const slide_clipper_params *p, const double *u0, uint64_t seed,
                           int32_t *nodes_out, int32_t cap, slide_clipper_solution *sol, double *u_out);

/* ---- measurement support ------------------------------------------------------------------------ */
/* Issue-rate micro-benchmark on the handle's device (about 10 ms): warp instructions per second of
 * independent LOP3 chains (the ALU / logic pipe the search kernels are bound by), of IMAD chains (FMA
 * pipe) and of both interleaved (every scheduler issuing every cycle).  bench.py's roofline uses them
 * as the measured compute peaks; MEASURED_PEAKS.json only holds HBM and tensor-core figures. */
int slide_pr_measure_issue_peaks(slide_pr_handle *h, double *alu_winst_per_s, double *fma_winst_per_s, double *mixed_winst_per_s);

/* ---- SlideGraph entry points ------------------------------------------------------------------ */
/* Observation::delaunayTriangulation (clipper_semantic_object/src/triangulation/observation.cpp:13-88,
 * qhull "Qt Qbb Qc Qz Q12 d"): Delaunay triangulation of n points (xy2: n x 2) on the host.  tri3_out
 * receives 3 vertex ids per triangle (counter-clockwise), capacity cap triangles; *n_tri the total.
 * Same triangle SET as qhull for points in general position; the triangle / vertex ORDER is not qhull's. */
int slide_pr_delaunay(const double *xy2, int32_t n, int32_t *tri3_out, int64_t cap, int64_t *n_tri);

/* rosparams sloam/place_recognition_slidegraph/ * (PR.cpp:64-75) */
typedef struct slide_pr_slidegraph_params {
  double  sigma;                          /* 0.1 */
  double  epsilon;                        /* 0.1 */
  double  matching_threshold;             /* descriptor_matching_threshold, 0.1 */
  int32_t num_inliers_threshold;          /* num_inliners_threshold, 10 */
  int32_t min_num_map_objects_to_start;   /* 20 */
  int32_t use_class_signature;            /* 0 = the reference (no label check: the TODO at SC.cpp:114,186);
                                             1 = matched triangles must also agree label by label */
  int32_t reserved;
  uint64_t seed;                          /* seed of the deterministic u0 that stands in for std::random_device */
} slide_pr_slidegraph_params;
void slide_pr_slidegraph_default_params(slide_pr_slidegraph_params *p);

typedef struct slide_pr_sc_info {
  int32_t found;
  int32_t n_inliers;                /* associations selected by CLIPPER */
  int32_t n_triangles_model, n_triangles_data;
  int64_t n_triangle_matches;
  int64_t n_associations;           /* 3 per triangle match (SC.cpp:102-105) */
  int64_t nnz_upper;                /* non-zeros of the affinity matrix (upper triangle) */
  double  score;                    /* CLIPPER objective of the solution */
  float   delaunay_ms;              /* host */
  float   match_ms, affinity_ms, solve_ms;   /* device */
} slide_pr_sc_info;

/* semantic_clipper::run_semantic_clipper (semantic_clipper.h:38, semantic_clipper.cpp:140-275):
 * Delaunay triangles of both maps' (x, y) -> descriptor matching -> CLIPPER on the matched points ->
 * 2-D Kabsch of the selected pairs.  tris_*6 (t x 6 coordinates) replace the internal triangulation
 * when given (e.g. qhull's own facets in the caller's order); u0 (n_associations doubles) replaces the
 * seeded initial vector.  tf16 (row-major 4x4) is written only when found: the reference's
 * tfFromQuery2Ref, which maps MODEL (reference map) points onto DATA (query map) points -- the
 * caller inverts it (PR.cpp:622-623).  Returns SLIDE_PR_OK / SLIDE_PR_NOT_FOUND. */
int slide_pr_run_semantic_clipper(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7, int32_t n_qry,
                                  const slide_pr_slidegraph_params *sp, const double *tris_model6, int32_t t_model,
                                  const double *tris_data6, int32_t t_data, const double *u0, int64_t u0_len, double *tf16,
                                  slide_pr_sc_info *info /* may be NULL */);

/* PlaceRecognition::findInterLoopClosureWithClipper (PR.h:109-112, PR.cpp:541-630): drops objects at
 * exactly (0, 0), size gate, run_semantic_clipper, inverse of the result.  tf16: tfFromQueryToRef. */
int slide_pr_find_inter_loop_closure_with_clipper(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7,
                                                  int32_t n_qry, const slide_pr_slidegraph_params *sp, double *tf16,
                                                  slide_pr_sc_info *info /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* SLIDE_PR_H */
