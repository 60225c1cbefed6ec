// spr_kernels.h -- launch wrappers of the sm_100a kernels (spr_kernels.cu, spr_kernels_aux.cu).
#pragma once
#include <cuda_runtime.h>

#include "spr_types.h"

#define SPR_WARP_CHUNKS 32  // chunks per group (one warp of the exact kernel)

// Sharding granule = a DOUBLE GROUP (64 chunks: 32 chunks + their 32 continuations along the row,
// one work item of the bound kernel).  Double groups are dealt round-robin to the shards; the
// k-th local group of shard si (of sc) is group spr_shard_group(k, si, sc) of the chunk range.
SPR_HD int spr_shard_group(int k_local, int si, int sc) { return (((k_local >> 1) * sc + si) << 1) | (k_local & 1); }
static inline int spr_shard_local_groups(int n_groups, int si, int sc) {  // n_groups is even
  const int n_dg = n_groups / 2;
  return n_dg > si ? 2 * ((n_dg - si + sc - 1) / sc) : 0;
}

// One pass of the lattice search = one (label, direction) pair over a range of chunks.
struct SprLaunch {
  uint32_t chunk_begin, chunk_end;  // chunks of direction `dir` scored by this pass (multiple of 32 apart)
  uint32_t n_chunks_total;          // size of the chunk list (indexes the global counters)
  int32_t  label;                   // label bucket probed by this pass; -1: no queries at all
  uint32_t dir;                     // bitmap direction of every chunk in [chunk_begin, chunk_end)
  uint32_t row_begin, row_end;      // row band of the plane whose tables this pass stages in shared memory (multiples of 8;
                                    //   the whole plane: 0 .. R); hits outside the band belong to another pass
  uint32_t tab_rank_lo, tab_cells;  // label-relative rank of the band's first marked cell, marked cells in the band
  uint32_t tab_refs;                // landmarks of the label
  uint32_t tab_cell_base, tab_ref_base; // absolute rank of the label's first cell / first landmark row (alignment of the bulk copies)
  int32_t  stage_reftab;            // 1: the label's landmark table is staged too; 0: read in place
  int32_t  first, last;             // first / last pass over these chunks: counters start at 0 / are reduced
  int32_t  shard_index, shard_count;
  void    *gcnt;                    // device: per-hypothesis inlier counters carried between passes
                                    //   [(yaw * n_chunks_total + chunk) * 32 + bit], u16 (u32 if nqp > 65535)
  unsigned long long *best_key;     // device: running max of spr_make_key
  unsigned long long *work_counter; // device: next work item (zeroed by the launcher)
  int32_t *counts_out;              // device, optional: [(ordinal - ord_begin) * n_yaw + iyaw]
  long long counts_cap;
  unsigned long long ord_begin;
  unsigned long long *stats;        // device, optional: [0] filter hits, [1] verified inliers,
                                    //                   [2] query groups probed, [3] query groups skipped
  // bound-and-verify (optional): upper bounds from spr_launch_bound_lattice.  A hypothesis is
  // verified only while its bound reaches the inlier count of the running best (*best_key).
  const uint32_t *ub_planes;        // device: [yaw][chunk / 32][ub_nplanes][32 lanes] bit planes of the bounds
  const uint32_t *item_ub;          // device: [yaw][chunk / 32] largest bound of the work item
  int32_t ub_nplanes;               // 0: exhaustive search
  int32_t pad;
  const uint32_t *cand_items;       // device, optional: work items (yaw * n_wg_local + k) to verify, built by
  const uint32_t *cand_count;       //   spr_launch_select_items; *cand_count entries
};

// One launch of the bound phase = one bitmap direction x a set of labels whose planes are staged
// together in shared memory, over a range of chunks.
#define SPR_BOUND_MAX_LABELS 8
#define SPR_SEED_SLOTS 32            // seeds per yaw: the best-bounded hypothesis of every 32nd work item column
struct SprBoundLaunch {
  uint32_t chunk_begin, chunk_end;  // chunks of direction `dir` (multiple of 32 apart)
  uint32_t n_chunks_total;
  uint32_t dir;
  int32_t  n_labels;
  int32_t  labels[SPR_BOUND_MAX_LABELS];
  int32_t  first, last;             // first / last launch over these chunks: planes start at 0 / item maxima are produced
  int32_t  shard_index, shard_count;
  uint32_t row_begin, row_end;      // rows of the planes staged in shared memory (row_end == row_begin: read in place)
  uint32_t *planes;                 // device: bit planes of the bounds (layout of SprLaunch::ub_planes)
  uint32_t *item_ub;                // device: [yaw][chunk / 32]
  unsigned long long *seed_key;     // device: [n_yaw][SPR_SEED_SLOTS] (bound + 1) << 40 | chunk * 32 + bit of the best-bounded hypothesis
  unsigned long long *work_counter; // device: next work item (zeroed by the caller)
  // refinement launches only: candidate double groups (local item numbers) from spr_launch_select_dgroups;
  // the kernel does nothing when there are fewer than refine_min of them
  const uint32_t *cand_items;
  const uint32_t *cand_count;
  uint32_t refine_min, pad;
};

enum { SPR_TABLES_GLOBAL = 0, SPR_TABLES_AUTO = 1 };  // AUTO: shared-memory-resident plane when it fits

// rotated query coordinates for every yaw: exact fp64 + fixed-point cell units in both layouts
// + per-group bounding boxes (PR.cpp:246-258)
cudaError_t spr_launch_rotate(const SprView &V, int32_t *qrotq_xy, int32_t *qrotq_yx, double *qrot, SprBox *gbox,
                              cudaStream_t st);

// one (label, direction) pass of the lattice search
int spr_score_smem_warps(const SprView &V, const SprLaunch &K, int tables_mode);  // 0: this pass reads its tables in place
// row bands of plane (label, dir) for the exact passes (band_rows == 0: read in place)
void spr_score_plan(const SprView &V, uint32_t dir, uint32_t label_cells, uint32_t label_refs, uint32_t *band_rows, int *stage_reftab);
cudaError_t spr_launch_score_lattice(const SprView &V, const SprLaunch &K, int tables_mode, int sm_count,
                                     cudaStream_t st, int *n_launches);

// bound phase of the bound-and-verify search (spr_kernels_bound.cu)
int spr_bound_planes(int nqp);                                   // bit planes needed for counts <= nqp (12 or 16)
void spr_bound_plan(const SprView &V, uint32_t dir, int n_active, size_t smem_budget, int *labels_per_launch, uint32_t *band_rows);
uint32_t spr_refine_band_rows(const SprView &V, uint32_t dir, size_t smem_budget);  // 0: refinement reads the variant planes in place
cudaError_t spr_launch_bound_lattice(const SprView &V, const SprBoundLaunch &B, int n_planes, int sm_count, cudaStream_t st,
                                     int *n_launches);
// work items of direction `B.dir` whose largest bound reaches the running best -> items[0 .. *count)
cudaError_t spr_launch_select_items(const SprView &V, const SprBoundLaunch &B, const unsigned long long *best_key,
                                    uint32_t *items, uint32_t *count, int sm_count, cudaStream_t st);
// double groups of direction `B.dir` with a group whose largest bound reaches the running best
cudaError_t spr_launch_select_dgroups(const SprView &V, const SprBoundLaunch &B, const unsigned long long *best_key,
                                      uint32_t *items, uint32_t *count, int sm_count, cudaStream_t st);
// half-cell variants of the occupancy bitmaps (V.vbitmap, `words` 32-bit words incl. 4 in front), built on the
// device from the raw reference rows; both kernels return at once unless counts[0] or counts[1] >= min_count
cudaError_t spr_launch_variant_planes(const SprView &V, uint32_t *vbuf, size_t words, const double *ref7, const int32_t *lab_of,
                                      int n_ref, double cell, double rc, double rc2, const uint32_t *counts, uint32_t min_count,
                                      int sm_count, cudaStream_t st);
// exact score of the best-bounded hypothesis of every yaw -> atomicMax on best_key
cudaError_t spr_launch_seed(const SprView &V, const unsigned long long *seed_key, unsigned long long *best_key,
                            cudaStream_t st);

// explicit hypothesis list (c, s, x, y), warp per hypothesis
cudaError_t spr_launch_score_list(const SprView &V, const double *hyps4, long long n, int32_t *counts_out,
                                  unsigned long long *best_key, int sm_count, cudaStream_t st);

// correspondences of one hypothesis by brute force over the raw maps, in the reference's own
// order (PR.cpp:281-357): match_ref[j] = first matching reference index or -1
cudaError_t spr_launch_extract(const double *ref7, int n_ref, const double *qry7, int n_qry, double c,
                               double s, double tx, double ty, double Tstar, double Sstar, double thr_dim,
                               int ignore_dim, int32_t *match_ref, cudaStream_t st);

// SlideGraph descriptor half (semantic_clipper.cpp:49-99); the matching is in spr_generate.cu
// labels3 (optional): vertex labels; sig receives them in sorted-descriptor order (class signature)
cudaError_t spr_launch_tri_desc(const double *tris6, const double *labels3, int t, double *desc, int32_t *perm,
                                double *sig, cudaStream_t st);

// issue-rate micro-benchmark: warp instructions / s of LOP3 chains (ALU pipe), IMAD chains (FMA pipe) and both interleaved
cudaError_t spr_measure_issue_peaks(int sm_count, double *alu, double *fma, double *mixed, cudaStream_t st);
