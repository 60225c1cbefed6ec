// spr_join_types.h -- POD records of the pair-join scorer shared by the host builder and the kernels (see spr_join.h).
#pragma once
#include <stdint.h>

#ifndef SPJ_WARPS
#define SPJ_WARPS 8
#endif
#define SPJ_THREADS (SPJ_WARPS * 32)
#ifndef SPJ_SLOT_BITS
#define SPJ_SLOT_BITS 12
#endif
#define SPJ_MAX_SLOTS (1 << SPJ_SLOT_BITS)   // lattice samples per block (slot index in the block's arg-max)
#define SPJ_TILE_WORDS SPJ_MAX_SLOTS         // shared-memory words of a block's counters: two arrays of nx * ((ny >> 1) + 1)
#ifndef SPJ_LIST
#define SPJ_LIST 256                         // per-warp list of (landmark, record) pairs that passed the filter, entries (>= 64)
#endif
#define SPJ_BANDS 4                          // coarse-cell bands of a landmark whose record ranges one pass gathers
#define SPJ_RANGE_WORDS (2 * 32 * SPJ_BANDS + 2)   // per warp: start of every range in the flat candidate sequence (+ end), first record
#ifndef SPJ_MIN_CTAS
#define SPJ_MIN_CTAS 4                        // resident CTAs per SM the kernel is compiled for (register budget)
#endif

// one reference landmark in the join order of direction d (label-major, then coarse cell)
struct SprJoinRef {
  double x, y, d1, d2, d3;
  uint32_t nbr_off, nbr_cnt;   // same-label landmarks with a LOWER reference index within 2 x reach: SprJoinView::nbr
};
struct SprJoinNbr { double x, y, d1, d2, d3; };

struct SprJoinBlock {          // 32 bytes
  uint32_t xi, yi;             // index into lat[] of the block's first x / y sample
  uint32_t nx, ny;             // samples along x / y; slot(i, j) = i * ny + j
  uint32_t ord0, row_stride;   // translation ordinal of sample (i, j) = ord0 + i * row_stride + j
  uint32_t dir;                // 0: long along y (join bands = coarse x), 1: long along x
  uint32_t ring;
};


struct alignas(16) SprJoinBox { float x0, x1, y0, y1; };   // bounding box of a query group, rounded outward

struct SprJoinView {
  const double *lat;           // lattice samples of every ring
  const double *qrot;          // [n_yaw][nqp][2] exact rotated query coordinates; NaN for padding entries
  const SprJoinBox *gbox;          // [n_yaw][n_groups] (x0, x1, y0, y1) of each query group, rounded outward
  const double *qdims;         // [nqp][3]
  const int32_t *glabel;       // [n_groups] label bucket of the group
  const double *qxy;           // [nqp][2]
  const int32_t *qlabel;       // [nqp] label bucket, -1: padding
  const double *cs;            // [n_yaw][2]
  int32_t nqp, n_groups, n_yaw, n_labels;
  const SprJoinRef *rec[2];    // per direction
  const double *xy[2];         // per direction: (x, y) of every record again, 16 bytes each: what the candidate filter reads
  const uint32_t *cell_start[2];  // per direction: [n_labels * n_cells + 1] first record of (label, cell)
  const SprJoinNbr *nbr;
  const double *labelbox;      // [n_labels][4] x0, x1, y0, y1 of the label's reference landmarks
  double gx0, gy0, inv_w;      // coarse grid: cell (cx, cy) = floor((x - gx0) * inv_w), floor((y - gy0) * inv_w)
  int32_t ncx, ncy;
  double Tstar, Sstar, thr_dim;
  int32_t ignore_dim, pad0;
  double reach;                // a match implies |r - (Rq + t)| < reach on both axes (threshold + rounding slack)
  double ireach;               // reach + accumulated drift of the lattice samples: index-range estimates
  double inv_step;
  const SprJoinBlock *blocks;
  uint32_t n_blocks, pad1;
};

struct SprJoinLaunch {
  unsigned long long *work_counter;   // zeroed by the caller
  unsigned long long *best_key;
  int32_t *counts_out;                // optional: [(ordinal - ord_begin) * n_yaw + iyaw]
  unsigned long long ord_begin, ord_end;   // slice of translation ordinals scored
  int32_t shard_index, shard_count;
};

