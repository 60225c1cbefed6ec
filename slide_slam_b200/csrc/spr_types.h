// spr_types.h -- POD types shared by the host index builder and the CUDA kernels.
//
// Vocabulary (follows the reference, place_recognition.cpp:98-387):
//   reference map / query map : the two object maps handed to MatchMaps
//   hypothesis                : one (yaw, x, y) sample of the search lattice
//   translation               : one (x, y) lattice sample; its "ordinal" is its position in the
//                               reference's enumeration order ring -> x -> y (PR.cpp:178,230,232)
//   chunk                     : 32 consecutive lattice samples along one axis at a fixed value of
//                               the other axis -- the unit one thread scores bit-parallel
//   occupancy bitmap          : per label, 1 bit per cell of edge c = match_xy_step_size_, set when
//                               a reference landmark of that label can be within match_threshold_
//                               of a point falling in the cell (conservative)
//   query group               : SPR_QGROUP consecutive query landmarks of one label in Morton
//                               order; has a bounding box per yaw so that a warp can skip it when
//                               none of its 1024 hypotheses can bring it over the label's cells
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SPR_HD __host__ __device__ __forceinline__
#else
#define SPR_HD inline
#endif

#define SPR_QGROUP 8
#define SPR_Q_SENTINEL (-1073741824)  // -2^30: fixed-point coordinate of a padding query (never inside the grid)

struct SprChunk {        // 32 bytes
  double   across;       // exact fp64 value of the fixed coordinate (x if dir == 0, y if dir == 1)
  uint32_t along_off;    // index into lat[] of the lattice sample under bit 0
  uint32_t valid;        // bit b set <=> sample b is a hypothesis to score
  uint32_t ord_base;     // translation ordinal of bit 0; ordinal(b) = ord_base + b * ord_stride
  uint32_t ord_stride;
  uint32_t dir;          // 0: bits run along y; 1: bits run along x
  uint32_t ring;
};

struct SprGrid {
  double  g0x, g0y;      // cell (0,0) covers [g0x, g0x + c) x [g0y, g0y + c)
  double  S;             // 2^F / c : metres -> fixed-point cell units
  int32_t F;             // fractional bits of the fixed-point cell coordinates
  int32_t GX, GY;        // cells along x and y
  int32_t R[2];          // rows of the plane for dir d (across cells + 2 zero rows)
  int32_t W[2];          // 32-bit words per row (32 pad bits in front, >= 64 behind)
  int32_t maxbit[2];     // clamp of the along bit offset: 32 * (W[d] - 2)
  uint32_t plane_words[2];
  uint32_t label_stride; // plane_words[0] + plane_words[1]
};

// rank tables of one (label, direction) plane, in global or shared memory
#define SPR_CELL_MULTI 0xffffu  // cellref value of a cell with more than one candidate landmark
struct SprTables {
  const uint32_t *bits;      // the plane
  const uint16_t *r16;       // marked cells of the row before each word
  const uint32_t *row_rank;  // rank of each row's first marked cell, relative to the label's first cell
  const uint16_t *cellref;   // [cells of the label] slot of the cell's only candidate, or SPR_CELL_MULTI
  const double   *reftab;    // [landmarks of the label][5] x, y, d1, d2, d3
  uint32_t W;                // words per row
  uint32_t cell_base;        // absolute rank (index into cand[d]) of the label's first cell
};

struct SprBox { int32_t x0, x1, y0, y1; };  // fixed-point, [x0, x1) x [y0, y1); empty if x0 >= x1

// one candidate of a marked cell: everything the exact test needs, in one 48-byte record.
// cand[rank] is the first (lowest reference index) candidate of the cell of that rank; further
// candidates of the same cell are chained through `next` (index into cand, 0 = end).
struct SprCand { double x, y, d1, d2, d3; uint32_t ref; uint32_t next; };

// Everything the scoring code reads.  Pointers are valid in the executing address space
// (device pointers for the kernels; host pointers for the test-only emulation).
struct SprView {
  const double   *lat;        // lattice samples of every ring (x arrays and y arrays)
  const SprChunk *chunks;
  uint32_t        n_chunks;
  int32_t         n_yaw;
  const double   *cs;         // [n_yaw][2] cos(yaw), sin(yaw) from host libm (PR.cpp:246-250)
  int32_t         nqp;        // query landmarks kept (label present in the reference), sorted by
                              // (label, Morton), every label segment padded to SPR_QGROUP
  int32_t         n_groups;   // nqp / SPR_QGROUP
  const int32_t  *qrotq_xy;   // [n_yaw][nqp][2] fixed-point cell coords (x, y) of the rotated query
  const int32_t  *qrotq_yx;   // [n_yaw][nqp][2] the same, swapped (y, x): layout read by dir-1 chunks
  const SprBox   *gbox;       // [n_yaw][n_groups] bounding box of each query group's fixed coords
  const double   *qrot;       // [n_yaw][nqp][2] exact fp64 (c*qx + (-s)*qy, s*qx + c*qy)
  const double   *qxy;        // [nqp][2] query x, y (search frame)
  const double   *qdims;      // [nqp][3]
  const int32_t  *label_gseg; // [n_labels + 1] group index boundaries of the label segments
  const int32_t  *qlabel;     // [nqp] label bucket of each sorted query; -1 for padding
  int32_t         n_labels;
  int32_t         n_ref;
  const SprBox   *labelbox;   // [n_labels] fixed-point bounds of the label's marked cells
  const uint32_t *bitmap;     // [n_labels][plane dir0 | plane dir1]
  const uint32_t *vbitmap;    // [n_labels][4 half-cell variants of plane dir0 | 4 variants of plane dir1] (refined bounds; device only)
  const SprCand  *cand[2];    // per plane direction d: first candidate of each marked cell by rank, then chained extras
  const uint16_t *rank16[2];  // per plane direction d: [n_labels][plane_words[d]] marked cells of the row before each word
  const uint32_t *row_rank[2];// per plane direction d: [n_labels][R[d]] rank of each row's first marked cell (label-relative)
  const uint16_t *cellref[2]; // per plane direction d: [n_cells] candidate slot per marked cell (rank order)
  const uint32_t *cell_base[2];// per plane direction d: [n_labels + 1] rank of each label's first cell
  const double   *reftab;     // [landmarks][5] x, y, d1, d2, d3, label-major
  const uint32_t *ref_base;   // [n_labels + 1] first row of each label in reftab
  SprGrid         grid;
  double          Tstar;      // sqrt(d2) < match_threshold_  <=>  d2 < Tstar   (PR.cpp:332-333)
  double          Sstar;      // (sum / 3) < thr_dim          <=>  sum < Sstar  (PR.cpp:329,338)
  double          thr_dim;
  int32_t         ignore_dim;
  int32_t         pad;
};

// best-hypothesis key: max() picks the larger inlier count, ties go to the smaller canonical
// hypothesis index (the reference's strict '>' first-wins rule, PR.cpp:361).  0 == "none".
#define SPR_KEY_IDX_BITS 40
#define SPR_KEY_IDX_MASK ((1ULL << SPR_KEY_IDX_BITS) - 1ULL)
SPR_HD unsigned long long spr_make_key(uint32_t count, unsigned long long hyp_idx) {
  return ((unsigned long long)(count + 1u) << SPR_KEY_IDX_BITS) | (SPR_KEY_IDX_MASK - hyp_idx);
}
SPR_HD int32_t spr_key_count(unsigned long long key) { return (int32_t)(key >> SPR_KEY_IDX_BITS) - 1; }
SPR_HD long long spr_key_index(unsigned long long key) {
  return (long long)(SPR_KEY_IDX_MASK - (key & SPR_KEY_IDX_MASK));
}
