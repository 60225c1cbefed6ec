// spr_host.h -- host side of the place-recognition search: lattice enumeration, chunking,
// label bucketing, occupancy bitmaps and candidate lists, closed-form refinement.
// Pure C++ (no CUDA) so the index structures can be unit-tested on a CPU-only box.
#pragma once
#include <cstdint>
#include <string>
#include <new>
#include <vector>

#include "../../include/slide_pr.h"
#include "spr_types.h"
#include "spr_join_types.h"

namespace spr {

// Host buffers that are copied to the device are allocated through a pair of hooks so that the
// library can make them page-locked (true asynchronous DMA, no staging copy); the default hooks
// are malloc / free (tests, emulation).
extern void *(*g_upload_alloc)(size_t bytes);
extern void (*g_upload_free)(void *p);
template <class T>
struct UploadAlloc {
  using value_type = T;
  UploadAlloc() = default;
  template <class U> UploadAlloc(const UploadAlloc<U> &) {}
  T *allocate(size_t n) {
    void *p = g_upload_alloc(n * sizeof(T));
    if (!p) throw std::bad_alloc();
    return static_cast<T *>(p);
  }
  void deallocate(T *p, size_t) { g_upload_free(p); }
  template <class U> bool operator==(const UploadAlloc<U> &) const { return true; }
  template <class U> bool operator!=(const UploadAlloc<U> &) const { return false; }
};
template <class T> using uvec = std::vector<T, UploadAlloc<T>>;

struct Lattice {
  int status = 0;                 // 0 ok, SLIDE_PR_SANITY_RETURN on PR.cpp:169-175
  int rings = 0;
  double ox = 0, oy = 0;          // outer_loop_step_size_{x,y}
  std::vector<double> yaw;        // PR.cpp:136-146
  uvec<double> cs;         // cos, sin per yaw (libm)
  uvec<double> lat;        // per ring: xs then ys
  struct Ring {
    uint32_t x_off, nx, y_off, ny;  // into lat
    int ixl, ixh, iyl, iyh;         // inner (already searched) box as closed index ranges; empty if l > h
    uint64_t ord_base;              // ordinal of the ring's first translation
    uint64_t count;                 // translations in the ring
    uint32_t chunk_begin, chunk_end;    // emission order (before grouping by direction)
    uint32_t dbegin[2] = {0, 0}, dend[2] = {0, 0};  // ring_major: this ring's chunks of direction d
  };
  std::vector<Ring> ring;
  uint64_t n_translations = 0;
  double drift = 0;               // largest deviation of a sample from (first sample of its axis) + index * step, over all rings
  uvec<SprChunk> chunks;   // grouped by direction (see dir_begin / Ring::dbegin), warp-padded
  uvec<SprChunk> scratch;  // regrouping buffer, kept for its capacity
  std::vector<int32_t> succ;      // scratch of build_lattice (kept for their capacity)
  std::vector<uint32_t> sort_pos, sort_key;
  std::vector<uint8_t> sort_second;
  std::vector<uint32_t> dg_bits;  // [chunks / 64] valid bits (lattice translations) of every double group
  bool ring_major = false;
  bool has_chunks = false;        // false after a rings_only build
  uint32_t dir_begin[2] = {0, 0}, dir_end[2] = {0, 0};  // !ring_major: all chunks of direction d
};

// PR.cpp:136-241.  yaw_half is match_yaw_half_range_ (inter) or the intra value.
// trans_begin/trans_end (end < 0: none) restrict the valid bits to a range of ordinals.
// rings_only: yaw table, ring geometry and lattice samples only, no chunks (all the pair-join scorer needs).
int build_lattice(const slide_pr_params &p, double half_x, double half_y, double yaw_half,
                  int64_t trans_begin, int64_t trans_end, bool ring_major, Lattice &L, std::string &err,
                  bool rings_only = false);

// translation (x, y) of a canonical ordinal; false if out of range
bool translation_of(const Lattice &L, uint64_t ordinal, double *x, double *y, int *ring);

struct RefIndex {
  std::vector<double> labels;       // distinct finite reference labels, ascending
  SprGrid grid{};
  uvec<uint32_t> bitmap;     // [n_labels][plane0 | plane1]
  uvec<SprCand> cand[2];      // per plane direction d: [n_cells] first candidate per cell rank, then chained extras
  uvec<uint16_t> rank16[2];   // per plane direction d: [n_labels][plane_words[d]] marked cells of the row before the word
  uvec<uint32_t> row_rank[2]; // per plane direction d: [n_labels][R[d]] rank (index into cand[d]) of the row's first marked cell
  uvec<uint16_t> cellref[2];  // per plane direction d: [n_cells] slot of the cell's only candidate in its label's table, or SPR_CELL_MULTI
  uvec<uint32_t> cell_base[2]; // per plane direction d: [n_labels + 1] rank of each label's first cell
  uvec<double> reftab;        // [n_ref kept][5] x, y, d1, d2, d3, label-major
  uvec<uint32_t> ref_base;    // [n_labels + 1] first row of each label in reftab
  uvec<SprBox> labelbox;     // [n_labels] fixed-point bounds of the label's marked cells
  double Tstar = 0, Sstar = 0;
  double mark_rc = 0, mark_rc2 = 0; // radius (cells) / squared radius of the marking predicate (0: nothing matches)
  int n_ref = 0;
  double reach_limit = 0;           // largest reach covered by grid.F (set by build_ref_grid)
  // carried from build_ref_bitmaps to build_ref_ranks
  struct Entry { uint32_t ref; int32_t nx, ny; };  // landmark `ref` marks cell (nx, ny); ascending landmark order
  std::vector<Entry> entries;
  uvec<int32_t> lab_of;       // [n_ref] label bucket, -1: NaN label
  std::vector<uint32_t> slot_of_ref; // [n_ref] position inside its label
};

// reach: max |coordinate| any transformed query point or translation can take (metres);
// fixes the fixed-point format.  Returns SLIDE_PR_OK or an error code.
int build_ref_index(const slide_pr_params &p, const double *ref7, int n_ref, double reach,
                    RefIndex &R, std::string &err);
// The stages of build_ref_index: labels + grid (all build_query_set needs), the occupancy bitmaps
// (all the bound phase of the search needs) and, per bitmap direction, the rank tables / candidate
// records of the exact verification.
int build_ref_grid(const slide_pr_params &p, const double *ref7, int n_ref, double reach,
                   RefIndex &R, std::string &err);
int build_ref_marks(const slide_pr_params &p, const double *ref7, int n_ref, RefIndex &R, std::string &err);
int build_ref_ranks(const double *ref7, int d, RefIndex &R, std::string &err);

struct QuerySet {
  int nq = 0;                       // kept queries (label present in the reference)
  int nqp = 0;                      // with every label segment padded to SPR_QGROUP
  std::vector<int32_t> orig;        // [nqp] original index, -1 for padding
  uvec<double> qxy;          // [nqp][2]
  uvec<double> qdims;        // [nqp][3]
  uvec<int32_t> qlabel;      // [nqp] label bucket, -1 for padding
  uvec<int32_t> label_gseg;  // [n_labels + 1] group (SPR_QGROUP queries) boundaries
};
int build_query_set(const RefIndex &R, const double *qry7, int n_qry, QuerySet &Q, std::string &err);
int build_query_set(const std::vector<double> &labels, const double *qry7, int n_qry, QuerySet &Q, std::string &err);
void unique_labels(const double *rows7, int n, std::vector<double> &labels);

// thresholds: sqrt(d2) < thr <=> d2 < Tstar ; (s / 3) < thr_dim <=> s < Sstar
double sqrt_threshold(double thr);
double div3_threshold(double thr_dim);

// PlaceRecognition::solveLSQ (PR.cpp:632-695) / getxyzYawfromTF (PR.cpp:697-711)
void solve_lsq(const double *tgt3, const double *src3, int k, double *xyz_yaw4, double *tf16);
void xyz_yaw_from_tf(const double *tf16, double *xyz_yaw4);
// semantic_clipper::estimate_tf (SC.cpp:122-138): 2-D Kabsch a -> b, tf9 row-major
void estimate_tf(const double *a2, const double *b2, int k, double *tf9);
void svd_jacobi(const double *A, int n, double *U, double *S, double *V);
void mat4_mul(const double *A, const double *B, double *C);
bool mat4_rigid_inverse(const double *A, double *Ainv);


// ------------------------------------------------------------------------------------------
// pair-join scorer (spr_join.h): reference landmarks binned by label and coarse grid cell in both
// join directions, lower-index neighbours of every landmark, and the lattice cut into blocks
// ------------------------------------------------------------------------------------------
struct JoinRef {
  std::vector<double> labels;        // distinct finite reference labels, ascending
  uvec<SprJoinRef> rec[2];
  uvec<double> xy[2];                // [records][2]: the coordinates alone, for the candidate filter
  uvec<uint32_t> cell_start[2];      // [n_labels * n_cells + 1] first record of (label, cell) (+ one spare entry)
  uvec<SprJoinNbr> nbr;
  uvec<double> labelbox;             // [n_labels][4]
  double gx0 = 0, gy0 = 0, w = 1, inv_w = 1;
  int ncx = 1, ncy = 1;
  double Tstar = 0, Sstar = 0;
  double reach = 0;                  // a match implies |dx|, |dy| < reach (at ordinary coordinate magnitudes; see max_abs)
  double max_abs = 0;                // largest |coordinate| of the map
  int n_ref = 0;
};
int build_join_ref(const slide_pr_params &p, const double *ref7, int n_ref, JoinRef &J, std::string &err);
// blocks of the lattice L (needs L.ring / L.lat); *drift: L.drift, the largest deviation of a sample from
// first sample + index * step over all rings
int build_join_blocks(const Lattice &L, double step, uvec<SprJoinBlock> &blocks, double *drift, std::string &err);

}  // namespace spr
