// spr_host.cpp -- host index builder for the B200 place-recognition search (see spr_host.h).
#include "spr_host.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>

namespace spr {

// ------------------------------------------------------------------------------------------
// thresholds
// ------------------------------------------------------------------------------------------
void *(*g_upload_alloc)(size_t) = [](size_t n) -> void * { return std::malloc(n ? n : 1); };
void (*g_upload_free)(void *) = [](void *p) { std::free(p); };

double sqrt_threshold(double thr) {
  // smallest double T with sqrt(T) >= thr, so that (sqrt(d2) < thr) == (d2 < T)  [PR.cpp:332-333]
  if (!(thr > 0)) return 0.0;  // also NaN: nothing passes
  if (std::isinf(thr)) return thr;
  double t = thr * thr;
  for (int i = 0; i < 64 && std::sqrt(t) >= thr; i++) t = std::nextafter(t, -HUGE_VAL);
  for (int i = 0; i < 64 && std::sqrt(std::nextafter(t, HUGE_VAL)) < thr; i++) t = std::nextafter(t, HUGE_VAL);
  return std::nextafter(t, HUGE_VAL);
}

double div3_threshold(double thr) {
  // smallest double S with (S / 3) >= thr, so that ((sum / 3) < thr) == (sum < S)  [PR.cpp:329,338]
  if (!(thr > 0)) return 0.0;
  if (std::isinf(thr)) return thr;
  double t = thr * 3.0;
  for (int i = 0; i < 64 && t / 3 >= thr; i++) t = std::nextafter(t, -HUGE_VAL);
  for (int i = 0; i < 64 && std::nextafter(t, HUGE_VAL) / 3 < thr; i++) t = std::nextafter(t, HUGE_VAL);
  return std::nextafter(t, HUGE_VAL);
}

// ------------------------------------------------------------------------------------------
// lattice + chunks
// ------------------------------------------------------------------------------------------
namespace {

struct RectEmitter {
  Lattice &L;
  int ring;
  uint32_t x_off, y_off;
  const double *xs, *ys;
  int64_t tb, te;  // ordinal filter, te < 0: none
  size_t &n_out;   // chunks emitted so far (L.chunks / succ are sized to capacity while emitting)

  std::vector<int32_t> *succ = nullptr;  // optional: index of the chunk that continues chunk i along its row, or -1
  std::vector<int32_t> prev;             // scratch: chunks of the previous block of the current rectangle

  int push(double across, uint32_t along_off, int n, uint64_t ord0, uint64_t stride, uint32_t dir) {
    // restrict to ordinals in [tb, te)
    int b_lo = 0, b_hi = n;
    if (tb > 0 || te >= 0) {
      const int64_t o0 = (int64_t)ord0, st = (int64_t)stride;
      if (tb > o0) b_lo = (int)std::min<int64_t>(n, (tb - o0 + st - 1) / st);
      if (te >= 0) b_hi = (te <= o0) ? 0 : (int)std::min<int64_t>(n, (te - o0 + st - 1) / st);
    }
    if (b_lo >= b_hi) return -1;
    SprChunk c;
    c.across = across;
    c.along_off = along_off;
    const uint32_t hi_mask = b_hi >= 32 ? 0xffffffffu : ((1u << b_hi) - 1u);
    const uint32_t lo_mask = b_lo <= 0 ? 0u : ((1u << b_lo) - 1u);
    c.valid = hi_mask & ~lo_mask;
    c.ord_base = (uint32_t)ord0;
    c.ord_stride = (uint32_t)stride;
    c.dir = dir;
    c.ring = (uint32_t)ring;
    if (n_out == L.chunks.size()) {  // grow geometrically; the vectors keep their size across calls
      const size_t cap = std::max<size_t>(2 * n_out, 4096);
      L.chunks.resize(cap);
      L.scratch.reserve(cap + 128);  // the two buffers swap roles every call: grow them together
      if (succ) succ->resize(cap);
    }
    L.chunks[n_out] = c;
    if (succ) (*succ)[n_out] = -1;
    return (int)n_out++;
  }

  // rectangle ix in [ix0, ix1), iy in [iy0, iy1); ordinal(ix, iy) = ord0 + (ix-ix0)*row_stride + (iy-iy0)
  void rect(int ix0, int ix1, int iy0, int iy1, uint64_t ord0, uint64_t row_stride) {
    const int w = ix1 - ix0, h = iy1 - iy0;
    if (w <= 0 || h <= 0) return;
    const int64_t cost_y = (int64_t)w * ((h + 31) / 32);
    const int64_t cost_x = (int64_t)h * ((w + 31) / 32);
    if (cost_y <= cost_x) {
      // bits along y; consecutive chunks = consecutive x rows of the same y block
      prev.assign((size_t)w, -1);
      for (int by = iy0; by < iy1; by += 32) {
        const int n = std::min(32, iy1 - by);
        for (int ix = ix0; ix < ix1; ix++) {
          const int i = push(xs[ix], y_off + (uint32_t)by, n, ord0 + (uint64_t)(ix - ix0) * row_stride + (uint64_t)(by - iy0), 1, 0);
          if (succ && i >= 0 && prev[ix - ix0] >= 0) (*succ)[prev[ix - ix0]] = i;
          prev[ix - ix0] = i;
        }
      }
    } else {
      prev.assign((size_t)h, -1);
      for (int bx = ix0; bx < ix1; bx += 32) {
        const int n = std::min(32, ix1 - bx);
        for (int iy = iy0; iy < iy1; iy++) {
          const int i = push(ys[iy], x_off + (uint32_t)bx, n, ord0 + (uint64_t)(bx - ix0) * row_stride + (uint64_t)(iy - iy0), row_stride, 1);
          if (succ && i >= 0 && prev[iy - iy0] >= 0) (*succ)[prev[iy - iy0]] = i;
          prev[iy - iy0] = i;
        }
      }
    }
  }
};

}  // namespace

// One axis of a ring's lattice: the samples start, start + step, ... <= end by repeated fp64 addition, exactly
// as the reference's loops (PR.cpp:230-232), written to out[0 .. cap).  In the shadow of the addition chain:
// the closed index range [*lo, *hi] of the samples inside [in_lo, in_hi] (left untouched when there is none) and
// the largest deviation of a sample from start + i * step (folded into *drift).  Returns the count, cap + 1 when
// cap does not suffice.
static uint32_t accumulate_samples(double start, double end, double step, double *out, uint32_t cap, double in_lo, double in_hi,
                                   int *lo, int *hi, double *drift) {
  uint32_t n = 0;
  int l = -1, h = -1;
  double d0 = 0.0, d1 = 0.0;
  for (double v = start; v <= end; v += step) {
    if (n >= cap) return cap + 1;
    out[n] = v;
    if (v >= in_lo && v <= in_hi) { if (l < 0) l = (int)n; h = (int)n; }
    const double dev = std::fabs(v - (start + (double)n * step));
    if (n & 1u) d1 = std::max(d1, dev); else d0 = std::max(d0, dev);
    n++;
  }
  if (l >= 0) { *lo = l; *hi = h; }
  *drift = std::max(*drift, std::max(d0, d1));
  return n;
}

int build_lattice(const slide_pr_params &p, double half_x, double half_y, double yaw_half,
                  int64_t trans_begin, int64_t trans_end, bool ring_major, Lattice &L, std::string &err,
                  bool rings_only) {
  // reset, keeping the vectors' capacity across calls
  L.status = 0; L.rings = 0; L.ox = L.oy = 0; L.n_translations = 0; L.ring_major = false; L.drift = 0;
  L.yaw.clear(); L.cs.clear(); L.lat.clear(); L.ring.clear();
  // L.chunks keeps its size while chunks are emitted (n_emitted tracks the fill level); every exit
  // before the emission is complete leaves it empty
  size_t n_emitted = 0;
  struct ChunkGuard {
    Lattice &L; bool done = false;
    ~ChunkGuard() { if (!done) L.chunks.clear(); }
  } chunk_guard{L};
  L.dir_begin[0] = L.dir_begin[1] = L.dir_end[0] = L.dir_end[1] = 0;
  const double step = p.match_xy_step_size;
  if (!(step > 0) || !std::isfinite(step)) { err = "match_xy_step_size must be positive and finite"; return SLIDE_PR_ERR_INVALID; }
  if (!std::isfinite(half_x) || !std::isfinite(half_y)) { err = "non-finite search half range"; return SLIDE_PR_ERR_NONFINITE; }
  // yaw candidates, PR.cpp:136-146
  if (p.disable_yaw_search) {
    L.yaw.push_back(0.0);
  } else {
    const double ys = p.match_yaw_angle_step_size;
    if (!std::isfinite(yaw_half) || !std::isfinite(ys)) { err = "non-finite yaw range"; return SLIDE_PR_ERR_NONFINITE; }
    if (yaw_half > 0 && !(ys > 0)) { err = "match_yaw_angle_step_size must be positive"; return SLIDE_PR_ERR_INVALID; }
    for (double yaw_raw = -yaw_half; yaw_raw < yaw_half; yaw_raw += ys) {
      L.yaw.push_back(yaw_raw);
      if (L.yaw.size() > (1u << 20)) { err = "more than 2^20 yaw candidates"; return SLIDE_PR_ERR_UNSUPPORTED; }
    }
  }
  L.cs.resize(2 * L.yaw.size());
  for (size_t a = 0; a < L.yaw.size(); a++) {  // libm, as PR.cpp:246-250
    L.cs[2 * a] = std::cos(L.yaw[a]);
    L.cs[2 * a + 1] = std::sin(L.yaw[a]);
  }
  // ring geometry, PR.cpp:154-175
  const double outer = 10 * step;
  const double steps_d = std::min(half_x, half_y) / outer;
  const double steps_c = std::ceil(steps_d);
  if (steps_c > 1e6) { err = "more than 1e6 search rings"; return SLIDE_PR_ERR_UNSUPPORTED; }
  const int rings = (int)steps_c;
  L.rings = rings;
  L.ox = half_x / static_cast<double>(rings);
  L.oy = half_y / static_cast<double>(rings);
  if (L.ox < step || L.oy < step) { L.status = SLIDE_PR_SANITY_RETURN; return SLIDE_PR_OK; }

  uint64_t ord = 0;
  std::vector<int32_t> &succ = L.succ;  // along-successor of every emitted chunk (pairing for the bound kernel)
  if (succ.size() < L.chunks.size()) succ.resize(L.chunks.size());
  for (int k = 0; k < rings; k++) {
    const double kd = static_cast<double>(k);
    const double x_right_prev = kd * L.ox, x_left_prev = -kd * L.ox;        // PR.cpp:204,210
    const double x_pos_end = (kd + 1) * L.ox, x_neg_start = -(kd + 1) * L.ox;  // :205-206
    const double y_right_prev = kd * L.oy, y_left_prev = -kd * L.oy;
    const double y_pos_end = (kd + 1) * L.oy, y_neg_start = -(kd + 1) * L.oy;
    Lattice::Ring R;
    // the samples come from repeated fp64 addition, exactly as the reference's loops; the arrays are
    // sized from the (generous) expected count first so that the loops write through a pointer
    const double est_x = (x_pos_end - x_neg_start) / step + 4.0, est_y = (y_pos_end - y_neg_start) / step + 4.0;
    if (!(est_x < 4194304.0) || !(est_y < 4194304.0)) { err = "more than 2^22 lattice samples per axis"; return SLIDE_PR_ERR_UNSUPPORTED; }
    const size_t lat0 = L.lat.size();
    L.lat.resize(lat0 + (size_t)est_x + (size_t)est_y);
    double *lp = L.lat.data() + lat0;
    R.x_off = (uint32_t)lat0;
    // samples, already-searched centre box as closed index ranges (PR.cpp:238-239) and drift, in one pass per axis
    R.ixl = 0; R.ixh = -1; R.iyl = 0; R.iyh = -1;
    uint32_t n = accumulate_samples(x_neg_start, x_pos_end, step, lp, (uint32_t)est_x, x_left_prev, x_right_prev, &R.ixl, &R.ixh, &L.drift);  // PR.cpp:230
    if (n > (uint32_t)est_x) { err = "internal: lattice sample estimate too small"; return SLIDE_PR_ERR_INTERNAL; }
    R.nx = n;
    R.y_off = R.x_off + n;
    lp += n;
    if (y_neg_start == x_neg_start && y_pos_end == x_pos_end && y_left_prev == x_left_prev && y_right_prev == x_right_prev &&
        n <= (uint32_t)est_y) {
      // square search range (the usual case, PR.cpp:777-782): the y loop adds the same numbers in the same order
      std::memcpy(lp, lp - n, (size_t)n * sizeof(double));
      R.iyl = R.ixl; R.iyh = R.ixh;
    } else {
      n = accumulate_samples(y_neg_start, y_pos_end, step, lp, (uint32_t)est_y, y_left_prev, y_right_prev, &R.iyl, &R.iyh, &L.drift);         // PR.cpp:232
      if (n > (uint32_t)est_y) { err = "internal: lattice sample estimate too small"; return SLIDE_PR_ERR_INTERNAL; }
    }
    R.ny = n;
    L.lat.resize((size_t)R.y_off + n);
    // lat may have been reallocated: take pointers now
    const double *xs = L.lat.data() + R.x_off, *ys = L.lat.data() + R.y_off;
    const int n_in_x = R.ixh >= R.ixl ? R.ixh - R.ixl + 1 : 0;
    const int n_in_y = R.iyh >= R.iyl ? R.iyh - R.iyl + 1 : 0;
    const bool has_box = n_in_x > 0 && n_in_y > 0;
    R.ord_base = ord;
    R.count = (uint64_t)R.nx * R.ny - (has_box ? (uint64_t)n_in_x * n_in_y : 0);
    R.chunk_begin = (uint32_t)n_emitted;
    RectEmitter E{L, k, R.x_off, R.y_off, xs, ys, trans_begin, trans_end, n_emitted};
    E.succ = ring_major ? nullptr : &succ;
    const int nx = (int)R.nx, ny = (int)R.ny;
    if (rings_only) {
      // ring geometry and samples only
    } else if (!has_box) {
      E.rect(0, nx, 0, ny, ord, (uint64_t)ny);
    } else {
      const uint64_t cnt_in = (uint64_t)(ny - n_in_y);
      const uint64_t base_b = ord + (uint64_t)R.ixl * ny;
      const uint64_t base_c = base_b + (uint64_t)n_in_x * cnt_in;
      E.rect(0, R.ixl, 0, ny, ord, (uint64_t)ny);                                   // x below the box
      E.rect(R.ixl, R.ixh + 1, 0, R.iyl, base_b, cnt_in);                           // y below the box
      E.rect(R.ixl, R.ixh + 1, R.iyh + 1, ny, base_b + (uint64_t)R.iyl, cnt_in);     // y above the box
      E.rect(R.ixh + 1, nx, 0, ny, base_c, (uint64_t)ny);                           // x above the box
    }
    R.chunk_end = (uint32_t)n_emitted;
    ord += R.count;
    L.ring.push_back(R);
    if (ord >= (1ull << 32)) { err = "more than 2^32 lattice translations"; return SLIDE_PR_ERR_UNSUPPORTED; }
  }
  L.chunks.resize(n_emitted);
  chunk_guard.done = true;
  L.n_translations = ord;
  if (ord * (uint64_t)std::max<size_t>(L.yaw.size(), 1) >= (1ull << SPR_KEY_IDX_BITS)) {
    err = "more than 2^40 hypotheses"; return SLIDE_PR_ERR_UNSUPPORTED;
  }
  L.has_chunks = !rings_only;
  if (rings_only) {
    L.dg_bits.clear();
    L.ring_major = ring_major;
    return SLIDE_PR_OK;
  }
  // Group the chunks by direction (every warp = 32 consecutive chunks probes ONE bitmap plane),
  // padding each group to a whole number of warps with empty chunks.  ring_major keeps the
  // chunks of a ring together (needed by the anytime budget, PR.cpp:181-191).
  uvec<SprChunk> &out = L.scratch;
  if (ring_major) out.clear();
  out.reserve(L.chunks.size() + 128 * (ring_major ? L.ring.size() + 1 : 2));
  auto pad32 = [&out](uint32_t d) {  // whole double groups (the sharding granule)
    SprChunk z{};
    z.dir = d;
    while (out.size() % 64) out.push_back(z);
  };
  if (ring_major) {
    for (Lattice::Ring &R : L.ring)
      for (uint32_t d = 0; d < 2; d++) {
        R.dbegin[d] = (uint32_t)out.size();
        for (uint32_t c = R.chunk_begin; c < R.chunk_end; c++)
          if (L.chunks[c].dir == d) out.push_back(L.chunks[c]);
        pad32(d);
        R.dend[d] = (uint32_t)out.size();
      }
  } else {
    // Chunks are paired along their row (chunk + the chunk that continues it 32 samples further:
    // the bound kernel probes both with three bitmap words instead of four) and the pairs are
    // bucketed by the 32 x 32-sample tile the first chunk starts in (counting sort, emission
    // order kept inside a bucket).  A work item of the bound kernel is a DOUBLE GROUP of 64
    // chunks = [32 first chunks][their 32 continuations, empty where a row ends]; the 32 chunks
    // of each half lie side by side across -- neighbouring bitmap rows at the same word column
    // (mostly distinct shared-memory banks with the odd row pitch) and a compact patch for the
    // visibility test.  The exact kernel sees two ordinary groups of 32 chunks.
    const double tile = 32.0 * step;
    const double lo = -std::max(std::fabs(half_x), std::fabs(half_y)) - tile;
    const size_t nb1 = (size_t)std::floor((-2.0 * lo) / tile) + 2;
    if (nb1 > 20000) { err = "search range spans more than 20000 chunk tiles per axis"; return SLIDE_PR_ERR_UNSUPPORTED; }
    const size_t n_buckets = nb1 * nb1;
    const double inv_tile = 1.0 / tile;
    const size_t n_all = L.chunks.size();
    std::vector<uint32_t> &pos = L.sort_pos, &key = L.sort_key;  // scratch kept across calls
    std::vector<uint8_t> &second = L.sort_second;
    pos.assign(2 * (n_buckets + 1), 0u);
    key.resize(n_all);                        // bucket of a pair's first chunk
    second.assign(n_all, 0);                  // chunk is the continuation of an earlier one
    const double *lat = L.lat.data();
    const double nb1_max = (double)(nb1 - 1);
    for (size_t i = 0; i < n_all; i++) {
      if (second[i]) continue;
      const SprChunk &c = L.chunks[i];
      if (succ[i] >= 0) second[(size_t)succ[i]] = 1;
      // tile indices: (v - lo) / tile >= 1 by construction of lo
      const double ta = std::min((lat[c.along_off] - lo) * inv_tile, nb1_max);
      const double tc = std::min((c.across - lo) * inv_tile, nb1_max);
      const uint32_t bi = (uint32_t)(ta > 0 ? ta : 0) * (uint32_t)nb1 + (uint32_t)(tc > 0 ? tc : 0);
      key[i] = bi;
      pos[(size_t)c.dir * (n_buckets + 1) + bi + 1]++;
    }
    size_t start[2], units[2];
    for (uint32_t d = 0; d < 2; d++) {
      uint32_t *pd = pos.data() + (size_t)d * (n_buckets + 1);
      for (size_t b2 = 0; b2 < n_buckets; b2++) pd[b2 + 1] += pd[b2];
      units[d] = pd[n_buckets];
      start[d] = d == 0 ? 0 : start[0] + ((units[0] + 31) / 32) * 64;
    }
    const size_t total = start[1] + ((units[1] + 31) / 32) * 64;
    if (out.size() != total) out.resize(total);  // usually the size of the previous call
    SprChunk empty[2] = {SprChunk{}, SprChunk{}};  // empty chunks (valid == 0) carry their direction
    empty[1].dir = 1;
    for (size_t i = 0; i < n_all; i++) {
      if (second[i]) continue;
      const uint32_t d = L.chunks[i].dir;
      const size_t u = pos[(size_t)d * (n_buckets + 1) + key[i]]++;
      const size_t slot = start[d] + (u / 32) * 64 + (u % 32);
      out[slot] = L.chunks[i];
      out[slot + 32] = succ[i] >= 0 ? L.chunks[(size_t)succ[i]] : empty[d];
    }
    for (uint32_t d = 0; d < 2; d++) {
      L.dir_begin[d] = (uint32_t)start[d];
      L.dir_end[d] = (uint32_t)(d == 0 ? start[1] : total);
      for (size_t u = units[d]; u < ((units[d] + 31) / 32) * 32; u++) {  // unused lanes of the last double group
        const size_t slot = start[d] + (u / 32) * 64 + (u % 32);
        out[slot] = empty[d];
        out[slot + 32] = empty[d];
      }
    }
  }
  L.ring_major = ring_major;
  L.chunks.swap(out);
  L.dg_bits.assign(L.chunks.size() / 64, 0u);  // every direction / ring segment is padded to whole double groups
  for (size_t g = 0; g < L.dg_bits.size(); g++) {
    uint32_t bits = 0;
    for (size_t c = 64 * g; c < 64 * g + 64; c++) bits += (uint32_t)__builtin_popcount(L.chunks[c].valid);
    L.dg_bits[g] = bits;
  }
  return SLIDE_PR_OK;
}

bool translation_of(const Lattice &L, uint64_t ordinal, double *x, double *y, int *ring_out) {
  for (size_t k = 0; k < L.ring.size(); k++) {
    const Lattice::Ring &R = L.ring[k];
    if (ordinal < R.ord_base || ordinal >= R.ord_base + R.count) continue;
    uint64_t o = ordinal - R.ord_base;
    const double *xs = L.lat.data() + R.x_off, *ys = L.lat.data() + R.y_off;
    const int n_in_x = R.ixh >= R.ixl ? R.ixh - R.ixl + 1 : 0;
    const int n_in_y = R.iyh >= R.iyl ? R.iyh - R.iyl + 1 : 0;
    const bool has_box = n_in_x > 0 && n_in_y > 0;
    int ix, iy;
    if (!has_box) {
      ix = (int)(o / R.ny); iy = (int)(o % R.ny);
    } else {
      const uint64_t a = (uint64_t)R.ixl * R.ny, cnt_in = R.ny - n_in_y, b = (uint64_t)n_in_x * cnt_in;
      if (o < a) { ix = (int)(o / R.ny); iy = (int)(o % R.ny); }
      else if (o < a + b) {
        o -= a; ix = R.ixl + (int)(o / cnt_in); iy = (int)(o % cnt_in);
        if (iy >= R.iyl) iy += n_in_y;
      } else { o -= a + b; ix = R.ixh + 1 + (int)(o / R.ny); iy = (int)(o % R.ny); }
    }
    *x = xs[ix]; *y = ys[iy];
    if (ring_out) *ring_out = (int)k;
    return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------
// reference-map index: label buckets, occupancy bitmaps, candidate lists
// ------------------------------------------------------------------------------------------
int build_ref_index(const slide_pr_params &p, const double *ref7, int n_ref, double reach,
                    RefIndex &R, std::string &err) {
  int rc = build_ref_grid(p, ref7, n_ref, reach, R, err);
  if (rc == SLIDE_PR_OK) rc = build_ref_marks(p, ref7, n_ref, R, err);
  for (int d = 0; d < 2 && rc == SLIDE_PR_OK; d++) rc = build_ref_ranks(ref7, d, R, err);
  return rc;
}

// Stage 1a: label buckets and the grid / fixed-point format (all the query set needs).
int build_ref_grid(const slide_pr_params &p, const double *ref7, int n_ref, double reach,
                   RefIndex &R, std::string &err) {
  R.labels.clear();  // the vectors are reused across calls (no re-allocation / page faults)
  R.n_ref = n_ref;
  const double c = p.match_xy_step_size, thr = p.match_threshold;
  R.Tstar = sqrt_threshold(thr);
  R.Sstar = div3_threshold(p.match_threshold_dimension);
  double minx = HUGE_VAL, maxx = -HUGE_VAL, miny = HUGE_VAL, maxy = -HUGE_VAL;
  for (int i = 0; i < n_ref; i++) {
    const double *r = ref7 + 7 * (size_t)i;
    if (!std::isfinite(r[1]) || !std::isfinite(r[2])) { err = "non-finite reference coordinate"; return SLIDE_PR_ERR_NONFINITE; }
    minx = std::min(minx, r[1]); maxx = std::max(maxx, r[1]);
    miny = std::min(miny, r[2]); maxy = std::max(maxy, r[2]);
    if (r[0] == r[0]) R.labels.push_back(r[0] + 0.0);  // NaN labels never compare equal (PR.cpp:306)
  }
  std::sort(R.labels.begin(), R.labels.end());
  R.labels.erase(std::unique(R.labels.begin(), R.labels.end()), R.labels.end());
  const int n_labels = (int)R.labels.size();
  if (n_ref == 0) { minx = maxx = miny = maxy = 0; }

  const bool matchable = thr > 0 && std::isfinite(thr);
  if (thr > 0 && std::isinf(thr)) { err = "infinite match_threshold is not supported"; return SLIDE_PR_ERR_UNSUPPORTED; }
  const double rad_m = matchable ? thr * (1.0 + 1e-9) + 1e-9 : 0.0;  // covers fp64 rounding of the reference's test
  SprGrid &G = R.grid;
  G.g0x = minx - rad_m - 2 * c;
  G.g0y = miny - rad_m - 2 * c;
  const double ex = (maxx + rad_m + 2 * c - G.g0x) / c, ey = (maxy + rad_m + 2 * c - G.g0y) / c;
  if (!(ex < 1e6) || !(ey < 1e6)) { err = "reference map spans more than 1e6 lattice steps"; return SLIDE_PR_ERR_UNSUPPORTED; }
  G.GX = (int32_t)std::ceil(ex) + 1;
  G.GY = (int32_t)std::ceil(ey) + 1;
  // fixed-point format: every |coordinate - g0| and every translation must stay below 2^30 units
  const double far = reach + std::max({std::fabs(G.g0x), std::fabs(G.g0y), std::fabs(G.g0x + G.GX * c),
                                       std::fabs(G.g0y + G.GY * c)}) + 4 * c;
  const double cells = far / c + 2.0;
  int F = 16;
  while (F > 0 && cells * std::ldexp(1.0, F) >= 1073741824.0) F--;
  if (F < 6) { err = "search region too large relative to match_xy_step_size for the fixed-point cell format"; return SLIDE_PR_ERR_UNSUPPORTED; }
  G.F = F;
  G.S = std::ldexp(1.0, F) / c;
  // the largest reach this fixed-point format still covers: a cached index stays valid for any query map /
  // search range up to it (the grid itself does not depend on the reach)
  R.reach_limit = reach + ((std::ldexp(1.0, 30 - F) - 3.0) * c - far);
  if (!(R.reach_limit > reach)) R.reach_limit = reach;
  // plane d: rows = across cells + 2 zero rows; bits = 32 pad + along cells + >= 64 pad
  const int across[2] = {G.GX, G.GY}, along[2] = {G.GY, G.GX};
  for (int d = 0; d < 2; d++) {
    G.R[d] = (across[d] + 2 + 7) & ~7;  // 2 zero rows + padding rows: plane sizes stay multiples of 32 bytes (bulk copies)
    int W = (32 + along[d] + 32 + 31) / 32 + 1;
    if ((W & 1) == 0) W++;  // odd row pitch: consecutive rows fall in different shared-memory banks
    G.W[d] = W;
    G.maxbit[d] = 32 * (W - 2);
    const uint64_t words = (uint64_t)G.R[d] * (uint64_t)W;
    if (words >= (1ull << 27)) { err = "occupancy bitmap too large (step too fine for this map extent)"; return SLIDE_PR_ERR_UNSUPPORTED; }
    G.plane_words[d] = (uint32_t)words;
  }
  G.label_stride = G.plane_words[0] + G.plane_words[1];
  const uint64_t total_words = (uint64_t)G.label_stride * (uint64_t)std::max(n_labels, 1);
  if (total_words >= (1ull << 28)) { err = "occupancy bitmaps exceed 1 GiB (step too fine for this map extent)"; return SLIDE_PR_ERR_UNSUPPORTED; }
  return SLIDE_PR_OK;
}

// Stage 1b: marks the occupancy bitmaps of both directions, fills the per-label landmark tables
// and the label boxes (needs build_ref_grid).
int build_ref_marks(const slide_pr_params &p, const double *ref7, int n_ref, RefIndex &R, std::string &err) {
  const double c = p.match_xy_step_size, thr = p.match_threshold;
  const int n_labels = (int)R.labels.size();
  const bool matchable = thr > 0 && std::isfinite(thr);
  const double rad_m = matchable ? thr * (1.0 + 1e-9) + 1e-9 : 0.0;  // covers fp64 rounding of the reference's test
  SprGrid &G = R.grid;
  const int F = G.F;
  const uint64_t total_words = (uint64_t)G.label_stride * (uint64_t)std::max(n_labels, 1);
  R.bitmap.assign((size_t)total_words, 0u);
  // per-label bounds of the marked cells (empty until a cell is marked)
  struct CellBounds { int x0, x1, y0, y1; };
  std::vector<CellBounds> cb((size_t)std::max(n_labels, 1), CellBounds{INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN});

  // label bucket and slot (position inside its label, ascending index) of every landmark; the
  // compact per-label landmark table [x, y, d1, d2, d3] a cell's 16-bit reference points into
  uvec<int32_t> &lab_of = R.lab_of;
  std::vector<uint32_t> &slot_of_ref = R.slot_of_ref;
  lab_of.assign((size_t)std::max(n_ref, 1), -1);
  slot_of_ref.assign((size_t)std::max(n_ref, 1), 0u);
  R.ref_base.assign((size_t)n_labels + 1, 0u);
  {
    std::vector<uint32_t> per((size_t)std::max(n_labels, 1), 0u);
    for (int i = 0; i < n_ref; i++) {
      const double lv = ref7[7 * (size_t)i];
      if (!(lv == lv)) continue;
      const int l = (int)(std::lower_bound(R.labels.begin(), R.labels.end(), lv + 0.0) - R.labels.begin());
      lab_of[i] = l;
      slot_of_ref[i] = per[l]++;
    }
    for (int l = 0; l < n_labels; l++) R.ref_base[l + 1] = R.ref_base[l] + per[l];
    R.reftab.assign(5 * (size_t)std::max<uint32_t>(R.ref_base[n_labels], 1) + 8, 0.0);  // + slack for aligned bulk copies
    for (int i = 0; i < n_ref; i++) {
      if (lab_of[i] < 0) continue;
      const double *r = ref7 + 7 * (size_t)i;
      double *t = R.reftab.data() + 5 * (size_t)(R.ref_base[lab_of[i]] + slot_of_ref[i]);
      t[0] = r[1]; t[1] = r[2]; t[2] = r[4]; t[3] = r[5]; t[4] = r[6];
    }
  }
  // mark every cell whose (slightly dilated) box a landmark's match disc touches; entries are
  // produced in ascending landmark order, which is the order a cell's candidates must keep
  std::vector<RefIndex::Entry> &entries = R.entries;
  entries.clear();
  R.mark_rc = R.mark_rc2 = 0.0;
  entries.reserve((size_t)n_ref * 12);
  if (matchable) {
    const double eps_cells = 3.0 / std::ldexp(1.0, F) + 1e-9 / c;  // fixed-point truncation + lattice drift
    const double rc = rad_m / c + eps_cells;
    const double rc2 = rc * rc * (1.0 + 1e-12);
    R.mark_rc = rc; R.mark_rc2 = rc2;
    for (int i = 0; i < n_ref; i++) {
      const int l = lab_of[i];
      if (l < 0) continue;
      const double *r = ref7 + 7 * (size_t)i;
      const double ux = (r[1] - G.g0x) / c, uy = (r[2] - G.g0y) / c;
      const int x0 = (int)std::floor(ux - rc), x1 = (int)std::floor(ux + rc);
      const int y0 = (int)std::floor(uy - rc), y1 = (int)std::floor(uy + rc);
      if (x0 < 0 || x1 >= G.GX || y0 < 0 || y1 >= G.GY) { err = "internal: match disc leaves the grid"; return SLIDE_PR_ERR_INTERNAL; }
      uint32_t *pl0 = R.bitmap.data() + (size_t)l * G.label_stride;
      uint32_t *pl1 = pl0 + G.plane_words[0];
      // squared distance from the landmark to the cell's box along y, once per column
      double dy2[64];
      const int ny_n = y1 - y0 + 1;
      const bool small = ny_n <= 64;
      for (int ny = y0; small && ny <= y1; ny++) {
        const double dy = std::fmax(std::fmax((double)ny - uy, 0.0), uy - (double)(ny + 1));
        dy2[ny - y0] = dy * dy;
      }
      for (int nx = x0; nx <= x1; nx++) {
        const double dx = std::fmax(std::fmax((double)nx - ux, 0.0), ux - (double)(nx + 1));
        const double dx2 = dx * dx;
        for (int ny = y0; ny <= y1; ny++) {
          double d2;
          if (small) {
            d2 = dx2 + dy2[ny - y0];
          } else {
            const double dy = std::fmax(std::fmax((double)ny - uy, 0.0), uy - (double)(ny + 1));
            d2 = dx2 + dy * dy;
          }
          if (d2 > rc2) continue;
          pl0[(size_t)(nx + 1) * G.W[0] + ((uint32_t)(ny + 32) >> 5)] |= 1u << ((ny + 32) & 31);
          pl1[(size_t)(ny + 1) * G.W[1] + ((uint32_t)(nx + 32) >> 5)] |= 1u << ((nx + 32) & 31);
          cb[l].x0 = std::min(cb[l].x0, nx); cb[l].x1 = std::max(cb[l].x1, nx);
          cb[l].y0 = std::min(cb[l].y0, ny); cb[l].y1 = std::max(cb[l].y1, ny);
          entries.push_back({(uint32_t)i, nx, ny});
        }
      }
    }
  }
  if (entries.size() >= (1ull << 32)) { err = "candidate lists exceed 2^32 entries"; return SLIDE_PR_ERR_UNSUPPORTED; }
  R.labelbox.resize((size_t)std::max(n_labels, 1));
  for (int l = 0; l < std::max(n_labels, 1); l++) {
    if (cb[l].x0 > cb[l].x1) { R.labelbox[l] = SprBox{0, -(1 << 30), 0, -(1 << 30)}; continue; }  // empty: never visible
    R.labelbox[l] = SprBox{cb[l].x0 << F, (cb[l].x1 + 1) << F, cb[l].y0 << F, (cb[l].y1 + 1) << F};
  }
  return SLIDE_PR_OK;
}

// Stage 2 (per bitmap direction): rank tables, per-cell landmark references and the chained
// candidate records -- what the exact verification reads.
int build_ref_ranks(const double *ref7, int d, RefIndex &R, std::string &err) {
  const SprGrid &G = R.grid;
  const int n_labels = (int)R.labels.size();
  const uvec<int32_t> &lab_of = R.lab_of;
  const std::vector<uint32_t> &slot_of_ref = R.slot_of_ref;
  const std::vector<RefIndex::Entry> &entries = R.entries;
  {
    // rank tables of direction d, label-major == rank order of the marked cells:
    //   row_rank[l][row]  = rank of the first marked cell of the row, relative to the label's first cell
    //   rank16[l][word]   = marked cells of the same row before the word
    //   cell_base[l]      = absolute rank of the label's first cell
    R.rank16[d].assign((size_t)G.plane_words[d] * (size_t)std::max(n_labels, 1) + 8, 0);
    R.row_rank[d].assign((size_t)G.R[d] * (size_t)std::max(n_labels, 1) + 1, 0u);
    R.cell_base[d].assign((size_t)n_labels + 1, 0u);
    uint32_t running = 0;
    for (int l = 0; l < n_labels; l++) {
      const uint32_t *pl = R.bitmap.data() + (size_t)l * G.label_stride + (d ? G.plane_words[0] : 0u);
      uint16_t *r16 = R.rank16[d].data() + (size_t)l * G.plane_words[d];
      uint32_t *rr = R.row_rank[d].data() + (size_t)l * G.R[d];
      const uint32_t label_start = running;
      R.cell_base[d][l] = label_start;
      for (int row = 0; row < G.R[d]; row++) {
        rr[row] = running - label_start;
        uint32_t in_row = 0;
        const uint32_t *prow = pl + (size_t)row * G.W[d];
        uint16_t *rrow = r16 + (size_t)row * G.W[d];
        for (int w = 0; w < G.W[d]; w++) {
          rrow[w] = (uint16_t)in_row;
          in_row += (uint32_t)__builtin_popcount(prow[w]);
        }
        if (in_row > 0xffffu) { err = "more than 65535 marked cells in one bitmap row"; return SLIDE_PR_ERR_UNSUPPORTED; }
        running += in_row;
      }
    }
    R.cell_base[d][n_labels] = running;
    const size_t n_cells = running;
    // cand[d][rank] = first candidate of the cell of that rank; extra candidates of a cell are
    // appended behind the n_cells first ones and chained in ascending landmark order.
    // cellref[d][rank] = slot (inside the label's landmark table) of the cell's only candidate,
    // or SPR_CELL_MULTI when the cell has several candidates (then cand[d] is walked).
    uvec<SprCand> &cand = R.cand[d];
    cand.resize(std::max<size_t>(entries.size(), 1));  // every slot below is overwritten
    cand[0] = SprCand{0, 0, 0, 0, 0, 0u, 0u};
    uvec<uint16_t> &cellref = R.cellref[d];
    cellref.assign(n_cells + 16, 0);  // + slack for aligned bulk copies
    std::vector<uint32_t> tail(n_cells, 0xffffffffu);  // last record of each cell's chain
    size_t extra = n_cells;
    for (const RefIndex::Entry &e : entries) {
      const int l = lab_of[e.ref];
      const uint32_t row = (uint32_t)((d ? e.ny : e.nx) + 1), bit = (uint32_t)((d ? e.nx : e.ny) + 32);
      const size_t wi = (size_t)row * G.W[d] + (bit >> 5);
      const uint32_t *pl = R.bitmap.data() + (size_t)l * G.label_stride + (d ? G.plane_words[0] : 0u);
      const uint32_t rank = R.cell_base[d][l] + R.row_rank[d][(size_t)l * G.R[d] + row] +
                            R.rank16[d][(size_t)l * G.plane_words[d] + wi] +
                            (uint32_t)__builtin_popcount(pl[wi] & ((1u << (bit & 31u)) - 1u));
      const double *r = ref7 + 7 * (size_t)e.ref;
      const SprCand c{r[1], r[2], r[4], r[5], r[6], e.ref, 0u};
      if (tail[rank] == 0xffffffffu) {
        const uint32_t slot = slot_of_ref[e.ref];
        cellref[rank] = slot < SPR_CELL_MULTI ? (uint16_t)slot : (uint16_t)SPR_CELL_MULTI;
        cand[rank] = c;
        tail[rank] = rank;
      } else {
        cellref[rank] = (uint16_t)SPR_CELL_MULTI;
        cand[tail[rank]].next = (uint32_t)extra;
        cand[extra] = c;  // c.next == 0
        tail[rank] = (uint32_t)extra++;
      }
    }
    if (extra != std::max<size_t>(entries.size(), n_cells) && !(entries.empty() && extra == 0)) {
      if (extra != entries.size()) { err = "internal: rank / cell list mismatch"; return SLIDE_PR_ERR_INTERNAL; }
    }
  }
  return SLIDE_PR_OK;
}

// ------------------------------------------------------------------------------------------
// query set: drop labels absent from the reference, sort by (label, Morton cell)
// ------------------------------------------------------------------------------------------
static inline uint32_t part1by1(uint32_t x) {
  x &= 0xffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

int build_query_set(const RefIndex &R, const double *qry7, int n_qry, QuerySet &Q, std::string &err) {
  return build_query_set(R.labels, qry7, n_qry, Q, err);
}

// distinct non-NaN labels of a map, ascending (-0.0 and 0.0 are one label: PR.cpp:306 compares with ==).
// Maps have a handful of labels: a small sorted vector with binary search instead of sorting all rows.
void unique_labels(const double *rows7, int n, std::vector<double> &labels) {
  labels.clear();
  double last = 0.0;
  bool have_last = false;
  for (int i = 0; i < n; i++) {
    const double l = rows7[7 * (size_t)i];
    if (!(l == l)) continue;  // NaN labels never compare equal
    if (have_last && l == last) continue;
    last = l; have_last = true;
    auto it = std::lower_bound(labels.begin(), labels.end(), l + 0.0);
    if (it == labels.end() || *it != l + 0.0) labels.insert(it, l + 0.0);
  }
}

int build_query_set(const std::vector<double> &labels, const double *qry7, int n_qry, QuerySet &Q, std::string &err) {
  Q.nq = Q.nqp = 0;  // the (page-locked) vectors keep their capacity across calls
  const int n_labels = (int)labels.size();
  // sort key: label bucket, Morton code, original index -- one 64-bit integer per kept query
  // (20 bits of label bucket, 22 bits of index: slide_pr_prepare limits n_qry to 2^22)
  std::vector<uint64_t> items;
  items.reserve((size_t)n_qry);
  if (n_labels >= (1 << 20)) { err = "more than 2^20 distinct labels"; return SLIDE_PR_ERR_UNSUPPORTED; }
  double minx = HUGE_VAL, maxx = -HUGE_VAL, miny = HUGE_VAL, maxy = -HUGE_VAL;
  for (int j = 0; j < n_qry; j++) {
    const double *q = qry7 + 7 * (size_t)j;
    if (!std::isfinite(q[1]) || !std::isfinite(q[2])) { err = "non-finite query coordinate"; return SLIDE_PR_ERR_NONFINITE; }
    minx = std::min(minx, q[1]); maxx = std::max(maxx, q[1]);
    miny = std::min(miny, q[2]); maxy = std::max(maxy, q[2]);
  }
  const double sx = maxx > minx ? 65535.0 / (maxx - minx) : 0.0, sy = maxy > miny ? 65535.0 / (maxy - miny) : 0.0;
  for (int j = 0; j < n_qry; j++) {
    const double *q = qry7 + 7 * (size_t)j;
    if (!(q[0] == q[0])) continue;
    const double lab = q[0] + 0.0;
    auto it = std::lower_bound(labels.begin(), labels.end(), lab);
    if (it == labels.end() || *it != lab) continue;  // no reference object can ever match it
    // Morton code of the 11 high bits of each 16-bit coordinate (22 bits): groups stay spatially compact
    const uint32_t mx = (uint32_t)((q[1] - minx) * sx) >> 5, my = (uint32_t)((q[2] - miny) * sy) >> 5;
    const uint64_t morton = part1by1(mx) | (part1by1(my) << 1);
    items.push_back(((uint64_t)(it - labels.begin()) << 44) | (morton << 22) | (uint64_t)j);
  }
  // stable LSD radix sort on the (label, Morton) bits: the items are generated in ascending index order
  {
    int key_bits = 22;
    for (int nl = n_labels; nl > 1; nl >>= 1) key_bits++;
    key_bits += 1;
    std::vector<uint64_t> tmp(items.size());
    uint32_t hist[2048];
    for (int shift = 22; shift < 22 + key_bits; shift += 11) {
      std::memset(hist, 0, sizeof(hist));
      for (const uint64_t it : items) hist[(it >> shift) & 2047u]++;
      uint32_t run = 0;
      for (uint32_t &h : hist) { const uint32_t c = h; h = run; run += c; }
      for (const uint64_t it : items) tmp[hist[(it >> shift) & 2047u]++] = it;
      items.swap(tmp);
    }
  }
  Q.nq = (int)items.size();
  // every label segment is padded to a whole number of query groups; padding entries carry
  // qlabel = -1 and get the sentinel fixed-point coordinate (never inside the grid)
  std::vector<int> per_label(n_labels, 0);
  for (const uint64_t it : items) per_label[(size_t)(it >> 44)]++;
  Q.label_gseg.assign(n_labels + 1, 0);
  for (int l = 0; l < n_labels; l++) Q.label_gseg[l + 1] = Q.label_gseg[l] + (per_label[l] + SPR_QGROUP - 1) / SPR_QGROUP;
  Q.nqp = Q.label_gseg[n_labels] * SPR_QGROUP;
  const size_t n = (size_t)std::max(Q.nqp, 1);
  Q.orig.assign(n, -1);
  Q.qlabel.assign(n, -1);
  Q.qxy.assign(2 * n, 0.0);
  Q.qdims.assign(3 * n, 0.0);
  std::vector<int> fill(n_labels, 0);
  for (const uint64_t it : items) {
    const int l = (int)(it >> 44), j = (int)(it & ((1u << 22) - 1u));
    const size_t s = (size_t)Q.label_gseg[l] * SPR_QGROUP + (size_t)fill[l]++;
    const double *q = qry7 + 7 * (size_t)j;
    Q.orig[s] = j;
    Q.qlabel[s] = l;
    Q.qxy[2 * s] = q[1]; Q.qxy[2 * s + 1] = q[2];
    Q.qdims[3 * s] = q[4]; Q.qdims[3 * s + 1] = q[5]; Q.qdims[3 * s + 2] = q[6];
  }
  return SLIDE_PR_OK;
}

// ------------------------------------------------------------------------------------------
// pair-join scorer (spr_join.h)
// ------------------------------------------------------------------------------------------
int build_join_ref(const slide_pr_params &p, const double *ref7, int n_ref, JoinRef &J, std::string &err) {
  J.n_ref = n_ref;
  J.labels.clear();
  const double step = p.match_xy_step_size, thr = p.match_threshold;
  if (!(step > 0) || !std::isfinite(step)) { err = "match_xy_step_size must be positive and finite"; return SLIDE_PR_ERR_INVALID; }
  if (thr > 0 && std::isinf(thr)) { err = "infinite match_threshold is not supported"; return SLIDE_PR_ERR_UNSUPPORTED; }
  J.Tstar = sqrt_threshold(thr);
  J.Sstar = div3_threshold(p.match_threshold_dimension);
  const bool matchable = thr > 0 && std::isfinite(thr);
  J.reach = matchable ? thr * (1.0 + 1e-9) + 1e-9 : 0.0;  // covers the fp64 rounding of the reference's test
  // pass 1: extent of the map and the label bucket of every landmark.  Maps carry a handful of classes: the
  // distinct labels are collected in order of appearance (linear search) and ranked afterwards; maps with
  // many labels take the sorted-vector path.
  double minx = HUGE_VAL, maxx = -HUGE_VAL, miny = HUGE_VAL, maxy = -HUGE_VAL;
  std::vector<int32_t> lab((size_t)std::max(n_ref, 1), -1), cxs((size_t)std::max(n_ref, 1), 0), cys((size_t)std::max(n_ref, 1), 0);
  constexpr int kSeenMax = 16;
  double seen[kSeenMax];
  int n_seen = 0;
  bool few_labels = true;
  for (int i = 0; i < n_ref; i++) {
    const double *r = ref7 + 7 * (size_t)i;
    if (!std::isfinite(r[1]) || !std::isfinite(r[2])) { err = "non-finite reference coordinate"; return SLIDE_PR_ERR_NONFINITE; }
    minx = std::min(minx, r[1]); maxx = std::max(maxx, r[1]);
    miny = std::min(miny, r[2]); maxy = std::max(maxy, r[2]);
    if (!few_labels || !(r[0] == r[0])) continue;  // NaN labels never compare equal (PR.cpp:306)
    int k = 0;
    while (k < n_seen && seen[k] != r[0]) k++;      // -0.0 == 0.0: one label
    if (k == n_seen) {
      if (n_seen == kSeenMax) { few_labels = false; continue; }
      seen[n_seen++] = r[0] + 0.0;
    }
    lab[i] = k;
  }
  if (few_labels) {
    int order[kSeenMax], rank[kSeenMax];
    for (int k = 0; k < n_seen; k++) order[k] = k;
    std::sort(order, order + n_seen, [&](int a, int b) { return seen[a] < seen[b]; });
    for (int k = 0; k < n_seen; k++) { rank[order[k]] = k; J.labels.push_back(seen[order[k]]); }
    for (int i = 0; i < n_ref; i++) if (lab[i] >= 0) lab[i] = rank[lab[i]];
  } else {
    unique_labels(ref7, n_ref, J.labels);
    for (int i = 0; i < n_ref; i++) {
      const double l = ref7[7 * (size_t)i];
      lab[i] = l == l ? (int)(std::lower_bound(J.labels.begin(), J.labels.end(), l + 0.0) - J.labels.begin()) : -1;
    }
  }
  const int n_labels = (int)J.labels.size();
  if (n_ref == 0) { minx = maxx = miny = maxy = 0; }
  // coarse grid: cells of 8 lattice steps (a block is about 10 steps wide), not smaller than the reach,
  // and few enough that the per-label cell tables stay small
  double w = std::max(8.0 * step, J.reach);
  for (;;) {
    const double cx = std::floor((maxx - minx) / w) + 1.0, cy = std::floor((maxy - miny) / w) + 1.0;
    if (cx * cy * std::max(n_labels, 1) <= 16777216.0 && cx < 30000.0 && cy < 30000.0) break;
    w *= 2.0;
    if (!std::isfinite(w)) { err = "reference map extent is not finite"; return SLIDE_PR_ERR_NONFINITE; }
  }
  J.w = w; J.inv_w = 1.0 / w; J.gx0 = minx; J.gy0 = miny;
  auto cell_of = [&](double v, double g0, int n) {  // the kernel evaluates the same expression (monotone in v)
    const double f = std::floor((v - g0) * J.inv_w);
    return (int)std::min(std::max(f, 0.0), (double)(n - 1));
  };
  J.ncx = (int)std::floor((maxx - minx) * J.inv_w) + 1;
  J.ncy = (int)std::floor((maxy - miny) * J.inv_w) + 1;
  const size_t n_cells = (size_t)J.ncx * (size_t)J.ncy;
  // pass 2: cell and label bounding box of every landmark
  J.labelbox.assign(4 * (size_t)std::max(n_labels, 1), 0.0);
  for (int l = 0; l < n_labels; l++) {
    J.labelbox[4 * (size_t)l] = HUGE_VAL; J.labelbox[4 * (size_t)l + 1] = -HUGE_VAL;
    J.labelbox[4 * (size_t)l + 2] = HUGE_VAL; J.labelbox[4 * (size_t)l + 3] = -HUGE_VAL;
  }
  size_t n_kept = 0;
  for (int i = 0; i < n_ref; i++) {
    if (lab[i] < 0) continue;
    const double *r = ref7 + 7 * (size_t)i;
    cxs[i] = cell_of(r[1], J.gx0, J.ncx); cys[i] = cell_of(r[2], J.gy0, J.ncy);
    double *lb = J.labelbox.data() + 4 * (size_t)lab[i];
    lb[0] = std::min(lb[0], r[1]); lb[1] = std::max(lb[1], r[1]);
    lb[2] = std::min(lb[2], r[2]); lb[3] = std::max(lb[3], r[2]);
    n_kept++;
  }
  // counting sort into both join orders (label, band, along cell); ascending reference index inside a cell.
  // Counts go two slots up, so that after the prefix sum cs[key + 1] is the key's insertion cursor and, once
  // every landmark is placed, cs[0 .. n_keys] is the table of first records (cs[n_keys + 1] repeats the total).
  std::vector<uint32_t> pos0, pos1;  // record position of landmark i in each order
  pos0.assign((size_t)std::max(n_ref, 1), 0u); pos1 = pos0;
  for (int d = 0; d < 2; d++) {
    uvec<uint32_t> &cs = J.cell_start[d];
    cs.assign((size_t)std::max(n_labels, 1) * n_cells + 2, 0u);
    std::vector<uint32_t> &pos = d == 0 ? pos0 : pos1;
    for (int i = 0; i < n_ref; i++) {   // pos holds the landmark's key until it is placed
      if (lab[i] < 0) continue;
      const size_t c = d == 0 ? (size_t)cxs[i] * (size_t)J.ncy + (size_t)cys[i] : (size_t)cys[i] * (size_t)J.ncx + (size_t)cxs[i];
      pos[i] = (uint32_t)((size_t)lab[i] * n_cells + c);
      cs[(size_t)pos[i] + 2]++;
    }
    for (size_t k = 1; k < cs.size(); k++) cs[k] += cs[k - 1];
    for (int i = 0; i < n_ref; i++) if (lab[i] >= 0) pos[i] = cs[(size_t)pos[i] + 1]++;
    J.rec[d].assign(std::max<size_t>(n_kept, 1), SprJoinRef{});
    J.xy[d].assign(2 * std::max<size_t>(n_kept, 1), 0.0);
  }
  // same-label landmarks with a lower reference index that can match the same point: within 2 x reach
  J.nbr.clear();
  // (+ the rounding of the reference's test at the magnitude of the map's coordinates)
  const double mag = std::max({std::fabs(minx), std::fabs(maxx), std::fabs(miny), std::fabs(maxy)});
  const double r2 = 2.0 * J.reach * (1.0 + 1e-9) + 1e-9 + 64.0 * DBL_EPSILON * mag;
  J.max_abs = mag;
  const int span = (int)std::floor(r2 * J.inv_w) + 1;
  const uvec<uint32_t> &cs0 = J.cell_start[0];
  std::vector<int32_t> by_pos0(std::max<size_t>(n_kept, 1), -1);  // landmark at each record position of order 0
  for (int i = 0; i < n_ref; i++) if (lab[i] >= 0) by_pos0[pos0[i]] = i;
  for (size_t at = 0; at < n_kept; at++) {   // in record order: the cell table is read front to back
    const int i = by_pos0[at];
    const double *r = ref7 + 7 * (size_t)i;
    const uint32_t off = (uint32_t)J.nbr.size();
    if (matchable) {
      for (int cx = std::max(cxs[i] - span, 0); cx <= std::min(cxs[i] + span, J.ncx - 1); cx++) {
        const int cy0 = std::max(cys[i] - span, 0), cy1 = std::min(cys[i] + span, J.ncy - 1);
        const size_t base = (size_t)lab[i] * n_cells + (size_t)cx * (size_t)J.ncy;
        for (uint32_t k = cs0[base + cy0]; k < cs0[base + cy1 + 1]; k++) {
          const int o = by_pos0[k];
          if (o >= i) continue;  // only landmarks the reference meets earlier (PR.cpp:299: input order)
          const double *q = ref7 + 7 * (size_t)o;
          if (std::fabs(q[1] - r[1]) <= r2 && std::fabs(q[2] - r[2]) <= r2) J.nbr.push_back(SprJoinNbr{q[1], q[2], q[4], q[5], q[6]});
        }
      }
    }
    const uint32_t cnt = (uint32_t)J.nbr.size() - off;
    const SprJoinRef rec{r[1], r[2], r[4], r[5], r[6], off, cnt};
    J.rec[0][at] = rec;
    J.rec[1][pos1[i]] = rec;
    J.xy[0][2 * at] = r[1]; J.xy[0][2 * at + 1] = r[2];
    J.xy[1][2 * (size_t)pos1[i]] = r[1]; J.xy[1][2 * (size_t)pos1[i] + 1] = r[2];
  }
  if (J.nbr.empty()) J.nbr.push_back(SprJoinNbr{});
  return SLIDE_PR_OK;
}

int build_join_blocks(const Lattice &L, double step, uvec<SprJoinBlock> &blocks, double *drift, std::string &err) {
  blocks.clear();
  // a rectangle ix in [ix0, ix1), iy in [iy0, iy1) of ring k; ordinal(ix, iy) = ord0 + (ix - ix0) * row_stride + (iy - iy0)
  auto rect = [&](const Lattice::Ring &R, uint32_t k, int ix0, int ix1, int iy0, int iy1, uint64_t ord0, uint64_t row_stride) {
    const int w = ix1 - ix0, h = iy1 - iy0;
    if (w <= 0 || h <= 0) return;
    // tiles of bw x bh samples: the short side in pieces of at most 32, the long side as far as the
    // counters of a block reach
    int bw, bh;
    auto ok = [](int nx, int ny) {   // the kernel's limits: slot index of the arg-max, the two counter arrays
      return nx * ny <= SPJ_MAX_SLOTS && 2 * nx * ((ny >> 1) + 1) <= SPJ_TILE_WORDS;
    };
    if (w <= h) {
      const int pieces = (w + 31) / 32;
      bw = (w + pieces - 1) / pieces;
      bh = std::min(h, SPJ_MAX_SLOTS / bw);
      while (bh > 1 && !ok(bw, bh)) bh--;
    } else {
      const int pieces = (h + 31) / 32;
      bh = (h + pieces - 1) / pieces;
      bw = std::min(w, SPJ_MAX_SLOTS / bh);
      while (bw > 1 && !ok(bw, bh)) bw--;
    }
    for (int bx = ix0; bx < ix1; bx += bw)
      for (int by = iy0; by < iy1; by += bh) {
        SprJoinBlock B;
        B.nx = (uint32_t)std::min(bw, ix1 - bx); B.ny = (uint32_t)std::min(bh, iy1 - by);
        B.xi = R.x_off + (uint32_t)bx; B.yi = R.y_off + (uint32_t)by;
        B.ord0 = (uint32_t)(ord0 + (uint64_t)(bx - ix0) * row_stride + (uint64_t)(by - iy0));
        B.row_stride = (uint32_t)row_stride;
        B.dir = B.ny >= B.nx ? 0u : 1u;
        B.ring = k;
        blocks.push_back(B);
      }
  };
  for (size_t k = 0; k < L.ring.size(); k++) {
    const Lattice::Ring &R = L.ring[k];
    const int nx = (int)R.nx, ny = (int)R.ny;
    const int n_in_x = R.ixh >= R.ixl ? R.ixh - R.ixl + 1 : 0, n_in_y = R.iyh >= R.iyl ? R.iyh - R.iyl + 1 : 0;
    const uint64_t ord = R.ord_base;
    if (!(n_in_x > 0 && n_in_y > 0)) {
      rect(R, (uint32_t)k, 0, nx, 0, ny, ord, (uint64_t)ny);
    } else {  // the four rectangles around the already-searched box, in the reference's ordinal layout (build_lattice)
      const uint64_t cnt_in = (uint64_t)(ny - n_in_y);
      const uint64_t base_b = ord + (uint64_t)R.ixl * ny;
      const uint64_t base_c = base_b + (uint64_t)n_in_x * cnt_in;
      rect(R, (uint32_t)k, 0, R.ixl, 0, ny, ord, (uint64_t)ny);
      rect(R, (uint32_t)k, R.ixl, R.ixh + 1, 0, R.iyl, base_b, cnt_in);
      rect(R, (uint32_t)k, R.ixl, R.ixh + 1, R.iyh + 1, ny, base_b + (uint64_t)R.iyl, cnt_in);
      rect(R, (uint32_t)k, R.ixh + 1, nx, 0, ny, base_c, (uint64_t)ny);
    }
    if (blocks.size() > (1u << 24)) { err = "more than 2^24 lattice blocks"; return SLIDE_PR_ERR_UNSUPPORTED; }
  }
  if (drift) *drift = L.drift;
  return SLIDE_PR_OK;
}

// ------------------------------------------------------------------------------------------
// closed-form refinement (PlaceRecognition::solveLSQ, PR.cpp:632-695)
// ------------------------------------------------------------------------------------------
namespace {
struct Givens { double c, s; };  // [[c, s], [-s, c]]
inline Givens compose(Givens a, Givens b) { return {a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c}; }
inline Givens inverse(Givens a) { return {a.c, -a.s}; }
inline void left(double *M, int n, int p, int q, Givens g) {
  for (int i = 0; i < n; i++) {
    const double x = M[p * n + i], y = M[q * n + i];
    M[p * n + i] = g.c * x + g.s * y;
    M[q * n + i] = g.c * y - g.s * x;
  }
}
inline void right(double *M, int n, int p, int q, Givens g) {
  for (int i = 0; i < n; i++) {
    const double x = M[i * n + p], y = M[i * n + q];
    M[i * n + p] = g.c * x - g.s * y;
    M[i * n + q] = g.s * x + g.c * y;
  }
}
// rotation diagonalising the symmetric 2x2 [[x, y], [y, z]]
inline Givens symmetric_schur(double x, double y, double z) {
  if (2.0 * std::fabs(y) < DBL_MIN) return {1.0, 0.0};
  const double tau = (x - z) / (2.0 * std::fabs(y));
  const double w = std::sqrt(tau * tau + 1.0);
  const double t = tau > 0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double n = 1.0 / std::sqrt(t * t + 1.0);
  const double sgn = t > 0 ? 1.0 : -1.0;
  return {n, -sgn * (y / std::fabs(y)) * std::fabs(t) * n};
}
}  // namespace

// Two-sided (Kogbetliantz) Jacobi SVD, n <= 3, row-major: A = U diag(S) V^T, S descending.
// Zero off-diagonal blocks are never rotated, so a planar (z == 0) cross-covariance keeps
// U(2,2) = V(2,2) = 1 -- the property the reference's reflection test (PR.cpp:680) sees with
// Eigen::JacobiSVD.
void svd_jacobi(const double *A, int n, double *U, double *S, double *V) {
  double M[9];
  double scale = 0;
  for (int i = 0; i < n * n; i++) scale = std::max(scale, std::fabs(A[i]));
  if (scale == 0) scale = 1;
  for (int i = 0; i < n * n; i++) { M[i] = A[i] / scale; U[i] = 0; V[i] = 0; }
  for (int i = 0; i < n; i++) U[i * n + i] = V[i * n + i] = 1;
  double big = 0;
  for (int i = 0; i < n; i++) big = std::max(big, std::fabs(M[i * n + i]));
  bool again = true;
  for (int sweep = 0; sweep < 200 && again; sweep++) {
    again = false;
    for (int p = 1; p < n; p++)
      for (int q = 0; q < p; q++) {
        const double tol = std::max(DBL_MIN, 2.0 * DBL_EPSILON * big);
        if (!(std::fabs(M[p * n + q]) > tol || std::fabs(M[q * n + p]) > tol)) continue;
        again = true;
        const double m00 = M[p * n + p], m01 = M[p * n + q], m10 = M[q * n + p], m11 = M[q * n + q];
        Givens sym{1.0, 0.0};  // first make the 2x2 block symmetric
        const double d = m10 - m01;
        if (std::fabs(d) >= DBL_MIN) {
          const double u = (m00 + m11) / d, h = std::sqrt(1.0 + u * u);
          sym = {u / h, 1.0 / h};
        }
        const double s00 = sym.c * m00 + sym.s * m10, s01 = sym.c * m01 + sym.s * m11;
        const double s11 = sym.c * m11 - sym.s * m01;
        const Givens jr = symmetric_schur(s00, s01, s11);
        const Givens jl = compose(sym, inverse(jr));
        left(M, n, p, q, jl);
        right(U, n, p, q, inverse(jl));
        right(M, n, p, q, jr);
        right(V, n, p, q, jr);
        big = std::max(big, std::max(std::fabs(M[p * n + p]), std::fabs(M[q * n + q])));
      }
  }
  for (int i = 0; i < n; i++) {
    const double a = M[i * n + i];
    S[i] = std::fabs(a) * scale;
    if (a < 0) for (int r = 0; r < n; r++) U[r * n + i] = -U[r * n + i];
  }
  for (int i = 0; i < n; i++) {
    int pos = i;
    for (int k = i + 1; k < n; k++) if (S[k] > S[pos]) pos = k;
    if (S[pos] == 0) break;
    if (pos == i) continue;
    std::swap(S[i], S[pos]);
    for (int r = 0; r < n; r++) { std::swap(U[r * n + i], U[r * n + pos]); std::swap(V[r * n + i], V[r * n + pos]); }
  }
}

void xyz_yaw_from_tf(const double *tf16, double *xyz_yaw4) {  // PR.cpp:697-711
  xyz_yaw4[0] = tf16[3]; xyz_yaw4[1] = tf16[7]; xyz_yaw4[2] = tf16[11];
  xyz_yaw4[3] = std::atan2(tf16[4], tf16[0]);
}

static void mul_abt3(const double *A, const double *B, double *C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}

void solve_lsq(const double *tgt3, const double *src3, int k, double *xyz_yaw4, double *tf16) {
  double cs[3] = {0, 0, 0}, ct[3] = {0, 0, 0};
  for (int i = 0; i < k; i++)
    for (int d = 0; d < 3; d++) { cs[d] += src3[3 * i + d]; ct[d] += tgt3[3 * i + d]; }
  for (int d = 0; d < 3; d++) { cs[d] /= (double)k; ct[d] /= (double)k; }      // PR.cpp:655-658
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < k; i++)                                                   // PR.cpp:665-671
    for (int a = 0; a < 3; a++) {
      const double sa = src3[3 * i + a] - cs[a];
      for (int b = 0; b < 3; b++) H[a * 3 + b] += sa * (tgt3[3 * i + b] - ct[b]);
    }
  double U[9], S[3], V[9], Rm[9];
  svd_jacobi(H, 3, U, S, V);                                                    // PR.cpp:674
  mul_abt3(V, U, Rm);                                                           // PR.cpp:678
  const double det = Rm[0] * (Rm[4] * Rm[8] - Rm[5] * Rm[7]) - Rm[1] * (Rm[3] * Rm[8] - Rm[5] * Rm[6]) +
                     Rm[2] * (Rm[3] * Rm[7] - Rm[4] * Rm[6]);
  if (det < 0) {                                                                // PR.cpp:680-686
    double U2[9], S2[3], V2[9];
    svd_jacobi(Rm, 3, U2, S2, V2);
    for (int r = 0; r < 3; r++) V2[r * 3 + 2] = -V2[r * 3 + 2];
    mul_abt3(V2, U2, Rm);
  }
  for (int i = 0; i < 16; i++) tf16[i] = 0;
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) tf16[a * 4 + b] = Rm[a * 3 + b];
    tf16[a * 4 + 3] = ct[a] - (Rm[a * 3] * cs[0] + Rm[a * 3 + 1] * cs[1] + Rm[a * 3 + 2] * cs[2]);  // PR.cpp:689
  }
  tf16[15] = 1;
  xyz_yaw_from_tf(tf16, xyz_yaw4);
}

void estimate_tf(const double *a, const double *b, int k, double *tf9) {  // SC.cpp:122-138
  double ca[2] = {0, 0}, cb[2] = {0, 0};
  for (int i = 0; i < k; i++) { ca[0] += a[2 * i]; ca[1] += a[2 * i + 1]; cb[0] += b[2 * i]; cb[1] += b[2 * i + 1]; }
  ca[0] /= (double)k; ca[1] /= (double)k; cb[0] /= (double)k; cb[1] /= (double)k;
  double H[4] = {0, 0, 0, 0};
  for (int i = 0; i < k; i++) {
    const double ax = a[2 * i] - ca[0], ay = a[2 * i + 1] - ca[1];
    const double bx = b[2 * i] - cb[0], by = b[2 * i + 1] - cb[1];
    H[0] += ax * bx; H[1] += ax * by; H[2] += ay * bx; H[3] += ay * by;
  }
  double U[4], S[2], V[4];
  svd_jacobi(H, 2, U, S, V);
  double R[4] = {V[0] * U[0] + V[1] * U[1], V[0] * U[2] + V[1] * U[3], V[2] * U[0] + V[3] * U[1], V[2] * U[2] + V[3] * U[3]};  // V U^T
  if (R[0] * R[3] - R[1] * R[2] < 0) { R[1] = -R[1]; R[3] = -R[3]; }  // R.col(1) *= -1
  tf9[0] = R[0]; tf9[1] = R[1]; tf9[2] = cb[0] - (R[0] * ca[0] + R[1] * ca[1]);
  tf9[3] = R[2]; tf9[4] = R[3]; tf9[5] = cb[1] - (R[2] * ca[0] + R[3] * ca[1]);
  tf9[6] = 0; tf9[7] = 0; tf9[8] = 1;
}

void mat4_mul(const double *A, const double *B, double *C) {
  double T[16];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double acc = 0;
      for (int k = 0; k < 4; k++) acc += A[i * 4 + k] * B[k * 4 + j];
      T[i * 4 + j] = acc;
    }
  std::memcpy(C, T, sizeof(T));
}

bool mat4_rigid_inverse(const double *A, double *Ainv) {
  // [R t; 0 1]^-1 = [R^T  -R^T t; 0 1]
  double T[16] = {0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[i * 4 + j] = A[j * 4 + i];
  for (int i = 0; i < 3; i++) T[i * 4 + 3] = -(T[i * 4] * A[3] + T[i * 4 + 1] * A[7] + T[i * 4 + 2] * A[11]);
  T[15] = 1;
  std::memcpy(Ainv, T, sizeof(T));
  return true;
}

}  // namespace spr
