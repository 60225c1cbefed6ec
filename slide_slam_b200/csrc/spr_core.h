// spr_core.h -- the per-thread scoring primitives (host + device).
//
// The same inline functions are compiled into the CUDA kernels (spr_kernels.cu) and into the
// test-only single-thread emulation (tests/emu), which lets the index structures and the
// decision arithmetic be checked against the CPU oracle without a GPU.  The emulation is a
// test harness, not a fallback: the product library never executes these on the host.
//
// Decision arithmetic follows place_recognition.cpp:246-357 in fp64 with every multiply and add
// rounded separately (the reference is built without FMA): __dmul_rn/__dadd_rn on the device,
// plain operators under -ffp-contract=off on the host.
#pragma once
#include <math.h>

#include "spr_types.h"

#if defined(__CUDA_ARCH__)
#define SPR_DADD(a, b) __dadd_rn((a), (b))
#define SPR_DSUB(a, b) __dsub_rn((a), (b))
#define SPR_DMUL(a, b) __dmul_rn((a), (b))
#define SPR_FUNNEL_R(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#define SPR_POPC(x) __popc(x)
#define SPR_FFS(x) __ffs(x)
#define SPR_UMIN(a, b) min((uint32_t)(a), (uint32_t)(b))
#else
#define SPR_DADD(a, b) ((a) + (b))
#define SPR_DSUB(a, b) ((a) - (b))
#define SPR_DMUL(a, b) ((a) * (b))
static inline uint32_t spr_funnel_r_host(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31u;
  return s ? ((lo >> s) | (hi << (32u - s))) : lo;
}
#define SPR_FUNNEL_R(lo, hi, s) spr_funnel_r_host((lo), (hi), (s))
#define SPR_POPC(x) __builtin_popcount(x)
#define SPR_FFS(x) __builtin_ffs((int)(x))
#define SPR_UMIN(a, b) ((uint32_t)(a) < (uint32_t)(b) ? (uint32_t)(a) : (uint32_t)(b))
#endif

// metres -> fixed-point cell units, rounded down.  The host guarantees that for every
// translation t and rotated query q:  |fx(t)| + |fx(q - g0)| < 2^30.
SPR_HD int32_t spr_fx(double v, double S) { return (int32_t)floor(SPR_DMUL(v, S)); }

// Rotated query coordinates, PR.cpp:257-258 (first two terms, left to right):
//   rx = c*qx + (-s)*qy ; ry = s*qx + c*qy
SPR_HD void spr_rotate(double c, double s, double qx, double qy, double *rx, double *ry) {
  const double ms = -s;
  *rx = SPR_DADD(SPR_DMUL(c, qx), SPR_DMUL(ms, qy));
  *ry = SPR_DADD(SPR_DMUL(s, qx), SPR_DMUL(c, qy));
}

// One query group (SPR_QGROUP sorted queries) under yaw a: exact rotated coordinates, their
// fixed-point cell coordinates in both layouts, and the group's fixed-point bounding box.
// Padding queries (qlabel < 0) get the sentinel coordinate and do not widen the box.
SPR_HD void spr_rotate_group(const double *cs, const double *qxy, const int32_t *qlabel, const SprGrid &G,
                             int32_t nqp, int32_t n_groups, int32_t a, int32_t g, int32_t *qrotq_xy,
                             int32_t *qrotq_yx, double *qrot, SprBox *gbox) {
  const double c = cs[2 * a], s = cs[2 * a + 1];
  SprBox box = {1 << 30, -(1 << 30), 1 << 30, -(1 << 30)};
  for (int k = 0; k < SPR_QGROUP; k++) {
    const int32_t js = g * SPR_QGROUP + k;
    const size_t qi = (size_t)a * (size_t)nqp + (size_t)js;
    int32_t fxq = SPR_Q_SENTINEL, fyq = SPR_Q_SENTINEL;
    double rx = 0.0, ry = 0.0;
    if (qlabel[js] >= 0) {
      spr_rotate(c, s, qxy[2 * (size_t)js], qxy[2 * (size_t)js + 1], &rx, &ry);
      fxq = spr_fx(SPR_DSUB(rx, G.g0x), G.S);
      fyq = spr_fx(SPR_DSUB(ry, G.g0y), G.S);
      box.x0 = fxq < box.x0 ? fxq : box.x0;
      box.x1 = fxq + 1 > box.x1 ? fxq + 1 : box.x1;
      box.y0 = fyq < box.y0 ? fyq : box.y0;
      box.y1 = fyq + 1 > box.y1 ? fyq + 1 : box.y1;
    }
    qrot[2 * qi] = rx; qrot[2 * qi + 1] = ry;
    qrotq_xy[2 * qi] = fxq; qrotq_xy[2 * qi + 1] = fyq;
    qrotq_yx[2 * qi] = fyq; qrotq_yx[2 * qi + 1] = fxq;
  }
  gbox[(size_t)a * (size_t)n_groups + (size_t)g] = box;
}

// PR.cpp:310-313,332-333 for one (query, reference) pair under translation (tx, ty).
SPR_HD bool spr_distance_match(double rx, double ry, double tx, double ty, double refx,
                               double refy, double Tstar) {
  const double xt = SPR_DADD(rx, tx);  // third term of PR.cpp:257-258; the "/ 1.0" is exact
  const double yt = SPR_DADD(ry, ty);
  const double dx = SPR_DSUB(refx, xt);
  const double dy = SPR_DSUB(refy, yt);
  const double d2 = SPR_DADD(SPR_DMUL(dx, dx), SPR_DMUL(dy, dy));
  return d2 < Tstar;  // <=> sqrt(d2) < match_threshold_ (sqrt is monotone, correctly rounded)
}

// PR.cpp:315-339: dimension rule, keyed on the REFERENCE object's d2 == 0 && d3 == 0.
SPR_HD bool spr_dimension_match(double r1, double r2, double r3, const double *qd, double thr_dim,
                                double Sstar) {
  const double a0 = fabs(SPR_DSUB(r1, qd[0]));
  if (r2 == 0 && r3 == 0) return a0 < thr_dim;
  const double a1 = fabs(SPR_DSUB(r2, qd[1]));
  const double a2 = fabs(SPR_DSUB(r3, qd[2]));
  const double sum = SPR_DADD(SPR_DADD(a0, a1), a2);  // 0 + a0 is exact
  return sum < Sstar;                                  // <=> sum / 3 < thr_dim
}

// Biased fixed-point origin of a chunk: adding a query's fixed coordinates and shifting by F
// gives directly the bitmap row (across + 1 zero row) and bit offset (along + 32 pad bits).
SPR_HD int32_t spr_bias_across(int32_t aq0, int32_t F) { return aq0 + (1 << F); }
SPR_HD int32_t spr_bias_along(int32_t bq0, int32_t F) { return bq0 + (32 << F); }

// Bit-parallel occupancy probe: 32 consecutive lattice samples along the chunk's axis against
// the bitmap row that the across coordinate selects.  a / b are the biased fixed-point sums
// (chunk + rotated query).  Out-of-grid sums clamp onto all-zero rows / words (unsigned min:
// negative values wrap to the upper clamp).
SPR_HD uint32_t spr_probe(const uint32_t *plane, uint32_t W, uint32_t Rm1, uint32_t maxbit, int32_t F,
                          int32_t a, int32_t b, uint32_t valid) {
  const uint32_t row = SPR_UMIN((uint32_t)(a >> F), Rm1);    // rows 0 and R-1 are all-zero
  const uint32_t bit = SPR_UMIN((uint32_t)(b >> F), maxbit);  // word 0 and words >= maxbit/32 are all-zero
  const uint32_t *p = plane + (row * W + (bit >> 5));
  return SPR_FUNNEL_R(p[0], p[1], bit) & valid;  // the funnel shift uses the low 5 bits of `bit`
}

// Row and bit offset (inside the chunk's own plane) of the cell under bit 0 of a probe that hit
// (no clamping: a hit implies the sums are inside the grid).  Bit b of the chunk is `bit + b`.
SPR_HD void spr_cell_of(int32_t F, int32_t a, int32_t b, uint32_t *row, uint32_t *bit) {
  *row = (uint32_t)(a >> F);
  *bit = (uint32_t)(b >> F);
}

// rank tables of plane (l, d) in global memory
SPR_HD SprTables spr_global_tables(const SprView &V, uint32_t d, int32_t l) {
  const SprGrid &G = V.grid;
  SprTables t;
  t.bits = V.bitmap + ((size_t)l * G.label_stride + (d ? G.plane_words[0] : 0u));
  t.r16 = V.rank16[d] + (size_t)l * G.plane_words[d];
  t.row_rank = V.row_rank[d] + (size_t)l * (size_t)G.R[d];
  t.cell_base = V.cell_base[d][l];
  t.cellref = V.cellref[d] + t.cell_base;
  t.reftab = V.reftab + 5 * (size_t)V.ref_base[l];
  t.W = (uint32_t)G.W[d];
  return t;
}

// May query group `g` (box of its fixed coords) land on label box `lb` for any translation of a
// patch whose unbiased fixed-point chunk origins span [X0, X1] x [Y0, Y1]?  (X1 / Y1 include the
// 32-sample extent of the chunks.)
SPR_HD bool spr_group_visible(const SprBox &g, const SprBox &lb, int32_t X0, int32_t X1, int32_t Y0, int32_t Y1) {
  return (g.x1 > lb.x0 - X1) && (g.x0 < lb.x1 - X0) && (g.y1 > lb.y0 - Y1) && (g.y0 < lb.y1 - Y0);
}

// Exact test of one marked cell, given its label-relative rank in the plane described by T
// (direction d): the reference's predicate (PR.cpp:299-355) on the cell's candidate landmarks in
// ascending reference order.  The common single-candidate cell is a 16-bit slot into the label's
// compact landmark table; cells with several candidates walk the chained records of cand[d].
SPR_HD bool spr_verify_rank(const SprView &V, const SprTables &T, uint32_t d, uint32_t rank, double rx, double ry,
                            double tx, double ty, const double *qd) {
  const uint32_t slot = T.cellref[rank];
  if (slot != SPR_CELL_MULTI) {
    const double *r = T.reftab + 5 * (size_t)slot;
    return spr_distance_match(rx, ry, tx, ty, r[0], r[1], V.Tstar) &&
           (V.ignore_dim || spr_dimension_match(r[2], r[3], r[4], qd, V.thr_dim, V.Sstar));
  }
  uint32_t k = T.cell_base + rank;
  for (;;) {
    const SprCand &c = V.cand[d][k];
    if (spr_distance_match(rx, ry, tx, ty, c.x, c.y, V.Tstar) &&
        (V.ignore_dim || spr_dimension_match(c.d1, c.d2, c.d3, qd, V.thr_dim, V.Sstar)))
      return true;
    if (c.next == 0u) return false;
    k = c.next;
  }
}

// Exact verification of one occupied cell (row, bit) of the plane described by T (direction d).
SPR_HD bool spr_verify_cell(const SprView &V, const SprTables &T, uint32_t d, uint32_t row, uint32_t bit, double rx,
                            double ry, double tx, double ty, const double *qd) {
  const uint32_t wi = row * T.W + (bit >> 5);
  const uint32_t rank = T.row_rank[row] + T.r16[wi] + (uint32_t)SPR_POPC(T.bits[wi] & ((1u << (bit & 31u)) - 1u));
  return spr_verify_rank(V, T, d, rank, rx, ry, tx, ty, qd);
}

// Exact verification of all filter hits H of ONE query landmark against the 32 consecutive cells
// starting at (row, bit) of the plane described by T (= the 32 translations of a chunk of
// direction d).  The translation of bit b is (across, along[b]) (dir 0) or (along[b], across)
// (dir 1), exactly the reference's accumulated lattice values.  Returns the mask of hypotheses
// for which the landmark is an inlier.
SPR_HD uint32_t spr_verify_mask(const SprView &V, const SprTables &T, uint32_t d, uint32_t row, uint32_t bit, uint32_t H,
                                double rx, double ry, double across, const double *along, const double *qd) {
  const uint32_t wi = row * T.W + (bit >> 5);
  const uint32_t w0 = T.bits[wi], w1 = T.bits[wi + 1];
  const uint32_t off = bit & 31u;
  // rank of the cell under chunk bit b = rank0 + marked cells among the window's bits below b
  const uint32_t rank0 = T.row_rank[row] + T.r16[wi] + (uint32_t)SPR_POPC(w0 & ((1u << off) - 1u));
  const uint32_t win = SPR_FUNNEL_R(w0, w1, off);  // the 32 cells of the chunk (before the valid mask)
  // the across coordinate is the same for every bit of the chunk: x' (dir 0) or y' (dir 1) of
  // PR.cpp:257-258 is computed once; only the along coordinate changes from bit to bit.
  // dx*dx + dy*dy is evaluated as pa2 + pb2 in either direction (fp64 addition commutes).
  const double pa = SPR_DADD(d ? ry : rx, across);
  const double rb = d ? rx : ry;
  const double *ref_a = T.reftab + (d ? 1 : 0), *ref_b = T.reftab + (d ? 0 : 1);
  uint32_t P = 0u;
  while (H) {
    const int b = SPR_FFS(H) - 1;
    H &= H - 1;
    const uint32_t rank = rank0 + (uint32_t)SPR_POPC(win & ((1u << b) - 1u));
    const uint32_t slot = T.cellref[rank];
    const double t = along[b];
    bool ok;
    if (slot != SPR_CELL_MULTI) {
      const uint32_t o = 5u * slot;
      const double da = SPR_DSUB(ref_a[o], pa);
      const double db = SPR_DSUB(ref_b[o], SPR_DADD(rb, t));
      ok = SPR_DADD(SPR_DMUL(da, da), SPR_DMUL(db, db)) < V.Tstar;  // PR.cpp:332-333
      if (ok && !V.ignore_dim) {
        const double *r = T.reftab + o;
        ok = spr_dimension_match(r[2], r[3], r[4], qd, V.thr_dim, V.Sstar);  // PR.cpp:334-339
      }
    } else {  // several candidate landmarks: chained records in ascending reference order
      ok = spr_verify_rank(V, T, d, rank, rx, ry, d ? t : across, d ? across : t, qd);
    }
    if (ok) P |= 1u << b;
  }
  return P;
}

// Occupancy test of a single point (general hypothesis lists): fixed-point cell of
// (xt - g0x, yt - g0y); returns false when the cell is outside the grid or unmarked.
SPR_HD bool spr_point_cell(const SprView &V, int32_t l, double xt, double yt, uint32_t *row, uint32_t *bitpos) {
  const SprGrid &G = V.grid;
  const double ux = SPR_DMUL(SPR_DSUB(xt, G.g0x), G.S), uy = SPR_DMUL(SPR_DSUB(yt, G.g0y), G.S);
  if (!(ux >= 0.0 && uy >= 0.0 && ux < 1073741824.0 && uy < 1073741824.0)) return false;
  const int32_t cx = (int32_t)ux >> G.F, cy = (int32_t)uy >> G.F;
  if (cx >= G.GX || cy >= G.GY) return false;
  const uint32_t bit = (uint32_t)(cy + 32);
  const uint32_t wloc = (uint32_t)(cx + 1) * (uint32_t)G.W[0] + (bit >> 5);
  const uint32_t word = V.bitmap[(size_t)l * G.label_stride + wloc];
  *row = (uint32_t)(cx + 1);
  *bitpos = bit;
  return (word >> (bit & 31u)) & 1u;
}

// ---------------------------------------------------------------------------------------------
// SlideGraph descriptor half (semantic_clipper.cpp:41-108)
// ---------------------------------------------------------------------------------------------
// compute_triangle_diff's descriptor: vertex-to-centroid distances sorted ascending, plus the
// argsort permutation (SC.cpp:66-90).  tri = [x0,y0,x1,y1,x2,y2].  Arithmetic is fp64,
// left-to-right, not fused (the contract of oracle/slide_oracle.c).
SPR_HD void spr_triangle_descriptor(const double *t, double *desc3, int32_t *perm3) {
  const double cx = SPR_DADD(SPR_DADD(t[0], t[2]), t[4]) / 3.0;
  const double cy = SPR_DADD(SPR_DADD(t[1], t[3]), t[5]) / 3.0;
  double d[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double dx = SPR_DSUB(t[2 * i], cx), dy = SPR_DSUB(t[2 * i + 1], cy);
    d[i] = sqrt(SPR_DADD(SPR_DMUL(dx, dx), SPR_DMUL(dy, dy)));
  }
  // stable 3-element argsort (std::sort on 3 elements is an insertion sort, SC.cpp:41-46)
  int i0 = 0, i1 = 1, i2 = 2;
  if (d[i1] < d[i0]) { const int x = i0; i0 = i1; i1 = x; }
  if (d[i2] < d[i1]) {
    const int x = i1; i1 = i2; i2 = x;
    if (d[i1] < d[i0]) { const int y = i0; i0 = i1; i1 = y; }
  }
  desc3[0] = d[i0]; desc3[1] = d[i1]; desc3[2] = d[i2];
  perm3[0] = i0; perm3[1] = i1; perm3[2] = i2;
}

// SC.cpp:92-99: sqrt(sum (dm - dd)^2) < threshold
SPR_HD bool spr_descriptor_match(const double *dm, const double *dd, double threshold) {
  const double e0 = SPR_DSUB(dm[0], dd[0]), e1 = SPR_DSUB(dm[1], dd[1]), e2 = SPR_DSUB(dm[2], dd[2]);
  const double s = SPR_DADD(SPR_DADD(SPR_DMUL(e0, e0), SPR_DMUL(e1, e1)), SPR_DMUL(e2, e2));
  return sqrt(s) < threshold;
}
