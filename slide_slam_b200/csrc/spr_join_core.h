// spr_join_core.h -- per-thread code of the pair-join scorer (host + device), see spr_join.h.
//
// Compiled into spr_join.cu and into the test-only single-thread emulation (tests/emu), which runs the
// same functions over the host-built structures and compares every counter with the CPU oracle on a box
// without a GPU.  The emulation is a test harness: the product never executes these on the host.
#pragma once
#include <math.h>

#include "spr_core.h"
#include "spr_join_types.h"

#if defined(__CUDA_ARCH__)
#define SPJ_LD(p) __ldg(p)
#define SPJ_ADD(p, v) atomicAdd((p), (v))
#else
#define SPJ_LD(p) (*(p))
#define SPJ_ADD(p, v) (*(p) += (v))
#endif

struct SpjBlock {
  const double *xs, *ys;   // the block's lattice samples
  int nx, ny, nwy;         // nwy: words per row of a counter array ((ny >> 1) + 1)
  int stride;              // words per counter array (nx * nwy)
  double X0, X1, Y0, Y1;   // first / last sample per axis
  double tx0, tx1, ty0, ty1;   // translations of the block +- reach
  uint32_t dir;
};

SPR_HD SpjBlock spj_block(const SprJoinView &V, const SprJoinBlock &blk) {
  SpjBlock B;
  B.xs = V.lat + blk.xi; B.ys = V.lat + blk.yi;
  B.nx = (int)blk.nx; B.ny = (int)blk.ny; B.nwy = (B.ny >> 1) + 1;
  B.stride = B.nx * B.nwy;
  B.X0 = SPJ_LD(B.xs); B.X1 = SPJ_LD(B.xs + B.nx - 1);
  B.Y0 = SPJ_LD(B.ys); B.Y1 = SPJ_LD(B.ys + B.ny - 1);
  B.tx0 = SPR_DSUB(B.X0, V.reach); B.tx1 = SPR_DADD(B.X1, V.reach);
  B.ty0 = SPR_DSUB(B.Y0, V.reach); B.ty1 = SPR_DADD(B.Y1, V.reach);
  B.dir = blk.dir;
  return B;
}

// Can some translation of block B bring a landmark of query group `box` within reach of a reference
// landmark of its label (bounding box lb: x0, x1, y0, y1)?
SPR_HD bool spj_visible(const SpjBlock &B, const SprJoinBox &box, const double *lb) {
  return (double)box.x1 + B.tx1 >= SPJ_LD(lb) && (double)box.x0 + B.tx0 <= SPJ_LD(lb + 1) &&
         (double)box.y1 + B.ty1 >= SPJ_LD(lb + 2) && (double)box.y0 + B.ty0 <= SPJ_LD(lb + 3);
}

// Coarse cells that can hold a reference landmark matching the rotated query (rx, ry) under some
// translation of block B: bands [*b0, *b1] (rows of cells along the block's long axis) x cells [*a0, *a1]
// of every band.  Empty: *b0 > *b1.  The cell of a coordinate is the same monotone expression the host
// used to bin the landmarks.
SPR_HD void spj_cells(const SprJoinView &V, const SpjBlock &B, double rx, double ry, int *b0, int *b1, int *a0, int *a1) {
  double fx0 = floor(SPR_DMUL(SPR_DSUB(SPR_DADD(rx, B.tx0), V.gx0), V.inv_w)), fx1 = floor(SPR_DMUL(SPR_DSUB(SPR_DADD(rx, B.tx1), V.gx0), V.inv_w));
  double fy0 = floor(SPR_DMUL(SPR_DSUB(SPR_DADD(ry, B.ty0), V.gy0), V.inv_w)), fy1 = floor(SPR_DMUL(SPR_DSUB(SPR_DADD(ry, B.ty1), V.gy0), V.inv_w));
  fx0 = fmax(fx0, 0.0); fy0 = fmax(fy0, 0.0);
  fx1 = fmin(fx1, (double)(V.ncx - 1)); fy1 = fmin(fy1, (double)(V.ncy - 1));
  if (!(fx0 <= fx1) || !(fy0 <= fy1)) { *b0 = 0; *b1 = -1; *a0 = 0; *a1 = -1; return; }
  const int cx0 = (int)fx0, cx1 = (int)fx1, cy0 = (int)fy0, cy1 = (int)fy1;
  *b0 = B.dir ? cy0 : cx0; *b1 = B.dir ? cy1 : cx1;
  *a0 = B.dir ? cx0 : cy0; *a1 = B.dir ? cx1 : cy1;
}

// Cheap filter of a (query, reference landmark) pair: can they match under a translation of the block at all?
SPR_HD bool spj_near(const SpjBlock &B, double rx, double ry, double px, double py) {
  const double ux = SPR_DSUB(px, rx), uy = SPR_DSUB(py, ry);
  return ux >= B.tx0 && ux <= B.tx1 && uy >= B.ty0 && uy <= B.ty1;
}

// One (query, reference landmark) pair against block B: dimension rule, lattice samples within reach, the
// reference's exact test on each, first-match attribution; hits are added to the counters in `tile`:
// two arrays of B.stride words, two u16 counters per word; word n of a row of array ay holds samples
// 2n - ay and 2n - ay + 1, so the two samples a pair can hit in a row share ONE word of one array.
SPR_HD void spj_pair(const SprJoinView &V, const SpjBlock &B, double rx, double ry, const double *qd, const SprJoinRef *rec, uint32_t *tile) {
#if defined(__CUDA_ARCH__)
  // the 48-byte record in three 16-byte loads
  const double2 r0 = __ldg(reinterpret_cast<const double2 *>(rec)), r1 = __ldg(reinterpret_cast<const double2 *>(rec) + 1);
  const double2 r2 = __ldg(reinterpret_cast<const double2 *>(rec) + 2);
  const double px = r0.x, py = r0.y, rd1 = r1.x, rd2 = r1.y, rd3 = r2.x;
  const unsigned long long nb = (unsigned long long)__double_as_longlong(r2.y);
  const uint32_t nbr_off = (uint32_t)nb, nbr_cnt = (uint32_t)(nb >> 32);
#else
  const double px = rec->x, py = rec->y, rd1 = rec->d1, rd2 = rec->d2, rd3 = rec->d3;
  const uint32_t nbr_off = rec->nbr_off, nbr_cnt = rec->nbr_cnt;
#endif
  if (!V.ignore_dim && !spr_dimension_match(rd1, rd2, rd3, qd, V.thr_dim, V.Sstar)) return;   // PR.cpp:315-339
  // index ranges from the regular spacing (ireach covers the drift of the accumulated samples), then the
  // exact test on the samples themselves
  const double ux = SPR_DSUB(SPR_DSUB(px, rx), B.X0), uy = SPR_DSUB(SPR_DSUB(py, ry), B.Y0);
  int i0 = (int)ceil(SPR_DMUL(SPR_DSUB(ux, V.ireach), V.inv_step)), i1 = (int)floor(SPR_DMUL(SPR_DADD(ux, V.ireach), V.inv_step));
  int j0 = (int)ceil(SPR_DMUL(SPR_DSUB(uy, V.ireach), V.inv_step)), j1 = (int)floor(SPR_DMUL(SPR_DADD(uy, V.ireach), V.inv_step));
  i0 = i0 > 0 ? i0 : 0; j0 = j0 > 0 ? j0 : 0;
  i1 = i1 < B.nx - 1 ? i1 : B.nx - 1; j1 = j1 < B.ny - 1 ? j1 : B.ny - 1;
  if (i0 > i1 || j0 > j1) return;
  if (i1 - i0 <= 1 && j1 - j0 <= 1) {
    // the usual case (threshold <= step): at most 2 x 2 samples, evaluated without branches; the two
    // samples of a row share a word of array ay
    const double Xa = SPJ_LD(B.xs + i0), Xb = SPJ_LD(B.xs + i1), Ya = SPJ_LD(B.ys + j0), Yb = SPJ_LD(B.ys + j1);
    const double dxa = SPR_DSUB(px, SPR_DADD(rx, Xa)), dxb = SPR_DSUB(px, SPR_DADD(rx, Xb));   // PR.cpp:257,310
    const double dya = SPR_DSUB(py, SPR_DADD(ry, Ya)), dyb = SPR_DSUB(py, SPR_DADD(ry, Yb));   // PR.cpp:258,312
    const double xa2 = SPR_DMUL(dxa, dxa), xb2 = SPR_DMUL(dxb, dxb), ya2 = SPR_DMUL(dya, dya), yb2 = SPR_DMUL(dyb, dyb);
    const bool two_i = i1 > i0, two_j = j1 > j0;
    uint32_t hits = (SPR_DADD(xa2, ya2) < V.Tstar ? 1u : 0u) | (two_j && SPR_DADD(xa2, yb2) < V.Tstar ? 2u : 0u) |   // PR.cpp:332-333
                    (two_i && SPR_DADD(xb2, ya2) < V.Tstar ? 4u : 0u) | (two_i && two_j && SPR_DADD(xb2, yb2) < V.Tstar ? 8u : 0u);
    if (hits && nbr_cnt) {
      // a landmark with a lower reference index that matches too is the reference's first match
      for (int b = 0; b < 4; b++) {
        if (!((hits >> b) & 1u)) continue;
        const double X = (b & 2) ? Xb : Xa, Y = (b & 1) ? Yb : Ya;
        for (uint32_t k = 0; k < nbr_cnt; k++) {
          const SprJoinNbr *n = V.nbr + nbr_off + k;
          if (spr_distance_match(rx, ry, X, Y, SPJ_LD(&n->x), SPJ_LD(&n->y), V.Tstar) &&
              (V.ignore_dim || spr_dimension_match(SPJ_LD(&n->d1), SPJ_LD(&n->d2), SPJ_LD(&n->d3), qd, V.thr_dim, V.Sstar))) {
            hits &= ~(1u << b);
            break;
          }
        }
      }
    }
    const int ay = j0 & 1;
    uint32_t *w = tile + ay * B.stride + i0 * B.nwy + ((j0 + ay) >> 1);
    const uint32_t row_a = (hits & 1u) | ((hits & 2u) << 15), row_b = ((hits >> 2) & 1u) | ((hits & 8u) << 13);
    if (row_a) SPJ_ADD(w, row_a);
    if (row_b) SPJ_ADD(w + B.nwy, row_b);
    return;
  }
  for (int i = i0; i <= i1; i++) {   // general case: one counter update per hit, array 0
    const double X = SPJ_LD(B.xs + i);
    const double dx = SPR_DSUB(px, SPR_DADD(rx, X));
    const double dx2 = SPR_DMUL(dx, dx);
    for (int j = j0; j <= j1; j++) {
      const double Y = SPJ_LD(B.ys + j);
      const double dy = SPR_DSUB(py, SPR_DADD(ry, Y));
      if (!(SPR_DADD(dx2, SPR_DMUL(dy, dy)) < V.Tstar)) continue;
      bool earlier = false;
      for (uint32_t k = 0; k < nbr_cnt && !earlier; k++) {
        const SprJoinNbr *n = V.nbr + nbr_off + k;
        earlier = spr_distance_match(rx, ry, X, Y, SPJ_LD(&n->x), SPJ_LD(&n->y), V.Tstar) &&
                  (V.ignore_dim || spr_dimension_match(SPJ_LD(&n->d1), SPJ_LD(&n->d2), SPJ_LD(&n->d3), qd, V.thr_dim, V.Sstar));
      }
      if (!earlier) SPJ_ADD(tile + i * B.nwy + (j >> 1), 1u << (16 * (j & 1)));
    }
  }
}

// One query landmark against block B, one thread (the emulation's order of work; the kernel spreads the
// same pairs over the lanes of a warp).
SPR_HD void spj_vote(const SprJoinView &V, const SpjBlock &B, int l, double rx, double ry, const double *qd, uint32_t *tile) {
  int b0, b1, a0, a1;
  spj_cells(V, B, rx, ry, &b0, &b1, &a0, &a1);
  const uint32_t *cstart = V.cell_start[B.dir] + (size_t)l * (size_t)(V.ncx * V.ncy);
  const SprJoinRef *rec = V.rec[B.dir];
  const int pitch = B.dir ? V.ncx : V.ncy;
  for (int band = b0; band <= b1; band++) {
    const uint32_t r_begin = SPJ_LD(cstart + band * pitch + a0), r_end = SPJ_LD(cstart + band * pitch + a1 + 1);
    for (uint32_t r = r_begin; r < r_end; r++)
      if (spj_near(B, rx, ry, SPJ_LD(&rec[r].x), SPJ_LD(&rec[r].y))) spj_pair(V, B, rx, ry, qd, rec + r, tile);
  }
}

// Inlier count of sample (i, j) of the block: its halves in the two counter arrays.
SPR_HD uint32_t spj_total(const uint32_t *tile, const SpjBlock &B, int i, int j) {
  const uint32_t w0 = tile[i * B.nwy + (j >> 1)], w1 = tile[B.stride + i * B.nwy + ((j + 1) >> 1)];
  return ((w0 >> (16 * (j & 1))) & 0xffffu) + ((w1 >> (16 * ((j + 1) & 1))) & 0xffffu);
}

// Slots [*s_lo, *s_hi) of block blk whose translation ordinals lie in [ord_begin, ord_end) (ordinals grow with the slot).
SPR_HD void spj_slice(const SprJoinBlock &blk, unsigned long long ord_begin, unsigned long long ord_end, int *s_lo, int *s_hi) {
  const int n_slots = (int)(blk.nx * blk.ny);
  const unsigned long long o0 = blk.ord0, rs = blk.row_stride, ny = blk.ny, nx = blk.nx;
  *s_lo = 0; *s_hi = n_slots;
  if (ord_begin > o0) {
    unsigned long long i = (ord_begin - o0) / rs, rem = (ord_begin - o0) % rs;
    if (rem >= ny) { i++; rem = 0; }
    *s_lo = i >= nx ? n_slots : (int)(i * ny + rem);
  }
  if (ord_end <= o0) {
    *s_hi = 0;
  } else {
    unsigned long long i = (ord_end - o0) / rs, rem = (ord_end - o0) % rs;
    if (rem >= ny) { i++; rem = 0; }
    *s_hi = i >= nx ? n_slots : (int)(i * ny + rem);
  }
  if (*s_hi < *s_lo) *s_hi = *s_lo;
}
