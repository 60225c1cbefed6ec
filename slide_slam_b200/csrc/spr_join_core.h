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
  int nx, ny, nwy;         // nwy: micro-tiles per row of an array
  double X0, X1, Y0, Y1;   // first / last sample per axis
  uint32_t dir;
};

SPR_HD SpjBlock spj_block(const SprJoinView &V, const SprJoinBlock &blk) {
  SpjBlock B;
  B.xs = V.lat + blk.xi; B.ys = V.lat + blk.yi;
  B.nx = (int)blk.nx; B.ny = (int)blk.ny; B.nwy = (B.ny >> 1) + 1;
  B.X0 = SPJ_LD(B.xs); B.X1 = SPJ_LD(B.xs + B.nx - 1);
  B.Y0 = SPJ_LD(B.ys); B.Y1 = SPJ_LD(B.ys + B.ny - 1);
  B.dir = blk.dir;
  return B;
}

// Can some translation of block B bring a landmark of query group `box` within reach of a reference
// landmark of its label (bounding box lb: x0, x1, y0, y1)?
SPR_HD bool spj_visible(const SprJoinView &V, const SpjBlock &B, const SprJoinBox &box, const double *lb) {
  const double tx0 = SPR_DSUB(B.X0, V.reach), tx1 = SPR_DADD(B.X1, V.reach);
  const double ty0 = SPR_DSUB(B.Y0, V.reach), ty1 = SPR_DADD(B.Y1, V.reach);
  return (double)box.x1 + tx1 >= SPJ_LD(lb) && (double)box.x0 + tx0 <= SPJ_LD(lb + 1) &&
         (double)box.y1 + ty1 >= SPJ_LD(lb + 2) && (double)box.y0 + ty0 <= SPJ_LD(lb + 3);
}

// One query landmark (rotated coordinates rx, ry; dimensions qd; label bucket l) against block B:
// every reference landmark of the label that can match under some translation of the block, every
// lattice sample of the block for which it does.  Returns true when something was added to `tile`
// (four arrays of SPJ_MAX_WORDS micro-tiles).
SPR_HD bool spj_vote(const SprJoinView &V, const SpjBlock &B, int l, double rx, double ry, const double *qd, uint32_t *tile) {
  // reference landmarks that can match lie in  [rotated query + block translations] +- reach
  const double bx0 = SPR_DSUB(SPR_DADD(rx, B.X0), V.reach), bx1 = SPR_DADD(SPR_DADD(rx, B.X1), V.reach);
  const double by0 = SPR_DSUB(SPR_DADD(ry, B.Y0), V.reach), by1 = SPR_DADD(SPR_DADD(ry, B.Y1), V.reach);
  // coarse cells of that box: the same monotone expression the host used to bin the landmarks
  double fx0 = floor(SPR_DMUL(SPR_DSUB(bx0, V.gx0), V.inv_w)), fx1 = floor(SPR_DMUL(SPR_DSUB(bx1, V.gx0), V.inv_w));
  double fy0 = floor(SPR_DMUL(SPR_DSUB(by0, V.gy0), V.inv_w)), fy1 = floor(SPR_DMUL(SPR_DSUB(by1, V.gy0), V.inv_w));
  fx0 = fmax(fx0, 0.0); fy0 = fmax(fy0, 0.0);
  fx1 = fmin(fx1, (double)(V.ncx - 1)); fy1 = fmin(fy1, (double)(V.ncy - 1));
  if (!(fx0 <= fx1) || !(fy0 <= fy1)) return false;
  const int cx0 = (int)fx0, cx1 = (int)fx1, cy0 = (int)fy0, cy1 = (int)fy1;
  const uint32_t *cstart = V.cell_start[B.dir] + (size_t)l * (size_t)(V.ncx * V.ncy);
  const SprJoinRef *rec = V.rec[B.dir];
  // bands: rows of coarse cells along the block's long axis; the cells [lo, hi] of a band are contiguous records
  const int b0 = B.dir ? cy0 : cx0, b1 = B.dir ? cy1 : cx1;
  const int a0 = B.dir ? cx0 : cy0, a1 = B.dir ? cx1 : cy1;
  const int pitch = B.dir ? V.ncx : V.ncy;
  bool voted = false;
  for (int band = b0; band <= b1; band++) {
    const uint32_t r_begin = SPJ_LD(cstart + band * pitch + a0), r_end = SPJ_LD(cstart + band * pitch + a1 + 1);
    for (uint32_t r = r_begin; r < r_end; r++) {
      const double px = SPJ_LD(&rec[r].x), py = SPJ_LD(&rec[r].y);
      if (px < bx0 || px > bx1 || py < by0 || py > by1) continue;
      if (!V.ignore_dim && !spr_dimension_match(SPJ_LD(&rec[r].d1), SPJ_LD(&rec[r].d2), SPJ_LD(&rec[r].d3), qd, V.thr_dim, V.Sstar)) continue;   // PR.cpp:315-339
      const uint32_t nbr_off = SPJ_LD(&rec[r].nbr_off), nbr_cnt = SPJ_LD(&rec[r].nbr_cnt);
      // lattice samples within reach of  (p - rotated query): index ranges from the regular spacing
      // (ireach covers the drift of the accumulated samples), then the exact test on the samples themselves
      const double ux = SPR_DSUB(SPR_DSUB(px, rx), B.X0), uy = SPR_DSUB(SPR_DSUB(py, ry), B.Y0);
      int i0 = (int)ceil(SPR_DMUL(SPR_DSUB(ux, V.ireach), V.inv_step)), i1 = (int)floor(SPR_DMUL(SPR_DADD(ux, V.ireach), V.inv_step));
      int j0 = (int)ceil(SPR_DMUL(SPR_DSUB(uy, V.ireach), V.inv_step)), j1 = (int)floor(SPR_DMUL(SPR_DADD(uy, V.ireach), V.inv_step));
      i0 = i0 > 0 ? i0 : 0; j0 = j0 > 0 ? j0 : 0;
      i1 = i1 < B.nx - 1 ? i1 : B.nx - 1; j1 = j1 < B.ny - 1 ? j1 : B.ny - 1;
      if (i0 > i1 || j0 > j1) continue;
      const bool small = i1 - i0 <= 1 && j1 - j0 <= 1;   // the usual case (threshold <= step): one micro-tile
      uint32_t pat = 0u;
      for (int i = i0; i <= i1; i++) {
        const double X = SPJ_LD(B.xs + i);
        const double dx = SPR_DSUB(px, SPR_DADD(rx, X));       // PR.cpp:257,310
        const double dx2 = SPR_DMUL(dx, dx);
        for (int j = j0; j <= j1; j++) {
          const double Y = SPJ_LD(B.ys + j);
          const double dy = SPR_DSUB(py, SPR_DADD(ry, Y));     // PR.cpp:258,312
          if (!(SPR_DADD(dx2, SPR_DMUL(dy, dy)) < V.Tstar)) continue;   // PR.cpp:332-333
          // a landmark with a lower reference index that matches too is the reference's first match
          bool earlier = false;
          for (uint32_t k = 0; k < nbr_cnt && !earlier; k++) {
            const SprJoinNbr *n = V.nbr + nbr_off + k;
            earlier = spr_distance_match(rx, ry, X, Y, SPJ_LD(&n->x), SPJ_LD(&n->y), V.Tstar) &&
                      (V.ignore_dim || spr_dimension_match(SPJ_LD(&n->d1), SPJ_LD(&n->d2), SPJ_LD(&n->d3), qd, V.thr_dim, V.Sstar));
          }
          if (earlier) continue;
          if (small) {
            pat |= 1u << (8 * (2 * (i - i0) + (j - j0)));
          } else {
            SPJ_ADD(tile + (i >> 1) * B.nwy + (j >> 1), 1u << (8 * (2 * (i & 1) + (j & 1))));
            voted = true;
          }
        }
      }
      if (pat) {
        // array (ax, ay) tiles the block with micro-tiles starting at odd (1) or even (0) samples
        const int ax = i0 & 1, ay = j0 & 1;
        SPJ_ADD(tile + (2 * ax + ay) * SPJ_MAX_WORDS + ((i0 + ax) >> 1) * B.nwy + ((j0 + ay) >> 1), pat);
        voted = true;
      }
    }
  }
  return voted;
}

// Fold: the four samples (2m + di, 2n + dj) of micro-tile w = (m, n) of array 0 collect their u8 counters
// from all four arrays and add them to the u16 totals (each total is owned by one w).
SPR_HD void spj_fold(const uint32_t *tile, int w, const SpjBlock &B, uint16_t *tot) {
  const int m = w / B.nwy, n = w - m * B.nwy;
  const uint32_t *t0 = tile + w, *t1 = t0 + SPJ_MAX_WORDS, *t2 = t0 + 2 * SPJ_MAX_WORDS, *t3 = t0 + 3 * SPJ_MAX_WORDS;
  const uint32_t a00 = t0[0];
  const uint32_t a01 = t1[0], a01n = t1[1];            // (m, n), (m, n + 1); one word past a row / the array is read but not used
  const uint32_t a10 = t2[0], a10m = t2[B.nwy];
  const uint32_t a11 = t3[0], a11n = t3[1], a11m = t3[B.nwy], a11mn = t3[B.nwy + 1];
  // byte 2 * di + dj of a00; array (0,1): word n + dj, byte 2 * di + (1 - dj); array (1,0): word m + di,
  // byte 2 * (1 - di) + dj; array (1,1): both shifted
  const uint32_t c00 = (a00 & 0xffu) + ((a01 >> 8) & 0xffu) + ((a10 >> 16) & 0xffu) + (a11 >> 24);
  const uint32_t c01 = ((a00 >> 8) & 0xffu) + (a01n & 0xffu) + (a10 >> 24) + ((a11n >> 16) & 0xffu);
  const uint32_t c10 = ((a00 >> 16) & 0xffu) + (a01 >> 24) + (a10m & 0xffu) + ((a11m >> 8) & 0xffu);
  const uint32_t c11 = (a00 >> 24) + ((a01n >> 16) & 0xffu) + ((a10m >> 8) & 0xffu) + (a11mn & 0xffu);
  const int i = 2 * m, j = 2 * n;
  if (i < B.nx) {
    if (j < B.ny && c00) tot[i * B.ny + j] += (uint16_t)c00;
    if (j + 1 < B.ny && c01) tot[i * B.ny + j + 1] += (uint16_t)c01;
  }
  if (i + 1 < B.nx) {
    if (j < B.ny && c10) tot[(i + 1) * B.ny + j] += (uint16_t)c10;
    if (j + 1 < B.ny && c11) tot[(i + 1) * B.ny + j + 1] += (uint16_t)c11;
  }
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
