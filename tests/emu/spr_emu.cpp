// spr_emu.cpp -- TEST-ONLY single-thread emulation of the lattice scoring kernel.
//
// Runs the exact per-thread code of the CUDA kernel (slide_slam_b200/csrc/spr_core.h) over the
// host-built index structures, one "thread" (chunk) at a time, so that the chunking, ordinals,
// bitmaps, candidate lists, query groups and fixed-point probing can be checked against the CPU
// oracle on a box without a GPU.  It is compiled only by tests/ and is not part of the product
// library.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../slide_slam_b200/csrc/spr_core.h"
#include "../../slide_slam_b200/csrc/spr_host.h"
#include "../../slide_slam_b200/csrc/spr_join_core.h"

extern "C" int spr_emu_match_maps(const slide_pr_params *p, const double *ref7, int n_ref,
                                  const double *qry7, int n_qry, double half_x, double half_y,
                                  long long trans_begin, long long trans_end, int *counts_out,
                                  long long counts_cap, int *best_count, long long *best_index,
                                  long long *hyps_scored, long long *filter_hits, char *errbuf, int errcap) {
  std::string err;
  spr::Lattice L;
  const double yaw_half = p->inter_loop_closure ? p->match_yaw_half_range : p->match_yaw_half_range_intra;
  int rc = spr::build_lattice(*p, half_x, half_y, yaw_half, trans_begin, trans_end, /*ring_major: exercised by the odd-start slices of the tests*/ (trans_begin & 1) != 0, L, err);
  auto fail = [&](int code) { if (errbuf && errcap > 0) { strncpy(errbuf, err.c_str(), errcap - 1); errbuf[errcap - 1] = 0; } return code; };
  if (rc != SLIDE_PR_OK) return fail(rc);
  *best_count = -10000; *best_index = -1; *hyps_scored = 0; *filter_hits = 0;
  if (L.status == SLIDE_PR_SANITY_RETURN) return SLIDE_PR_SANITY_RETURN;
  double qrad = 0;
  for (int j = 0; j < n_qry; j++) qrad = std::max(qrad, std::hypot(qry7[7 * j + 1], qry7[7 * j + 2]));
  spr::RefIndex R;
  rc = spr::build_ref_index(*p, ref7, n_ref, qrad + std::max(std::fabs(half_x), std::fabs(half_y)) + p->match_xy_step_size, R, err);
  if (rc != SLIDE_PR_OK) return fail(rc);
  spr::QuerySet Q;
  rc = spr::build_query_set(R, qry7, n_qry, Q, err);
  if (rc != SLIDE_PR_OK) return fail(rc);
  const int n_yaw = (int)L.yaw.size(), nqp = Q.nqp, n_groups = nqp / SPR_QGROUP;
  const size_t nrot = (size_t)std::max(1, n_yaw * nqp);
  std::vector<double> qrot(2 * nrot);
  std::vector<int32_t> qxy_fx(2 * nrot), qyx_fx(2 * nrot);
  std::vector<SprBox> gbox((size_t)std::max(1, n_yaw * n_groups));
  for (int a = 0; a < n_yaw; a++)
    for (int g = 0; g < n_groups; g++)
      spr_rotate_group(L.cs.data(), Q.qxy.data(), Q.qlabel.data(), R.grid, nqp, n_groups, a, g, qxy_fx.data(),
                       qyx_fx.data(), qrot.data(), gbox.data());
  SprView V{};
  V.lat = L.lat.data(); V.chunks = L.chunks.data(); V.n_chunks = (uint32_t)L.chunks.size();
  V.n_yaw = n_yaw; V.cs = L.cs.data(); V.nqp = nqp; V.n_groups = n_groups;
  V.qrotq_xy = qxy_fx.data(); V.qrotq_yx = qyx_fx.data(); V.gbox = gbox.data(); V.qrot = qrot.data();
  V.qxy = Q.qxy.data(); V.qdims = Q.qdims.data(); V.label_gseg = Q.label_gseg.data(); V.qlabel = Q.qlabel.data();
  V.n_labels = (int)R.labels.size(); V.n_ref = n_ref; V.labelbox = R.labelbox.data();
  V.bitmap = R.bitmap.data();
  for (int d = 0; d < 2; d++) { V.rank16[d] = R.rank16[d].data(); V.row_rank[d] = R.row_rank[d].data(); V.cand[d] = R.cand[d].data();
    V.cellref[d] = R.cellref[d].data(); V.cell_base[d] = R.cell_base[d].data(); }
  V.reftab = R.reftab.data(); V.ref_base = R.ref_base.data();
  V.grid = R.grid; V.Tstar = R.Tstar; V.Sstar = R.Sstar; V.thr_dim = p->match_threshold_dimension;
  V.ignore_dim = p->ignore_dimension;
  const SprGrid &G = V.grid;
  unsigned long long best = 0;
  long long scored = 0, hits = 0;
  const long long tb = trans_begin < 0 ? 0 : trans_begin;
  for (uint32_t ci = 0; ci < V.n_chunks; ci++) {
    const SprChunk &ch = V.chunks[ci];
    const int d = (int)ch.dir;
    const int32_t aq0 = spr_fx(ch.across, G.S), bq0 = spr_fx(V.lat[ch.along_off], G.S);
    const int32_t aqb = spr_bias_across(aq0, G.F), bqb = spr_bias_along(bq0, G.F);
    // this chunk's own window (the kernel uses the union over the warp's 32 chunks: weaker skip)
    const int32_t X0 = d ? bq0 : aq0, X1 = d ? bq0 + (32 << G.F) : aq0;
    const int32_t Y0 = d ? aq0 : bq0, Y1 = d ? aq0 : bq0 + (32 << G.F);
    const int32_t *qfx = d ? V.qrotq_yx : V.qrotq_xy;
    for (int a = 0; a < n_yaw; a++) {
      uint32_t cnt[32] = {0};
      for (int l = 0; l < V.n_labels; l++) {
        const uint32_t *plane = V.bitmap + (size_t)l * G.label_stride + (d ? G.plane_words[0] : 0);
        for (int g = V.label_gseg[l]; g < V.label_gseg[l + 1]; g++) {
          if (!spr_group_visible(V.gbox[(size_t)a * n_groups + g], V.labelbox[l], X0, X1, Y0, Y1)) continue;
          for (int k = 0; k < SPR_QGROUP; k++) {
            const int s = g * SPR_QGROUP + k;
            const size_t qi = (size_t)a * nqp + s;
            const int32_t asum = aqb + qfx[2 * qi], bsum = bqb + qfx[2 * qi + 1];
            const uint32_t H = spr_probe(plane, (uint32_t)G.W[d], (uint32_t)G.R[d] - 1u, (uint32_t)G.maxbit[d], G.F, asum, bsum, ch.valid);
            if (!H) continue;
            hits += __builtin_popcount(H);
            uint32_t row, bit;
            spr_cell_of(G.F, asum, bsum, &row, &bit);
            uint32_t P = spr_verify_mask(V, spr_global_tables(V, (uint32_t)d, l), (uint32_t)d, row, bit, H, qrot[2 * qi],
                                         qrot[2 * qi + 1], ch.across, V.lat + ch.along_off, V.qdims + 3 * (size_t)s);
            while (P) { cnt[SPR_FFS(P) - 1]++; P &= P - 1; }
          }
        }
      }
      for (int b = 0; b < 32; b++) {
        if (!((ch.valid >> b) & 1u)) continue;
        const unsigned long long ord = (unsigned long long)ch.ord_base + (unsigned long long)b * ch.ord_stride;
        const unsigned long long h = ord * (unsigned long long)n_yaw + (unsigned long long)a;
        const unsigned long long key = spr_make_key(cnt[b], h);
        if (key > best) best = key;
        scored++;
        if (counts_out) {
          const long long slot = ((long long)ord - tb) * n_yaw + a;
          if (slot >= 0 && slot < counts_cap) counts_out[slot] = (int)cnt[b];
        }
      }
    }
  }
  if (best) { *best_count = spr_key_count(best); *best_index = spr_key_index(best); }
  *hyps_scored = scored; *filter_hits = hits;
  return SLIDE_PR_OK;
}

// The pair-join scorer (spr_join.cu) run one thread at a time over the host-built structures: same blocks,
// same visibility test, same per-pair code and counters, same scan.
extern "C" int spr_emu_join_match_maps(const slide_pr_params *p, const double *ref7, int n_ref,
                                       const double *qry7, int n_qry, double half_x, double half_y,
                                       long long trans_begin, long long trans_end, int *counts_out,
                                       long long counts_cap, int *best_count, long long *best_index,
                                       long long *hyps_scored, long long *filter_hits, char *errbuf, int errcap) {
  std::string err;
  spr::Lattice L;
  const double yaw_half = p->inter_loop_closure ? p->match_yaw_half_range : p->match_yaw_half_range_intra;
  int rc = spr::build_lattice(*p, half_x, half_y, yaw_half, 0, -1, false, L, err, true);
  auto fail = [&](int code) { if (errbuf && errcap > 0) { strncpy(errbuf, err.c_str(), errcap - 1); errbuf[errcap - 1] = 0; } return code; };
  if (rc != SLIDE_PR_OK) return fail(rc);
  *best_count = -10000; *best_index = -1; *hyps_scored = 0; *filter_hits = 0;
  if (L.status == SLIDE_PR_SANITY_RETURN) return SLIDE_PR_SANITY_RETURN;
  spr::JoinRef J;
  if ((rc = spr::build_join_ref(*p, ref7, n_ref, J, err)) != SLIDE_PR_OK) return fail(rc);
  spr::uvec<SprJoinBlock> blocks;
  double drift = 0;
  if ((rc = spr::build_join_blocks(L, p->match_xy_step_size, blocks, &drift, err)) != SLIDE_PR_OK) return fail(rc);
  spr::QuerySet Q;
  if ((rc = spr::build_query_set(J.labels, qry7, n_qry, Q, err)) != SLIDE_PR_OK) return fail(rc);
  const int n_yaw = (int)L.yaw.size(), nqp = Q.nqp, n_groups = nqp / SPR_QGROUP;
  std::vector<int32_t> glabel((size_t)std::max(n_groups, 1), 0);
  for (int g = 0; g < n_groups; g++) glabel[g] = Q.qlabel[(size_t)g * SPR_QGROUP];
  std::vector<double> qrot(2 * (size_t)std::max(1, n_yaw * nqp));
  std::vector<SprJoinBox> gbox((size_t)std::max(1, n_yaw * n_groups));
  for (int a = 0; a < n_yaw; a++)
    for (int g = 0; g < n_groups; g++) {   // spr_join_rotate_kernel
      SprJoinBox box = {INFINITY, -INFINITY, INFINITY, -INFINITY};
      for (int k = 0; k < SPR_QGROUP; k++) {
        const int js = g * SPR_QGROUP + k;
        double rx = NAN, ry = NAN;
        if (Q.qlabel[js] >= 0) {
          spr_rotate(L.cs[2 * a], L.cs[2 * a + 1], Q.qxy[2 * (size_t)js], Q.qxy[2 * (size_t)js + 1], &rx, &ry);
          box.x0 = std::min(box.x0, std::nextafterf((float)rx, -INFINITY)); box.x1 = std::max(box.x1, std::nextafterf((float)rx, INFINITY));
          box.y0 = std::min(box.y0, std::nextafterf((float)ry, -INFINITY)); box.y1 = std::max(box.y1, std::nextafterf((float)ry, INFINITY));
        }
        qrot[2 * ((size_t)a * nqp + js)] = rx; qrot[2 * ((size_t)a * nqp + js) + 1] = ry;
      }
      gbox[(size_t)a * n_groups + g] = box;
    }
  SprJoinView V{};
  V.lat = L.lat.data(); V.qrot = qrot.data(); V.gbox = gbox.data(); V.qdims = Q.qdims.data(); V.glabel = glabel.data();
  V.qxy = Q.qxy.data(); V.qlabel = Q.qlabel.data(); V.cs = L.cs.data();
  V.nqp = nqp; V.n_groups = n_groups; V.n_yaw = n_yaw; V.n_labels = (int)J.labels.size();
  for (int d = 0; d < 2; d++) { V.rec[d] = J.rec[d].data(); V.xy[d] = J.xy[d].data(); V.cell_start[d] = J.cell_start[d].data(); }
  V.nbr = J.nbr.data(); V.labelbox = J.labelbox.data();
  V.gx0 = J.gx0; V.gy0 = J.gy0; V.inv_w = J.inv_w; V.ncx = J.ncx; V.ncy = J.ncy;
  V.Tstar = J.Tstar; V.Sstar = J.Sstar; V.thr_dim = p->match_threshold_dimension; V.ignore_dim = p->ignore_dimension;
  double qmag = 0.0;
  for (int j = 0; j < n_qry; j++) qmag = std::max(qmag, std::hypot(qry7[7 * (size_t)j + 1], qry7[7 * (size_t)j + 2]));
  V.reach = J.reach + 64.0 * 2.220446049250313e-16 * (J.max_abs + qmag + std::max(std::fabs(half_x), std::fabs(half_y)));
  V.ireach = V.reach + 2.0 * drift + 1e-9; V.inv_step = 1.0 / p->match_xy_step_size;
  V.blocks = blocks.data(); V.n_blocks = (uint32_t)blocks.size();
  const unsigned long long ob = trans_begin < 0 ? 0 : (unsigned long long)trans_begin;
  const unsigned long long oe = trans_end < 0 ? L.n_translations : std::min<unsigned long long>((unsigned long long)trans_end, L.n_translations);
  unsigned long long best = 0;
  long long scored = 0;
  std::vector<uint32_t> tile(SPJ_TILE_WORDS);
  for (int a = 0; a < n_yaw; a++)
    for (const SprJoinBlock &blk : blocks) {
      const SpjBlock B = spj_block(V, blk);
      const int n_slots = B.nx * B.ny;
      if (n_slots > SPJ_MAX_SLOTS || 2 * B.stride > SPJ_TILE_WORDS) { err = "block exceeds the kernel's limits"; return fail(SLIDE_PR_ERR_INTERNAL); }
      std::fill(tile.begin(), tile.end(), 0xdeadbeefu);   // only the words the kernel zeroes may be relied on
      std::fill(tile.begin(), tile.begin() + 2 * B.stride, 0u);
      for (int g = 0; g < n_groups; g++) {
        if (!spj_visible(B, gbox[(size_t)a * n_groups + g], V.labelbox + 4 * (size_t)glabel[g])) continue;
        for (int k = 0; k < SPR_QGROUP; k++) {
          const int js = g * SPR_QGROUP + k;
          const double rx = qrot[2 * ((size_t)a * nqp + js)], ry = qrot[2 * ((size_t)a * nqp + js) + 1];
          if (rx == rx) spj_vote(V, B, glabel[g], rx, ry, Q.qdims.data() + 3 * (size_t)js, tile.data());
        }
      }
      int s_lo, s_hi;
      spj_slice(blk, ob, oe, &s_lo, &s_hi);
      for (int s = s_lo; s < s_hi; s++) {
        const int i = s / B.ny, j = s - i * B.ny;
        const unsigned long long ord = (unsigned long long)blk.ord0 + (unsigned long long)i * blk.row_stride + (unsigned long long)j;
        if (ord < ob || ord >= oe) { err = "slot outside the slice"; return fail(SLIDE_PR_ERR_INTERNAL); }
        const uint32_t cnt = spj_total(tile.data(), B, i, j);
        const unsigned long long key = spr_make_key(cnt, ord * (unsigned long long)n_yaw + (unsigned long long)a);
        if (key > best) best = key;
        scored++;
        if (counts_out) {
          const long long slot = (long long)(ord - ob) * n_yaw + a;
          if (slot >= 0 && slot < counts_cap) counts_out[slot] = (int)cnt;
        }
      }
    }
  if (best) { *best_count = spr_key_count(best); *best_index = spr_key_index(best); }
  *hyps_scored = scored;
  return SLIDE_PR_OK;
}

// lattice enumeration through the product's host builder (chunks -> translations), for tests
extern "C" long long spr_emu_lattice(const slide_pr_params *p, double half_x, double half_y, double *tx, double *ty,
                                     long long cap, int *n_yaw, double *yaw, int yaw_cap, int *status) {
  std::string err;
  spr::Lattice L;
  const double yaw_half = p->inter_loop_closure ? p->match_yaw_half_range : p->match_yaw_half_range_intra;
  int rc = spr::build_lattice(*p, half_x, half_y, yaw_half, 0, -1, false, L, err);
  *status = rc != SLIDE_PR_OK ? rc : L.status;
  if (rc != SLIDE_PR_OK || L.status != 0) return -1;
  *n_yaw = (int)L.yaw.size();
  for (int i = 0; i < *n_yaw && i < yaw_cap; i++) yaw[i] = L.yaw[i];
  // every translation must be covered by exactly one chunk bit
  std::vector<int> seen((size_t)L.n_translations, 0);
  for (const SprChunk &c : L.chunks)
    for (int b = 0; b < 32; b++) {
      if (!((c.valid >> b) & 1u)) continue;
      const unsigned long long ord = (unsigned long long)c.ord_base + (unsigned long long)b * c.ord_stride;
      if (ord >= L.n_translations) return -2;
      seen[ord]++;
      const double along = L.lat[c.along_off + b];
      if ((long long)ord < cap) { tx[ord] = c.dir ? along : c.across; ty[ord] = c.dir ? c.across : along; }
    }
  for (size_t i = 0; i < seen.size(); i++) if (seen[i] != 1) return -3;
  // translation_of must agree
  for (unsigned long long o = 0; o < L.n_translations && (long long)o < cap; o += 1 + L.n_translations / 5000) {
    double x, y; int ring;
    if (!spr::translation_of(L, o, &x, &y, &ring) || x != tx[o] || y != ty[o]) return -4;
  }
  return (long long)L.n_translations;
}
