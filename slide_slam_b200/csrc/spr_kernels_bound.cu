// spr_kernels_bound.cu -- upper bounds for the bound-and-verify lattice search.
//
// MatchMaps only returns the best hypothesis (place_recognition.cpp:361-386).  The number of query
// landmarks whose occupancy-bitmap probe hits -- the conservative filter that precedes every exact
// test in spr_kernels.cu -- is an UPPER BOUND of a hypothesis' inlier count, and it costs a third of
// the exact count (no hit queue, no fp64).  The search therefore runs in two phases:
//   1. this file: the bound of every hypothesis (same probes, same visibility culling as the exact
//      kernel), kept bit-sliced in registers -- plane i holds bit i of the 32 counters a lane owns,
//      adding a probe result is a handful of LOP3 -- and stored as bit planes; the best-bounded
//      hypothesis of every yaw is scored exactly (spr_seed_kernel) to seed the running best;
//   2. spr_score_lattice_kernel verifies exactly only the hypotheses whose bound reaches the
//      running best (a plane-wise compare builds the lane's candidate mask), so the winner -- max
//      count, ties to the smallest canonical index -- is the one the exhaustive search finds.
// Several labels are handled per launch: their bitmap planes are staged together in shared
// memory with TMA bulk copies (as many as fit), the counters never leave the registers in between.
#include <cuda_runtime.h>
#include <stdint.h>

#include "spr_core.h"
#include "spr_kernels.h"

#define SPR_FULL 0xffffffffu
#define SPR_SMEM_LIMIT (227 * 1024)
#ifndef SPB_THREADS
#define SPB_THREADS 768
#endif
#ifndef SPB_GLOBAL_MINB
#define SPB_GLOBAL_MINB 3   // resident 256-thread CTAs per SM of the variant that reads the planes in place
#endif

__device__ __forceinline__ uint32_t spb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void spb_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spb_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void spb_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(spb_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void spb_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(spb_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(spb_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ bool spb_mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(spb_smem_addr(bar)), "r"(parity)
               : "memory");
  return ok != 0u;
}

// full adder on 32 independent bit positions
__device__ __forceinline__ void spb_fa(uint32_t a, uint32_t b, uint32_t c, uint32_t &s, uint32_t &cy) {
  s = a ^ b ^ c;
  cy = (a & b) | (c & (a ^ b));
}

// add plane `c` (weight 2^from) into the binary counter P
template <int PLANES>
__device__ __forceinline__ void spb_ripple(uint32_t (&P)[PLANES], uint32_t c, int from) {
#pragma unroll
  for (int i = 0; i < PLANES; i++) {
    if (i < from) continue;
    const uint32_t t = P[i] & c;
    P[i] ^= c;
    c = t;
  }
}

// Work item = (yaw, DOUBLE GROUP): lane k owns chunk k of the group (32 lattice samples along its
// row) and chunk k of the next group, which continues it for another 32 samples (empty where the
// row ends).  One probe = three consecutive bitmap words of the row and two funnel shifts = the
// filter bits of 64 hypotheses.
// REFINE: second, tighter bound of the candidate double groups only (B.cand_items): the probe
// picks one of four bitmap variants by the half-cell the point falls in (spr_variant_mark_kernel),
// planes read in place; the counters start from zero and replace the first bound.
template <int PLANES, bool SMEM_TAB, bool REFINE>
__global__ void __launch_bounds__(SMEM_TAB ? SPB_THREADS : 256, SMEM_TAB ? 1 : SPB_GLOBAL_MINB)
spr_bound_lattice_kernel(const __grid_constant__ SprView V, const __grid_constant__ SprBoundLaunch B, const int n_dg_local,
                         const long long n_items) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const SprGrid &G = V.grid;
  const int32_t F = G.F;
  const uint32_t d = B.dir;
  const uint32_t W = (uint32_t)G.W[d], maxbit2 = (uint32_t)G.maxbit[d] + 32u;
  const uint32_t W4 = W * 4u;
  const uint32_t PW4 = G.plane_words[d] * 4u;
  // Row band [row_begin, row_end) of the planes handled by this launch (the whole plane unless it
  // does not fit in shared memory).  Staged as band_rows rows + one all-zero row per label; rows
  // outside the band clamp onto the zero row (unsigned min), so the probe code is unchanged.
  const uint32_t band_rows = SMEM_TAB ? B.row_end - B.row_begin : (uint32_t)G.R[d];
  const uint32_t Rm1 = SMEM_TAB ? band_rows : (uint32_t)G.R[d] - 1u;
  const uint32_t BW = ((band_rows + 1u) * W + 3u) & ~3u;  // words per staged label (16-byte aligned bulk-copy targets)
  const int32_t row_shift = SMEM_TAB ? (int32_t)(B.row_begin << F) : 0;

  if (REFINE && __ldg(B.cand_count) < B.refine_min) return;  // few candidates: not worth refining (uniform for the grid)
  // staged planes: the launch's labels, or (refinement) the four half-cell variants of its one label
  const uint32_t NP = REFINE ? 4u : (uint32_t)B.n_labels;
  if (SMEM_TAB) {  // one TMA bulk copy per plane
    // [4 zero words][planes x BW words][4 zero words][mbarrier]: a probe reads the word before and
    // the word after its base word, also for the first row of the first plane / the last zero row
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 4 + (size_t)NP * BW + 4);
    if (threadIdx.x == 0) {
      spb_mbar_init(bar, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < NP * W; i += blockDim.x)
      smem[4 + (size_t)(i / W) * BW + (size_t)band_rows * W + (i % W)] = 0u;
    if (threadIdx.x < 4) { smem[threadIdx.x] = 0u; smem[4 + (size_t)NP * BW + threadIdx.x] = 0u; }
    if (threadIdx.x < NP * 4u) {  // alignment words behind each plane's zero row
      const uint32_t k = threadIdx.x >> 2, j = threadIdx.x & 3u;
      if ((band_rows + 1u) * W + j < BW) smem[4 + (size_t)k * BW + (band_rows + 1u) * W + j] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      spb_mbar_expect_tx(bar, NP * band_rows * W * 4u);
      for (uint32_t k = 0; k < NP; k++) {
        const uint32_t *src = REFINE ? V.vbitmap + (4 * ((size_t)B.labels[0] * G.label_stride + (d ? G.plane_words[0] : 0u)) + (size_t)k * G.plane_words[d])
                                     : V.bitmap + ((size_t)B.labels[k] * G.label_stride + (d ? G.plane_words[0] : 0u));
        spb_bulk_g2s(smem + 4 + (size_t)k * BW, src + (size_t)B.row_begin * W, band_rows * W * 4u, bar);
      }
    }
    for (uint32_t spin = 0; !spb_mbar_try_wait(bar, 0u); spin++)
      if (spin > 200000000u) __trap();  // a lost copy must not hang the GPU
  }
  // across range (fixed point) covered by the band: plane row r holds across cell r - 1
  const int32_t band_lo = SMEM_TAB ? ((int32_t)B.row_begin - 1) << F : -(1 << 30);
  const int32_t band_hi = SMEM_TAB ? ((int32_t)B.row_end - 1) << F : (1 << 30);
  const int32_t *q_fx = d ? V.qrotq_yx : V.qrotq_xy;
  const uint32_t n_wg_total = B.n_chunks_total / SPR_WARP_CHUNKS;

  for (;;) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(B.work_counter, 1ull);
    item = __shfl_sync(SPR_FULL, item, 0);
    if (REFINE) {
      const uint32_t n_cand = __ldg(B.cand_count);
      if (n_cand < B.refine_min || item >= (long long)n_cand) break;  // few candidates: not worth refining
      item = (long long)__ldg(B.cand_items + item);
    } else if (item >= n_items) {
      break;
    }
    const int a = (int)(item / n_dg_local);
    const int dg = B.shard_index + (int)(item % n_dg_local) * B.shard_count;
    const uint32_t cidx = B.chunk_begin + (uint32_t)dg * (2 * SPR_WARP_CHUNKS) + (uint32_t)lane;  // first chunk; + 32: its continuation
    const uint4 c0 = __ldcs(reinterpret_cast<const uint4 *>(V.chunks + cidx));
    const double across = __hiloint2double((int)c0.y, (int)c0.x);
    const uint32_t along_off = c0.z, validA = c0.w;
    const uint32_t validB = __ldcs(reinterpret_cast<const uint4 *>(V.chunks + cidx + SPR_WARP_CHUNKS)).w;
    const int32_t aq0 = spr_fx(across, G.S);
    const int32_t bq0 = spr_fx(__ldg(V.lat + along_off), G.S);
    const int32_t aqb = spr_bias_across(aq0, F) - row_shift, bqb2 = spr_bias_along(bq0, F) + (32 << F);
    const int32_t big = 1 << 30;
    const bool live = validA != 0u;  // a continuation exists only behind a first chunk
    const int32_t ext = (validB ? 64 : 32) << F;
    const int32_t lx0 = d ? bq0 : aq0, lx1 = d ? bq0 + ext : aq0;
    const int32_t ly0 = d ? aq0 : bq0, ly1 = d ? aq0 : bq0 + ext;
    const int32_t X0 = __reduce_min_sync(SPR_FULL, live ? lx0 : big), X1 = __reduce_max_sync(SPR_FULL, live ? lx1 : -big);
    const int32_t Y0 = __reduce_min_sync(SPR_FULL, live ? ly0 : big), Y1 = __reduce_max_sync(SPR_FULL, live ? ly1 : -big);

    // bit-sliced counters: PA[i] / PB[i] = bit i of the 32 bounds of the lane's first / second chunk
    uint32_t PA[PLANES], PB[PLANES];
    uint32_t *gpA = B.planes + (((size_t)a * n_wg_total + cidx / SPR_WARP_CHUNKS) * PLANES) * 32 + lane;
    uint32_t *gpB = gpA + (size_t)PLANES * 32;
    if (B.first) {
#pragma unroll
      for (int i = 0; i < PLANES; i++) PA[i] = PB[i] = 0u;
    } else {
#pragma unroll
      for (int i = 0; i < PLANES; i++) { PA[i] = __ldcs(gpA + i * 32); PB[i] = __ldcs(gpB + i * 32); }
    }
    // carry-save state: planes 0..1 absorb four probe results per step and emit one plane of weight
    // 4; two of those make one of weight 8 (into plane 2), ... up to weight 32, which ripples on
    uint32_t p4A = 0u, p8A = 0u, p16A = 0u, p4B = 0u, p8B = 0u, p16B = 0u;
    bool have4 = false, have8 = false, have16 = false;

    if (X0 <= X1) {
      for (int k = 0; k < B.n_labels; k++) {
        const int l = B.labels[k];
        const int g0 = V.label_gseg[l], g1 = V.label_gseg[l + 1];
        if (g0 >= g1) continue;
        const uint32_t *bits = SMEM_TAB ? smem + 4 + (REFINE ? (size_t)0 : (size_t)k * BW)
                               : REFINE ? V.vbitmap + (4 * ((size_t)l * G.label_stride + (d ? G.plane_words[0] : 0u)))
                                        : V.bitmap + ((size_t)l * G.label_stride + (d ? G.plane_words[0] : 0u));
        SprBox lb = V.labelbox[l];
        if (d == 0) { lb.x0 = max(lb.x0, band_lo); lb.x1 = min(lb.x1, band_hi); }  // marked cells inside the band
        else        { lb.y0 = max(lb.y0, band_lo); lb.y1 = min(lb.y1, band_hi); }
        if (lb.x0 >= lb.x1 || lb.y0 >= lb.y1) continue;
        const int32_t tx_lo = lb.x0 - X1, tx_hi = lb.x1 - X0, ty_lo = lb.y0 - Y1, ty_hi = lb.y1 - Y0;
        const int4 *gbp = reinterpret_cast<const int4 *>(V.gbox) + ((size_t)a * (size_t)V.n_groups + (size_t)g0);
        const int4 *qgp = reinterpret_cast<const int4 *>(q_fx + 2 * (size_t)a * (size_t)V.nqp) + (size_t)g0 * (SPR_QGROUP / 2);
        for (int gb = g0; gb < g1; gb += 32) {
          bool vis = false;
          if (gb + lane < g1) {
            const int4 box = __ldg(gbp + (gb - g0) + lane);  // (x0, x1, y0, y1)
            vis = box.y > tx_lo && box.x < tx_hi && box.w > ty_lo && box.z < ty_hi;
          }
          uint32_t vm = __ballot_sync(SPR_FULL, vis);
          while (vm) {
            const int kk = __ffs(vm) - 1;
            vm &= vm - 1;
            const int4 *q = qgp + (size_t)(gb - g0 + kk) * (SPR_QGROUP / 2);
            int4 v[SPR_QGROUP / 2];
#pragma unroll
            for (int u = 0; u < SPR_QGROUP / 2; u++) v[u] = __ldg(q + u);
#pragma unroll
            for (int hq = 0; hq < 2; hq++) {  // four queries at a time
              uint32_t HA[4], HB[4];
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const int4 vv = v[2 * hq + (u >> 1)];
                const int32_t qa_ = (u & 1) ? vv.z : vv.x, qb_ = (u & 1) ? vv.w : vv.y;
                const uint32_t row = min((uint32_t)((aqb + qa_) >> F), Rm1);     // rows 0 and Rm1 are all-zero
                // bit = plane bit of the SECOND chunk's first sample (the first chunk starts 32 bits
                // earlier, possibly before the row: the word in front of a row is a zero pad word)
                const uint32_t bit = min((uint32_t)((bqb2 + qb_) >> F), maxbit2);  // words >= maxbit/32 are all-zero
                // byte offset with two multiply-adds (FMA pipe) instead of LEA.HI + LEA (ALU pipe, the busy one).
                // (Measured and rejected: the shifts as IMAD.HI / IMAD.WIDE by 2^(32-F) -- same instruction
                // count, 17 % slower: the high-word multiply is a quarter-rate instruction on sm_100.)
                uint32_t boff;
                asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(boff) : "r"(bit >> 5), "r"(row * W4));
                if (REFINE) {  // variant = 2 * (upper half of the cell across) + (upper half of the cell along)
                  const uint32_t var = ((uint32_t)((aqb + qa_) >> (F - 1)) & 1u) * 2u + ((uint32_t)((bqb2 + qb_) >> (F - 1)) & 1u);
                  boff += var * (SMEM_TAB ? BW * 4u : PW4);
                }
                const uint32_t *p = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(bits) + boff);
                const uint32_t w0 = p[-1], w1 = p[0], w2 = p[1];
                HA[u] = __funnelshift_r(w0, w1, bit);
                HB[u] = __funnelshift_r(w1, w2, bit);
              }
              uint32_t tA, tB, f4A, f4B;
              spb_fa(PA[0], HA[0], HA[1], PA[0], tA);
              spb_fa(PA[0], HA[2], HA[3], PA[0], tB);
              spb_fa(PA[1], tA, tB, PA[1], f4A);
              spb_fa(PB[0], HB[0], HB[1], PB[0], tA);
              spb_fa(PB[0], HB[2], HB[3], PB[0], tB);
              spb_fa(PB[1], tA, tB, PB[1], f4B);
              if (!have4) { p4A = f4A; p4B = f4B; have4 = true; continue; }
              uint32_t e8A, e8B;
              spb_fa(PA[2], p4A, f4A, PA[2], e8A);
              spb_fa(PB[2], p4B, f4B, PB[2], e8B);
              have4 = false;
              if (!have8) { p8A = e8A; p8B = e8B; have8 = true; continue; }
              uint32_t s16A, s16B;
              spb_fa(PA[3], p8A, e8A, PA[3], s16A);
              spb_fa(PB[3], p8B, e8B, PB[3], s16B);
              have8 = false;
              if (!have16) { p16A = s16A; p16B = s16B; have16 = true; continue; }
              uint32_t cA, cB;
              spb_fa(PA[4], p16A, s16A, PA[4], cA);
              spb_fa(PB[4], p16B, s16B, PB[4], cB);
              have16 = false;
              spb_ripple<PLANES>(PA, cA, 5);
              spb_ripple<PLANES>(PB, cB, 5);
            }
          }
        }
      }
    }
    // pending planes back into the binary counters
    if (have4) { spb_ripple<PLANES>(PA, p4A, 2); spb_ripple<PLANES>(PB, p4B, 2); }
    if (have8) { spb_ripple<PLANES>(PA, p8A, 3); spb_ripple<PLANES>(PB, p8B, 3); }
    if (have16) { spb_ripple<PLANES>(PA, p16A, 4); spb_ripple<PLANES>(PB, p16B, 4); }
#pragma unroll
    for (int i = 0; i < PLANES; i++) {
      PA[i] &= validA;
      PB[i] &= validB;
      __stcs(gpA + i * 32, PA[i]);
      __stcs(gpB + i * 32, PB[i]);
    }
    if (B.last) {
      // largest bound of each chunk (MSB-first narrowing of the candidate bits), then of each group
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const uint32_t valid = half ? validB : validA;
        uint32_t cand = valid, val = 0u;
#pragma unroll
        for (int i = PLANES - 1; i >= 0; i--) {
          const uint32_t t = cand & (half ? PB[i] : PA[i]);
          if (t) { cand = t; val |= 1u << i; }
        }
        const bool lv = valid != 0u;
        const uint32_t packed = lv ? ((val << 5) | (uint32_t)(__ffs(cand) - 1)) : 0u;
        const uint32_t wmax = __reduce_max_sync(SPR_FULL, packed);
        const uint32_t who = __ballot_sync(SPR_FULL, packed == wmax && lv);
        const uint32_t grp = cidx / SPR_WARP_CHUNKS + (uint32_t)half;
        if (lane == 0) B.item_ub[(size_t)a * n_wg_total + grp] = wmax >> 5;
        if (!REFINE && who && lane == __ffs(who) - 1)
          atomicMax(B.seed_key + (size_t)a * SPR_SEED_SLOTS + grp % SPR_SEED_SLOTS,
                    ((unsigned long long)(val + 1u) << SPR_KEY_IDX_BITS) |
                        ((unsigned long long)(cidx + (uint32_t)half * SPR_WARP_CHUNKS) * 32ull + (unsigned long long)(packed & 31u)));
      }
    }
    __syncwarp();
  }
}

// Exact score of the best-bounded hypotheses -- SPR_SEED_SLOTS per yaw candidate, one per residue
// class of the work-item columns -- (one CTA each, the warps stride over the query landmarks): seeds the running best of the verification phase.  Same
// decision code as the hypothesis-list scorer.
__global__ void __launch_bounds__(256)
spr_seed_kernel(const __grid_constant__ SprView V, const unsigned long long *__restrict__ seed_key,
                unsigned long long *best_key) {
  __shared__ int s_cnt;
  const int a = blockIdx.x / SPR_SEED_SLOTS;
  const unsigned long long sk = seed_key[blockIdx.x];
  if (sk == 0ull) return;  // uniform for the CTA
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const unsigned long long loc = sk & SPR_KEY_IDX_MASK;
  const SprChunk ch = V.chunks[loc >> 5];
  const int b = (int)(loc & 31ull);
  const double t = V.lat[ch.along_off + b];
  const double tx = ch.dir ? t : ch.across, ty = ch.dir ? ch.across : t;
  int cnt = 0;
  for (int js = threadIdx.x; js < V.nqp; js += blockDim.x) {
    const int l = V.qlabel[js];
    if (l < 0) continue;  // padding
    const size_t qi = (size_t)a * (size_t)V.nqp + (size_t)js;
    const double rx = V.qrot[2 * qi], ry = V.qrot[2 * qi + 1];
    const double xt = SPR_DADD(rx, tx), yt = SPR_DADD(ry, ty);
    uint32_t row, bit;
    if (spr_point_cell(V, l, xt, yt, &row, &bit) &&
        spr_verify_cell(V, spr_global_tables(V, 0u, l), 0u, row, bit, rx, ry, tx, ty, V.qdims + 3 * (size_t)js))
      cnt++;
  }
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) cnt += __shfl_xor_sync(SPR_FULL, cnt, dlt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long ord = (unsigned long long)ch.ord_base + (unsigned long long)b * ch.ord_stride;
    atomicMax(best_key, spr_make_key((uint32_t)s_cnt, ord * (unsigned long long)V.n_yaw + (unsigned long long)a));
  }
}

cudaError_t spr_launch_seed(const SprView &V, const unsigned long long *seed_key, unsigned long long *best_key,
                            cudaStream_t st) {
  if (V.n_yaw <= 0 || V.nqp <= 0) return cudaSuccess;
  spr_seed_kernel<<<V.n_yaw * SPR_SEED_SLOTS, 256, 0, st>>>(V, seed_key, best_key);
  return cudaGetLastError();
}

// Candidate work items of the verification phase: those of this shard / direction whose largest
// bound reaches the running best (seeded by spr_seed_kernel).  Warp-aggregated append.
__global__ void __launch_bounds__(256)
spr_select_items_kernel(const SprBoundLaunch B, const int n_wg_local, const long long n_items, const int n_yaw,
                        const unsigned long long *__restrict__ best_key, uint32_t *__restrict__ items, uint32_t *count) {
  const long long bc = (long long)(*best_key >> SPR_KEY_IDX_BITS) - 1;
  const uint32_t tau = bc > 0 ? (uint32_t)bc : 0u;
  const uint32_t n_wg_total = B.n_chunks_total / SPR_WARP_CHUNKS;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n_items; i0 += stride) {
    const long long i = i0 + lane;
    bool take = false;
    if (i < n_items) {
      const int a = (int)(i / n_wg_local);
      const int wg = spr_shard_group((int)(i % n_wg_local), B.shard_index, B.shard_count);
      take = B.item_ub[(size_t)a * n_wg_total + B.chunk_begin / SPR_WARP_CHUNKS + (uint32_t)wg] >= tau;
    }
    const uint32_t m = __ballot_sync(SPR_FULL, take);
    if (m) {
      uint32_t base = 0u;
      if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(m));
      base = __shfl_sync(SPR_FULL, base, 0);
      if (take) items[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    }
  }
}

cudaError_t spr_launch_select_items(const SprView &V, const SprBoundLaunch &B, const unsigned long long *best_key,
                                    uint32_t *items, uint32_t *count, int sm_count, cudaStream_t st) {
  if (B.chunk_end <= B.chunk_begin || V.n_yaw <= 0) return cudaSuccess;
  const int n_wg = (int)((B.chunk_end - B.chunk_begin) / SPR_WARP_CHUNKS);
  const int sc = B.shard_count > 1 ? B.shard_count : 1;
  const int si = B.shard_count > 1 ? B.shard_index : 0;
  const int n_wg_local = spr_shard_local_groups(n_wg, si, sc);
  if (n_wg_local <= 0) return cudaSuccess;
  SprBoundLaunch B2 = B;
  B2.shard_index = si;
  B2.shard_count = sc;
  const long long n_items = (long long)n_wg_local * V.n_yaw;
  const long long want = (n_items + 255) / 256;
  spr_select_items_kernel<<<(int)(want < sm_count * 4 ? want : sm_count * 4), 256, 0, st>>>(B2, n_wg_local, n_items, V.n_yaw, best_key, items, count);
  return cudaGetLastError();
}

// Candidate double groups for the refinement: either of the two groups has a bound >= the running best.
__global__ void __launch_bounds__(256)
spr_select_dgroups_kernel(const SprBoundLaunch B, const int n_dg_local, const long long n_items,
                          const unsigned long long *__restrict__ best_key, uint32_t *__restrict__ items, uint32_t *count) {
  const long long bc = (long long)(*best_key >> SPR_KEY_IDX_BITS) - 1;
  const uint32_t tau = bc > 0 ? (uint32_t)bc : 0u;
  const uint32_t n_wg_total = B.n_chunks_total / SPR_WARP_CHUNKS;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n_items; i0 += stride) {
    const long long i = i0 + lane;
    bool take = false;
    if (i < n_items) {
      const int a = (int)(i / n_dg_local);
      const int dg = B.shard_index + (int)(i % n_dg_local) * B.shard_count;
      const uint32_t *ub = B.item_ub + (size_t)a * n_wg_total + B.chunk_begin / SPR_WARP_CHUNKS + 2u * (uint32_t)dg;
      take = ub[0] >= tau || ub[1] >= tau;
    }
    const uint32_t m = __ballot_sync(SPR_FULL, take);
    if (m) {
      uint32_t base = 0u;
      if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(m));
      base = __shfl_sync(SPR_FULL, base, 0);
      if (take) items[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    }
  }
}

cudaError_t spr_launch_select_dgroups(const SprView &V, const SprBoundLaunch &B, const unsigned long long *best_key,
                                      uint32_t *items, uint32_t *count, int sm_count, cudaStream_t st) {
  if (B.chunk_end <= B.chunk_begin || V.n_yaw <= 0) return cudaSuccess;
  const int n_wg = (int)((B.chunk_end - B.chunk_begin) / SPR_WARP_CHUNKS);
  const int sc = B.shard_count > 1 ? B.shard_count : 1;
  const int si = B.shard_count > 1 ? B.shard_index : 0;
  const int n_dg_local = spr_shard_local_groups(n_wg, si, sc) / 2;
  if (n_dg_local <= 0) return cudaSuccess;
  SprBoundLaunch B2 = B;
  B2.shard_index = si;
  B2.shard_count = sc;
  const long long n_items = (long long)n_dg_local * V.n_yaw;
  const long long want = (n_items + 255) / 256;
  spr_select_dgroups_kernel<<<(int)(want < sm_count * 4 ? want : sm_count * 4), 256, 0, st>>>(B2, n_dg_local, n_items, best_key, items, count);
  return cudaGetLastError();
}

// Half-cell variants of the occupancy bitmaps, built on the device.  Variant (sa, sb) of a cell is
// marked when a landmark's match disc touches the (slightly dilated) half-cell box
// [n + s/2, n + (s+1)/2] on both axes -- the same predicate as spr::build_ref_marks on a quarter
// of the cell, so a probe that knows the half-cell its point falls in gets a tighter bound.
__global__ void __launch_bounds__(256)
spr_variant_clear_kernel(uint4 *__restrict__ buf, size_t n16, const uint32_t *__restrict__ counts, uint32_t min_count) {
  if (counts[0] < min_count && counts[1] < min_count) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(128)
spr_variant_mark_kernel(const __grid_constant__ SprView V, uint32_t *__restrict__ vb, const double *__restrict__ ref7,
                        const int32_t *__restrict__ lab_of, int n_ref, double cell, double rc, double rc2,
                        const uint32_t *__restrict__ counts, uint32_t min_count) {
  if (counts[0] < min_count && counts[1] < min_count) return;
  const SprGrid &G = V.grid;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ref; i += gridDim.x * blockDim.x) {
    const int l = lab_of[i];
    if (l < 0) continue;
    const double ux = (ref7[7 * (size_t)i + 1] - G.g0x) / cell, uy = (ref7[7 * (size_t)i + 2] - G.g0y) / cell;
    const int x0 = (int)floor(ux - rc), x1 = (int)floor(ux + rc), y0 = (int)floor(uy - rc), y1 = (int)floor(uy + rc);
    uint32_t *pl0 = vb + 4 * ((size_t)l * G.label_stride);
    uint32_t *pl1 = pl0 + 4 * (size_t)G.plane_words[0];
    for (int nx = x0; nx <= x1; nx++)
      for (int ny = y0; ny <= y1; ny++) {
        if (nx < 0 || nx >= G.GX || ny < 0 || ny >= G.GY) continue;  // cannot happen (grid margins), kept as a guard
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const int sx = v >> 1, sy = v & 1;
          const double bx0 = (double)nx + 0.5 * sx, by0 = (double)ny + 0.5 * sy;
          const double dx = fmax(fmax(bx0 - ux, 0.0), ux - (bx0 + 0.5));
          const double dy = fmax(fmax(by0 - uy, 0.0), uy - (by0 + 0.5));
          if (dx * dx + dy * dy > rc2) continue;
          // plane dir 0: across = x, along = y -> variant 2 * sx + sy; plane dir 1: across = y, along = x -> 2 * sy + sx
          atomicOr(pl0 + (size_t)(2 * sx + sy) * G.plane_words[0] + (size_t)(nx + 1) * G.W[0] + ((uint32_t)(ny + 32) >> 5),
                   1u << ((ny + 32) & 31));
          atomicOr(pl1 + (size_t)(2 * sy + sx) * G.plane_words[1] + (size_t)(ny + 1) * G.W[1] + ((uint32_t)(nx + 32) >> 5),
                   1u << ((nx + 32) & 31));
        }
      }
  }
}

cudaError_t spr_launch_variant_planes(const SprView &V, uint32_t *vbuf, size_t words, const double *ref7, const int32_t *lab_of,
                                      int n_ref, double cell, double rc, double rc2, const uint32_t *counts, uint32_t min_count,
                                      int sm_count, cudaStream_t st) {
  const size_t n16 = words / 4;
  const long long want = (long long)((n16 + 255) / 256);
  spr_variant_clear_kernel<<<(int)(want < sm_count * 8 ? (want > 0 ? want : 1) : sm_count * 8), 256, 0, st>>>(
      reinterpret_cast<uint4 *>(vbuf), n16, counts, min_count);
  if (n_ref > 0)
    spr_variant_mark_kernel<<<(n_ref + 127) / 128, 128, 0, st>>>(V, vbuf + 4, ref7, lab_of, n_ref, cell, rc, rc2, counts, min_count);
  return cudaGetLastError();
}

int spr_bound_planes(int nqp) { return nqp < 4096 ? 12 : 16; }

// Splits the planes of direction `dir` for the bound launches: as many whole label planes per
// launch as fit in shared memory, or -- when one plane does not fit -- one label per launch and
// row bands (multiples of 4 rows: 16-byte aligned bulk copies).  band_rows == 0: planes stay in
// global memory (a single row does not fit).
// smem_budget: 0 = the whole shared memory of an SM; the handle passes a smaller value (read once
// from SLIDE_PR_BOUND_SMEM at slide_pr_create, a test hook) to shrink the planner's budget and drop
// the query-count threshold, so that small maps exercise the label-batch / row-band / in-place paths
void spr_bound_plan(const SprView &V, uint32_t dir, int n_active, size_t smem_budget, int *labels_per_launch, uint32_t *band_rows) {
  const size_t W4 = (size_t)V.grid.W[dir] * 4, R = (size_t)V.grid.R[dir];
  size_t limit = SPR_SMEM_LIMIT - 64;
  int min_nqp = 2048;
  if (smem_budget > 0 && smem_budget < limit) { limit = smem_budget; min_nqp = 0; }
  if ((R + 1) * W4 + 64 <= limit) {
    int per = (int)((limit - 48) / ((R + 1) * W4 + 16));
    if (per > SPR_BOUND_MAX_LABELS) per = SPR_BOUND_MAX_LABELS;
    if (per > n_active) per = n_active;
    // spread the labels evenly over the launches (5 labels, 4 fit -> 3 + 2)
    const int launches = (n_active + per - 1) / per;
    *labels_per_launch = (n_active + launches - 1) / launches;
    *band_rows = (uint32_t)R;
    return;
  }
  *labels_per_launch = 1;
  size_t max_rows = limit / W4;
  // Row bands cost one launch (and one pass over the counters of every work item) per band: they
  // pay off when a work item has many query landmarks to probe; small query maps against a large
  // reference map (streaming submap queries) read the planes in place through L1 / L2 instead.
  if (max_rows < 9 || V.nqp < min_nqp) { *band_rows = 0; *labels_per_launch = SPR_BOUND_MAX_LABELS; return; }
  max_rows = (max_rows - 1) & ~(size_t)3;
  const size_t bands = (R + max_rows - 1) / max_rows;
  *band_rows = (uint32_t)((((R + bands - 1) / bands) + 3) & ~(size_t)3);
}

// Row bands for the refinement launches (four variant planes of one label staged per launch);
// 0: read the variant planes in place.
uint32_t spr_refine_band_rows(const SprView &V, uint32_t dir, size_t smem_budget) {
  const size_t W4 = (size_t)V.grid.W[dir] * 4, R = (size_t)V.grid.R[dir];
  size_t limit = SPR_SMEM_LIMIT - 64;
  int min_nqp = 2048;
  if (smem_budget > 0 && smem_budget < limit) { limit = smem_budget; min_nqp = 0; }  // as in spr_bound_plan
  if (V.nqp < min_nqp) return 0;  // few query landmarks per work item: not worth one launch per band
  size_t max_rows = (limit - 48) / (4 * (W4 + 4));
  if (max_rows < 9) return 0;
  max_rows = (max_rows - 1) & ~(size_t)3;
  if (max_rows >= R) return (uint32_t)R;
  const size_t bands = (R + max_rows - 1) / max_rows;
  return (uint32_t)((((R + bands - 1) / bands) + 3) & ~(size_t)3);
}

template <int PLANES>
static cudaError_t spb_launch(const SprView &V, const SprBoundLaunch &B, int n_dg_local, long long n_items, int sm_count,
                              cudaStream_t st) {
  const size_t BW4 = (((size_t)(B.row_end - B.row_begin + 1) * (size_t)V.grid.W[B.dir] + 3) & ~(size_t)3) * 4;
  const size_t n_staged = B.cand_items ? 4 : (size_t)B.n_labels;  // refinement: the four variants of one label
  const size_t smem = 16 + n_staged * BW4 + 16 + 16;  // 4 zero words in front and behind + the mbarrier
  if (B.cand_items && B.row_end > B.row_begin && smem <= SPR_SMEM_LIMIT && B.n_labels == 1) {
    cudaError_t e = cudaFuncSetAttribute(spr_bound_lattice_kernel<PLANES, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long want = (n_items + SPB_THREADS / 32 - 1) / (SPB_THREADS / 32);
    spr_bound_lattice_kernel<PLANES, true, true><<<(int)(want < sm_count ? want : sm_count), SPB_THREADS, smem, st>>>(V, B, n_dg_local, n_items);
  } else if (!B.cand_items && B.row_end > B.row_begin && smem <= SPR_SMEM_LIMIT) {
    cudaError_t e = cudaFuncSetAttribute(spr_bound_lattice_kernel<PLANES, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long want = (n_items + SPB_THREADS / 32 - 1) / (SPB_THREADS / 32);
    spr_bound_lattice_kernel<PLANES, true, false><<<(int)(want < sm_count ? want : sm_count), SPB_THREADS, smem, st>>>(V, B, n_dg_local, n_items);
  } else {
    const long long want = (n_items + 7) / 8, cap = (long long)sm_count * SPB_GLOBAL_MINB;
    if (B.cand_items)
      spr_bound_lattice_kernel<PLANES, false, true><<<(int)(want < cap ? want : cap), 256, 0, st>>>(V, B, n_dg_local, n_items);
    else
      spr_bound_lattice_kernel<PLANES, false, false><<<(int)(want < cap ? want : cap), 256, 0, st>>>(V, B, n_dg_local, n_items);
  }
  return cudaGetLastError();
}

cudaError_t spr_launch_bound_lattice(const SprView &V, const SprBoundLaunch &B, int n_planes, int sm_count, cudaStream_t st,
                                     int *n_launches) {
  if (B.chunk_end <= B.chunk_begin || V.n_yaw <= 0 || B.n_labels <= 0) return cudaSuccess;
  const int n_wg = (int)((B.chunk_end - B.chunk_begin) / SPR_WARP_CHUNKS);
  const int sc = B.shard_count > 1 ? B.shard_count : 1;
  const int si = B.shard_count > 1 ? B.shard_index : 0;
  const int n_dg_local = spr_shard_local_groups(n_wg, si, sc) / 2;
  if (n_dg_local <= 0) return cudaSuccess;
  SprBoundLaunch B2 = B;
  B2.shard_index = si;
  B2.shard_count = sc;
  const long long n_items = (long long)n_dg_local * V.n_yaw;
  if (n_launches) (*n_launches)++;
  return n_planes == 12 ? spb_launch<12>(V, B2, n_dg_local, n_items, sm_count, st)
                        : spb_launch<16>(V, B2, n_dg_local, n_items, sm_count, st);
}
