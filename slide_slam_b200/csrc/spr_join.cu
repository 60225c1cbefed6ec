// spr_join.cu -- pair-join scorer (see spr_join.h): exact inlier counts of every lattice hypothesis,
// accumulated from the (query, reference) landmark pairs that can match.  sm_100a.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "spr_core.h"
#include "spr_join.h"
#include "spr_join_core.h"

#define SPJ_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// rotated query coordinates (exact fp64, PR.cpp:246-258) and the bounding box of every query
// group under every yaw candidate, rounded outward to float
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spr_join_rotate_kernel(const SprJoinView V, double *__restrict__ qrot, SprJoinBox *__restrict__ gbox) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)V.n_yaw * V.n_groups) return;
  const int a = (int)(idx / V.n_groups), g = (int)(idx % V.n_groups);
  const double c = V.cs[2 * a], s = V.cs[2 * a + 1];
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  float4 box = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);  // x0, x1, y0, y1
#pragma unroll
  for (int k = 0; k < SPR_QGROUP; k++) {
    const int js = g * SPR_QGROUP + k;
    double rx = qnan, ry = qnan;
    if (V.qlabel[js] >= 0) {
      spr_rotate(c, s, V.qxy[2 * (size_t)js], V.qxy[2 * (size_t)js + 1], &rx, &ry);
      box.x = fminf(box.x, __double2float_rd(rx));
      box.y = fmaxf(box.y, __double2float_ru(rx));
      box.z = fminf(box.z, __double2float_rd(ry));
      box.w = fmaxf(box.w, __double2float_ru(ry));
    }
    const size_t qi = (size_t)a * (size_t)V.nqp + (size_t)js;
    reinterpret_cast<double2 *>(qrot)[qi] = make_double2(rx, ry);
  }
  *reinterpret_cast<float4 *>(gbox + idx) = box;
}

cudaError_t spr_launch_join_rotate(const SprJoinView &V, double *qrot, SprJoinBox *gbox, cudaStream_t st) {
  const long long n = (long long)V.n_yaw * V.n_groups;
  if (n <= 0) return cudaSuccess;
  spr_join_rotate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(V, qrot, gbox);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// the scorer
// ---------------------------------------------------------------------------------------------
// A query landmark as the lanes of its warp see it while its candidate pairs are processed.
struct SpjQuery { double rx, ry, d1, d2, d3; int32_t label, pad; };

// 32 query landmarks (a "quad" of four visible groups, one landmark per lane) against block B.  Every lane
// looks up the record ranges of the coarse-cell bands its landmark can reach; a prefix sum over all ranges of
// the warp turns them into ONE flat sequence of candidate (landmark, record) pairs, which the lanes then walk
// side by side (a lane finds the range of its candidate by binary search in the prefix table): filter (can
// the pair match under a translation of the block at all?), compaction of the survivors into the pair list,
// exact tests and counter updates -- the lanes stay busy although the landmarks have different numbers of
// candidates.
__device__ __forceinline__ void spj_pairs(const SprJoinView &V, const SpjBlock &B, const SprJoinRef *rec, const SpjQuery *sq,
                                          const uint32_t *list, uint32_t n_pairs, int lane, uint32_t *tile) {
  for (uint32_t w = (uint32_t)lane; w < n_pairs; w += 32) {
    const uint32_t e = list[w];
    const SpjQuery &q = sq[e & 31u];
    spj_pair(V, B, q.rx, q.ry, &q.d1, rec + (e >> 5), tile);
  }
}

__device__ __forceinline__ void spj_quad(const SprJoinView &V, const SpjBlock &B, const double2 *__restrict__ qr, int g, int lane,
                                         bool active, SpjQuery *sq, uint32_t *list, uint32_t *pre, uint32_t *beg, uint32_t *tile) {
  int b0 = 0, b1 = -1, a0 = 0, a1 = -1;
  const uint32_t *cstart = V.cell_start[B.dir];
  if (active) {
    const int js = g * SPR_QGROUP + (lane & 7);
    const double2 q = __ldg(qr + js);
    if (q.x == q.x) {   // not a padding entry
      const double *qd = V.qdims + 3 * (size_t)js;
      const int l = __ldg(V.glabel + g);
      sq[lane].rx = q.x; sq[lane].ry = q.y;
      sq[lane].d1 = __ldg(qd); sq[lane].d2 = __ldg(qd + 1); sq[lane].d3 = __ldg(qd + 2);
      sq[lane].label = l;
      spj_cells(V, B, q.x, q.y, &b0, &b1, &a0, &a1);
      cstart += (size_t)l * (size_t)(V.ncx * V.ncy);
    }
  }
  const int pitch = B.dir ? V.ncx : V.ncy;
  const SprJoinRef *rec = V.rec[B.dir];
  const double2 *xy = reinterpret_cast<const double2 *>(V.xy[B.dir]);
  uint32_t n_pairs = 0u;
  for (int bb = b0; __any_sync(SPJ_FULL, bb <= b1); bb += SPJ_BANDS) {
    // record ranges of the lane's next SPJ_BANDS bands, and their places in the warp's flat candidate sequence
    uint32_t bg[SPJ_BANDS], cnt[SPJ_BANDS], t = 0u;
#pragma unroll
    for (int k = 0; k < SPJ_BANDS; k++) {
      uint32_t s = 0u, e = 0u;
      if (bb + k <= b1) {
        s = __ldg(cstart + (bb + k) * pitch + a0);
        e = __ldg(cstart + (bb + k) * pitch + a1 + 1);
      }
      bg[k] = s; cnt[k] = e - s; t += e - s;
    }
    uint32_t incl = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t u = __shfl_up_sync(SPJ_FULL, incl, d);
      if (lane >= d) incl += u;
    }
    const uint32_t total = __shfl_sync(SPJ_FULL, incl, 31);
    uint32_t at = incl - t;
#pragma unroll
    for (int k = 0; k < SPJ_BANDS; k++) {
      pre[lane * SPJ_BANDS + k] = at;
      beg[lane * SPJ_BANDS + k] = bg[k];
      at += cnt[k];
    }
    if (lane == 31) pre[32 * SPJ_BANDS] = total;
    __syncwarp();
    for (uint32_t w0 = 0; w0 < total; w0 += 64) {
      // two candidates per lane and step (w and w + 32): their range searches and coordinate loads are independent of
      // each other, which hides part of the latency of the dependent shared-memory and global loads; consecutive
      // lanes still take consecutive candidates (neighbouring records, coalesced loads)
      const uint32_t wa = w0 + (uint32_t)lane, wb = wa + 32u;
      // -> the last range that starts at or before the candidate (empty ranges share their start with the next one);
      // the search stays inside the table for any w, as pre[32 * SPJ_BANDS] = total
      const uint32_t *ra = pre, *rb = pre;
#pragma unroll
      for (uint32_t step = 16 * SPJ_BANDS; step >= 1; step >>= 1) {
        if (ra[step] <= wa) ra += step;
        if (rb[step] <= wb) rb += step;
      }
      uint32_t ea = 0u, eb = 0u;
      bool oka = false, okb = false;
      double2 pa = make_double2(0.0, 0.0), pb = pa;
      // entry = (record << 5) | lane of the query landmark; beg = pre + 32 * SPJ_BANDS + 1
      if (wa < total) { ea = ((ra[32 * SPJ_BANDS + 1] + (wa - ra[0])) << 5) | ((uint32_t)(ra - pre) / SPJ_BANDS); pa = __ldg(xy + (ea >> 5)); }
      if (wb < total) { eb = ((rb[32 * SPJ_BANDS + 1] + (wb - rb[0])) << 5) | ((uint32_t)(rb - pre) / SPJ_BANDS); pb = __ldg(xy + (eb >> 5)); }
      if (wa < total) { const SpjQuery &q = sq[ea & 31u]; oka = spj_near(B, q.rx, q.ry, pa.x, pa.y); }
      if (wb < total) { const SpjQuery &q = sq[eb & 31u]; okb = spj_near(B, q.rx, q.ry, pb.x, pb.y); }
      const uint32_t ma = __ballot_sync(SPJ_FULL, oka), mb = __ballot_sync(SPJ_FULL, okb);
      const uint32_t below = (1u << lane) - 1u, na = (uint32_t)__popc(ma);
      if (oka) list[n_pairs + (uint32_t)__popc(ma & below)] = ea;
      if (okb) list[n_pairs + na + (uint32_t)__popc(mb & below)] = eb;
      n_pairs += na + (uint32_t)__popc(mb);
      if (n_pairs + 64u > (uint32_t)SPJ_LIST) {   // the pair list may not take another step: exact tests and counter updates now
        __syncwarp();
        spj_pairs(V, B, rec, sq, list, n_pairs, lane, tile);
        __syncwarp();
        n_pairs = 0u;
      }
    }
    __syncwarp();   // the range tables are rewritten by the next pass
  }
  __syncwarp();
  spj_pairs(V, B, rec, sq, list, n_pairs, lane, tile);
  __syncwarp();
}

// WARPS x 32 threads per CTA, CTAS resident CTAs per SM: 8 x 4 for ordinary query maps; 4 x 7 when the query map is small
// (few quads per work item: more items in flight per SM pay more than wide CTAs).
template <int WARPS, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
spr_join_score_kernel(const __grid_constant__ SprJoinView V, const __grid_constant__ SprJoinLaunch K,
                      const uint32_t n_blocks_local, const unsigned long long n_items) {
  // dynamic shared memory (more than the 48 KB a static allocation may take): counters, staged query
  // landmarks, per-warp work lists, visible groups
  extern __shared__ __align__(16) unsigned char spj_smem[];
  SpjQuery (*s_q)[32] = reinterpret_cast<SpjQuery (*)[32]>(spj_smem);
  uint32_t *s_tile = reinterpret_cast<uint32_t *>(spj_smem + WARPS * 32 * sizeof(SpjQuery));
  uint32_t (*s_list)[SPJ_LIST] = reinterpret_cast<uint32_t (*)[SPJ_LIST]>(s_tile + SPJ_TILE_WORDS);
  uint32_t (*s_rng)[SPJ_RANGE_WORDS] = reinterpret_cast<uint32_t (*)[SPJ_RANGE_WORDS]>(s_tile + SPJ_TILE_WORDS + WARPS * SPJ_LIST);
  uint16_t *s_vis = reinterpret_cast<uint16_t *>(s_tile + SPJ_TILE_WORDS + WARPS * (SPJ_LIST + SPJ_RANGE_WORDS));
  __shared__ uint32_t s_nvis, s_next;
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_red[WARPS];
  constexpr int THREADS = WARPS * 32, SEG_GROUPS = THREADS * 4;   // query groups whose visibility is tested per pass
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (;;) {
    if (tid == 0) s_item = atomicAdd(K.work_counter, 1ull);
    __syncthreads();   // also: the previous item's scan has finished with s_tile / s_red
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const int a = (int)(item / n_blocks_local);
    const uint32_t b = (uint32_t)K.shard_index + (uint32_t)(item % n_blocks_local) * (uint32_t)K.shard_count;
    const SprJoinBlock blk = V.blocks[b];
    const SpjBlock B = spj_block(V, blk);
    for (int w = tid; w < (2 * B.stride + 3) / 4; w += THREADS) reinterpret_cast<uint4 *>(s_tile)[w] = make_uint4(0u, 0u, 0u, 0u);

    const SprJoinBox *gb = V.gbox + (size_t)a * (size_t)V.n_groups;
    const double2 *qr = reinterpret_cast<const double2 *>(V.qrot) + (size_t)a * (size_t)V.nqp;

    for (int seg0 = 0; seg0 < V.n_groups; seg0 += SEG_GROUPS) {
      if (tid == 0) { s_nvis = 0u; s_next = 0u; }
      __syncthreads();   // counters zeroed (first segment); the previous segment's warps are done with s_vis
      // groups of the segment that some translation of the block brings over their label's landmarks
      const int seg_passes = min(SEG_GROUPS / THREADS, (V.n_groups - seg0 + THREADS - 1) / THREADS);
      for (int k = 0; k < seg_passes; k++) {
        const int g = seg0 + k * THREADS + tid;
        bool vis = false;
        if (g < V.n_groups) {
          const float4 bx = __ldg(reinterpret_cast<const float4 *>(gb + g));
          const SprJoinBox box = {bx.x, bx.y, bx.z, bx.w};
          vis = spj_visible(B, box, V.labelbox + 4 * (size_t)__ldg(V.glabel + g));
        }
        const uint32_t m = __ballot_sync(SPJ_FULL, vis);
        uint32_t base = 0u;
        if (lane == 0 && m) base = atomicAdd(&s_nvis, (uint32_t)__popc(m));
        base = __shfl_sync(SPJ_FULL, base, 0);
        if (vis) s_vis[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = (uint16_t)(k * THREADS + tid);
      }
      __syncthreads();
      const uint32_t nvis = s_nvis, n_quads = (nvis + 3u) / 4u;
      // the warps take quads of visible groups (32 query landmarks) until none is left
      for (;;) {
        uint32_t quad = 0u;
        if (lane == 0) quad = atomicAdd(&s_next, 1u);
        quad = __shfl_sync(SPJ_FULL, quad, 0);
        if (quad >= n_quads) break;
        const uint32_t gi = quad * 4u + ((uint32_t)lane >> 3);
        const bool active = gi < nvis;
        spj_quad(V, B, qr, active ? seg0 + (int)s_vis[gi] : 0, lane, active, s_q[warp], s_list[warp], s_rng[warp],
                 s_rng[warp] + 32 * SPJ_BANDS + 1, s_tile);
      }
      __syncthreads();   // every warp is done with the segment's list (and, after the last segment, with its counter updates)
    }
    if (V.n_groups <= 0) __syncthreads();   // the zeroed counters

    // ---- scan: slots of the block inside the requested slice of ordinals; max count, smallest ordinal
    int s_lo, s_hi;
    spj_slice(blk, K.ord_begin, K.ord_end, &s_lo, &s_hi);
    uint32_t best = 0u;   // (count + 1) << SPJ_SLOT_BITS | (SPJ_MAX_SLOTS - 1 - slot): max count, then smallest slot = smallest ordinal
    if (!K.counts_out && s_lo == 0 && s_hi == B.nx * B.ny) {
      // the whole block (the usual case): two samples per step -- word n of array 0 holds samples (2n, 2n + 1) of its
      // row, array 1 contributes the high half of its word n to sample 2n and the low half of word n + 1 to 2n + 1
      const float inv_nwy = 1.0f / (float)B.nwy;
      for (int w = tid; w < B.stride; w += THREADS) {
        const int i = __float2int_rd(((float)w + 0.5f) * inv_nwy), n = w - i * B.nwy, j = 2 * n;   // exact: w < 2^13
        if (j >= B.ny) continue;   // the spare word of an even-length row
        const uint32_t w0 = s_tile[w], w1 = s_tile[B.stride + w];
        const int s = i * B.ny + j;
        const uint32_t c0 = (w0 & 0xffffu) + (w1 >> 16);
        best = max(best, ((c0 + 1u) << SPJ_SLOT_BITS) | (uint32_t)(SPJ_MAX_SLOTS - 1 - s));
        if (j + 1 < B.ny) {
          const uint32_t c1 = (w0 >> 16) + (s_tile[B.stride + w + 1] & 0xffffu);
          best = max(best, ((c1 + 1u) << SPJ_SLOT_BITS) | (uint32_t)(SPJ_MAX_SLOTS - 2 - s));
        }
      }
    } else {
      const float inv_ny = 1.0f / (float)B.ny;
      for (int s = s_lo + tid; s < s_hi; s += THREADS) {
        const int i = __float2int_rd(((float)s + 0.5f) * inv_ny), j = s - i * B.ny;   // exact: s < 2^13
        const uint32_t c = spj_total(s_tile, B, i, j);
        best = max(best, ((c + 1u) << SPJ_SLOT_BITS) | (uint32_t)(SPJ_MAX_SLOTS - 1 - s));
        if (K.counts_out) {
          const unsigned long long ord = (unsigned long long)blk.ord0 + (unsigned long long)i * blk.row_stride + (unsigned long long)j;
          K.counts_out[(ord - K.ord_begin) * (unsigned long long)V.n_yaw + (unsigned long long)a] = (int32_t)c;
        }
      }
    }
    best = __reduce_max_sync(SPJ_FULL, best);
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int w = 1; w < WARPS; w++) best = max(best, s_red[w]);
      if (best) {
        const int s = SPJ_MAX_SLOTS - 1 - (int)(best & (SPJ_MAX_SLOTS - 1));
        const uint32_t c = (best >> SPJ_SLOT_BITS) - 1u;
        const int i = s / B.ny, j = s - i * B.ny;
        const unsigned long long ord = (unsigned long long)blk.ord0 + (unsigned long long)i * blk.row_stride + (unsigned long long)j;
        atomicMax(K.best_key, spr_make_key(c, ord * (unsigned long long)V.n_yaw + (unsigned long long)a));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// explicit hypothesis lists (c, s, x, y) -- e.g. the 2-D Kabsch fits of matched triangles -- scored with the
// MatchMaps predicate (PR.cpp:272-357) through the landmark bins: warp per hypothesis, lanes stride over the
// query landmarks; a landmark counts when ANY reference landmark of its label matches (the reference stops at
// the first one it finds)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spr_join_score_list_kernel(const SprJoinView V, const double *__restrict__ hyps4, long long n, int32_t *__restrict__ counts_out,
                           unsigned long long *best_key) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const size_t n_cells = (size_t)V.ncx * (size_t)V.ncy;
  unsigned long long best = 0ull;
  for (long long h = warp0; h < n; h += n_warps) {
    const double c = hyps4[4 * h], s = hyps4[4 * h + 1], tx = hyps4[4 * h + 2], ty = hyps4[4 * h + 3];
    int cnt = 0;
    for (int js = lane; js < V.nqp; js += 32) {
      const int l = __ldg(V.qlabel + js);
      if (l < 0) continue;  // padding
      double rx, ry;
      spr_rotate(c, s, __ldg(V.qxy + 2 * (size_t)js), __ldg(V.qxy + 2 * (size_t)js + 1), &rx, &ry);
      const double xt = SPR_DADD(rx, tx), yt = SPR_DADD(ry, ty);
      // coarse cells within reach of the transformed landmark (+ the rounding of the test at this magnitude)
      const double reach = V.reach + 1e-12 * (fabs(xt) + fabs(yt));
      double fx0 = floor(SPR_DMUL(SPR_DSUB(SPR_DSUB(xt, reach), V.gx0), V.inv_w)), fx1 = floor(SPR_DMUL(SPR_DSUB(SPR_DADD(xt, reach), V.gx0), V.inv_w));
      double fy0 = floor(SPR_DMUL(SPR_DSUB(SPR_DSUB(yt, reach), V.gy0), V.inv_w)), fy1 = floor(SPR_DMUL(SPR_DSUB(SPR_DADD(yt, reach), V.gy0), V.inv_w));
      fx0 = fmax(fx0, 0.0); fy0 = fmax(fy0, 0.0);
      fx1 = fmin(fx1, (double)(V.ncx - 1)); fy1 = fmin(fy1, (double)(V.ncy - 1));
      if (!(fx0 <= fx1) || !(fy0 <= fy1)) continue;
      const uint32_t *cstart = V.cell_start[0] + (size_t)l * n_cells;
      const double *qd = V.qdims + 3 * (size_t)js;
      const double qdl[3] = {__ldg(qd), __ldg(qd + 1), __ldg(qd + 2)};
      bool hit = false;
      for (int cx = (int)fx0; cx <= (int)fx1 && !hit; cx++) {
        const uint32_t r_end = __ldg(cstart + cx * V.ncy + (int)fy1 + 1);
        for (uint32_t r = __ldg(cstart + cx * V.ncy + (int)fy0); r < r_end && !hit; r++) {
          const SprJoinRef *rec = V.rec[0] + r;
          hit = spr_distance_match(rx, ry, tx, ty, __ldg(&rec->x), __ldg(&rec->y), V.Tstar) &&
                (V.ignore_dim || spr_dimension_match(__ldg(&rec->d1), __ldg(&rec->d2), __ldg(&rec->d3), qdl, V.thr_dim, V.Sstar));
        }
      }
      cnt += hit ? 1 : 0;
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) cnt += __shfl_xor_sync(SPJ_FULL, cnt, dlt);
    if (lane == 0) {
      if (counts_out) counts_out[h] = cnt;
      const unsigned long long key = spr_make_key((uint32_t)cnt, (unsigned long long)h);
      best = key > best ? key : best;
    }
  }
  if (lane == 0 && best != 0ull) atomicMax(best_key, best);
}

cudaError_t spr_launch_join_score_list(const SprJoinView &V, const double *hyps4, long long n, int32_t *counts_out,
                                       unsigned long long *best_key, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long want = (n + 7) / 8;
  const long long cap = (long long)sm_count * 8;
  spr_join_score_list_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(V, hyps4, n, counts_out, best_key);
  return cudaGetLastError();
}

template <int WARPS, int CTAS>
static cudaError_t spj_launch(const SprJoinView &V, const SprJoinLaunch &L, uint32_t n_local, unsigned long long n_items, int sm_count, cudaStream_t st) {
  constexpr size_t smem = WARPS * 32 * sizeof(SpjQuery) + (SPJ_TILE_WORDS + WARPS * (SPJ_LIST + SPJ_RANGE_WORDS)) * sizeof(uint32_t) +
                          WARPS * 32 * 4 * sizeof(uint16_t);
  // the opt-in to more than 48 KB of dynamic shared memory and the residency query are per device: done once
  static int per_sm_of_device[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int per_sm = dev >= 0 && dev < 64 ? per_sm_of_device[dev] : 0;
  if (per_sm == 0) {
    e = cudaFuncSetAttribute(spr_join_score_kernel<WARPS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spr_join_score_kernel<WARPS, CTAS>, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (dev >= 0 && dev < 64) per_sm_of_device[dev] = per_sm;
  }
  const unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)per_sm;
  spr_join_score_kernel<WARPS, CTAS><<<(unsigned)(n_items < cap ? n_items : cap), WARPS * 32, smem, st>>>(V, L, n_local, n_items);
  return cudaGetLastError();
}

cudaError_t spr_launch_join_score(const SprJoinView &V, const SprJoinLaunch &K, int sm_count, cudaStream_t st) {
  const int sc = K.shard_count > 1 ? K.shard_count : 1, si = K.shard_count > 1 ? K.shard_index : 0;
  const uint32_t n_local = V.n_blocks > (uint32_t)si ? (V.n_blocks - (uint32_t)si + (uint32_t)sc - 1) / (uint32_t)sc : 0u;
  const unsigned long long n_items = (unsigned long long)n_local * (unsigned long long)(V.n_yaw > 0 ? V.n_yaw : 0);
  if (n_items == 0) return cudaSuccess;
  SprJoinLaunch L = K;
  L.shard_index = si; L.shard_count = sc;
  // small query maps (at most 16 quads per work item): narrow CTAs, more of them
  if (V.n_groups <= 64) return spj_launch<4, 7>(V, L, n_local, n_items, sm_count, st);
  return spj_launch<SPJ_WARPS, SPJ_MIN_CTAS>(V, L, n_local, n_items, sm_count, st);
}
