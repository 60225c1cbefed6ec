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
__global__ void __launch_bounds__(SPJ_THREADS, 4)
spr_join_score_kernel(const __grid_constant__ SprJoinView V, const __grid_constant__ SprJoinLaunch K,
                      const uint32_t n_blocks_local, const unsigned long long n_items) {
  __shared__ uint32_t s_tile[4 * SPJ_MAX_WORDS];
  __shared__ uint32_t s_tot[SPJ_MAX_SLOTS / 2];    // u16 totals, two per word
  __shared__ uint16_t s_vis[SPJ_SEG_GROUPS];
  __shared__ uint32_t s_nvis;
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_red[SPJ_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (;;) {
    if (tid == 0) s_item = atomicAdd(K.work_counter, 1ull);
    __syncthreads();   // also: the previous item's scan has finished with s_tot / s_red
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const int a = (int)(item / n_blocks_local);
    const uint32_t b = (uint32_t)K.shard_index + (uint32_t)(item % n_blocks_local) * (uint32_t)K.shard_count;
    const SprJoinBlock blk = V.blocks[b];
    const SpjBlock B = spj_block(V, blk);
    const int n_slots = B.nx * B.ny;
    const int n_words = ((B.nx >> 1) + 1) * B.nwy;
    for (int w = tid; w < n_words; w += SPJ_THREADS) {
      s_tile[w] = 0u; s_tile[SPJ_MAX_WORDS + w] = 0u; s_tile[2 * SPJ_MAX_WORDS + w] = 0u; s_tile[3 * SPJ_MAX_WORDS + w] = 0u;
    }
    for (int w = tid; w < (n_slots + 1) / 2; w += SPJ_THREADS) s_tot[w] = 0u;

    const SprJoinBox *gb = V.gbox + (size_t)a * (size_t)V.n_groups;
    const double2 *qr = reinterpret_cast<const double2 *>(V.qrot) + (size_t)a * (size_t)V.nqp;

    for (int seg0 = 0; seg0 < V.n_groups; seg0 += SPJ_SEG_GROUPS) {
      if (tid == 0) s_nvis = 0u;
      __syncthreads();   // s_tile / s_tot zeroed (first segment); s_vis free
      // groups of the segment that some translation of the block brings over their label's landmarks
#pragma unroll
      for (int k = 0; k < SPJ_SEG_GROUPS / SPJ_THREADS; k++) {
        const int g = seg0 + k * SPJ_THREADS + tid;
        bool vis = false;
        if (g < V.n_groups) {
          const float4 bx = __ldg(reinterpret_cast<const float4 *>(gb + g));
          const SprJoinBox box = {bx.x, bx.y, bx.z, bx.w};
          vis = spj_visible(V, B, box, V.labelbox + 4 * (size_t)__ldg(V.glabel + g));
        }
        const uint32_t m = __ballot_sync(SPJ_FULL, vis);
        uint32_t base = 0u;
        if (lane == 0 && m) base = atomicAdd(&s_nvis, (uint32_t)__popc(m));
        base = __shfl_sync(SPJ_FULL, base, 0);
        if (vis) s_vis[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = (uint16_t)(k * SPJ_THREADS + tid);
      }
      __syncthreads();
      const int nvis = (int)s_nvis;
      // rounds of SPJ_THREADS query landmarks: every landmark adds at most 1 to a counter, so the u8
      // counters of the micro-tiles cannot wrap before they are folded into the totals
      for (int gbase = 0; gbase < nvis; gbase += SPJ_THREADS / SPR_QGROUP) {
        const int gi = gbase + (tid >> 3);
        bool voted = false;
        if (gi < nvis) {
          const int g = seg0 + (int)s_vis[gi];
          const int js = g * SPR_QGROUP + (tid & 7);
          const double2 q = __ldg(qr + js);
          if (q.x == q.x) {   // not a padding entry
            const double *qd = V.qdims + 3 * (size_t)js;
            const double qdl[3] = {__ldg(qd), __ldg(qd + 1), __ldg(qd + 2)};
            voted = spj_vote(V, B, __ldg(V.glabel + g), q.x, q.y, qdl, s_tile);
          }
        }
        if (!__syncthreads_or(voted)) continue;
        // fold: the four samples of micro-tile (m, n) of array 0 collect their bytes from all four arrays
        for (int w = tid; w < n_words; w += SPJ_THREADS) spj_fold(s_tile, w, B, reinterpret_cast<uint16_t *>(s_tot));
        __syncthreads();
        for (int w = tid; w < n_words; w += SPJ_THREADS) {
          s_tile[w] = 0u; s_tile[SPJ_MAX_WORDS + w] = 0u; s_tile[2 * SPJ_MAX_WORDS + w] = 0u; s_tile[3 * SPJ_MAX_WORDS + w] = 0u;
        }
        __syncthreads();
      }
    }
    if (V.n_groups <= 0) __syncthreads();   // the zeroed totals

    // ---- scan: slots of the block inside the requested slice of ordinals; max count, smallest ordinal
    int s_lo, s_hi;
    spj_slice(blk, K.ord_begin, K.ord_end, &s_lo, &s_hi);
    const uint16_t *tot = reinterpret_cast<const uint16_t *>(s_tot);
    uint32_t best = 0u;   // (count + 1) << 12 | (4095 - slot): max count, then smallest slot = smallest ordinal
    for (int s = s_lo + tid; s < s_hi; s += SPJ_THREADS) {
      const uint32_t c = tot[s];
      best = max(best, ((c + 1u) << 12) | (uint32_t)(SPJ_MAX_SLOTS - 1 - s));
      if (K.counts_out) {
        const int i = s / B.ny, j = s - i * B.ny;
        const unsigned long long ord = (unsigned long long)blk.ord0 + (unsigned long long)i * blk.row_stride + (unsigned long long)j;
        K.counts_out[(ord - K.ord_begin) * (unsigned long long)V.n_yaw + (unsigned long long)a] = (int32_t)c;
      }
    }
    best = __reduce_max_sync(SPJ_FULL, best);
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int w = 1; w < SPJ_WARPS; w++) best = max(best, s_red[w]);
      if (best) {
        const int s = SPJ_MAX_SLOTS - 1 - (int)(best & (SPJ_MAX_SLOTS - 1));
        const uint32_t c = (best >> 12) - 1u;
        const int i = s / B.ny, j = s - i * B.ny;
        const unsigned long long ord = (unsigned long long)blk.ord0 + (unsigned long long)i * blk.row_stride + (unsigned long long)j;
        atomicMax(K.best_key, spr_make_key(c, ord * (unsigned long long)V.n_yaw + (unsigned long long)a));
      }
    }
  }
}

cudaError_t spr_launch_join_score(const SprJoinView &V, const SprJoinLaunch &K, int sm_count, cudaStream_t st) {
  const int sc = K.shard_count > 1 ? K.shard_count : 1, si = K.shard_count > 1 ? K.shard_index : 0;
  const uint32_t n_local = V.n_blocks > (uint32_t)si ? (V.n_blocks - (uint32_t)si + (uint32_t)sc - 1) / (uint32_t)sc : 0u;
  const unsigned long long n_items = (unsigned long long)n_local * (unsigned long long)(V.n_yaw > 0 ? V.n_yaw : 0);
  if (n_items == 0) return cudaSuccess;
  SprJoinLaunch L = K;
  L.shard_index = si; L.shard_count = sc;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spr_join_score_kernel, SPJ_THREADS, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const unsigned long long cap = (unsigned long long)sm_count * (unsigned long long)per_sm;
  spr_join_score_kernel<<<(unsigned)(n_items < cap ? n_items : cap), SPJ_THREADS, 0, st>>>(V, L, n_local, n_items);
  return cudaGetLastError();
}
