// spr_kernels.cu -- sm_100a kernels of the SlideMatch lattice search.
//
// Replaces the five nested loops of PlaceRecognition::MatchMaps (place_recognition.cpp:178-372).
// Work decomposition:
//   * one THREAD owns one chunk = 32 consecutive lattice translations along one axis, for one
//     yaw candidate; a WARP owns 32 chunks (1024 hypotheses) and is the unit of scheduling: warps
//     pull (yaw, 32-chunk) work items from a global counter, there is no block-level barrier;
//   * query landmarks come in groups of 8 (one label, Morton order) with a bounding box per yaw;
//     a group that cannot reach the label's occupied cells from any of the warp's 1024
//     translations is skipped with four integer compares;
//   * for every remaining query landmark the thread reads two words of the label's occupancy
//     bitmap and obtains the 32 hypotheses' filter bits with one funnel shift (spr_probe);
//   * set bits ("filter hits", well under 1 % of the probes) are compacted through a per-warp
//     shared-memory queue (warp prefix-sum over popcounts) and verified 32 at a time in exact,
//     non-fused fp64 against the cell's candidates (spr_verify_cell) -- the reference's own
//     predicate, so every hypothesis gets its exact inlier count;
//   * per-hypothesis counters live in shared memory; the best (count, canonical index) key is
//     reduced with shuffles and one 64-bit atomicMax per warp.
// This is integer/bit and fp64 ALU work on L1/L2-resident data; there is no GEMM in it, so no
// tensor-core path (BASELINE.json north_star).
#include <cuda_runtime.h>
#include <stdint.h>

#include "spr_core.h"
#include "spr_kernels.h"

#ifndef SPR_BLOCK
#define SPR_BLOCK 256   // threads per CTA (8 independent warps)
#endif
#ifndef SPR_MINB
#define SPR_MINB 4      // CTAs per SM the register allocation is tuned for
#endif
#define SPR_WARPS (SPR_BLOCK / 32)
#define SPR_QHALF 4              // queries probed per push
#define SPR_QCAP (32 * SPR_QHALF + 32)  // per-warp hit queue, records (3 words each): residual < 32 + one push
// Per-hypothesis inlier counters, [32 lanes][32 bits] per warp.  Counts are bounded by the number
// of query landmarks, so they are packed two per word (pitch 17 words per lane) unless the query
// map has more than 65535 landmarks (pitch 33 words per lane).  Shared memory is kept small on
// purpose: what the CTAs do not take stays L1 cache for the occupancy bitmaps.
#define SPR_CNT_WORDS(CNT32) ((CNT32) ? 32 * 33 : 32 * 17)
#define SPR_WARP_SMEM(CNT32) (SPR_CNT_WORDS(CNT32) + 3 * SPR_QCAP)  // words per warp

template <bool CNT32> __device__ __forceinline__ void spr_cnt_zero(uint32_t *cnt, int lane) {
  if (CNT32) {
#pragma unroll
    for (int b = 0; b < 32; b++) cnt[lane * 33 + b] = 0u;
  } else {
#pragma unroll
    for (int w = 0; w < 16; w++) cnt[lane * 17 + w] = 0u;
  }
}
template <bool CNT32> __device__ __forceinline__ void spr_cnt_inc(uint32_t *cnt, int owner, int b) {
  if (CNT32) atomicAdd(&cnt[owner * 33 + b], 1u);
  else atomicAdd(&cnt[owner * 17 + (b >> 1)], 1u << ((b & 1) << 4));
}
template <bool CNT32> __device__ __forceinline__ uint32_t spr_cnt_get(const uint32_t *cnt, int lane, int b) {
  if (CNT32) return cnt[lane * 33 + b];
  return (cnt[lane * 17 + (b >> 1)] >> ((b & 1) << 4)) & 0xffffu;
}
#define SPR_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// rotate: one thread per (yaw, query group)
// ---------------------------------------------------------------------------------------------
__global__ void spr_rotate_kernel(SprView V, int32_t *__restrict__ qrotq_xy, int32_t *__restrict__ qrotq_yx,
                                  double *__restrict__ qrot, SprBox *__restrict__ gbox) {
  const long long n = (long long)V.n_yaw * V.n_groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    spr_rotate_group(V.cs, V.qxy, V.qlabel, V.grid, V.nqp, V.n_groups, (int)(i / V.n_groups), (int)(i % V.n_groups),
                     qrotq_xy, qrotq_yx, qrot, gbox);
}

cudaError_t spr_launch_rotate(const SprView &V, int32_t *qrotq_xy, int32_t *qrotq_yx, double *qrot, SprBox *gbox,
                              cudaStream_t st) {
  const long long n = (long long)V.n_yaw * V.n_groups;
  if (n <= 0) return cudaSuccess;
  const int block = 128;
  const long long want = (n + block - 1) / block;
  spr_rotate_kernel<<<(int)(want > 148 * 16 ? 148 * 16 : want), block, 0, st>>>(V, qrotq_xy, qrotq_yx, qrot, gbox);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// lattice scoring
// ---------------------------------------------------------------------------------------------
struct WarpState {
  uint32_t *cnt;    // inlier counters of the warp's 1024 hypotheses (spr_cnt_*)
  uint32_t *q_meta; // [SPR_QCAP] pending records: query index js << 5 | owner lane
  uint32_t *q_code; // [SPR_QCAP] bit address (own plane) of the cell under bit 0 of the probe
  uint32_t *q_mask; // [SPR_QCAP] the probe's hit bits
  int qcount;       // warp-uniform
};

// One queued record = all filter hits of ONE query landmark on ONE chunk (lane): verify its bits
// in exact fp64 and add the survivors to the owner's counters.
template <bool CNT32>
__device__ __forceinline__ void spr_verify_record(const SprView &V, uint32_t *cnt, int l, int a, int js, int owner,
                                                  uint32_t code0, uint32_t H, uint32_t o_dir, uint32_t o_off,
                                                  double o_across, unsigned long long &n_inl) {
  const size_t qi = (size_t)a * (size_t)V.nqp + (size_t)js;
  const double2 r = __ldg(reinterpret_cast<const double2 *>(V.qrot) + qi);
  uint32_t P = spr_verify_mask(V, o_dir, l, code0, H, r.x, r.y, o_across, V.lat + o_off, V.qdims + 3 * (size_t)js);
  n_inl += (unsigned long long)__popc(P);
  while (P) {
    const int b = __ffs(P) - 1;
    P &= P - 1;
    spr_cnt_inc<CNT32>(cnt, owner, b);
  }
}

// Verify up to 32 queued records, one per lane.  Chunk parameters of the owning lane come
// through shuffles; every lane executes the shuffles.
template <bool CNT32>
__device__ __forceinline__ void spr_drain32(const SprView &V, WarpState &ws, int l, int a, int lane, uint32_t dir,
                                            uint32_t along_off, double across, unsigned long long &n_inl) {
  const int n = ws.qcount < 32 ? ws.qcount : 32;
  const bool active = lane < n;
  const int slot = ws.qcount - n + lane;
  const uint32_t meta = active ? ws.q_meta[slot] : 0u;
  const uint32_t code0 = active ? ws.q_code[slot] : 0u;
  const uint32_t H = active ? ws.q_mask[slot] : 0u;
  const int owner = meta & 31;
  const uint32_t o_dir = __shfl_sync(SPR_FULL, dir, owner);
  const uint32_t o_off = __shfl_sync(SPR_FULL, along_off, owner);
  const double o_across = __shfl_sync(SPR_FULL, across, owner);
  if (active) spr_verify_record<CNT32>(V, ws.cnt, l, a, (int)(meta >> 5), owner, code0, H, o_dir, o_off, o_across, n_inl);
  ws.qcount -= n;
  __syncwarp();
}

template <int VARIANT, bool WRITE_COUNTS, bool STATS, bool CNT32>
__global__ void __launch_bounds__(SPR_BLOCK, SPR_MINB)
spr_score_lattice_kernel(const SprView V, const SprLaunch K, const int n_wg_local, const long long n_items) {
  extern __shared__ uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpState ws;
  ws.cnt = smem + warp * SPR_WARP_SMEM(CNT32);
  ws.q_meta = ws.cnt + SPR_CNT_WORDS(CNT32);
  ws.q_code = ws.q_meta + SPR_QCAP;
  ws.q_mask = ws.q_code + SPR_QCAP;
  ws.qcount = 0;
  const SprGrid &G = V.grid;
  const int32_t F = G.F;
  unsigned long long best = 0ull, n_hits = 0ull, n_inl = 0ull, n_probed = 0ull, n_skipped = 0ull;

  for (;;) {
    // per-warp dynamic scheduling, yaw-major: warps running at the same time share qrotq[a][*]
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(K.work_counter, 1ull);
    item = __shfl_sync(SPR_FULL, item, 0);
    if (item >= n_items) break;
    const int a = (int)(item / n_wg_local);
    const int wg = K.shard_index + (int)(item % n_wg_local) * K.shard_count;
    const uint32_t cidx = K.chunk_begin + (uint32_t)wg * SPR_WARP_CHUNKS + (uint32_t)lane;
    double across = 0.0;
    uint32_t along_off = 0u, valid = 0u, ord_base = 0u, ord_stride = 0u, dir = 0u;
    if (cidx < K.chunk_end) {
      const SprChunk ch = V.chunks[cidx];
      across = ch.across; along_off = ch.along_off; valid = ch.valid;
      ord_base = ch.ord_base; ord_stride = ch.ord_stride; dir = ch.dir;
    }
    spr_cnt_zero<CNT32>(ws.cnt, lane);
    const int32_t aq0 = spr_fx(across, G.S);
    const int32_t bq0 = spr_fx(__ldg(V.lat + along_off), G.S);
    const int32_t aqb = spr_bias_across(aq0, F), bqb = spr_bias_along(bq0, F);
    // patch window of the warp's 1024 translations (unbiased fixed point)
    const int32_t big = 1 << 30;
    const bool live = valid != 0u;
    const int32_t lx0 = dir ? bq0 : aq0, lx1 = dir ? bq0 + (32 << F) : aq0;
    const int32_t ly0 = dir ? aq0 : bq0, ly1 = dir ? aq0 : bq0 + (32 << F);
    const int32_t X0 = __reduce_min_sync(SPR_FULL, live ? lx0 : big), X1 = __reduce_max_sync(SPR_FULL, live ? lx1 : -big);
    const int32_t Y0 = __reduce_min_sync(SPR_FULL, live ? ly0 : big), Y1 = __reduce_max_sync(SPR_FULL, live ? ly1 : -big);
    __syncwarp();
    const uint32_t W = (uint32_t)G.W[dir], Rm1 = (uint32_t)G.R[dir] - 1u, maxbit = (uint32_t)G.maxbit[dir];
    const int2 *__restrict__ qa2 =
        reinterpret_cast<const int2 *>((dir ? V.qrotq_yx : V.qrotq_xy) + 2 * (size_t)a * (size_t)V.nqp);
    const int4 *__restrict__ qa = reinterpret_cast<const int4 *>(qa2);

    if (X0 <= X1) {  // at least one live lane
      for (int l = 0; l < V.n_labels; l++) {
        // this lane's bitmap plane as a materialised 64-bit pointer: one IMAD.WIDE per probe
        const uint32_t *plane = V.bitmap + ((size_t)l * G.label_stride + (dir ? G.plane_words[0] : 0u));
        asm volatile("" : "+l"(plane));
        const SprBox lb = V.labelbox[l];
        // a group is visible iff gx1 > tx_lo && gx0 < tx_hi && gy1 > ty_lo && gy0 < ty_hi
        const int32_t tx_lo = lb.x0 - X1, tx_hi = lb.x1 - X0, ty_lo = lb.y0 - Y1, ty_hi = lb.y1 - Y0;
        const int g0 = V.label_gseg[l], g1 = V.label_gseg[l + 1];
        const int4 *gbp = reinterpret_cast<const int4 *>(V.gbox) + ((size_t)a * (size_t)V.n_groups + (size_t)g0);
        const int4 *qgp = qa + (size_t)g0 * (SPR_QGROUP / 2);
        for (int g = g0; g < g1; g++, gbp++, qgp += SPR_QGROUP / 2) {
          const int4 box = __ldg(gbp);  // (x0, x1, y0, y1), same address for the whole warp
          if (!(box.y > tx_lo && box.x < tx_hi && box.w > ty_lo && box.z < ty_hi)) {
            if (STATS) n_skipped++;
            continue;
          }
          if (STATS) n_probed++;
          // the group is probed in two halves of SPR_QHALF queries: at most 32 * SPR_QHALF new
          // records per push, so the queue (drained below 32 after every push) cannot overflow
#pragma unroll 1
          for (int hq = 0; hq < SPR_QGROUP / SPR_QHALF; hq++) {
            uint32_t H[SPR_QHALF];
#pragma unroll
            for (int u = 0; u < SPR_QHALF / 2; u++) {
              const int4 v = __ldg(qgp + hq * (SPR_QHALF / 2) + u);  // two queries: (across, along) x 2
              H[2 * u] = spr_probe(plane, W, Rm1, maxbit, F, aqb + v.x, bqb + v.y, SPR_FULL);
              H[2 * u + 1] = spr_probe(plane, W, Rm1, maxbit, F, aqb + v.z, bqb + v.w, SPR_FULL);
            }
            uint32_t any = 0u;
#pragma unroll
            for (int u = 0; u < SPR_QHALF; u++) any |= H[u];
            any &= valid;
            if (__ballot_sync(SPR_FULL, any != 0u) == 0u) continue;
            // one record per (lane, query) with hits
            int n = 0;
#pragma unroll
            for (int u = 0; u < SPR_QHALF; u++) { H[u] &= valid; n += H[u] != 0u; }
            if (STATS) {
#pragma unroll
              for (int u = 0; u < SPR_QHALF; u++) n_hits += (unsigned long long)__popc(H[u]);
            }
            const int js0 = g * SPR_QGROUP + hq * SPR_QHALF;
            if (VARIANT == SPR_VARIANT_DIRECT) {  // test variant: every lane verifies its own hits
#pragma unroll
              for (int u = 0; u < SPR_QHALF; u++) {
                if (H[u] == 0u) continue;
                const int2 q = __ldg(qa2 + js0 + u);
                spr_verify_record<CNT32>(V, ws.cnt, l, a, js0 + u, lane, spr_cell_code(W, F, aqb + q.x, bqb + q.y), H[u], dir,
                                         along_off, across, n_inl);
              }
              continue;
            }
            // warp inclusive prefix sum of the per-lane record counts
            int incl = n;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
              const int t = __shfl_up_sync(SPR_FULL, incl, dlt);
              if (lane >= dlt) incl += t;
            }
            const int total = __shfl_sync(SPR_FULL, incl, 31);
            int pos = ws.qcount + incl - n;
#pragma unroll
            for (int u = 0; u < SPR_QHALF; u++) {
              if (H[u] == 0u) continue;
              const int2 q = __ldg(qa2 + js0 + u);  // L1 hit: loaded a moment ago by the probe
              ws.q_meta[pos] = ((uint32_t)(js0 + u) << 5) | (uint32_t)lane;
              ws.q_code[pos] = spr_cell_code(W, F, aqb + q.x, bqb + q.y);
              ws.q_mask[pos] = H[u];
              pos++;
            }
            ws.qcount += total;
            __syncwarp();
            while (ws.qcount >= 32) spr_drain32<CNT32>(V, ws, l, a, lane, dir, along_off, across, n_inl);
          }
        }
        // the queue only ever holds hits of the current label
        while (ws.qcount > 0) spr_drain32<CNT32>(V, ws, l, a, lane, dir, along_off, across, n_inl);
      }
      __syncwarp();
      // each lane scans the 32 hypotheses of its chunk
      uint32_t v = valid;
      while (v) {
        const int b = __ffs(v) - 1;
        v &= v - 1;
        const uint32_t c = spr_cnt_get<CNT32>(ws.cnt, lane, b);
        const unsigned long long ord = (unsigned long long)ord_base + (unsigned long long)b * ord_stride;
        const unsigned long long key = spr_make_key(c, ord * (unsigned long long)V.n_yaw + (unsigned long long)a);
        best = key > best ? key : best;
        if (WRITE_COUNTS) {
          const long long slot = ((long long)ord - (long long)K.ord_begin) * V.n_yaw + a;
          if (slot >= 0 && slot < K.counts_cap) K.counts_out[slot] = (int32_t)c;
        }
      }
    }
    __syncwarp();
  }
  // warp max of the 64-bit key, then one atomic per warp
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) {
    const unsigned long long o = __shfl_xor_sync(SPR_FULL, best, dlt);
    best = o > best ? o : best;
  }
  if (lane == 0 && best != 0ull) atomicMax(K.best_key, best);
  if (STATS) {
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
      n_hits += __shfl_xor_sync(SPR_FULL, n_hits, dlt);
      n_inl += __shfl_xor_sync(SPR_FULL, n_inl, dlt);
    }
    if (lane == 0) {
      atomicAdd(K.stats, n_hits); atomicAdd(K.stats + 1, n_inl);
      atomicAdd(K.stats + 2, n_probed); atomicAdd(K.stats + 3, n_skipped);
    }
  }
}

template <int VARIANT, bool CNT32>
static cudaError_t launch_variant(const SprView &V, const SprLaunch &K, int n_wg_local, long long n_items,
                                  int grid, cudaStream_t st) {
  const bool wc = K.counts_out != nullptr, stt = K.stats != nullptr;
  const size_t smem = (size_t)SPR_WARPS * SPR_WARP_SMEM(CNT32) * sizeof(uint32_t);
#define SPR_GO(WC, ST)                                                                                            \
  do {                                                                                                            \
    cudaError_t e = cudaFuncSetAttribute(spr_score_lattice_kernel<VARIANT, WC, ST, CNT32>,                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
    if (e != cudaSuccess) return e;                                                                               \
    spr_score_lattice_kernel<VARIANT, WC, ST, CNT32><<<grid, SPR_BLOCK, smem, st>>>(V, K, n_wg_local, n_items);    \
  } while (0)
  if (wc && stt) SPR_GO(true, true);
  else if (wc) SPR_GO(true, false);
  else if (stt) SPR_GO(false, true);
  else SPR_GO(false, false);
#undef SPR_GO
  return cudaGetLastError();
}

cudaError_t spr_launch_score_lattice(const SprView &V, const SprLaunch &K, int variant, int sm_count,
                                     cudaStream_t st, int *n_launches) {
  if (K.chunk_end <= K.chunk_begin || V.n_yaw <= 0) return cudaSuccess;
  const uint32_t n_chunks = K.chunk_end - K.chunk_begin;
  const int n_wg = (int)((n_chunks + SPR_WARP_CHUNKS - 1) / SPR_WARP_CHUNKS);
  const int sc = K.shard_count > 1 ? K.shard_count : 1;
  const int si = K.shard_count > 1 ? K.shard_index : 0;
  const int n_wg_local = n_wg > si ? (n_wg - si + sc - 1) / sc : 0;
  if (n_wg_local <= 0) return cudaSuccess;
  SprLaunch K2 = K;
  K2.shard_index = si;
  K2.shard_count = sc;
  const long long n_items = (long long)n_wg_local * V.n_yaw;
  const long long max_grid = (long long)sm_count * SPR_MINB;
  const long long want = (n_items + SPR_WARPS - 1) / SPR_WARPS;
  const int grid = (int)(want < max_grid ? want : max_grid);
  cudaError_t e = cudaMemsetAsync(K.work_counter, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  if (n_launches) (*n_launches)++;
  const bool cnt32 = V.nqp > 65535;  // a count can reach the number of query landmarks
  if (variant == SPR_VARIANT_DIRECT)
    return cnt32 ? launch_variant<SPR_VARIANT_DIRECT, true>(V, K2, n_wg_local, n_items, grid, st)
                 : launch_variant<SPR_VARIANT_DIRECT, false>(V, K2, n_wg_local, n_items, grid, st);
  return cnt32 ? launch_variant<SPR_VARIANT_QUEUED, true>(V, K2, n_wg_local, n_items, grid, st)
               : launch_variant<SPR_VARIANT_QUEUED, false>(V, K2, n_wg_local, n_items, grid, st);
}

// ---------------------------------------------------------------------------------------------
