// spr_kernels.cu -- sm_100a kernels of the SlideMatch lattice search.
//
// Replaces the five nested loops of PlaceRecognition::MatchMaps (place_recognition.cpp:178-372).
// Work decomposition:
//   * one THREAD owns one chunk = 32 consecutive lattice translations along one axis, for one
//     yaw candidate; a warp owns 32 chunks (1024 hypotheses), a CTA 8 warps;
//   * for every query landmark the thread reads two words of the label's occupancy bitmap and
//     obtains the 32 hypotheses' filter bits with one funnel shift (spr_probe);
//   * set bits ("filter hits", a few % of the probes) are compacted through a per-warp
//     shared-memory queue (warp prefix-sum over popcounts) and verified 32 at a time in exact,
//     non-fused fp64 against the cell's candidate list (spr_verify_cell) -- the reference's own
//     predicate, so every hypothesis gets its exact inlier count;
//   * per-hypothesis counters live in shared memory; the best (count, canonical index) key is
//     reduced with shuffles and one 64-bit atomicMax per warp and work item.
// This is integer/bit and fp64 ALU work on L1/L2-resident data; there is no GEMM in it, so no
// tensor-core path (BASELINE.json north_star).
#include <cuda_runtime.h>
#include <stdint.h>

#include "spr_core.h"
#include "spr_kernels.h"

#define SPR_BLOCK 256
#define SPR_WARPS (SPR_BLOCK / 32)
#define SPR_QCAP 512            // per-warp hit queue, records
#define SPR_UNROLL 4

// ---------------------------------------------------------------------------------------------
// rotate: qrot[a][s] = R(yaw_a) * q_s in the reference's operation order, plus fixed-point cells
// ---------------------------------------------------------------------------------------------
__global__ void spr_rotate_kernel(SprView V, int32_t *__restrict__ qrotq, double *__restrict__ qrot) {
  const long long n = (long long)V.n_yaw * V.nq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i / V.nq), s = (int)(i % V.nq);
    double rx, ry;
    spr_rotate(V.cs[2 * a], V.cs[2 * a + 1], V.qxy[2 * (size_t)s], V.qxy[2 * (size_t)s + 1], &rx, &ry);
    reinterpret_cast<double2 *>(qrot)[i] = make_double2(rx, ry);
    reinterpret_cast<int2 *>(qrotq)[i] =
        make_int2(spr_fx(SPR_DSUB(rx, V.grid.g0x), V.grid.S), spr_fx(SPR_DSUB(ry, V.grid.g0y), V.grid.S));
  }
}

cudaError_t spr_launch_rotate(const SprView &V, int32_t *qrotq, double *qrot, cudaStream_t st) {
  const long long n = (long long)V.n_yaw * V.nq;
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  const int grid = (int)((n + block - 1) / block > 148 * 8 ? 148 * 8 : (n + block - 1) / block);
  spr_rotate_kernel<<<grid, block, 0, st>>>(V, qrotq, qrot);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// lattice scoring
// ---------------------------------------------------------------------------------------------
struct WarpState {
  uint32_t *cnt;    // [32 bits][32 lanes] inlier counters of the warp's 1024 hypotheses
  uint32_t *queue;  // [SPR_QCAP] pending filter hits: js << 10 | owner lane << 5 | bit
  int qcount;       // warp-uniform
};

// Verify up to 32 queued hits, one per lane.  Chunk parameters of the owning lane come through
// shuffles; every lane executes the shuffles.
__device__ __forceinline__ void spr_drain32(const SprView &V, WarpState &ws, int a, int lane,
                                            const SprChunk &ch, int32_t aq0, int32_t bq0,
                                            unsigned long long &n_inl) {
  const int n = ws.qcount < 32 ? ws.qcount : 32;
  const bool active = lane < n;
  const uint32_t rec = active ? ws.queue[ws.qcount - n + lane] : 0u;
  const int owner = (rec >> 5) & 31, b = rec & 31;
  const int js = (int)(rec >> 10);
  const int32_t o_aq0 = __shfl_sync(0xffffffffu, aq0, owner);
  const int32_t o_bq0 = __shfl_sync(0xffffffffu, bq0, owner);
  const uint32_t o_dir = __shfl_sync(0xffffffffu, ch.dir, owner);
  const uint32_t o_off = __shfl_sync(0xffffffffu, ch.along_off, owner);
  const double o_across = __shfl_sync(0xffffffffu, ch.across, owner);
  if (active) {
    const SprGrid &G = V.grid;
    const int2 qq = __ldg(reinterpret_cast<const int2 *>(V.qrotq) + (size_t)a * V.nq + js);
    const int32_t na = (o_aq0 + (o_dir ? qq.y : qq.x)) >> G.F;
    const int32_t nb = ((o_bq0 + (o_dir ? qq.x : qq.y)) >> G.F) + b;
    const int32_t nx = o_dir ? nb : na, ny = o_dir ? na : nb;
    const double along = __ldg(V.lat + o_off + b);
    const double tx = o_dir ? along : o_across, ty = o_dir ? o_across : along;
    const double2 r = __ldg(reinterpret_cast<const double2 *>(V.qrot) + (size_t)a * V.nq + js);
    int32_t first;
    if (spr_verify_cell(V, __ldg(V.qlabel + js), nx, ny, r.x, r.y, tx, ty, V.qdims + 3 * (size_t)js, &first)) {
      atomicAdd(&ws.cnt[b * 32 + owner], 1u);
      n_inl++;
    }
  }
  ws.qcount -= n;
  __syncwarp();
}

template <int VARIANT, bool WRITE_COUNTS, bool STATS>
__global__ void __launch_bounds__(SPR_BLOCK)
spr_score_lattice_kernel(SprView V, SprLaunch K, int n_groups_local, long long n_items) {
  extern __shared__ uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpState ws;
  ws.cnt = smem + warp * 1024;
  ws.queue = smem + SPR_WARPS * 1024 + warp * SPR_QCAP;
  ws.qcount = 0;
  const SprGrid &G = V.grid;
  unsigned long long best = 0ull, n_hits = 0ull, n_inl = 0ull;

  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    // yaw-major order: neighbouring CTAs work on the same yaw and share qrotq[a][*] in L2/L1
    const int a = (int)(item / n_groups_local);
    const int g = K.shard_index + (int)(item % n_groups_local) * K.shard_count;
    const uint32_t cidx = K.chunk_begin + (uint32_t)g * SPR_BLOCK + threadIdx.x;
    SprChunk ch;
    if (cidx < K.chunk_end) {
      ch = V.chunks[cidx];
    } else {
      ch.across = 0.0; ch.along_off = 0; ch.valid = 0; ch.ord_base = 0; ch.ord_stride = 0; ch.dir = 0; ch.ring = 0;
    }
#pragma unroll
    for (int b = 0; b < 32; b++) ws.cnt[b * 32 + lane] = 0u;
    __syncwarp();
    const int d = (int)ch.dir;
    const int32_t aq0 = spr_fx(ch.across, G.S);
    const int32_t bq0 = spr_fx(__ldg(V.lat + ch.along_off), G.S);
    const int32_t W = G.W[d], R = G.R[d], maxbit = G.maxbit[d];
    const int2 *__restrict__ qa = reinterpret_cast<const int2 *>(V.qrotq) + (size_t)a * V.nq;
    // skip the whole item when the warp has nothing to score (tail of the chunk list)
    if (__ballot_sync(0xffffffffu, ch.valid != 0u) != 0u) {
      for (int l = 0; l < V.n_labels; l++) {
        const uint32_t *__restrict__ plane = V.bitmap + (size_t)l * G.label_stride + (d ? G.plane_words[0] : 0u);
        const int s1 = V.label_seg[l + 1];
        for (int s0 = V.label_seg[l]; s0 < s1; s0 += SPR_UNROLL) {
          uint32_t H[SPR_UNROLL];
          int32_t na[SPR_UNROLL], nb[SPR_UNROLL];
#pragma unroll
          for (int u = 0; u < SPR_UNROLL; u++) {
            const int s = s0 + u < s1 ? s0 + u : s1 - 1;
            const int2 qq = __ldg(qa + s);
            H[u] = spr_probe(plane, W, R, maxbit, G.F, aq0 + (d ? qq.y : qq.x), bq0 + (d ? qq.x : qq.y),
                             s0 + u < s1 ? ch.valid : 0u, &na[u], &nb[u]);
          }
          int n = 0;
#pragma unroll
          for (int u = 0; u < SPR_UNROLL; u++) n += __popc(H[u]);
          if (STATS) n_hits += (unsigned long long)n;
          if (__ballot_sync(0xffffffffu, n != 0) == 0u) continue;
          if (VARIANT == SPR_VARIANT_DIRECT) {
#pragma unroll
            for (int u = 0; u < SPR_UNROLL; u++) {
              uint32_t h = H[u];
              while (h) {
                const int b = __ffs(h) - 1;
                h &= h - 1;
                int32_t first;
                if (spr_verify_hit(V, ch, l, a, s0 + u, na[u], nb[u], b, &first)) {
                  ws.cnt[b * 32 + lane]++;
                  n_inl++;
                }
              }
            }
          } else {
            // warp inclusive prefix sum of the per-lane hit counts
            int incl = n;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, incl, dlt);
              if (lane >= dlt) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total > SPR_QCAP) {
              // pathological density: verify in place (atomics: queued hits may target our columns)
#pragma unroll
              for (int u = 0; u < SPR_UNROLL; u++) {
                uint32_t h = H[u];
                while (h) {
                  const int b = __ffs(h) - 1;
                  h &= h - 1;
                  int32_t first;
                  if (spr_verify_hit(V, ch, l, a, s0 + u, na[u], nb[u], b, &first)) {
                    atomicAdd(&ws.cnt[b * 32 + lane], 1u);
                    n_inl++;
                  }
                }
              }
              continue;
            }
            while (ws.qcount + total > SPR_QCAP) spr_drain32(V, ws, a, lane, ch, aq0, bq0, n_inl);
            int pos = ws.qcount + incl - n;
#pragma unroll
            for (int u = 0; u < SPR_UNROLL; u++) {
              uint32_t h = H[u];
              while (h) {
                const int b = __ffs(h) - 1;
                h &= h - 1;
                ws.queue[pos++] = ((uint32_t)(s0 + u) << 10) | ((uint32_t)lane << 5) | (uint32_t)b;
              }
            }
            ws.qcount += total;
            __syncwarp();
            while (ws.qcount >= 32) spr_drain32(V, ws, a, lane, ch, aq0, bq0, n_inl);
          }
        }
      }
      if (VARIANT != SPR_VARIANT_DIRECT)
        while (ws.qcount > 0) spr_drain32(V, ws, a, lane, ch, aq0, bq0, n_inl);
      __syncwarp();
      // each lane scans the 32 hypotheses of its chunk
      uint32_t v = ch.valid;
      while (v) {
        const int b = __ffs(v) - 1;
        v &= v - 1;
        const uint32_t c = ws.cnt[b * 32 + lane];
        const unsigned long long ord = (unsigned long long)ch.ord_base + (unsigned long long)b * ch.ord_stride;
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
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, dlt);
    best = o > best ? o : best;
  }
  if (lane == 0 && best != 0ull) atomicMax(K.best_key, best);
  if (STATS) {
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
      n_hits += __shfl_xor_sync(0xffffffffu, n_hits, dlt);
      n_inl += __shfl_xor_sync(0xffffffffu, n_inl, dlt);
    }
    if (lane == 0) { atomicAdd(K.stats, n_hits); atomicAdd(K.stats + 1, n_inl); }
  }
}

template <int VARIANT>
static cudaError_t launch_variant(const SprView &V, const SprLaunch &K, int n_groups_local, long long n_items,
                                  int grid, size_t smem, cudaStream_t st) {
  const bool wc = K.counts_out != nullptr, stt = K.stats != nullptr;
#define SPR_GO(WC, ST)                                                                                     \
  do {                                                                                                     \
    cudaError_t e = cudaFuncSetAttribute(spr_score_lattice_kernel<VARIANT, WC, ST>,                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (e != cudaSuccess) return e;                                                                        \
    spr_score_lattice_kernel<VARIANT, WC, ST><<<grid, SPR_BLOCK, smem, st>>>(V, K, n_groups_local, n_items); \
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
  const int n_groups = (int)((n_chunks + SPR_BLOCK - 1) / SPR_BLOCK);
  const int sc = K.shard_count > 1 ? K.shard_count : 1;
  const int si = K.shard_count > 1 ? K.shard_index : 0;
  const int n_groups_local = n_groups > si ? (n_groups - si + sc - 1) / sc : 0;
  if (n_groups_local <= 0) return cudaSuccess;
  SprLaunch K2 = K;
  K2.shard_index = si;
  K2.shard_count = sc;
  const long long n_items = (long long)n_groups_local * V.n_yaw;
  const size_t smem = (size_t)SPR_WARPS * (1024 + SPR_QCAP) * sizeof(uint32_t);
  const long long max_grid = (long long)sm_count * 4;
  const int grid = (int)(n_items < max_grid ? n_items : max_grid);
  if (n_launches) (*n_launches)++;
  if (variant == SPR_VARIANT_DIRECT) return launch_variant<SPR_VARIANT_DIRECT>(V, K2, n_groups_local, n_items, grid, smem, st);
  return launch_variant<SPR_VARIANT_QUEUED>(V, K2, n_groups_local, n_items, grid, smem, st);
}

// ---------------------------------------------------------------------------------------------
// explicit hypothesis list: warp per hypothesis, lanes stride over the query landmarks
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spr_score_list_kernel(SprView V, const double *__restrict__ hyps4, long long n, int32_t *__restrict__ counts_out,
                      unsigned long long *best_key) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long best = 0ull;
  for (long long h = warp0; h < n; h += n_warps) {
    const double c = hyps4[4 * h], s = hyps4[4 * h + 1], tx = hyps4[4 * h + 2], ty = hyps4[4 * h + 3];
    int cnt = 0;
    for (int js = lane; js < V.nq; js += 32) {
      double rx, ry;
      spr_rotate(c, s, V.qxy[2 * (size_t)js], V.qxy[2 * (size_t)js + 1], &rx, &ry);
      const double xt = SPR_DADD(rx, tx), yt = SPR_DADD(ry, ty);
      const int l = V.qlabel[js];
      int32_t nx, ny, first;
      if (spr_point_cell(V, l, xt, yt, &nx, &ny) &&
          spr_verify_cell(V, l, nx, ny, rx, ry, tx, ty, V.qdims + 3 * (size_t)js, &first))
        cnt++;
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, dlt);
    if (lane == 0) {
      if (counts_out) counts_out[h] = cnt;
      const unsigned long long key = spr_make_key((uint32_t)cnt, (unsigned long long)h);
      best = key > best ? key : best;
    }
  }
  if (lane == 0 && best != 0ull) atomicMax(best_key, best);
}

cudaError_t spr_launch_score_list(const SprView &V, const double *hyps4, long long n, int32_t *counts_out,
                                  unsigned long long *best_key, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long want = (n + 7) / 8;
  const long long cap = (long long)sm_count * 8;
  spr_score_list_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(V, hyps4, n, counts_out, best_key);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// correspondences of the winner: the reference's own double loop (PR.cpp:281-357), one thread
// per query object, reference objects in ascending order, first match wins.
// ---------------------------------------------------------------------------------------------
__global__ void spr_extract_kernel(const double *__restrict__ ref7, int n_ref, const double *__restrict__ qry7,
                                   int n_qry, double c, double s, double tx, double ty, double Tstar,
                                   double Sstar, double thr_dim, int ignore_dim, int32_t *__restrict__ match_ref) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_qry) return;
  const double *q = qry7 + 7 * (size_t)j;
  const double label = q[0];
  double rx, ry;
  spr_rotate(c, s, q[1], q[2], &rx, &ry);
  const double qd[3] = {q[4], q[5], q[6]};
  int32_t found = -1;
  for (int i = 0; i < n_ref; i++) {
    const double *r = ref7 + 7 * (size_t)i;
    if (r[0] != label) continue;                                         // PR.cpp:306
    if (!spr_distance_match(rx, ry, tx, ty, r[1], r[2], Tstar)) continue;  // PR.cpp:332
    if (!ignore_dim) {
      const double rd[3] = {r[4], r[5], r[6]};
      if (!spr_dimension_match(rd, qd, thr_dim, Sstar)) continue;        // PR.cpp:334-339
    }
    found = i;
    break;                                                               // PR.cpp:353
  }
  match_ref[j] = found;
}

cudaError_t spr_launch_extract(const double *ref7, int n_ref, const double *qry7, int n_qry, double c,
                               double s, double tx, double ty, double Tstar, double Sstar, double thr_dim,
                               int ignore_dim, int32_t *match_ref, cudaStream_t st) {
  if (n_qry <= 0) return cudaSuccess;
  const int block = 128;
  spr_extract_kernel<<<(n_qry + block - 1) / block, block, 0, st>>>(ref7, n_ref, qry7, n_qry, c, s, tx, ty, Tstar,
                                                                    Sstar, thr_dim, ignore_dim, match_ref);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// SlideGraph descriptor half: triangle descriptors + all-pairs matching
// (semantic_clipper.cpp:49-118).  The reference recomputes both descriptors for each of the
// T1 x T2 pairs; here they are built once per triangle, then every model triangle (one warp)
// sweeps the data descriptors 32 at a time and compacts its matches with ballot + popc so the
// output keeps the reference's order (model-major, data-minor).
// ---------------------------------------------------------------------------------------------
__global__ void spr_tri_desc_kernel(const double *__restrict__ tris6, int t, double *__restrict__ desc,
                                    int32_t *__restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t) return;
  double tri[6], d[3];
  int32_t p[3];
#pragma unroll
  for (int k = 0; k < 6; k++) tri[k] = tris6[6 * (size_t)i + k];
  spr_triangle_descriptor(tri, d, p);
#pragma unroll
  for (int k = 0; k < 3; k++) { desc[3 * (size_t)i + k] = d[k]; perm[3 * (size_t)i + k] = p[k]; }
}

template <bool FILL>
__global__ void __launch_bounds__(256)
spr_tri_match_kernel(const double *__restrict__ dm, int tm, const double *__restrict__ dd, int td, double thr,
                     unsigned long long *__restrict__ counts, const unsigned long long *__restrict__ offsets,
                     int32_t *__restrict__ model_idx, int32_t *__restrict__ data_idx, long long cap) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < tm; i += n_warps) {
    const double m[3] = {dm[3 * (size_t)i], dm[3 * (size_t)i + 1], dm[3 * (size_t)i + 2]};
    unsigned long long base = FILL ? offsets[i] : 0ull;
    for (int j0 = 0; j0 < td; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < td) {
        const double d[3] = {dd[3 * (size_t)j], dd[3 * (size_t)j + 1], dd[3 * (size_t)j + 2]};
        hit = spr_descriptor_match(m, d, thr);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (FILL && hit) {
        const unsigned long long pos = base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
        if ((long long)pos < cap) { model_idx[pos] = i; data_idx[pos] = j; }
      }
      base += (unsigned long long)__popc(mask);
    }
    if (!FILL && lane == 0) counts[i] = base;
  }
}

// exclusive prefix sum of n counters, one block (n <= a few 10^5 triangles)
__global__ void spr_scan_kernel(const unsigned long long *__restrict__ counts, int n,
                                unsigned long long *__restrict__ offsets, unsigned long long *__restrict__ total) {
  __shared__ unsigned long long tile[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0ull;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < n ? counts[i] : 0ull;
    tile[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const unsigned long long t = threadIdx.x >= d ? tile[threadIdx.x - d] : 0ull;
      __syncthreads();
      tile[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) offsets[i] = carry + tile[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += tile[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

cudaError_t spr_launch_tri_desc(const double *tris6, int t, double *desc, int32_t *perm, cudaStream_t st) {
  if (t <= 0) return cudaSuccess;
  spr_tri_desc_kernel<<<(t + 255) / 256, 256, 0, st>>>(tris6, t, desc, perm);
  return cudaGetLastError();
}

cudaError_t spr_launch_tri_match(const double *dm, int tm, const double *dd, int td, double thr,
                                 unsigned long long *counts, unsigned long long *offsets, unsigned long long *total,
                                 int32_t *model_idx, int32_t *data_idx, long long cap, bool fill, int sm_count,
                                 cudaStream_t st) {
  if (tm <= 0) return cudaSuccess;
  const int want = (tm + 7) / 8, capg = sm_count * 8;
  const int grid = want < capg ? want : capg;
  if (!fill) {
    spr_tri_match_kernel<false><<<grid, 256, 0, st>>>(dm, tm, dd, td, thr, counts, offsets, model_idx, data_idx, cap);
    spr_scan_kernel<<<1, 1024, 0, st>>>(counts, tm, offsets, total);
  } else {
    spr_tri_match_kernel<true><<<grid, 256, 0, st>>>(dm, tm, dd, td, thr, counts, offsets, model_idx, data_idx, cap);
  }
  return cudaGetLastError();
}
