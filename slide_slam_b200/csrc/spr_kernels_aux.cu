// spr_kernels_aux.cu -- the smaller sm_100a kernels around the lattice search: explicit
// hypothesis lists (warp per hypothesis), correspondence extraction of the winner, and the
// SlideGraph triangle-descriptor matching (semantic_clipper.cpp:49-118).
#include <cuda_runtime.h>
#include <stdint.h>

#include "spr_core.h"
#include "spr_kernels.h"

#define SPR_FULL 0xffffffffu

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
    for (int js = lane; js < V.nqp; js += 32) {
      const int l = V.qlabel[js];
      if (l < 0) continue;  // padding
      double rx, ry;
      spr_rotate(c, s, V.qxy[2 * (size_t)js], V.qxy[2 * (size_t)js + 1], &rx, &ry);
      const double xt = SPR_DADD(rx, tx), yt = SPR_DADD(ry, ty);
      uint32_t row, bit;
      if (spr_point_cell(V, l, xt, yt, &row, &bit) &&
          spr_verify_cell(V, spr_global_tables(V, 0u, l), 0u, row, bit, rx, ry, tx, ty, V.qdims + 3 * (size_t)js))
        cnt++;
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) cnt += __shfl_xor_sync(SPR_FULL, cnt, dlt);
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
// correspondences of the winner: the reference's own double loop (PR.cpp:281-357) over the raw
// maps, reference objects in ascending order, first match wins.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
spr_extract_kernel(const double *__restrict__ ref7, int n_ref, const double *__restrict__ qry7, int n_qry, double c,
                   double s, double tx, double ty, double Tstar, double Sstar, double thr_dim, int ignore_dim,
                   int32_t *__restrict__ match_ref) {
  // one CTA per query object: thread t tests reference objects t, t + 128, ... in ascending order
  // and stops at its first match; the smallest index over the CTA is the reference's "first
  // match, then break" (PR.cpp:299-355)
  __shared__ int s_first;
  const int j = blockIdx.x;
  if (threadIdx.x == 0) s_first = 0x7fffffff;
  __syncthreads();
  const double *q = qry7 + 7 * (size_t)j;
  const double label = q[0];
  double rx, ry;
  spr_rotate(c, s, q[1], q[2], &rx, &ry);
  const double qd[3] = {q[4], q[5], q[6]};
  int first = 0x7fffffff;
  for (int i = threadIdx.x; i < n_ref; i += blockDim.x) {
    const double *r = ref7 + 7 * (size_t)i;
    if (r[0] == label &&                                                         // PR.cpp:306
        spr_distance_match(rx, ry, tx, ty, r[1], r[2], Tstar) &&                 // PR.cpp:332
        (ignore_dim || spr_dimension_match(r[4], r[5], r[6], qd, thr_dim, Sstar))) {  // PR.cpp:334-339
      first = i;
      break;                                                                     // PR.cpp:353
    }
  }
  first = __reduce_min_sync(SPR_FULL, first);
  if ((threadIdx.x & 31) == 0 && first != 0x7fffffff) atomicMin(&s_first, first);
  __syncthreads();
  if (threadIdx.x == 0) match_ref[j] = s_first == 0x7fffffff ? -1 : s_first;
}

cudaError_t spr_launch_extract(const double *ref7, int n_ref, const double *qry7, int n_qry, double c,
                               double s, double tx, double ty, double Tstar, double Sstar, double thr_dim,
                               int ignore_dim, int32_t *match_ref, cudaStream_t st) {
  if (n_qry <= 0) return cudaSuccess;
  spr_extract_kernel<<<n_qry, 128, 0, st>>>(ref7, n_ref, qry7, n_qry, c, s, tx, ty, Tstar, Sstar, thr_dim, ignore_dim,
                                            match_ref);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// SlideGraph descriptor half: triangle descriptors + all-pairs matching
// (semantic_clipper.cpp:49-118).  The reference recomputes both descriptors for each of the
// T1 x T2 pairs; here they are built once per triangle, then every model triangle (one warp)
// sweeps the data descriptors 32 at a time and compacts its matches with ballot + popc so the
// output keeps the reference's order (model-major, data-minor).
// ---------------------------------------------------------------------------------------------
__global__ void spr_tri_desc_kernel(const double *__restrict__ tris6, const double *__restrict__ labels3, int t,
                                    double *__restrict__ desc, int32_t *__restrict__ perm, double *__restrict__ sig) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t) return;
  double tri[6], d[3];
  int32_t p[3];
#pragma unroll
  for (int k = 0; k < 6; k++) tri[k] = tris6[6 * (size_t)i + k];
  spr_triangle_descriptor(tri, d, p);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    desc[3 * (size_t)i + k] = d[k];
    perm[3 * (size_t)i + k] = p[k];
    // class signature: the vertex labels in sorted-descriptor order (the order in which the
    // vertices are paired when two triangles match, SC.cpp:102-105)
    if (labels3) sig[3 * (size_t)i + k] = labels3[3 * (size_t)i + p[k]];
  }
}

template <bool FILL>
__global__ void __launch_bounds__(256)
spr_tri_match_kernel(const double *__restrict__ dm, int tm, const double *__restrict__ dd, int td,
                     const double *__restrict__ sm, const double *__restrict__ sd, double thr,
                     unsigned long long *__restrict__ counts, const unsigned long long *__restrict__ offsets,
                     int32_t *__restrict__ model_idx, int32_t *__restrict__ data_idx, long long cap) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < tm; i += n_warps) {
    const double m[3] = {dm[3 * (size_t)i], dm[3 * (size_t)i + 1], dm[3 * (size_t)i + 2]};
    double ms[3] = {0.0, 0.0, 0.0};
    if (sm) { ms[0] = sm[3 * (size_t)i]; ms[1] = sm[3 * (size_t)i + 1]; ms[2] = sm[3 * (size_t)i + 2]; }
    unsigned long long base = FILL ? offsets[i] : 0ull;
    for (int j0 = 0; j0 < td; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < td) {
        const double d[3] = {dd[3 * (size_t)j], dd[3 * (size_t)j + 1], dd[3 * (size_t)j + 2]};
        hit = spr_descriptor_match(m, d, thr);
        // class-consistent pairs only: every paired vertex must carry the same label (the check the
        // reference leaves as a TODO, SC.cpp:114,186); labels compare as doubles like PR.cpp:306
        if (hit && sm)
          hit = ms[0] == sd[3 * (size_t)j] && ms[1] == sd[3 * (size_t)j + 1] && ms[2] == sd[3 * (size_t)j + 2];
      }
      const unsigned mask = __ballot_sync(SPR_FULL, hit);
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

cudaError_t spr_launch_tri_desc(const double *tris6, const double *labels3, int t, double *desc, int32_t *perm,
                                double *sig, cudaStream_t st) {
  if (t <= 0) return cudaSuccess;
  spr_tri_desc_kernel<<<(t + 255) / 256, 256, 0, st>>>(tris6, labels3, t, desc, perm, sig);
  return cudaGetLastError();
}

cudaError_t spr_launch_tri_match(const double *dm, int tm, const double *dd, int td, const double *sm, const double *sd, double thr,
                                 unsigned long long *counts, unsigned long long *offsets, unsigned long long *total,
                                 int32_t *model_idx, int32_t *data_idx, long long cap, bool fill, int sm_count,
                                 cudaStream_t st) {
  if (tm <= 0) return cudaSuccess;
  const int want = (tm + 7) / 8, capg = sm_count * 8;
  const int grid = want < capg ? want : capg;
  if (!fill) {
    spr_tri_match_kernel<false><<<grid, 256, 0, st>>>(dm, tm, dd, td, sm, sd, thr, counts, offsets, model_idx, data_idx, cap);
    spr_scan_kernel<<<1, 1024, 0, st>>>(counts, tm, offsets, total);
  } else {
    spr_tri_match_kernel<true><<<grid, 256, 0, st>>>(dm, tm, dd, td, sm, sd, thr, counts, offsets, model_idx, data_idx, cap);
  }
  return cudaGetLastError();
}
