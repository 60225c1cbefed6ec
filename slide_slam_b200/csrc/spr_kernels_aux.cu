// spr_kernels_aux.cu -- the smaller sm_100a kernels around the lattice search: explicit
// hypothesis lists (warp per hypothesis), correspondence extraction of the winner, the triangle
// descriptors of the SlideGraph half (semantic_clipper.cpp:49-99) and the issue-rate micro-benchmark.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

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
// SlideGraph descriptor half: triangle descriptors (semantic_clipper.cpp:49-99).  The reference
// recomputes both descriptors for each of the T1 x T2 pairs; here they are built once per triangle.
// The matching itself (binning, windowed sweep, radix sort into the reference's order) is in
// spr_generate.cu.
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

cudaError_t spr_launch_tri_desc(const double *tris6, const double *labels3, int t, double *desc, int32_t *perm,
                                double *sig, cudaStream_t st) {
  if (t <= 0) return cudaSuccess;
  spr_tri_desc_kernel<<<(t + 255) / 256, 256, 0, st>>>(tris6, labels3, t, desc, perm, sig);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Issue-rate micro-benchmark (bench.py's roofline denominator).  MEASURED_PEAKS.json only holds the
// HBM and bf16 tensor peaks; the search kernels are bound by the SM's instruction issue (integer /
// logic pipe), so the peak they are compared with is measured in the same run: independent
// dependency chains of LOP3 (ALU pipe), of IMAD (FMA pipe), and both interleaved (all four
// schedulers issuing every cycle), counted in warp instructions per second.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(1024) spr_issue_peak_kernel(uint32_t *out, int iters, uint32_t seed) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = seed + threadIdx.x * 8u + k; b[k] = seed * 3u + threadIdx.x + k; }
  const uint32_t c = seed | 1u, d = seed ^ 0x9e3779b9u;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (MODE == 0 || MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(c), "r"(d));
        if (MODE == 1 || MODE == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(c), "r"(d));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= a[k] ^ b[k];
  if (s == 0x12345678u) out[blockIdx.x] = s;  // keeps the chains alive
}

// warp instructions per second of the three instruction mixes; each launch runs ~1 ms
cudaError_t spr_measure_issue_peaks(int sm_count, double *alu, double *fma, double *mixed, cudaStream_t st) {
  uint32_t *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_out, 4096 * sizeof(uint32_t));
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sm_count * 2, iters = 4096;
  double *res[3] = {alu, fma, mixed};
  for (int mode = 0; mode < 3 && e == cudaSuccess; mode++) {
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {  // the first repetition warms the clocks up
      cudaEventRecord(e0, st);
      if (mode == 0) spr_issue_peak_kernel<0><<<grid, 1024, 0, st>>>(d_out, iters, 12345u + rep);
      else if (mode == 1) spr_issue_peak_kernel<1><<<grid, 1024, 0, st>>>(d_out, iters, 12345u + rep);
      else spr_issue_peak_kernel<2><<<grid, 1024, 0, st>>>(d_out, iters, 12345u + rep);
      cudaEventRecord(e1, st);
      e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) break;
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double insts = (double)grid * 32.0 /* warps */ * (double)iters * 32.0 * (mode == 2 ? 2.0 : 1.0);
      if (rep > 0 && ms > 0) best = std::max(best, insts / (ms * 1e-3));
    }
    *res[mode] = best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_out);
  return e != cudaSuccess ? e : cudaGetLastError();
}
