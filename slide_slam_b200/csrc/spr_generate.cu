// spr_generate.cu -- descriptor and hypothesis generation on the device (BASELINE.json north_star,
// subsystem 1; reference primitives: semantic_clipper.cpp:49-108 compute_triangle_diff, :111-118
// match_triangles, :122-138 estimate_tf).
//
// The reference tests all T_model x T_data triangle pairs and recomputes both descriptors for each.
// Here, per map pair:
//   1. descriptors once per triangle (spr_tri_desc_kernel, spr_kernels_aux.cu);
//   2. the data triangles are BINNED by the first descriptor component d0 (bin width >= the matching
//      threshold): histogram, one-CTA scan, scatter of compact records -- a counting sort by key;
//   3. a model triangle can only match data triangles whose d0 lies within the threshold of its own
//      (|e0| <= sqrt(e0^2 + e1^2 + e2^2) < thr), i.e. a contiguous window of the binned array: one warp
//      per model triangle sweeps that window, evaluates the reference's predicate exactly
//      (sum < T* <=> sqrt(sum) < thr), optionally the class signature, and appends the hits with one
//      atomic per warp and ballot compaction -- a single pass, no count pass;
//   4. the matches (keys model_idx * T_data + data_idx) are brought back into the reference's order
//      -- model-major, data-minor, SC.cpp:111-118 -- by an LSD radix sort over the significant bits;
//   5. one 2-D Kabsch fit per match (closed form of estimate_tf's 2 x 2 SVD) writes (c, s, x, y)
//      straight into the hypothesis buffer of the list scorer: no host round trip between matching
//      and scoring.
#include <cuda_runtime.h>
#include <stdint.h>

#include "spr_core.h"
#include "spr_generate.h"

#define GEN_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// LSD radix sort of 64-bit keys, 8-bit digits, stable
// ---------------------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 16                       // keys per thread and tile
#define RS_TILE (RS_THREADS * RS_ITEMS)

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long *__restrict__ keys, long long n, int shift, uint32_t *__restrict__ hist, int n_blocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0u;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE;
  for (int k = 0; k < RS_ITEMS; k++) {
    const long long i = base + (long long)k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];  // digit-major: one scan gives global offsets
}

// exclusive scan of n 32-bit counters by one CTA (n = 256 * tiles, or the descriptor bins)
__global__ void __launch_bounds__(1024) gen_scan_u32_kernel(uint32_t *__restrict__ a, long long n, uint32_t *__restrict__ total) {
  __shared__ uint32_t wtot[32];
  __shared__ uint32_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + threadIdx.x;
    const uint32_t v = i < n ? a[i] : 0u;
    uint32_t s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(GEN_FULL, s, d); if (lane >= d) s += t; }
    if (lane == 31) wtot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wtot[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(GEN_FULL, w, d); if (lane >= d) w += t; }
      wtot[lane] = w;
    }
    __syncthreads();
    const uint32_t before = carry + (warp ? wtot[warp - 1] : 0u) + s - v;
    if (i < n) a[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out, long long n, int shift,
                  const uint32_t *__restrict__ offs, int n_blocks) {
  // the tile is consumed in chunks of RS_THREADS consecutive keys; inside a chunk a key's rank among
  // the keys with the same digit is (same-digit keys in lower warps) + (same-digit lower lanes)
  __shared__ uint32_t running[256];
  __shared__ uint32_t wcnt[RS_THREADS / 32][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = offs[(size_t)threadIdx.x * n_blocks + blockIdx.x];
  const long long base = (long long)blockIdx.x * RS_TILE;
  for (int k = 0; k < RS_ITEMS; k++) {
    for (int w = 0; w < RS_THREADS / 32; w++) wcnt[w][threadIdx.x] = 0u;
    __syncthreads();
    const long long i = base + (long long)k * RS_THREADS + threadIdx.x;
    const bool in_range = i < n;
    const unsigned long long key = in_range ? in[i] : 0ull;
    const uint32_t dg = in_range ? (uint32_t)(key >> shift) & 255u : 256u;
    const uint32_t same = __match_any_sync(GEN_FULL, dg);
    const uint32_t rank_w = (uint32_t)__popc(same & ((1u << lane) - 1u));
    if (in_range && rank_w == 0u) wcnt[warp][dg] = (uint32_t)__popc(same);
    __syncthreads();
    if (in_range) {
      uint32_t before = running[dg];
      for (int w = 0; w < warp; w++) before += wcnt[w][dg];
      out[before + rank_w] = key;
    }
    __syncthreads();
    uint32_t add = 0u;
    for (int w = 0; w < RS_THREADS / 32; w++) add += wcnt[w][threadIdx.x];
    running[threadIdx.x] += add;
    __syncthreads();
  }
}

// Sorts keys[0 .. n) ascending on their low `bits` bits.  tmp: n keys; hist: 256 * tiles counters.
// The sorted keys end up in `keys` (an odd number of passes is followed by a copy).
cudaError_t spr_radix_sort_u64(unsigned long long *keys, unsigned long long *tmp, long long n, int bits, uint32_t *hist,
                               cudaStream_t st) {
  if (n <= 1) return cudaSuccess;
  const int n_blocks = (int)((n + RS_TILE - 1) / RS_TILE);
  unsigned long long *a = keys, *b = tmp;
  for (int shift = 0; shift < bits; shift += 8) {
    rs_hist_kernel<<<n_blocks, RS_THREADS, 0, st>>>(a, n, shift, hist, n_blocks);
    gen_scan_u32_kernel<<<1, 1024, 0, st>>>(hist, 256ll * n_blocks, nullptr);
    rs_scatter_kernel<<<n_blocks, RS_THREADS, 0, st>>>(a, b, n, shift, hist, n_blocks);
    unsigned long long *t = a; a = b; b = t;
  }
  if (a != keys) {
    cudaError_t e = cudaMemcpyAsync(keys, a, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

size_t spr_radix_sort_hist_words(long long n) { return 256 * (size_t)((n + RS_TILE - 1) / RS_TILE) + 256; }

// ---------------------------------------------------------------------------------------------
// binning of the data triangles by d0
// ---------------------------------------------------------------------------------------------
struct GenBinned {  // one data triangle in bin order
  double d[3];
  int32_t j;
  int32_t pad;
};

__device__ __forceinline__ uint32_t gen_bin_of(double d0, double inv_w, uint32_t n_bins) {
  const double b = d0 * inv_w;           // d0 >= 0; NaN descriptors (degenerate input) land in bin 0 and never match
  if (!(b > 0.0)) return 0u;
  return b >= (double)(n_bins - 1u) ? n_bins - 1u : (uint32_t)b;
}

__global__ void gen_bin_count_kernel(const double *__restrict__ desc, int t, double inv_w, uint32_t n_bins, uint32_t *__restrict__ cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < t) atomicAdd(&cnt[gen_bin_of(desc[3 * (size_t)j], inv_w, n_bins)], 1u);
}

__global__ void gen_bin_scatter_kernel(const double *__restrict__ desc, int t, double inv_w, uint32_t n_bins,
                                       const uint32_t *__restrict__ start, uint32_t *__restrict__ fill, GenBinned *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= t) return;
  const uint32_t b = gen_bin_of(desc[3 * (size_t)j], inv_w, n_bins);
  const uint32_t pos = start[b] + atomicAdd(&fill[b], 1u);
  GenBinned r;
  r.d[0] = desc[3 * (size_t)j]; r.d[1] = desc[3 * (size_t)j + 1]; r.d[2] = desc[3 * (size_t)j + 2];
  r.j = j; r.pad = 0;
  out[pos] = r;
}

// One warp per model triangle: sweep the window of bins around its d0, exact predicate, append.
__global__ void __launch_bounds__(256)
gen_match_kernel(const double *__restrict__ dm, int tm, const GenBinned *__restrict__ binned, const uint32_t *__restrict__ start,
                 uint32_t n_bins, double inv_w, double window, double Tstar, const double *__restrict__ sm,
                 const double *__restrict__ sd, unsigned long long t_data, unsigned long long *__restrict__ keys,
                 unsigned long long cap, unsigned long long *__restrict__ total) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < tm; i += n_warps) {
    const double m0 = dm[3 * (size_t)i], m1 = dm[3 * (size_t)i + 1], m2 = dm[3 * (size_t)i + 2];
    if (!(m0 == m0)) continue;  // NaN descriptor: sqrt(NaN) < thr is false for every pair
    const uint32_t b_lo = gen_bin_of(m0 - window, inv_w, n_bins), b_hi = gen_bin_of(m0 + window, inv_w, n_bins);
    const uint32_t p0 = start[b_lo], p1 = start[b_hi + 1u];  // start has n_bins + 1 entries
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (sm) { s0 = sm[3 * (size_t)i]; s1 = sm[3 * (size_t)i + 1]; s2 = sm[3 * (size_t)i + 2]; }
    for (uint32_t q0 = p0; q0 < p1; q0 += 32u) {
      const uint32_t q = q0 + (uint32_t)lane;
      bool hit = false;
      int32_t j = 0;
      if (q < p1) {
        const GenBinned r = binned[q];
        j = r.j;
        // SC.cpp:92-99: sqrt(sum (dm - dd)^2) < threshold, left to right, not fused
        const double e0 = SPR_DSUB(m0, r.d[0]), e1 = SPR_DSUB(m1, r.d[1]), e2 = SPR_DSUB(m2, r.d[2]);
        const double s = SPR_DADD(SPR_DADD(SPR_DMUL(e0, e0), SPR_DMUL(e1, e1)), SPR_DMUL(e2, e2));
        hit = s < Tstar;
        if (hit && sm)  // class signature: the vertices paired by the sorted order carry equal labels
          hit = s0 == sd[3 * (size_t)j] && s1 == sd[3 * (size_t)j + 1] && s2 == sd[3 * (size_t)j + 2];
      }
      const uint32_t mask = __ballot_sync(GEN_FULL, hit);
      if (mask) {
        unsigned long long base = 0ull;
        if (lane == 0) base = atomicAdd(total, (unsigned long long)__popc(mask));
        base = __shfl_sync(GEN_FULL, base, 0);
        if (hit) {
          const unsigned long long pos = base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
          if (pos < cap) keys[pos] = (unsigned long long)i * t_data + (unsigned long long)j;
        }
      }
    }
  }
}

__global__ void gen_unpack_kernel(const unsigned long long *__restrict__ keys, long long n, unsigned long long t_data,
                                  int32_t *__restrict__ model_idx, int32_t *__restrict__ data_idx) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  model_idx[k] = (int32_t)(keys[k] / t_data);
  data_idx[k] = (int32_t)(keys[k] % t_data);
}

// semantic_clipper::estimate_tf (SC.cpp:122-138) for the 3 vertex pairs of one triangle match,
// data (query) -> model (reference), vertices paired in sorted-descriptor order.  R = V U^T of the
// 2 x 2 cross-covariance H is the orthogonal polar factor of H^T, which has a closed form: a rotation by
// atan2(H01 - H10, H00 + H11) when det H >= 0, otherwise the reflection [[c, s], [s, -c]] with angle
// atan2(H01 + H10, H00 - H11), whose second column the reference negates (SC.cpp:130-132).
__global__ void gen_kabsch_kernel(const double *__restrict__ tris_model6, const double *__restrict__ tris_data6,
                                  const int32_t *__restrict__ perm_model, const int32_t *__restrict__ perm_data,
                                  const int32_t *__restrict__ model_idx, const int32_t *__restrict__ data_idx, long long n,
                                  double *__restrict__ hyps4, double *__restrict__ pts_model, double *__restrict__ pts_data) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int mi = model_idx[k], di = data_idx[k];
  double a[6], b[6];
#pragma unroll
  for (int v = 0; v < 3; v++) {
    const int pd = perm_data[3 * (size_t)di + v], pm = perm_model[3 * (size_t)mi + v];
    a[2 * v] = tris_data6[6 * (size_t)di + 2 * pd]; a[2 * v + 1] = tris_data6[6 * (size_t)di + 2 * pd + 1];
    b[2 * v] = tris_model6[6 * (size_t)mi + 2 * pm]; b[2 * v + 1] = tris_model6[6 * (size_t)mi + 2 * pm + 1];
  }
  if (pts_model) {  // matched point lists of run_semantic_clipper (SC.cpp:102-105): 3 per match, sorted order
#pragma unroll
    for (int v = 0; v < 6; v++) { pts_model[6 * (size_t)k + v] = b[v]; pts_data[6 * (size_t)k + v] = a[v]; }
  }
  if (!hyps4) return;
  const double cax = ((a[0] + a[2]) + a[4]) / 3.0, cay = ((a[1] + a[3]) + a[5]) / 3.0;
  const double cbx = ((b[0] + b[2]) + b[4]) / 3.0, cby = ((b[1] + b[3]) + b[5]) / 3.0;
  double h00 = 0, h01 = 0, h10 = 0, h11 = 0;
#pragma unroll
  for (int v = 0; v < 3; v++) {
    const double ax = a[2 * v] - cax, ay = a[2 * v + 1] - cay, bx = b[2 * v] - cbx, by = b[2 * v + 1] - cby;
    h00 += ax * bx; h01 += ax * by; h10 += ay * bx; h11 += ay * by;
  }
  const double det = h00 * h11 - h01 * h10;
  double cr, sr;
  if (det >= 0) { cr = h00 + h11; sr = h01 - h10; }
  else          { cr = h00 - h11; sr = h01 + h10; }
  const double nrm = sqrt(cr * cr + sr * sr);
  double c = 1.0, s = 0.0;
  if (nrm > 0) { c = cr / nrm; s = sr / nrm; }
  hyps4[4 * k] = c;
  hyps4[4 * k + 1] = s;
  hyps4[4 * k + 2] = cbx - (c * cax - s * cay);
  hyps4[4 * k + 3] = cby - (s * cax + c * cay);
}

// ---------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------
cudaError_t spr_launch_gen_bin(const double *desc_data, int t_data, double inv_w, uint32_t n_bins, uint32_t *bin_start /* n_bins + 2 */,
                               uint32_t *bin_fill /* n_bins + 1 */, void *binned, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(bin_start, 0, ((size_t)n_bins + 2) * 4, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(bin_fill, 0, ((size_t)n_bins + 1) * 4, st);
  if (e != cudaSuccess) return e;
  if (t_data > 0) gen_bin_count_kernel<<<(t_data + 255) / 256, 256, 0, st>>>(desc_data, t_data, inv_w, n_bins, bin_start);
  gen_scan_u32_kernel<<<1, 1024, 0, st>>>(bin_start, (long long)n_bins + 1, nullptr);  // entry n_bins = total
  if (t_data > 0)
    gen_bin_scatter_kernel<<<(t_data + 255) / 256, 256, 0, st>>>(desc_data, t_data, inv_w, n_bins, bin_start, bin_fill,
                                                                  static_cast<GenBinned *>(binned));
  return cudaGetLastError();
}

size_t spr_gen_binned_bytes(int t_data) { return (size_t)(t_data > 0 ? t_data : 1) * sizeof(GenBinned); }

cudaError_t spr_launch_gen_match(const double *desc_model, int t_model, const void *binned, const uint32_t *bin_start, uint32_t n_bins,
                                 double inv_w, double window, double Tstar, const double *sig_model, const double *sig_data,
                                 int t_data, unsigned long long *keys, unsigned long long cap, unsigned long long *total,
                                 int sm_count, cudaStream_t st) {
  if (t_model <= 0) return cudaSuccess;
  const int want = (t_model + 7) / 8, capg = sm_count * 8;
  gen_match_kernel<<<want < capg ? want : capg, 256, 0, st>>>(desc_model, t_model, static_cast<const GenBinned *>(binned), bin_start,
                                                              n_bins, inv_w, window, Tstar, sig_model, sig_data,
                                                              (unsigned long long)t_data, keys, cap, total);
  return cudaGetLastError();
}

cudaError_t spr_launch_gen_unpack(const unsigned long long *keys, long long n, int t_data, int32_t *model_idx, int32_t *data_idx,
                                  cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  gen_unpack_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(keys, n, (unsigned long long)t_data, model_idx, data_idx);
  return cudaGetLastError();
}

cudaError_t spr_launch_gen_kabsch(const double *tris_model6, const double *tris_data6, const int32_t *perm_model,
                                  const int32_t *perm_data, const int32_t *model_idx, const int32_t *data_idx, long long n,
                                  double *hyps4, double *pts_model, double *pts_data, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  gen_kabsch_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(tris_model6, tris_data6, perm_model, perm_data, model_idx, data_idx, n,
                                                            hyps4, pts_model, pts_data);
  return cudaGetLastError();
}
