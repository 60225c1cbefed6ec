// spr_clipper.cu -- the SlideGraph half on the GPU: CLIPPER's pairwise-consistency (affinity)
// scoring and its dense-clique solver.
//
// Replaces (CSO = backend/sloam/clipper_semantic_object of the reference)
//   clipper::CLIPPER::scorePairwiseConsistency        CSO/src/clipper.cpp:21-65
//   clipper::invariants::EuclideanDistance            CSO/src/invariants/euclidean_distance.cpp:13-30
//   clipper::CLIPPER::findDenseClique                 CSO/src/clipper.cpp:172-323
// The reference fills a DENSE m x m MatrixXd with an OpenMP loop over the m(m-1)/2 pairs and then
// takes its sparseView(); at SlideGraph sizes (m = 3 x triangle matches, 5e4 for two 2000-landmark
// maps) the dense matrix is 18 GB.  Here the affinity matrix is built directly in CSR form -- both
// triangles, rows in ascending column order, no diagonal (the reference adds the identity
// implicitly, clipper.cpp:59-60) -- by a tiled all-pairs kernel:
//   * one CTA owns 32 rows (8 warps x 4 rows), the column points stream through shared memory in
//     tiles of 256, a lane keeps its column in registers across the warp's 4 rows;
//   * a conservative fp32 prefilter (|l1 - l2| against epsilon + a rigorous rounding margin) rejects
//     most pairs; survivors are scored in fp64 with the reference's expression, exp() included;
//   * hits are compacted with ballots, so every row comes out in ascending column order without a
//     sort; a count pass sizes the rows, a one-CTA scan turns the counts into row offsets, the fill
//     pass writes (col, val).
// The solver is ONE persistent cooperative kernel: the projected-gradient ascent with its three
// nested loops, line search and homotopy on d runs entirely on the device; every SpMV is a
// warp-per-row gather over the symmetric CSR, every scalar (sum, norm, dot) a fixed-order tree
// reduction finished redundantly by every CTA after a grid barrier -- no host round trip per
// iteration, deterministic for a given grid.  The constraint matrix C is the pattern of M
// (clipper.cpp:62-64), so one CSR serves both products.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

#include "spr_clipper.h"

namespace cg = cooperative_groups;

#define CLP_FULL 0xffffffffu
#define CLP_ROWS_PER_WARP 4
#define CLP_WARPS 8
#define CLP_ROWS_PER_CTA (CLP_ROWS_PER_WARP * CLP_WARPS)
#define CLP_TILE 256

namespace {

struct ClpBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

// one association = its two points (fp64, up to 3 coordinates), their fp32 copies relative to the
// clouds' centroids, and the point ids of the distinctness constraint (clipper.cpp:34-37)
struct ClpAssoc {
  double a[3], b[3];
  float af[3], bf[3];
  int32_t ia, ib;
};  // 80 bytes

}  // namespace

struct SprClipper {
  int m = 0, dim = 0;
  long long nnz = 0;          // entries of the symmetric CSR (twice the reference's upper-triangular count)
  ClpBuf d_assoc, d_cnt, d_rowptr, d_col, d_val, d_vec, d_part, d_scal, d_D1, d_D2, d_A;
  std::vector<int32_t> A;     // the associations scored (m x 2), clipper.cpp:24-25
};

// ---------------------------------------------------------------------------------------------
// affinity
// ---------------------------------------------------------------------------------------------
__global__ void clp_gather_kernel(const double *__restrict__ D1, const double *__restrict__ D2, int dim,
                                  const int32_t *__restrict__ A, int m, double c1x, double c1y, double c1z, double c2x,
                                  double c2y, double c2z, ClpAssoc *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int ia = A[2 * i], ib = A[2 * i + 1];
  ClpAssoc r;
  const double c1[3] = {c1x, c1y, c1z}, c2[3] = {c2x, c2y, c2z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r.a[k] = k < dim ? D1[(size_t)dim * ia + k] : 0.0;   // points are the columns of a dim x n column-major matrix
    r.b[k] = k < dim ? D2[(size_t)dim * ib + k] : 0.0;
    r.af[k] = (float)(r.a[k] - c1[k]);
    r.bf[k] = (float)(r.b[k] - c2[k]);
  }
  r.ia = ia; r.ib = ib;
  out[i] = r;
}

struct ClpScoreArgs {
  double epsilon, sigma2, mindist, affinityeps;  // EuclideanDistance::Params, clipper::Params::affinityeps
  float prefilter;                               // epsilon + fp32 rounding margin
  int m;
};

// EuclideanDistance::operator() (euclidean_distance.cpp:13-30), fp64, not fused
__device__ __forceinline__ double clp_exact_score(const ClpAssoc &x, const ClpAssoc &y, const ClpScoreArgs &P) {
  double s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) {  // unused coordinates are 0: adding 0*0 leaves the sum bit-identical
    const double e1 = __dsub_rn(x.a[k], y.a[k]), e2 = __dsub_rn(x.b[k], y.b[k]);
    s1 = __dadd_rn(s1, __dmul_rn(e1, e1));
    s2 = __dadd_rn(s2, __dmul_rn(e2, e2));
  }
  const double l1 = sqrt(s1), l2 = sqrt(s2);
  if (P.mindist > 0 && (l1 < P.mindist || l2 < P.mindist)) return 0.0;
  const double c = fabs(__dsub_rn(l1, l2));
  if (!(c < P.epsilon)) return 0.0;
  return exp(__dmul_rn(__dmul_rn(-0.5, c), c) / P.sigma2);
}

// FILL == false: cnt[i] = entries of row i.  FILL == true: writes (col, val) at rowptr[i] + position.
template <bool FILL>
__global__ void __launch_bounds__(CLP_WARPS * 32)
clp_affinity_kernel(const ClpAssoc *__restrict__ assoc, const ClpScoreArgs P, unsigned long long *__restrict__ cnt,
                    const unsigned long long *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ val) {
  __shared__ ClpAssoc tile[CLP_TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = P.m;
  const int row0 = blockIdx.x * CLP_ROWS_PER_CTA + warp * CLP_ROWS_PER_WARP;
  ClpAssoc rows[CLP_ROWS_PER_WARP];
  unsigned long long pos[CLP_ROWS_PER_WARP];
#pragma unroll
  for (int r = 0; r < CLP_ROWS_PER_WARP; r++) {
    const int i = row0 + r;
    rows[r] = assoc[i < m ? i : m - 1];
    pos[r] = FILL && i < m ? rowptr[i] : 0ull;
  }
  for (int j0 = 0; j0 < m; j0 += CLP_TILE) {
    __syncthreads();
    {  // stage the tile: 80-byte records as 16-byte pieces, coalesced
      const int n_rec = min(CLP_TILE, m - j0);
      const uint4 *src = reinterpret_cast<const uint4 *>(assoc + j0);
      uint4 *dst = reinterpret_cast<uint4 *>(tile);
      for (int k = threadIdx.x; k < n_rec * 5; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    for (int jj = 0; jj < CLP_TILE && j0 + jj < m; jj += 32) {
      const int j = j0 + jj + lane;
      const bool in = j < m;
      const ClpAssoc &y = tile[in ? jj + lane : 0];
#pragma unroll
      for (int r = 0; r < CLP_ROWS_PER_WARP; r++) {
        const int i = row0 + r;
        double scr = 0.0;
        if (in && i < m && i != j && rows[r].ia != y.ia && rows[r].ib != y.ib) {   // clipper.cpp:34-37
          const float ax = rows[r].af[0] - y.af[0], ay = rows[r].af[1] - y.af[1], az = rows[r].af[2] - y.af[2];
          const float bx = rows[r].bf[0] - y.bf[0], by = rows[r].bf[1] - y.bf[1], bz = rows[r].bf[2] - y.bf[2];
          const float l1 = sqrtf(ax * ax + ay * ay + az * az), l2 = sqrtf(bx * bx + by * by + bz * bz);
          if (fabsf(l1 - l2) < P.prefilter) scr = clp_exact_score(rows[r], y, P);
        }
        const bool hit = scr > P.affinityeps;                                      // clipper.cpp:52-54
        const unsigned mask = __ballot_sync(CLP_FULL, hit);
        if (FILL && hit) {
          const unsigned long long q = pos[r] + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
          col[q] = j;
          val[q] = scr;
        }
        pos[r] += (unsigned long long)__popc(mask);
      }
    }
  }
  if (!FILL && lane == 0) {
#pragma unroll
    for (int r = 0; r < CLP_ROWS_PER_WARP; r++)
      if (row0 + r < m) cnt[row0 + r] = pos[r];
  }
}

// exclusive prefix sum of n counters by one CTA; offsets[n] = total
__global__ void __launch_bounds__(1024) clp_scan_kernel(const unsigned long long *__restrict__ counts, int n,
                                                        unsigned long long *__restrict__ offsets) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0ull;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < n ? counts[i] : 0ull;
    unsigned long long s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long t = __shfl_up_sync(CLP_FULL, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(CLP_FULL, w, d);
        if (lane >= d) w += t;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const unsigned long long before = carry + (warp ? warp_tot[warp - 1] : 0ull) + s - v;
    if (i < n) offsets[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// ---------------------------------------------------------------------------------------------
// dense-clique solver (clipper.cpp:172-323): one persistent cooperative kernel
// ---------------------------------------------------------------------------------------------
struct ClpSolveArgs {
  const unsigned long long *rowptr;
  const int32_t *col;
  const double *val;
  int n;
  double tol_u, tol_F, beta, eps;
  int maxiniters, maxoliters, maxlsiters, rescale_u0;
  double *vec;    // 9 vectors of length n (see clp_solve_kernel); u0 is uploaded into the third one
  double *uout;   // [n] the final u
  double *part;   // [2][gridDim.x][4] per-CTA partial sums, double-buffered
  double *scal;   // out: [0] F, [1] ifinal, [2] d, [3] line-search steps (diagnostic)
};

// CTA-wide sum of up to 4 values; the result is valid in thread 0 (fixed tree: deterministic)
template <int K>
__device__ __forceinline__ void clp_block_sum(double (&x)[K], double *smem /* [K][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x[k] += __shfl_xor_sync(CLP_FULL, x[k], d);
    if (lane == 0) smem[k * 32 + warp] = x[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      double v = lane < nw ? smem[k * 32 + lane] : 0.0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(CLP_FULL, v, d);
      x[k] = v;
    }
  }
  __syncthreads();
}

// grid-wide sum of K per-thread values: per-CTA partials -> grid barrier -> every CTA adds the
// partials of all CTAs in the same order, so all CTAs hold bit-identical totals.  The partials
// alternate between two buffers (`phase`), so one barrier per reduction is enough: a CTA can only
// overwrite a buffer after every CTA has passed the barrier of the reduction in between.
template <int K>
__device__ __forceinline__ void clp_grid_sum(cg::grid_group &grid, double (&x)[K], double *part, unsigned &phase, double *smem) {
  clp_block_sum<K>(x, smem);
  double *buf = part + (size_t)(phase & 1u) * gridDim.x * 4;
  phase++;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) buf[(size_t)blockIdx.x * 4 + k] = x[k];
  }
  grid.sync();
  double t[K];
#pragma unroll
  for (int k = 0; k < K; k++) t[k] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; k++) t[k] += __ldcg(buf + (size_t)b * 4 + k);
  }
  clp_block_sum<K>(t, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) smem[128 + k] = t[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) x[k] = smem[128 + k];
  __syncthreads();
}

// (M x)_i and (C x)_i (pattern sum) of row i by one warp; x_j = scale * xin[j]
__device__ __forceinline__ void clp_row_products(const ClpSolveArgs &A, int i, const double *xin, double scale, int lane,
                                                 double *mx, double *cx) {
  const unsigned long long b = A.rowptr[i], e = A.rowptr[i + 1];
  double sm = 0.0, sc = 0.0;
  for (unsigned long long k = b + lane; k < e; k += 32) {
    const double xj = __ldcg(xin + A.col[k]) * scale;
    sm += A.val[k] * xj;
    sc += xj;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sm += __shfl_xor_sync(CLP_FULL, sm, d);
    sc += __shfl_xor_sync(CLP_FULL, sc, d);
  }
  *mx = sm; *cx = sc;
}

__global__ void __launch_bounds__(256) clp_solve_kernel(const ClpSolveArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double smem[160];
  const int n = A.n;
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  // u, un: current / candidate point; v: unnormalised candidate (u0 on entry); g, gn: gradients;
  // Mx, Cx (Mxn, Cxn): M u and C u at the current (candidate) point
  double *u = A.vec, *un = u + n, *v = un + n, *g = v + n, *gn = g + n, *Mx = gn + n, *Cx = Mx + n, *Mxn = Cx + n, *Cxn = Mxn + n;
  unsigned phase = 0;

  auto products = [&](const double *xin, double scale, double *mxo, double *cxo) {
    for (int i = gwarp; i < n; i += nwarps) {
      double mx, cx;
      clp_row_products(A, i, xin, scale, lane, &mx, &cx);
      if (lane == 0) { mxo[i] = mx; cxo[i] = cx; }
    }
    grid.sync();
  };

  // u = M u0 + u0 (one power step, clipper.cpp:195-199), then u /= u.norm() (clipper.cpp:200)
  if (A.rescale_u0) {
    products(v, 1.0, Mxn, Cxn);
    for (int i = gtid; i < n; i += nthreads) u[i] = Mxn[i] + v[i];
  } else {
    for (int i = gtid; i < n; i += nthreads) u[i] = v[i];
  }
  double su;
  {
    double s[2] = {0.0, 0.0};
    for (int i = gtid; i < n; i += nthreads) { s[0] += u[i] * u[i]; s[1] += u[i]; }
    clp_grid_sum<2>(grid, s, A.part, phase, smem);
    const double nu = sqrt(s[0]);
    for (int i = gtid; i < n; i += nthreads) u[i] = u[i] / nu;
    su = 0.0;
  }
  grid.sync();
  {
    double s[1] = {0.0};
    for (int i = gtid; i < n; i += nthreads) s[0] += u[i];
    clp_grid_sum<1>(grid, s, A.part, phase, smem);
    su = s[0];
  }
  products(u, 1.0, Mx, Cx);

  // homotopy (clipper.cpp:203-212, 271-287): mean over {Cbu > eps and u > eps} of (M u + u) / Cbu,
  // Cbu = sum(u) - C u - u
  auto homotopy = [&](bool use_abs, double *out) -> bool {
    double s[2] = {0.0, 0.0};
    for (int i = gtid; i < n; i += nthreads) {
      const double cbu = su - Cx[i] - u[i];
      if (cbu > A.eps && u[i] > A.eps) {
        const double q = (Mx[i] + u[i]) / cbu;
        s[0] += use_abs ? fabs(q) : q;
        s[1] += 1.0;
      }
    }
    clp_grid_sum<2>(grid, s, A.part, phase, smem);
    if (s[1] == 0.0) return false;
    *out = s[0] / s[1];
    return true;
  };

  double d = 0.0;
  { double t; if (homotopy(false, &t)) d = t; }
  double F = 0.0;
  long long ls_steps = 0;
  int it;
  for (it = 0; it < A.maxoliters; ++it) {
    {  // gradF = (1 + d) u - d sum(u) + M u + d C u;  F = u . gradF   (clipper.cpp:222-223)
      double s[1] = {0.0};
      for (int i = gtid; i < n; i += nthreads) {
        const double gi = (1.0 + d) * u[i] - d * su + Mx[i] + Cx[i] * d;
        g[i] = gi;
        s[0] += u[i] * gi;
      }
      clp_grid_sum<1>(grid, s, A.part, phase, smem);
      F = s[0];
    }
    for (int j = 0; j < A.maxiniters; ++j) {
      double alpha = 1.0, Fnew = 0.0, deltaF = 0.0, deltau = 0.0, sun = 0.0;
      for (int k = 0; k < A.maxlsiters; ++k) {                                   // clipper.cpp:237-255
        ls_steps++;
        double s[2] = {0.0, 0.0};
        for (int i = gtid; i < n; i += nthreads) {                               // gradient step, projection on the positive orthant
          const double t = u[i] + alpha * g[i];
          const double w = t > 0.0 ? t : 0.0;
          v[i] = w;
          s[0] += w * w;
          s[1] += w;
        }
        clp_grid_sum<2>(grid, s, A.part, phase, smem);
        const double inv = s[0] > 0.0 ? 1.0 / sqrt(s[0]) : 1.0;                  // normalize(): a zero vector stays as it is
        sun = s[1] * inv;
        products(v, inv, Mxn, Cxn);
        double q[2] = {0.0, 0.0};
        for (int i = gtid; i < n; i += nthreads) {
          const double x = v[i] * inv;
          const double gi = (1.0 + d) * x - d * sun + Mxn[i] + Cxn[i] * d;       // clipper.cpp:241-244
          un[i] = x;
          gn[i] = gi;
          q[0] += x * gi;
          const double e = x - u[i];
          q[1] += e * e;
        }
        clp_grid_sum<2>(grid, q, A.part, phase, smem);
        Fnew = q[0];
        deltau = sqrt(q[1]);
        deltaF = Fnew - F;
        if (deltaF < -A.eps) alpha = alpha * A.beta;
        else break;
      }
      // accept the candidate (clipper.cpp:259-261): the buffers swap roles.  Every thread holds the
      // same scalars, so all CTAs take the same branches.
      F = Fnew;
      su = sun;
      { double *t = u; u = un; un = t; }
      { double *t = g; g = gn; gn = t; }
      { double *t = Mx; Mx = Mxn; Mxn = t; }
      { double *t = Cx; Cx = Cxn; Cxn = t; }
      if (deltau < A.tol_u || fabs(deltaF) < A.tol_F) break;                      // clipper.cpp:264
    }
    double dd;
    if (homotopy(true, &dd)) d += dd;                                            // clipper.cpp:271-287
    else break;
  }
  grid.sync();
  for (int i = gtid; i < n; i += nthreads) A.uout[i] = u[i];
  if (gtid == 0) { A.scal[0] = F; A.scal[1] = (double)it; A.scal[2] = d; A.scal[3] = (double)ls_steps; }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define CLP_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      err = std::string(#call) + ": " + cudaGetErrorString(e__);           \
      return SLIDE_PR_ERR_CUDA;                                            \
    }                                                                      \
  } while (0)

SprClipper *spr_clipper_create() { return new SprClipper(); }

void spr_clipper_destroy(SprClipper *c) {
  if (!c) return;
  for (ClpBuf *b : {&c->d_assoc, &c->d_cnt, &c->d_rowptr, &c->d_col, &c->d_val, &c->d_vec, &c->d_part, &c->d_scal, &c->d_D1,
                    &c->d_D2, &c->d_A})
    b->release();
  delete c;
}

int spr_clipper_size(const SprClipper *c, int *m, long long *nnz_sym) {
  if (m) *m = c->m;
  if (nnz_sym) *nnz_sym = c->nnz;
  return SLIDE_PR_OK;
}

const int32_t *spr_clipper_associations(const SprClipper *c) { return c->A.data(); }

int spr_clipper_score(SprClipper *c, const slide_clipper_params &p, const double *D1, int n1, const double *D2, int n2,
                      int dim, const int32_t *A_in, int m, const double *device_hint, int sm_count, cudaStream_t st,
                      long long *nnz_upper, float *kernel_ms, std::string &err) {
  const bool from_device = device_hint != nullptr;
  if (dim < 1 || dim > 3) { err = "the EuclideanDistance invariant is built for 1 to 3 dimensions"; return SLIDE_PR_ERR_UNSUPPORTED; }
  if (n1 < 0 || n2 < 0) { err = "negative point count"; return SLIDE_PR_ERR_INVALID; }
  c->m = 0; c->nnz = 0; c->dim = dim;
  if (nnz_upper) *nnz_upper = 0;
  if (kernel_ms) *kernel_ms = 0.f;
  // associations: given, or all-to-all (utils.h:60-70)
  if (!A_in) {
    if ((long long)n1 * n2 > 0x7fffffffLL / 2) { err = "all-to-all hypothesis exceeds 2^30 associations"; return SLIDE_PR_ERR_UNSUPPORTED; }
    m = n1 * n2;
    c->A.resize(2 * (size_t)m);
    for (int i = 0; i < n1; i++)
      for (int j = 0; j < n2; j++) { c->A[2 * ((size_t)j + (size_t)i * n2)] = i; c->A[2 * ((size_t)j + (size_t)i * n2) + 1] = j; }
  } else {
    if (m < 0) { err = "negative association count"; return SLIDE_PR_ERR_INVALID; }
    c->A.assign(A_in, A_in + 2 * (size_t)m);
    for (int i = 0; i < m; i++)
      if (c->A[2 * i] < 0 || c->A[2 * i] >= n1 || c->A[2 * i + 1] < 0 || c->A[2 * i + 1] >= n2) { err = "association index out of range"; return SLIDE_PR_ERR_INVALID; }
  }
  c->m = m;
  if (m == 0) return SLIDE_PR_OK;
  // centroids (fp32 prefilter works on centred coordinates) and the extent that scales its margin
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0}, R = 0;
  if (from_device) {
    // the points are already on the device (gathered per association by the generator); the caller knows a
    // centre for each cloud and a bound on the distance of any point from it (the maps' bounding boxes)
    for (int k = 0; k < 3; k++) { c1[k] = device_hint[k]; c2[k] = device_hint[3 + k]; }
    R = device_hint[6];
    if (!std::isfinite(R)) { err = "non-finite coordinate in a dataset"; return SLIDE_PR_ERR_NONFINITE; }
  } else {
    for (int k = 0; k < dim; k++) {
      for (int i = 0; i < n1; i++) c1[k] += D1[(size_t)dim * i + k];
      for (int i = 0; i < n2; i++) c2[k] += D2[(size_t)dim * i + k];
      c1[k] /= (double)std::max(n1, 1); c2[k] /= (double)std::max(n2, 1);
    }
    for (int k = 0; k < dim; k++) {
      for (int i = 0; i < n1; i++) {
        const double v = D1[(size_t)dim * i + k];
        if (!std::isfinite(v)) { err = "non-finite coordinate in dataset 1"; return SLIDE_PR_ERR_NONFINITE; }
        R = std::max(R, std::fabs(v - c1[k]));
      }
      for (int i = 0; i < n2; i++) {
        const double v = D2[(size_t)dim * i + k];
        if (!std::isfinite(v)) { err = "non-finite coordinate in dataset 2"; return SLIDE_PR_ERR_NONFINITE; }
        R = std::max(R, std::fabs(v - c2[k]));
      }
    }
  }
  const double *dD1 = D1, *dD2 = D2;
  if (!from_device) {
    CLP_CUDA(c->d_D1.ensure(std::max<size_t>((size_t)n1 * dim, 1) * 8));
    CLP_CUDA(c->d_D2.ensure(std::max<size_t>((size_t)n2 * dim, 1) * 8));
    CLP_CUDA(cudaMemcpyAsync(c->d_D1.p, D1, (size_t)n1 * dim * 8, cudaMemcpyHostToDevice, st));
    CLP_CUDA(cudaMemcpyAsync(c->d_D2.p, D2, (size_t)n2 * dim * 8, cudaMemcpyHostToDevice, st));
    dD1 = c->d_D1.as<double>(); dD2 = c->d_D2.as<double>();
  }
  CLP_CUDA(c->d_A.ensure((size_t)m * 8));
  CLP_CUDA(cudaMemcpyAsync(c->d_A.p, c->A.data(), (size_t)m * 8, cudaMemcpyHostToDevice, st));
  CLP_CUDA(c->d_assoc.ensure((size_t)m * sizeof(ClpAssoc)));
  CLP_CUDA(c->d_cnt.ensure((size_t)m * 8));
  CLP_CUDA(c->d_rowptr.ensure(((size_t)m + 1) * 8));
  cudaEvent_t e0, e1;
  CLP_CUDA(cudaEventCreate(&e0)); CLP_CUDA(cudaEventCreate(&e1));
  CLP_CUDA(cudaEventRecord(e0, st));
  clp_gather_kernel<<<(m + 255) / 256, 256, 0, st>>>(dD1, dD2, dim, c->d_A.as<int32_t>(), m, c1[0], c1[1], c1[2], c2[0], c2[1],
                                                    c2[2], c->d_assoc.as<ClpAssoc>());
  ClpScoreArgs S;
  S.epsilon = p.epsilon; S.sigma2 = p.sigma * p.sigma; S.mindist = p.mindist; S.affinityeps = p.affinityeps; S.m = m;
  // |fl32 path - exact| <= 44 * 2^-24 * R for either length difference (conversion, subtraction,
  // squares, sum, square root); 64 * 2^-24 * R is used, and the float threshold is rounded up
  const double margin = 64.0 * std::ldexp(1.0, -24) * R;
  float pf = (float)(p.epsilon + margin);
  if ((double)pf < p.epsilon + margin) pf = std::nextafterf(pf, INFINITY);
  if (!(p.epsilon == p.epsilon)) pf = 0.f;  // NaN epsilon: nothing is consistent
  S.prefilter = pf;
  const int grid = (m + CLP_ROWS_PER_CTA - 1) / CLP_ROWS_PER_CTA;
  clp_affinity_kernel<false><<<grid, CLP_WARPS * 32, 0, st>>>(c->d_assoc.as<ClpAssoc>(), S, c->d_cnt.as<unsigned long long>(), nullptr,
                                                              nullptr, nullptr);
  clp_scan_kernel<<<1, 1024, 0, st>>>(c->d_cnt.as<unsigned long long>(), m, c->d_rowptr.as<unsigned long long>());
  unsigned long long total = 0;
  CLP_CUDA(cudaMemcpyAsync(&total, c->d_rowptr.as<unsigned long long>() + m, 8, cudaMemcpyDeviceToHost, st));
  CLP_CUDA(cudaStreamSynchronize(st));
  c->nnz = (long long)total;
  CLP_CUDA(c->d_col.ensure(std::max<size_t>((size_t)total, 1) * 4));
  CLP_CUDA(c->d_val.ensure(std::max<size_t>((size_t)total, 1) * 8));
  clp_affinity_kernel<true><<<grid, CLP_WARPS * 32, 0, st>>>(c->d_assoc.as<ClpAssoc>(), S, nullptr, c->d_rowptr.as<unsigned long long>(),
                                                             c->d_col.as<int32_t>(), c->d_val.as<double>());
  CLP_CUDA(cudaEventRecord(e1, st));
  CLP_CUDA(cudaStreamSynchronize(st));
  CLP_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (kernel_ms) *kernel_ms = ms;
  if (nnz_upper) *nnz_upper = (long long)(total / 2);  // the matrix is symmetric with an empty diagonal
  (void)sm_count;
  return SLIDE_PR_OK;
}

int spr_clipper_get_csr(SprClipper *c, int64_t *row_ptr, int32_t *col, double *val, long long cap, cudaStream_t st,
                        std::string &err) {
  if (!row_ptr) { err = "row_ptr is NULL"; return SLIDE_PR_ERR_INVALID; }
  if (c->m == 0) { row_ptr[0] = 0; return SLIDE_PR_OK; }
  CLP_CUDA(cudaMemcpyAsync(row_ptr, c->d_rowptr.p, ((size_t)c->m + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (col && val) {
    if (cap < c->nnz) { err = "CSR capacity too small"; return SLIDE_PR_ERR_INVALID; }
    CLP_CUDA(cudaMemcpyAsync(col, c->d_col.p, (size_t)c->nnz * 4, cudaMemcpyDeviceToHost, st));
    CLP_CUDA(cudaMemcpyAsync(val, c->d_val.p, (size_t)c->nnz * 8, cudaMemcpyDeviceToHost, st));
  }
  CLP_CUDA(cudaStreamSynchronize(st));
  return SLIDE_PR_OK;
}

namespace {

// utils::findIndicesOfkLargest (utils.cpp:34-55): a min-heap of (value, index) pairs; the result
// lists the kept entries in descending (value, index) order
std::vector<int> k_largest(const double *x, int n, int k) {
  using T = std::pair<double, int>;
  if (k < 1) return {};
  std::priority_queue<T, std::vector<T>, std::greater<T>> q;
  for (int i = 0; i < n; i++) {
    if ((int)q.size() < k) q.push({x[i], i});
    else if (q.top().first < x[i]) { q.pop(); q.push({x[i], i}); }
  }
  std::vector<int> idx(q.size());  // the reference pops k times even from a smaller heap (undefined there)
  for (size_t i = idx.size(); i-- > 0;) { idx[i] = q.top().second; q.pop(); }
  return idx;
}

// Goldberg's densest subgraph by parametric min-cut (dsd::solve, dsd.cpp:167-326) on the subgraph
// induced by S.  Written as an iterative Dinic on a compact arc list; the bisection on the density
// g, its bounds and its stopping rule are the reference's, so the returned cut is the same set.
struct MaxFlow {
  struct Arc { int to; double cap; };
  std::vector<Arc> arcs;
  std::vector<std::vector<int>> adj;
  std::vector<int> level, it;
  explicit MaxFlow(int n) : adj(n), level(n), it(n) {}
  void add(int u, int v, double cuv, double cvu) {
    adj[u].push_back((int)arcs.size()); arcs.push_back({v, cuv});
    adj[v].push_back((int)arcs.size()); arcs.push_back({u, cvu});
  }
  bool bfs(int s, int t) {
    std::fill(level.begin(), level.end(), -1);
    std::vector<int> q{s};
    level[s] = 0;
    for (size_t h = 0; h < q.size(); h++)
      for (int a : adj[q[h]])
        if (arcs[a].cap > 0 && level[arcs[a].to] < 0) { level[arcs[a].to] = level[q[h]] + 1; q.push_back(arcs[a].to); }
    return level[t] >= 0;
  }
  double dfs(int s, int t) {  // one augmenting path along the level graph, iterative
    std::vector<int> path;
    int u = s;
    for (;;) {
      if (u == t) {
        double f = HUGE_VAL;
        for (int a : path) f = std::min(f, arcs[a].cap);
        for (int a : path) { arcs[a].cap -= f; arcs[a ^ 1].cap += f; }
        return f;
      }
      bool adv = false;
      for (int &k = it[u]; k < (int)adj[u].size(); k++) {
        const int a = adj[u][k];
        if (arcs[a].cap > 0 && level[arcs[a].to] == level[u] + 1) { path.push_back(a); u = arcs[a].to; adv = true; break; }
      }
      if (adv) continue;
      if (path.empty()) return 0;
      level[u] = -1;  // dead end
      const int a = path.back();
      path.pop_back();
      u = arcs[a ^ 1].to;
    }
  }
  void run(int s, int t) {
    while (bfs(s, t)) {
      std::fill(it.begin(), it.end(), 0);
      while (dfs(s, t) > 0) {}
    }
  }
  std::vector<char> source_side(int s) {
    std::vector<char> seen(adj.size(), 0);
    std::vector<int> q{s};
    seen[s] = 1;
    for (size_t h = 0; h < q.size(); h++)
      for (int a : adj[q[h]])
        if (arcs[a].cap > 0 && !seen[arcs[a].to]) { seen[arcs[a].to] = 1; q.push_back(arcs[a].to); }
    return seen;
  }
};

std::vector<int> densest_subgraph(int n, const std::vector<int> &S, const std::vector<int64_t> &rp, const std::vector<int32_t> &col,
                                  const std::vector<double> &val) {
  const int ns = (int)S.size();
  const long long m = (long long)ns * ns - ns;                                  // dsd.cpp:292
  std::vector<int> pos(n, -1);
  for (int a = 0; a < ns; a++) pos[S[a]] = a;
  std::vector<double> degree(n, 0.0);
  for (int i : S)
    for (int64_t k = rp[i]; k < rp[i + 1]; k++)
      if (pos[col[k]] >= 0) degree[i] += val[k];
  double L = 0, U = (double)(m / 2);                                            // dsd.cpp:191-192
  std::vector<char> final_cut(n + 2, 0);
  const double half = (double)(m / 2);
  while ((double)n * (n - 1) * (U - L) >= 1) {                                  // dsd.cpp:213
    const double g = (U + L) / 2;
    MaxFlow F(n + 2);
    const int src = 0, dst = n + 1;
    // the reference builds a COMPLETE weighted graph on S (zero weights included): zero arcs carry no flow
    for (int i : S)
      for (int64_t k = rp[i]; k < rp[i + 1]; k++)
        if (pos[col[k]] >= 0 && i < col[k]) F.add(i + 1, col[k] + 1, val[k], val[k]);
    for (int v = 0; v < n; v++) {                                               // dsd.cpp:21-37
      F.add(src, v + 1, half, 0.0);
      F.add(v + 1, dst, half + 2 * g - degree[v], 0.0);
    }
    F.run(src, dst);
    std::vector<char> cut = F.source_side(src);
    int cs = 0;
    for (char b : cut) cs += b;
    if (cs == 1) U = g;                                                         // dsd.cpp:222-229
    else { L = g; final_cut = cut; }
  }
  std::vector<int> nodes;
  for (int v = 0; v < n; v++)
    if (final_cut[v + 1]) nodes.push_back(v);
  return nodes;
}

}  // namespace

int spr_clipper_solve(SprClipper *c, const slide_clipper_params &p, const double *u0, int sm_count, cudaStream_t st,
                      int32_t *nodes_out, int32_t cap, slide_clipper_solution *sol, double *u_out, std::string &err) {
  std::memset(sol, 0, sizeof(*sol));
  const int n = c->m;
  if (n <= 0) return SLIDE_PR_OK;
  if (!u0) { err = "u0 is NULL"; return SLIDE_PR_ERR_INVALID; }
  CLP_CUDA(c->d_vec.ensure((size_t)n * 10 * 8));
  CLP_CUDA(c->d_scal.ensure(8 * 8));
  double *vec = c->d_vec.as<double>();
  CLP_CUDA(cudaMemcpyAsync(vec + 2 * (size_t)n, u0, (size_t)n * 8, cudaMemcpyHostToDevice, st));  // u0 -> v
  int per_sm = 0;
  CLP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, clp_solve_kernel, 256, 0));
  if (per_sm < 1) { err = "the solver kernel does not fit on an SM"; return SLIDE_PR_ERR_CUDA; }
  // enough warps for a row each, never more CTAs than can be co-resident (cooperative launch)
  int grid = std::min(sm_count * std::min(per_sm, 2), std::max(1, (n + 7) / 8));
  CLP_CUDA(c->d_part.ensure((size_t)grid * 4 * 2 * 8));
  ClpSolveArgs A;
  A.rowptr = c->d_rowptr.as<unsigned long long>(); A.col = c->d_col.as<int32_t>(); A.val = c->d_val.as<double>();
  A.n = n;
  A.tol_u = p.tol_u; A.tol_F = p.tol_F; A.beta = p.beta; A.eps = p.eps;
  A.maxiniters = p.maxiniters; A.maxoliters = p.maxoliters; A.maxlsiters = p.maxlsiters; A.rescale_u0 = p.rescale_u0;
  A.vec = vec; A.uout = vec + 9 * (size_t)n; A.part = c->d_part.as<double>(); A.scal = c->d_scal.as<double>();
  void *args[] = {&A};
  cudaEvent_t e0, e1;
  CLP_CUDA(cudaEventCreate(&e0)); CLP_CUDA(cudaEventCreate(&e1));
  CLP_CUDA(cudaEventRecord(e0, st));
  CLP_CUDA(cudaLaunchCooperativeKernel((void *)clp_solve_kernel, dim3(grid), dim3(256), args, 0, st));
  CLP_CUDA(cudaEventRecord(e1, st));
  std::vector<double> u(n);
  double scal[4] = {0, 0, 0, 0};
  CLP_CUDA(cudaMemcpyAsync(u.data(), A.uout, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CLP_CUDA(cudaMemcpyAsync(scal, A.scal, sizeof(scal), cudaMemcpyDeviceToHost, st));
  CLP_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  sol->score = scal[0]; sol->ifinal = (int32_t)scal[1]; sol->d = scal[2]; sol->line_search_steps = (int64_t)scal[3];
  sol->kernel_ms = ms;
  if (u_out) std::memcpy(u_out, u.data(), (size_t)n * 8);
  // rounding (clipper.cpp:293-316) on the host: O(n log n) on one vector
  std::vector<int> nodes;
  if (p.rounding == SLIDE_CLIPPER_ROUND_NONZERO) {
    for (int i = 0; i < n; i++) if (u[i] > 0.0) nodes.push_back(i);             // utils.cpp:59-69
  } else if (p.rounding == SLIDE_CLIPPER_ROUND_DSD) {
    std::vector<int> S;
    for (int i = 0; i < n; i++) if (u[i] > 0.0) S.push_back(i);
    std::vector<int64_t> rp((size_t)n + 1);
    std::vector<int32_t> col((size_t)std::max<long long>(c->nnz, 1));
    std::vector<double> val((size_t)std::max<long long>(c->nnz, 1));
    const int rc = spr_clipper_get_csr(c, rp.data(), col.data(), val.data(), c->nnz, st, err);
    if (rc != SLIDE_PR_OK) return rc;
    nodes = densest_subgraph(n, S, rp, col, val);
  } else {
    nodes = k_largest(u.data(), n, (int)std::round(scal[0]));                     // clipper.cpp:312-315
  }
  sol->n_nodes = (int32_t)std::min<size_t>(nodes.size(), (size_t)std::max(cap, 0));
  if (nodes_out) for (int i = 0; i < sol->n_nodes; i++) nodes_out[i] = nodes[i];
  if ((int)nodes.size() > cap) { err = "nodes_out capacity too small"; return SLIDE_PR_ERR_INVALID; }
  return SLIDE_PR_OK;
}
