// spr_kernels.cu -- sm_100a kernels of the SlideMatch lattice search, lattice engine (the library's
// second engine; the default search is the pair-join scorer, spr_join.cu).
//
// Replaces the five nested loops of PlaceRecognition::MatchMaps (place_recognition.cpp:178-372):
// the EXACT inlier count of hypotheses -- all of them in the exhaustive mode, the candidates that
// survive the bound phase (spr_kernels_bound.cu) in the bound-and-verify mode.
// Work decomposition:
//   * the search runs as one PASS per (label, bitmap direction): all CTAs work on the same
//     occupancy plane at the same time, so the plane and its rank tables are staged once per CTA
//     into SHARED MEMORY with TMA bulk copies when they fit (lazily, by the first warp that has
//     work; otherwise they are read through L1/L2); per-hypothesis inlier counters are carried
//     between the passes in HBM (2 bytes per hypothesis);
//   * one THREAD owns one chunk = 32 consecutive lattice translations along one axis, for one
//     yaw candidate; a WARP owns a group of 32 chunks (1024 hypotheses) and is the unit of
//     scheduling: warps pull (yaw, group) work items from a global counter -- or from the
//     candidate list of the verification phase -- there is no barrier after staging;
//   * in the verification phase a lane first rebuilds its candidate mask by comparing the bit planes
//     of the bounds with the running best and drops the other hypotheses;
//   * query landmarks come in groups of 8 (one label, Morton order) with a bounding box per yaw;
//     the visibility of 32 groups is tested at once (one group per lane, one ballot) against the
//     1024 translations of the warp, only the visible groups are walked;
//   * for every remaining query landmark the thread reads two words of the plane and obtains the
//     32 hypotheses' filter bits with one funnel shift (spr_probe);
//   * probes with set bits ("filter hits") are compacted through a per-warp shared-memory queue
//     (one record per (chunk, query), slots from ballots) and verified 32 records at a time in
//     exact, non-fused fp64 against the cells' candidates (spr_verify_mask) -- the reference's own
//     predicate, so every verified hypothesis gets its exact inlier count;
//   * after the last pass the best (count, canonical index) key is reduced with shuffles and one
//     64-bit atomicMax per warp.
// This is integer/bit and fp64 ALU work on shared-memory / L2-resident data; there is no GEMM in
// it, so no tensor-core path (BASELINE.json north_star).
#include <cuda_runtime.h>
#include <stdint.h>

#include "spr_core.h"
#include "spr_kernels.h"

#define SPR_QHALF 4                      // queries probed per push
#define SPR_QCAP (32 * SPR_QHALF + 32)   // per-warp queue, records: residual < 32 + one push
#define SPR_FULL 0xffffffffu
#define SPR_SMEM_LIMIT (227 * 1024)

// Per-hypothesis inlier counters of one warp's work item, [32 lanes][32 bits].  Counts are
// bounded by the number of query landmarks, so they are packed two per word (pitch 17 words per
// lane) unless the query map has more than 65535 landmarks (pitch 33 words per lane).
#define SPR_CNT_WORDS(CNT32) ((CNT32) ? 32 * 33 : 32 * 17)
#define SPR_WARP_WORDS(CNT32) (SPR_CNT_WORDS(CNT32) + 4 * SPR_QCAP)  // multiple of 4 words (16 B)

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

// Byte sizes / word offsets of the tables staged into the CTA's shared memory.  Every part is a
// multiple of 16 bytes and starts 16-byte aligned, as cp.async.bulk requires; the per-label slices
// of cellref / reftab start at arbitrary elements, so they are copied from the enclosing aligned
// window (`*_skip` leading elements are skipped by the pointer the kernel uses).
struct SprTabLayout {
  uint32_t reftab_w, bits_w, r16_w, rr_w, cellref_w, total_w;        // word offsets
  uint32_t reftab_b, bits_b, r16_b, rr_b, cellref_b;                 // bytes to copy
  uint32_t reftab_skip, cellref_skip;                                // leading elements (doubles / u16) to skip
};
// band_rows rows of the plane (a multiple of 8, starting at a multiple of 8) + one all-zero row behind them;
// cell_first / cells: absolute rank of the band's first marked cell and their number; refs == 0: the
// landmark table is not staged (read in place).
// zero_row: 0 for a whole plane (its own first / last rows are the zero rows).
__host__ __device__ static inline SprTabLayout spr_tab_layout(uint32_t W, uint32_t band_rows, uint32_t cell_first, uint32_t cells,
                                                              uint32_t ref_base, uint32_t refs, uint32_t zero_row) {
  SprTabLayout o;
  o.reftab_skip = 5u * (ref_base & 1u);
  o.cellref_skip = cell_first & 7u;
  o.reftab_b = refs ? ((5u * ((ref_base & 1u) + refs) * 8u) + 15u) & ~15u : 0u;
  o.bits_b = band_rows * W * 4u;
  o.r16_b = band_rows * W * 2u;
  o.rr_b = band_rows * 4u;
  o.cellref_b = (((cell_first & 7u) + cells) * 2u + 15u) & ~15u;
  uint32_t w = 0;
  o.reftab_w = w;  w += o.reftab_b >> 2;
  o.bits_w = w;    w += (o.bits_b >> 2) + (zero_row ? ((W + 3u) & ~3u) : 0u);   // + the zero row (kept 16-byte aligned)
  o.r16_w = w;     w += o.r16_b >> 2;
  o.rr_w = w;      w += o.rr_b >> 2;
  o.cellref_w = w; w += o.cellref_b >> 2;
  o.total_w = w + 4;                   // + the mbarrier (8 bytes, 16-byte slot)
  return o;
}

// --- TMA bulk copy (cp.async.bulk) + mbarrier helpers ---------------------------------------------
__device__ __forceinline__ uint32_t spr_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void spr_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spr_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void spr_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(spr_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void spr_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(spr_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(spr_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ bool spr_mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(spr_smem_addr(bar)), "r"(parity)
               : "memory");
  return ok != 0u;
}

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
  uint32_t *cnt;  // inlier counters of the warp's 1024 hypotheses (spr_cnt_*)
  uint4 *queue;   // [SPR_QCAP] pending records: x = query index js << 5 | owner lane, y = row,
                  //   z = bit offset of chunk bit 0 in the row, w = the probe's hit bits
  int qcount;     // warp-uniform
};

// Verify up to 32 queued records, one per lane.  A record = all filter hits of ONE query landmark
// on ONE chunk (lane).  Chunk parameters of the owning lane come through shuffles (executed by
// every lane); survivors are added to the owner's counters.
template <bool CNT32>
__device__ __forceinline__ void spr_drain32(const SprView &V, const SprTables &T, WarpState &ws, uint32_t d, int a,
                                            int lane, uint32_t along_off, double across, unsigned long long &n_inl) {
  const int n = ws.qcount < 32 ? ws.qcount : 32;
  const bool active = lane < n;
  uint4 rec = make_uint4(0u, 0u, 0u, 0u);
  if (active) rec = ws.queue[ws.qcount - n + lane];
  const int owner = rec.x & 31;
  const uint32_t o_off = __shfl_sync(SPR_FULL, along_off, owner);
  const double o_across = __shfl_sync(SPR_FULL, across, owner);
  if (active) {
    const int js = (int)(rec.x >> 5);
    const double2 r = __ldg(reinterpret_cast<const double2 *>(V.qrot) + ((size_t)a * (size_t)V.nqp + (size_t)js));
    uint32_t P = spr_verify_mask(V, T, d, rec.y, rec.z, rec.w, r.x, r.y, o_across, V.lat + o_off, V.qdims + 3 * (size_t)js);
    n_inl += (unsigned long long)__popc(P);
    while (P) {
      const int b = __ffs(P) - 1;
      P &= P - 1;
      spr_cnt_inc<CNT32>(ws.cnt, owner, b);
    }
  }
  ws.qcount -= n;
  __syncwarp();
}

template <int BLOCK, int MINB, bool SMEM_TAB, bool STATS, bool CNT32>
__global__ void __launch_bounds__(BLOCK, MINB)
spr_score_lattice_kernel(const SprView V, const SprLaunch K, const int n_wg_local, const long long n_items,
                         const uint32_t tab_bytes) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SprGrid &G = V.grid;
  const int32_t F = G.F;
  const uint32_t d = K.dir;
  const int l = K.label;
  const uint32_t W = (uint32_t)G.W[d], maxbit = (uint32_t)G.maxbit[d];
  // Row band [row_begin, row_end) of the plane handled by this pass: staged as band_rows rows + one all-zero
  // row; rows outside the band clamp onto the zero row (unsigned min), so the probe code is the same as for
  // a whole plane.  Tables read in place cover the whole plane (rows 0 and R - 1 are its zero rows).
  const uint32_t band_rows = SMEM_TAB ? K.row_end - K.row_begin : (uint32_t)G.R[d];
  const bool whole_plane = !SMEM_TAB || band_rows == (uint32_t)G.R[d];   // its rows 0 and R - 1 are the zero rows
  const uint32_t Rm1 = whole_plane ? (uint32_t)G.R[d] - 1u : band_rows;
  const int32_t row_shift = SMEM_TAB ? (int32_t)(K.row_begin << F) : 0;

  // tables of this pass' plane (bits, ranks, per-cell landmark slots, the label's landmark
  // table): a row band of them staged into shared memory once per CTA, or read in place
  SprTables T = spr_global_tables(V, d, l < 0 ? 0 : l);
  const SprTables GT = T;
  SprTabLayout Lo{};
  uint64_t *bar = nullptr;
  uint32_t *stage_flag = nullptr;
  bool staged = !SMEM_TAB;
  if (SMEM_TAB) {
    // The tables are staged by the first warp of the CTA that has work (a CTA of the verification
    // phase may find no candidate item at all): that warp arms an mbarrier with the byte count
    // and issues the TMA bulk copies (cp.async.bulk global -> shared); every warp with work
    // waits on the barrier's phase before its first probe.
    Lo = spr_tab_layout(W, band_rows, GT.cell_base + K.tab_rank_lo, K.tab_cells, V.ref_base[l], K.stage_reftab ? K.tab_refs : 0u, whole_plane ? 0u : 1u);
    bar = reinterpret_cast<uint64_t *>(smem + Lo.total_w - 4);
    stage_flag = smem + Lo.total_w - 2;
    if (threadIdx.x == 0) {
      spr_mbar_init(bar, 1u);
      *stage_flag = 0u;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!whole_plane)
      for (uint32_t i = threadIdx.x; i < ((W + 3u) & ~3u); i += blockDim.x) smem[Lo.bits_w + band_rows * W + i] = 0u;  // the zero row
    __syncthreads();
    T.bits = smem + Lo.bits_w;
    T.r16 = reinterpret_cast<uint16_t *>(smem + Lo.r16_w);
    T.row_rank = smem + Lo.rr_w;
    // label-relative ranks index the staged slice: it starts at the band's first cell
    T.cellref = reinterpret_cast<uint16_t *>(smem + Lo.cellref_w) + Lo.cellref_skip - K.tab_rank_lo;
    if (K.stage_reftab) T.reftab = reinterpret_cast<double *>(smem + Lo.reftab_w) + Lo.reftab_skip;
  }
  WarpState ws;
  ws.cnt = smem + (tab_bytes >> 2) + warp * SPR_WARP_WORDS(CNT32);
  ws.queue = reinterpret_cast<uint4 *>(ws.cnt + SPR_CNT_WORDS(CNT32));
  ws.qcount = 0;
  unsigned long long best = 0ull, n_hits = 0ull, n_inl = 0ull, n_probed = 0ull, n_skipped = 0ull;

  SprBox lb = l >= 0 ? V.labelbox[l] : SprBox{0, -(1 << 30), 0, -(1 << 30)};
  if (SMEM_TAB) {  // marked cells inside the band only (plane row r holds across cell r - 1): query groups that cannot reach it are skipped
    const int32_t band_lo = ((int32_t)K.row_begin - 1) << F, band_hi = ((int32_t)K.row_end - 1) << F;
    if (d == 0) { lb.x0 = max(lb.x0, band_lo); lb.x1 = min(lb.x1, band_hi); }
    else        { lb.y0 = max(lb.y0, band_lo); lb.y1 = min(lb.y1, band_hi); }
  }
  const int g0 = l >= 0 ? V.label_gseg[l] : 0, g1 = l >= 0 ? V.label_gseg[l + 1] : 0;
  const int32_t *q_fx = d ? V.qrotq_yx : V.qrotq_xy;

  const long long n_todo = K.cand_items ? (long long)__ldg(K.cand_count) : n_items;
  for (;;) {
    // per-warp dynamic scheduling, yaw-major: warps running at the same time share qrotq[a][*]
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(K.work_counter, 1ull);
    item = __shfl_sync(SPR_FULL, item, 0);
    if (item >= n_todo) break;
    if (K.cand_items) item = (long long)__ldg(K.cand_items + item);  // verification phase: candidate items only
    const int a = (int)(item / n_wg_local);
    const int wg = spr_shard_group((int)(item % n_wg_local), K.shard_index, K.shard_count);
    const uint32_t cidx = K.chunk_begin + (uint32_t)wg * SPR_WARP_CHUNKS + (uint32_t)lane;  // < chunk_end (padded)
    SprChunk ch;
    {
      const uint4 *cp = reinterpret_cast<const uint4 *>(V.chunks + cidx);
      const uint4 c0 = __ldcs(cp), c1 = __ldcs(cp + 1);
      ch.across = __hiloint2double((int)c0.y, (int)c0.x);
      ch.along_off = c0.z; ch.valid = c0.w; ch.ord_base = c1.x; ch.ord_stride = c1.y; ch.dir = c1.z; ch.ring = c1.w;
    }
    const double across = ch.across;
    const uint32_t along_off = ch.along_off;
    uint32_t valid = ch.valid;
    if (K.ub_nplanes) {
      // bound-and-verify: keep only the hypotheses whose upper bound reaches the running best
      // (>=: an equal count with a smaller canonical index would still win the tie).  The best only
      // grows, so a hypothesis dropped in one label pass stays dropped in the later ones.
      const long long bc = (long long)(*reinterpret_cast<volatile unsigned long long *>(K.best_key) >> SPR_KEY_IDX_BITS) - 1;
      const uint32_t tau = bc > 0 ? (uint32_t)bc : 0u;
      const size_t ii = (size_t)a * (size_t)(K.n_chunks_total / SPR_WARP_CHUNKS) + (size_t)(cidx / SPR_WARP_CHUNKS);
      if (__ldg(K.item_ub + ii) < tau) continue;
      const uint32_t *up = K.ub_planes + (ii * (size_t)K.ub_nplanes) * 32 + lane;
      uint32_t gt = 0u, eq = 0xffffffffu;  // plane-wise compare of the 32 bounds with tau, MSB first
      if ((tau >> K.ub_nplanes) != 0u) eq = 0u;
      for (int i = K.ub_nplanes - 1; i >= 0; i--) {
        const uint32_t p = __ldcs(up + i * 32);
        if ((tau >> i) & 1u) eq &= p;
        else gt |= eq & p;
      }
      valid &= gt | eq;
    }
    if (SMEM_TAB && !staged) {
      if (lane == 0 && atomicExch(stage_flag, 1u) == 0u) {
        spr_mbar_expect_tx(bar, Lo.reftab_b + Lo.bits_b + Lo.r16_b + Lo.rr_b + Lo.cellref_b);
        if (Lo.reftab_b) spr_bulk_g2s(smem + Lo.reftab_w, GT.reftab - Lo.reftab_skip, Lo.reftab_b, bar);
        spr_bulk_g2s(smem + Lo.bits_w, GT.bits + (size_t)K.row_begin * W, Lo.bits_b, bar);
        spr_bulk_g2s(smem + Lo.r16_w, GT.r16 + (size_t)K.row_begin * W, Lo.r16_b, bar);
        spr_bulk_g2s(smem + Lo.rr_w, GT.row_rank + K.row_begin, Lo.rr_b, bar);
        spr_bulk_g2s(smem + Lo.cellref_w, GT.cellref + K.tab_rank_lo - Lo.cellref_skip, Lo.cellref_b, bar);
      }
      for (uint32_t spin = 0; !spr_mbar_try_wait(bar, 0u); spin++)
        if (spin > 200000000u) __trap();  // a lost copy must not hang the GPU
      staged = true;
    }
    spr_cnt_zero<CNT32>(ws.cnt, lane);
    const int32_t aq0 = spr_fx(across, G.S);
    const int32_t bq0 = spr_fx(__ldg(V.lat + along_off), G.S);
    const int32_t aqb = spr_bias_across(aq0, F) - row_shift, bqb = spr_bias_along(bq0, F);
    // patch window of the warp's 1024 translations (unbiased fixed point)
    const int32_t big = 1 << 30;
    const bool live = valid != 0u;
    const int32_t lx0 = d ? bq0 : aq0, lx1 = d ? bq0 + (32 << F) : aq0;
    const int32_t ly0 = d ? aq0 : bq0, ly1 = d ? aq0 : bq0 + (32 << F);
    const int32_t X0 = __reduce_min_sync(SPR_FULL, live ? lx0 : big), X1 = __reduce_max_sync(SPR_FULL, live ? lx1 : -big);
    const int32_t Y0 = __reduce_min_sync(SPR_FULL, live ? ly0 : big), Y1 = __reduce_max_sync(SPR_FULL, live ? ly1 : -big);
    __syncwarp();

    if (X0 <= X1 && g0 < g1) {  // at least one live lane and at least one query of this label
      const int2 *__restrict__ qa2 = reinterpret_cast<const int2 *>(q_fx + 2 * (size_t)a * (size_t)V.nqp);
      // a group is visible iff gx1 > tx_lo && gx0 < tx_hi && gy1 > ty_lo && gy0 < ty_hi
      const int32_t tx_lo = lb.x0 - X1, tx_hi = lb.x1 - X0, ty_lo = lb.y0 - Y1, ty_hi = lb.y1 - Y0;
      const int4 *gbp = reinterpret_cast<const int4 *>(V.gbox) + ((size_t)a * (size_t)V.n_groups + (size_t)g0);
      const int4 *qgp = reinterpret_cast<const int4 *>(qa2) + (size_t)g0 * (SPR_QGROUP / 2);
      auto probe_group = [&](const int g, const int4 *__restrict__ qgp) {
        int4 vg[SPR_QGROUP / 2];
#pragma unroll
        for (int u = 0; u < SPR_QGROUP / 2; u++) vg[u] = __ldg(qgp + u);
        // the group is probed in halves of SPR_QHALF queries: at most 32 * SPR_QHALF new records
        // per push, so the queue (drained below 32 after every push) cannot overflow
#pragma unroll
        for (int hq = 0; hq < SPR_QGROUP / SPR_QHALF; hq++) {
          uint32_t H[SPR_QHALF];
          int4 v[SPR_QHALF / 2];  // two queries each: (across, along) x 2
#pragma unroll
          for (int u = 0; u < SPR_QHALF / 2; u++) v[u] = vg[hq * (SPR_QHALF / 2) + u];
#pragma unroll
          for (int u = 0; u < SPR_QHALF / 2; u++) {
            H[2 * u] = spr_probe(T.bits, W, Rm1, maxbit, F, aqb + v[u].x, bqb + v[u].y, SPR_FULL);
            H[2 * u + 1] = spr_probe(T.bits, W, Rm1, maxbit, F, aqb + v[u].z, bqb + v[u].w, SPR_FULL);
          }
          uint32_t any = 0u;
#pragma unroll
          for (int u = 0; u < SPR_QHALF; u++) any |= H[u];
          any &= valid;
          if (__ballot_sync(SPR_FULL, any != 0u) == 0u) continue;
          if (STATS) {
#pragma unroll
            for (int u = 0; u < SPR_QHALF; u++) n_hits += (unsigned long long)__popc(H[u] & valid);
          }
          const int js0 = g * SPR_QGROUP + hq * SPR_QHALF;
          // one record per (lane, query) with hits; the queue is unordered, so records are laid
          // out query-major and a lane's slot comes from one ballot per query (no shuffle scan)
          const uint32_t lt = (1u << lane) - 1u;
          int base = ws.qcount;
#pragma unroll
          for (int u = 0; u < SPR_QHALF; u++) {
            const uint32_t h = H[u] & valid;
            const uint32_t m = __ballot_sync(SPR_FULL, h != 0u);
            if (h != 0u) {
              const int qa_ = (u & 1) ? v[u >> 1].z : v[u >> 1].x, qb_ = (u & 1) ? v[u >> 1].w : v[u >> 1].y;
              uint32_t row, bit;
              spr_cell_of(F, aqb + qa_, bqb + qb_, &row, &bit);
              ws.queue[base + __popc(m & lt)] = make_uint4(((uint32_t)(js0 + u) << 5) | (uint32_t)lane, row, bit, h);
            }
            base += __popc(m);
          }
          ws.qcount = base;
          __syncwarp();
          while (ws.qcount >= 32) spr_drain32<CNT32>(V, T, ws, d, a, lane, along_off, across, n_inl);
        }
      };
      // visibility of 32 query groups at a time, one group per lane; only the visible ones are walked
      for (int gb = g0; gb < g1; gb += 32) {
        bool vis = false;
        if (gb + lane < g1) {
          const int4 box = __ldg(gbp + (gb - g0) + lane);  // (x0, x1, y0, y1)
          vis = box.y > tx_lo && box.x < tx_hi && box.w > ty_lo && box.z < ty_hi;
        }
        uint32_t vm = __ballot_sync(SPR_FULL, vis);
        if (STATS) { n_probed += __popc(vm); n_skipped += (g1 - gb < 32 ? g1 - gb : 32) - __popc(vm); }
        while (vm) {
          const int k = __ffs(vm) - 1;
          vm &= vm - 1;
          probe_group(gb + k, qgp + (size_t)(gb - g0 + k) * (SPR_QGROUP / 2));
        }
      }
      while (ws.qcount > 0) spr_drain32<CNT32>(V, T, ws, d, a, lane, along_off, across, n_inl);
    }
    __syncwarp();

    // merge with the counters of the earlier passes; after the last pass reduce to the best key
    uint32_t *gw = reinterpret_cast<uint32_t *>(K.gcnt) +
                   ((size_t)a * (size_t)K.n_chunks_total + (size_t)(cidx - lane)) * (CNT32 ? 32 : 16) +
                   (size_t)lane * (CNT32 ? 32 : 16);
    if (!K.first) {
#pragma unroll
      for (int w = 0; w < (CNT32 ? 32 : 16); w += 4) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(gw + w));
        uint32_t *c = ws.cnt + lane * (CNT32 ? 33 : 17) + w;
        c[0] += v.x; c[1] += v.y; c[2] += v.z; c[3] += v.w;
      }
    }
    if (!K.last) {
#pragma unroll
      for (int w = 0; w < (CNT32 ? 32 : 16); w += 4) {
        const uint32_t *c = ws.cnt + lane * (CNT32 ? 33 : 17) + w;
        __stcs(reinterpret_cast<uint4 *>(gw + w), make_uint4(c[0], c[1], c[2], c[3]));
      }
    } else {
      uint32_t v = valid;
      while (v) {
        const int b = __ffs(v) - 1;
        v &= v - 1;
        const uint32_t c = spr_cnt_get<CNT32>(ws.cnt, lane, b);
        const unsigned long long ord = (unsigned long long)ch.ord_base + (unsigned long long)b * ch.ord_stride;
        const unsigned long long key = spr_make_key(c, ord * (unsigned long long)V.n_yaw + (unsigned long long)a);
        best = key > best ? key : best;
        if (K.counts_out) {
          const long long slot = ((long long)ord - (long long)K.ord_begin) * V.n_yaw + a;
          if (slot >= 0 && slot < K.counts_cap) K.counts_out[slot] = (int32_t)c;
        }
      }
    }
    __syncwarp();
  }
  if (K.last) {  // warp max of the 64-bit key, then one atomic per warp
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
      const unsigned long long o = __shfl_xor_sync(SPR_FULL, best, dlt);
      best = o > best ? o : best;
    }
    if (lane == 0 && best != 0ull) atomicMax(K.best_key, best);
  }
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

template <int BLOCK, int MINB, bool SMEM_TAB, bool CNT32>
static cudaError_t spr_launch_cfg(const SprView &V, const SprLaunch &K, int n_wg_local, long long n_items, int grid,
                                  int threads, size_t smem, uint32_t tab_bytes, cudaStream_t st) {
  cudaError_t e;
  if (K.stats) {
    e = cudaFuncSetAttribute(spr_score_lattice_kernel<BLOCK, MINB, SMEM_TAB, true, CNT32>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    spr_score_lattice_kernel<BLOCK, MINB, SMEM_TAB, true, CNT32><<<grid, threads, smem, st>>>(V, K, n_wg_local, n_items, tab_bytes);
  } else {
    e = cudaFuncSetAttribute(spr_score_lattice_kernel<BLOCK, MINB, SMEM_TAB, false, CNT32>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    spr_score_lattice_kernel<BLOCK, MINB, SMEM_TAB, false, CNT32><<<grid, threads, smem, st>>>(V, K, n_wg_local, n_items, tab_bytes);
  }
  return cudaGetLastError();
}

// warps per CTA of the shared-memory-resident variant for this pass (its band, K.row_begin .. K.row_end);
// 0: the tables are read in place
int spr_score_smem_warps(const SprView &V, const SprLaunch &K, int tables_mode) {
  const bool cnt32 = V.nqp > 65535;
  const size_t warp_bytes = (size_t)(cnt32 ? SPR_WARP_WORDS(true) : SPR_WARP_WORDS(false)) * 4;
  if (tables_mode != SPR_TABLES_AUTO || K.label < 0 || K.tab_refs >= SPR_CELL_MULTI || K.row_end <= K.row_begin) return 0;
  const uint32_t tab = spr_tab_layout((uint32_t)V.grid.W[K.dir], K.row_end - K.row_begin, K.tab_cell_base + K.tab_rank_lo, K.tab_cells,
                                      K.tab_ref_base, K.stage_reftab ? K.tab_refs : 0u,
                                      K.row_end - K.row_begin == (uint32_t)V.grid.R[K.dir] ? 0u : 1u).total_w * 4u;
  if ((size_t)tab + 8 * warp_bytes > SPR_SMEM_LIMIT) return 0;
  const int smem_warps = (int)((SPR_SMEM_LIMIT - tab) / warp_bytes);
  return smem_warps > 24 ? 24 : smem_warps;
}

// Row bands of plane (label, dir) for the exact passes: the whole plane when its tables leave room for at
// least 16 warps' counters (24 on BASELINE config 2), otherwise bands of a multiple of 8 rows sized for 16
// warps with the landmark table staged only if it is small.  The per-cell slot table is assumed spread evenly
// over the rows with a factor 2 of slack; spr_score_smem_warps gives the exact verdict per band.
void spr_score_plan(const SprView &V, uint32_t dir, uint32_t label_cells, uint32_t label_refs, uint32_t *band_rows, int *stage_reftab) {
  const bool cnt32 = V.nqp > 65535;
  const size_t warp_bytes = (size_t)(cnt32 ? SPR_WARP_WORDS(true) : SPR_WARP_WORDS(false)) * 4;
  const size_t W = (size_t)V.grid.W[dir], R = (size_t)V.grid.R[dir];
  const size_t reftab_bytes = 40 * (size_t)label_refs + 32;
  const size_t whole = reftab_bytes + (R + 1) * W * 4 + R * W * 2 + R * 4 + 2 * (size_t)label_cells + 64;
  if (whole + 16 * warp_bytes <= SPR_SMEM_LIMIT) { *band_rows = (uint32_t)R; *stage_reftab = 1; return; }
  *stage_reftab = reftab_bytes <= 48 * 1024;
  const size_t budget = SPR_SMEM_LIMIT - 16 * warp_bytes - (*stage_reftab ? reftab_bytes : 0) - W * 4 - 128;
  const double per_row = (double)(W * 6 + 4) + 2.0 * 2.0 * (double)label_cells / (double)R;
  size_t rows = (size_t)((double)budget / per_row);
  rows &= ~(size_t)7;
  if (rows < 8) { *band_rows = 0; return; }   // a single row band does not fit: read in place
  const size_t bands = (R + rows - 1) / rows;
  *band_rows = (uint32_t)((((R + bands - 1) / bands) + 7) & ~(size_t)7);
}

cudaError_t spr_launch_score_lattice(const SprView &V, const SprLaunch &K, int tables_mode, int sm_count,
                                     cudaStream_t st, int *n_launches) {
  if (K.chunk_end <= K.chunk_begin || V.n_yaw <= 0) return cudaSuccess;
  const uint32_t n_chunks = K.chunk_end - K.chunk_begin;
  const int n_wg = (int)(n_chunks / SPR_WARP_CHUNKS);
  const int sc = K.shard_count > 1 ? K.shard_count : 1;
  const int si = K.shard_count > 1 ? K.shard_index : 0;
  const int n_wg_local = spr_shard_local_groups(n_wg, si, sc);
  if (n_wg_local <= 0) return cudaSuccess;
  SprLaunch K2 = K;
  K2.shard_index = si;
  K2.shard_count = sc;
  const long long n_items = (long long)n_wg_local * V.n_yaw;
  const bool cnt32 = V.nqp > 65535;  // a count can reach the number of query landmarks
  const size_t warp_bytes = (size_t)(cnt32 ? SPR_WARP_WORDS(true) : SPR_WARP_WORDS(false)) * 4;
  if (n_launches) (*n_launches)++;  // K.work_counter was zeroed by the caller (one memset for all passes)

  // shared-memory-resident band: one CTA per SM with as many warps as fit next to the tables
  const int smem_warps = spr_score_smem_warps(V, K, tables_mode);
  if (smem_warps >= 8) {
    const uint32_t tab = spr_tab_layout((uint32_t)V.grid.W[K.dir], K.row_end - K.row_begin, K.tab_cell_base + K.tab_rank_lo, K.tab_cells,
                                        K.tab_ref_base, K.stage_reftab ? K.tab_refs : 0u,
                                        K.row_end - K.row_begin == (uint32_t)V.grid.R[K.dir] ? 0u : 1u).total_w * 4u;
    const long long want = (n_items + smem_warps - 1) / smem_warps;
    const int grid = (int)(want < sm_count ? want : sm_count);
    const size_t smem = (size_t)tab + (size_t)smem_warps * warp_bytes;
    return cnt32 ? spr_launch_cfg<768, 1, true, true>(V, K2, n_wg_local, n_items, grid, smem_warps * 32, smem, tab, st)
                 : spr_launch_cfg<768, 1, true, false>(V, K2, n_wg_local, n_items, grid, smem_warps * 32, smem, tab, st);
  }
  const int warps = 8;
  const long long max_grid = (long long)sm_count * 4;
  const long long want = (n_items + warps - 1) / warps;
  const int grid = (int)(want < max_grid ? want : max_grid);
  const size_t smem = (size_t)warps * warp_bytes;
  return cnt32 ? spr_launch_cfg<256, 4, false, true>(V, K2, n_wg_local, n_items, grid, warps * 32, smem, 0u, st)
               : spr_launch_cfg<256, 4, false, false>(V, K2, n_wg_local, n_items, grid, warps * 32, smem, 0u, st);
}
