// spr_api.cu -- C-ABI of the B200 place-recognition search (include/slide_pr.h).
//
// Host control flow mirrors PlaceRecognition::findInterLoopClosure -> findTransformation ->
// MatchMaps -> solveLSQ (place_recognition.cpp:498-538, 736-945, 98-387, 632-695); the scoring
// itself only ever runs on the GPU (spr_join.cu: the default pair-join scorer; spr_kernels*.cu: the
// lattice kernels).  There is no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <climits>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/slide_pr.h"
#include "spr_clipper.h"
#include "spr_delaunay.h"
#include "spr_generate.h"
#include "spr_host.h"
#include "spr_join.h"
#include "spr_kernels.h"

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

std::string g_create_error;

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Page-locked allocation for the host vectors that are uploaded (spr::uvec): a 64-byte header
// remembers whether cudaHostAlloc succeeded, otherwise the block comes from malloc.
std::atomic<int> g_pageable_uploads{0};  // page-locking failed at least once: see slide_pr_prepare
void *pinned_alloc(size_t n) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, n + 64, cudaHostAllocPortable) == cudaSuccess && p) {
    *static_cast<uint64_t *>(p) = 1;
  } else {
    cudaGetLastError();
    p = std::malloc(n + 64);
    if (!p) return nullptr;
    *static_cast<uint64_t *>(p) = 0;
    if (g_pageable_uploads.fetch_add(1) == 0)
      std::fprintf(stderr, "[slide_pr] warning: cudaHostAlloc failed, uploads fall back to pageable memory (synchronous staging)\n");
  }
  return static_cast<char *>(p) + 64;
}
void pinned_free(void *q) {
  if (!q) return;
  void *p = static_cast<char *>(q) - 64;
  if (*static_cast<uint64_t *>(p) == 1) cudaFreeHost(p);
  else std::free(p);
}

// One persistent helper thread: runs one job at a time (submit, then wait).  The handle owns two,
// so that the lattice and the query set are built while the calling thread marks the bitmaps.
class Worker {
 public:
  Worker() : th_([this] { loop(); }) {}
  ~Worker() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_.notify_all();
    th_.join();
  }
  void submit(std::function<int()> job) {
    { std::lock_guard<std::mutex> g(m_); job_ = std::move(job); busy_ = true; }
    cv_.notify_all();
  }
  int wait() {  // result of the last submitted job (0 if none)
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [this] { return !busy_; });
    return result_;
  }
 private:
  void loop() {
    std::unique_lock<std::mutex> g(m_);
    for (;;) {
      cv_.wait(g, [this] { return stop_ || (busy_ && job_); });
      if (stop_) return;
      std::function<int()> job = std::move(job_);
      job_ = nullptr;
      g.unlock();
      const int r = job();
      g.lock();
      result_ = r;
      busy_ = false;
      cv_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::function<int()> job_;
  bool busy_ = false, stop_ = false;
  int result_ = 0;
  std::thread th_;
};

// SLIDE_PR_TRACE=1: per-phase host timings of every call on stderr (developer aid)
struct Trace {
  bool on = std::getenv("SLIDE_PR_TRACE") != nullptr;
  double t0 = 0, last = 0;
  std::string line;
  void start() { if (on) { t0 = last = now_ms(); line.clear(); } }
  void mark(const char *what) {
    if (!on) return;
    const double t = now_ms();
    char buf[96];
    std::snprintf(buf, sizeof(buf), " %s=%.3f", what, t - last);
    line += buf;
    last = t;
  }
  void flush(const char *call) { if (on) std::fprintf(stderr, "[slide_pr] %s total=%.3f ms:%s\n", call, now_ms() - t0, line.c_str()); }
};
thread_local Trace g_trace;

}  // namespace

// Everything derived from ONE reference map: the host index, its device copies and the rows it was built
// from.  The handle keeps one anonymous slot (maps handed over by value, reused when the same bytes come
// again) and a cache of slots keyed by robot id (the reference keeps one map per
// robot in databaseManager::robotMapDict_, databaseManager.h:99-102, and deep-copies it for every attempt,
// sloamNode.cpp:603-614).
struct RefSide {
  spr::RefIndex R;
  spr::uvec<double> cached_ref;      // the (shifted) reference rows the index in R was built from
  double cached_reach = -1.0;        // the index' fixed-point format covers |coordinates| up to this
  slide_pr_params cached_ref_p{};    // parameters the index depends on
  bool ref_index_valid = false;
  bool ranks_pending = false;        // stage 2 of the index (rank tables) not built / uploaded yet
  bool rows_valid = false;           // cached_ref holds the rows of the prepared reference map
  bool ref7_uploaded = false;        // ... and d_ref7 their device copy
  // pair-join scorer (spr_join.h): landmarks binned by label and coarse cell, both join directions
  spr::JoinRef J;
  bool join_valid = false;
  slide_pr_params join_p{};
  DevBuf dj_rec0, dj_rec1, dj_xy0, dj_xy1, dj_cs0, dj_cs1, dj_nbr, dj_labelbox;
  DevBuf d_labelbox, d_bitmap, d_rank16, d_rank16b, d_rowrank, d_rowrankb, d_cellref, d_cellrefb, d_cellbase, d_cellbaseb,
      d_reftab, d_refbase, d_cand, d_cand1, d_vbitmap, d_labof, d_ref7;
  // cache bookkeeping (unused by the anonymous slot)
  int64_t robot_id = -1;
  uint64_t version = 0;
  int n_rows = 0;
  bool shifted = false;              // rows are centroid-shifted (inter-robot mode, PR.cpp:752-765)
  double centroid[2] = {0, 0}, max_abs[2] = {0, 0};   // getCentroid / getMapBoundaries of the map (PR.cpp:713-734)
  uint64_t last_use = 0;
  void release() {
    for (DevBuf *b : {&d_labelbox, &d_bitmap, &d_rank16, &d_rank16b, &d_rowrank, &d_rowrankb, &d_cellref, &d_cellrefb, &d_cellbase,
                      &d_cellbaseb, &d_reftab, &d_refbase, &d_cand, &d_cand1, &d_vbitmap, &d_labof, &d_ref7,
                      &dj_rec0, &dj_rec1, &dj_xy0, &dj_xy1, &dj_cs0, &dj_cs1, &dj_nbr, &dj_labelbox})
      b->release();
  }
};

struct slide_pr_handle {
  slide_pr_params p{};
  int device = 0;
  int sm_count = 148;
  int tables_mode = SPR_TABLES_AUTO;
  bool force_exhaustive = false;  // env SLIDE_PR_EXHAUSTIVE=1
  int refine_min = 32;            // candidate double groups from which the bounds are refined (env SLIDE_PR_REFINE_MIN; < 0: never)
  bool refine_forced = false;     // env SLIDE_PR_REFINE_MIN given: refine whenever there are that many candidates
  size_t bound_smem = 0;          // env SLIDE_PR_BOUND_SMEM (test hook): shared-memory budget of the bound planner, 0 = all
  bool ring_major = false;        // lattice of the prepared problem is chunked ring by ring (anytime budget may bind)
  bool env_lattice = false;       // env SLIDE_PR_ENGINE=lattice: the bound-and-verify lattice kernels are the default search
  double Tstar = 0, Sstar = 0;    // thresholds of the prepared problem (spr::sqrt_threshold / div3_threshold)
  bool lattice_ready = false;     // index structures of the lattice kernels built for the prepared problem (built on demand)
  bool join_ready = false;        // structures of the pair-join scorer built for the prepared problem
  // pair-join scorer: query side, lattice blocks, device copies
  spr::QuerySet JQ;
  spr::uvec<int32_t> j_glabel;
  spr::uvec<SprJoinBlock> j_blocks;
  bool j_blocks_valid = false;
  double j_drift = 0;
  DevBuf dj_lat, dj_cs, dj_qxy, dj_qdims, dj_qlabel, dj_glabel, dj_qrot, dj_gbox, dj_blocks;
  SprJoinView JV{};
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // uploads that overlap the bound phase of a search
  cudaStream_t side_stream = nullptr;  // verification passes of the second bitmap direction
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy = nullptr, ev_prep = nullptr;  // ev_prep: prepare's uploads enqueued
  Worker worker_lattice, worker_query;  // helper threads of slide_pr_prepare
  bool bounds_valid = false;           // the bound planes of the last exhaustive = 2 search are still valid ...
  int bounds_shard_index = 0, bounds_shard_count = 0;  // ... for this shard of the prepared problem
  std::string err;
  // prepared problem (host side)
  bool prepared = false;
  double half_x = 0, half_y = 0, yaw_half = 0;
  int n_ref = 0, n_qry = 0;
  int64_t lat_tb = 0, lat_te = -1;
  double prepare_ms = 0;
  // streaming reuse: the reference-map index and the lattice are rebuilt only when
  // their inputs change (same map bytes / same search ranges), e.g. many submap queries against one
  // accumulated map.  `reuse` in the result tells which were reused.
  spr::uvec<double> qry_rows;       // copy of the query rows (source of the asynchronous upload)
  double lat_hx = 0, lat_hy = 0, lat_yaw_half = 0;
  slide_pr_params lat_p{};
  bool lattice_valid = false;
  int reuse_flags = 0;
  int64_t h2d_bytes = 0;
  spr::Lattice L;
  RefSide anon;                        // slot of maps handed over by value (slide_pr_prepare & co.)
  std::map<int64_t, std::unique_ptr<RefSide>> cache;   // slots keyed by robot id (slide_pr_map_cache_put)
  std::vector<std::unique_ptr<RefSide>> free_slots;   // dropped slots, kept with their page-locked and device buffers for reuse
  RefSide *rs = &anon;                 // the slot of the prepared problem
  uint64_t use_clock = 0;
  size_t cache_capacity = 64;          // least recently used slots beyond this many are dropped
  spr::QuerySet Q;
  // device side
  DevBuf d_lat, d_chunks, d_cs, d_qxy, d_qdims, d_labelseg, d_qlabel, d_gbox, d_gcnt, d_qrot, d_qrotq, d_qrotq_yx, d_work, d_ubplanes,
      d_itemub, d_seed, d_canditems, d_candcount, d_dgitems, d_dgcount, d_qry7, d_best, d_counts, d_match, d_stats, d_hyps, d_tri, d_tri_out;
  SprView V{};
  unsigned long long gen_cap = 0;      // capacity (entries) of the generator's match-key buffer, kept across calls
  SprClipper *clipper = nullptr;       // SlideGraph half: device-resident CLIPPER problem (created on first use)
  DevBuf d_batch_keys, d_batch_match;  // batched intra search: one best key / one correspondence row per candidate
  spr::uvec<unsigned long long> h_batch_keys;
  spr::uvec<int32_t> h_batch_match;
  spr::uvec<int32_t> h_match;          // page-locked D2H targets
  spr::uvec<unsigned long long> h_scalars;
};

#define SPR_CUDA(h, call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                       \
      return SLIDE_PR_ERR_CUDA;                                                             \
    }                                                                                       \
  } while (0)

// compute_budget_sec (PR.cpp:181-191) is tested before every ring, in whole seconds, with '>': a search
// that ends within the budget scores every ring and returns what the unlimited search returns.  The
// GPU search takes milliseconds where the reference takes its full budget, so the default
// bound-and-verify search stays in place whenever a (deliberately pessimistic: ~1/8 of the measured
// rate) estimate of the search time is below the budget; only then-unrealistic problems fall back to
// the ring-by-ring exhaustive search that can stop between rings.
static bool budget_may_bind(const slide_pr_params &p, double half_x, double half_y, double yaw_half, int n_qry) {
  if (!(p.compute_budget_sec > 0)) return false;
  const double step = p.match_xy_step_size;
  if (!(step > 0)) return true;
  const double n_trans = (2.0 * half_x / step + 1.0) * (2.0 * half_y / step + 1.0);
  const double n_yaw = p.disable_yaw_search ? 1.0 : (p.match_yaw_angle_step_size > 0 ? 2.0 * yaw_half / p.match_yaw_angle_step_size + 1.0 : 1e30);
  const double est_sec = n_trans * n_yaw * (double)std::max(n_qry, 1) / 2.0e12;
  return !(est_sec < p.compute_budget_sec);
}

template <typename T, typename A>
static int upload(slide_pr_handle *h, DevBuf &b, const std::vector<T, A> &v, cudaStream_t st) {
  SPR_CUDA(h, b.ensure(std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) SPR_CUDA(h, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  h->h2d_bytes += (int64_t)(v.size() * sizeof(T));
  return SLIDE_PR_OK;
}

static int upload_raw(slide_pr_handle *h, DevBuf &b, const void *src, size_t bytes, cudaStream_t st) {
  SPR_CUDA(h, b.ensure(std::max<size_t>(bytes, 8)));
  if (bytes) SPR_CUDA(h, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
  h->h2d_bytes += (int64_t)bytes;
  return SLIDE_PR_OK;
}

// A slot that leaves the cache keeps its buffers (page-locked host vectors, device allocations) in a small
// pool: allocating them afresh for every batch of maps costs more than the searches themselves.
static void retire_slot(slide_pr_handle *h, std::unique_ptr<RefSide> slot) {
  slot->rows_valid = false; slot->ref_index_valid = false; slot->ranks_pending = false;
  slot->join_valid = false; slot->ref7_uploaded = false;
  slot->robot_id = -1; slot->version = 0; slot->n_rows = 0;
  if (h->free_slots.size() < 16) h->free_slots.push_back(std::move(slot));
  else slot->release();
}

extern "C" {
#pragma GCC visibility push(default)

int slide_pr_abi_version(void) { return SLIDE_PR_ABI_VERSION; }

double slide_pr_deg2rad(double deg) { return deg * M_PI / 180.; }

void slide_pr_default_params(slide_pr_params *p) {  // PR.cpp:24-75
  std::memset(p, 0, sizeof(*p));
  p->compute_budget_sec = -1.0;
  p->dilation_factor = 1.2;
  p->match_xy_step_size = 0.5;
  p->match_yaw_half_range = slide_pr_deg2rad(180.);
  p->match_yaw_angle_step_size = slide_pr_deg2rad(2.0);
  p->match_threshold = 0.5;
  p->match_threshold_dimension = 1.0;
  p->match_x_half_range_intra = 5.0;
  p->match_y_half_range_intra = 5.0;
  p->match_yaw_half_range_intra = slide_pr_deg2rad(10.);
  p->disable_yaw_search = 0;
  p->ignore_dimension = 0;
  p->min_num_inliers = 5;
  p->use_lsq = 1;
  p->min_num_map_objects_to_start = 1;
  p->inter_loop_closure = 1;
  p->device = -1;
}

const char *slide_pr_last_error(const slide_pr_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int slide_pr_create(const slide_pr_params *p, slide_pr_handle **out) {
  if (!p || !out) { g_create_error = "null argument"; return SLIDE_PR_ERR_INVALID; }
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0) {
    g_create_error = std::string("no CUDA device (the place-recognition search has no CPU fallback): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return SLIDE_PR_ERR_CUDA;
  }
  spr::g_upload_alloc = pinned_alloc;  // uploaded host vectors are page-locked from here on
  spr::g_upload_free = pinned_free;
  slide_pr_handle *h = new slide_pr_handle();
  h->p = *p;
  int dev = p->device;
  if (dev < 0) cudaGetDevice(&dev);
  if (dev >= n_dev) { g_create_error = "device ordinal out of range"; delete h; return SLIDE_PR_ERR_INVALID; }
  h->device = dev;
  if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_prep, cudaEventDisableTiming)) != cudaSuccess) {
    g_create_error = std::string("CUDA init: ") + cudaGetErrorString(e);
    delete h;
    return SLIDE_PR_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  // SLIDE_PR_VARIANT=0 forces the global-memory table path (tests cover both paths)
  if (const char *v = std::getenv("SLIDE_PR_VARIANT")) h->tables_mode = std::atoi(v) == 0 ? SPR_TABLES_GLOBAL : SPR_TABLES_AUTO;
  // SLIDE_PR_EXHAUSTIVE=1 verifies every hypothesis exactly (no bound-and-verify pruning)
  if (const char *v = std::getenv("SLIDE_PR_EXHAUSTIVE")) h->force_exhaustive = std::atoi(v) != 0;
  if (const char *v = std::getenv("SLIDE_PR_REFINE_MIN")) { h->refine_min = std::atoi(v); h->refine_forced = true; }
  if (const char *v = std::getenv("SLIDE_PR_BOUND_SMEM")) { const long b = std::atol(v); if (b > 0) h->bound_smem = (size_t)b; }
  // SLIDE_PR_ENGINE=lattice: searches default to the bound-and-verify lattice kernels instead of the pair-join scorer (A/B)
  if (const char *v = std::getenv("SLIDE_PR_ENGINE")) h->env_lattice = std::strcmp(v, "lattice") == 0;
  *out = h;
  return SLIDE_PR_OK;
}

void slide_pr_destroy(slide_pr_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (DevBuf *b : {&h->d_lat, &h->d_chunks, &h->d_cs, &h->d_qxy, &h->d_qdims, &h->d_labelseg, &h->d_qlabel, &h->d_gbox, &h->d_gcnt, &h->d_qrot,
                    &h->d_qrotq, &h->d_qrotq_yx, &h->d_work, &h->d_qry7, &h->d_best, &h->d_counts, &h->d_match, &h->d_stats, &h->d_hyps,
                    &h->d_tri, &h->d_tri_out, &h->d_ubplanes, &h->d_itemub, &h->d_seed, &h->d_canditems, &h->d_candcount, &h->d_dgitems,
                    &h->d_dgcount, &h->dj_lat, &h->dj_cs, &h->dj_qxy, &h->dj_qdims, &h->dj_qlabel, &h->dj_glabel, &h->dj_qrot,
                    &h->dj_gbox, &h->dj_blocks, &h->d_batch_keys, &h->d_batch_match})
    b->release();
  h->anon.release();
  for (auto &kv : h->cache) kv.second->release();
  h->cache.clear();
  for (auto &f : h->free_slots) f->release();
  h->free_slots.clear();
  spr_clipper_destroy(h->clipper);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->ev_prep) cudaEventDestroy(h->ev_prep);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int slide_pr_set_params(slide_pr_handle *h, const slide_pr_params *p) {
  if (!h || !p) return SLIDE_PR_ERR_INVALID;
  const int dev = h->p.device;
  h->p = *p;
  h->p.device = dev;
  h->prepared = false;
  return SLIDE_PR_OK;
}

static int upload_lattice(slide_pr_handle *h, cudaStream_t st) {
  int rc;
  if ((rc = upload(h, h->d_lat, h->L.lat, st))) return rc;
  if ((rc = upload(h, h->d_chunks, h->L.chunks, st))) return rc;
  if ((rc = upload(h, h->d_cs, h->L.cs, st))) return rc;
  h->V.lat = h->d_lat.as<double>();
  h->V.chunks = h->d_chunks.as<SprChunk>();
  h->V.n_chunks = (uint32_t)h->L.chunks.size();
  h->V.n_yaw = (int32_t)h->L.yaw.size();
  h->V.cs = h->d_cs.as<double>();
  return SLIDE_PR_OK;
}

// Stage 2 of the reference index: rank tables / candidate records of both bitmap directions,
// built from the cached reference rows and uploaded on `st`.
static int finish_ranks(slide_pr_handle *h, cudaStream_t st) {
  int rc;
  for (int d = 0; d < 2; d++)
    if ((rc = spr::build_ref_ranks(h->rs->cached_ref.data(), d, h->rs->R, h->err)) != SLIDE_PR_OK) return rc;
  g_trace.mark("ref_ranks_build");
  if ((rc = upload(h, h->rs->d_rank16, h->rs->R.rank16[0], st))) return rc;
  if ((rc = upload(h, h->rs->d_rank16b, h->rs->R.rank16[1], st))) return rc;
  if ((rc = upload(h, h->rs->d_rowrank, h->rs->R.row_rank[0], st))) return rc;
  if ((rc = upload(h, h->rs->d_rowrankb, h->rs->R.row_rank[1], st))) return rc;
  if ((rc = upload(h, h->rs->d_cellref, h->rs->R.cellref[0], st))) return rc;
  if ((rc = upload(h, h->rs->d_cellrefb, h->rs->R.cellref[1], st))) return rc;
  if ((rc = upload(h, h->rs->d_cellbase, h->rs->R.cell_base[0], st))) return rc;
  if ((rc = upload(h, h->rs->d_cellbaseb, h->rs->R.cell_base[1], st))) return rc;
  if ((rc = upload(h, h->rs->d_cand, h->rs->R.cand[0], st))) return rc;
  if ((rc = upload(h, h->rs->d_cand1, h->rs->R.cand[1], st))) return rc;
  SprView &V = h->V;
  V.rank16[0] = h->rs->d_rank16.as<uint16_t>();
  V.rank16[1] = h->rs->d_rank16b.as<uint16_t>();
  V.row_rank[0] = h->rs->d_rowrank.as<uint32_t>();
  V.row_rank[1] = h->rs->d_rowrankb.as<uint32_t>();
  V.cellref[0] = h->rs->d_cellref.as<uint16_t>();
  V.cellref[1] = h->rs->d_cellrefb.as<uint16_t>();
  V.cell_base[0] = h->rs->d_cellbase.as<uint32_t>();
  V.cell_base[1] = h->rs->d_cellbaseb.as<uint32_t>();
  V.cand[0] = h->rs->d_cand.as<SprCand>();
  V.cand[1] = h->rs->d_cand1.as<SprCand>();
  h->rs->ranks_pending = false;
  g_trace.mark("ref_ranks_upload");
  return SLIDE_PR_OK;
}

// Which search a prepared problem gets by default: the pair-join scorer, unless the parameters / environment
// ask for the lattice kernels or the problem is outside the scorer's limits (u16 counters; the anytime
// budget works ring by ring).
static bool join_supported(const slide_pr_handle *h) { return !h->ring_major && h->n_qry <= 65535; }
static bool join_is_default(const slide_pr_handle *h) {
  return h->p.exhaustive_search == 0 && !h->force_exhaustive && !h->env_lattice && join_supported(h);
}
static int lattice_prepare(slide_pr_handle *h);
static int join_prepare(slide_pr_handle *h);

// slot: where the reference-map index lives (the anonymous slot or a cache entry).  trusted: ref7 IS the
// slot's cached rows (a cache entry at its current version), so the byte comparison is skipped.
// Copies the maps, uploads the raw rows and builds the structures of the default search; those of the
// other engine are built on demand by the search that needs them.
static int prepare_impl(slide_pr_handle *h, RefSide *slot, bool trusted, const double *ref7, int32_t n_ref, const double *qry7,
                        int32_t n_qry, double half_x, double half_y) {
  if (!h) return SLIDE_PR_ERR_INVALID;
  h->rs = slot;
  h->prepared = false;
  h->bounds_valid = false;
  h->lattice_ready = false;
  h->join_ready = false;
  if (n_ref < 0 || n_qry < 0 || (n_ref > 0 && !ref7) || (n_qry > 0 && !qry7)) { h->err = "bad map arguments"; return SLIDE_PR_ERR_INVALID; }
  if (n_qry >= (1 << 22)) { h->err = "more than 2^22 query landmarks"; return SLIDE_PR_ERR_UNSUPPORTED; }
  const double t0 = now_ms();
  SPR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // the previous prepare's uploads read the handle's page-locked vectors asynchronously: they must have
  // been consumed before those vectors are rebuilt (a search in between has already waited for them)
  SPR_CUDA(h, cudaEventSynchronize(h->ev_prep));
  g_trace.mark("prev_uploads_wait");
  h->half_x = half_x; h->half_y = half_y;
  h->h2d_bytes = 0;
  h->yaw_half = h->p.inter_loop_closure ? h->p.match_yaw_half_range : h->p.match_yaw_half_range_intra;
  h->n_ref = n_ref; h->n_qry = n_qry;
  h->ring_major = budget_may_bind(h->p, half_x, half_y, h->yaw_half, n_qry);
  h->lat_tb = 0; h->lat_te = -1;
  h->reuse_flags = 0;
  h->prepare_ms = 0;
  h->Tstar = spr::sqrt_threshold(h->p.match_threshold);
  h->Sstar = spr::div3_threshold(h->p.match_threshold_dimension);
  for (int j = 0; j < n_qry; j++)
    if (!std::isfinite(qry7[7 * (size_t)j + 1]) || !std::isfinite(qry7[7 * (size_t)j + 2])) { h->err = "non-finite query coordinate"; return SLIDE_PR_ERR_NONFINITE; }
  int rc;
  // rows of the reference map: page-locked host copy + device copy, kept while the same bytes come again
  const bool same_rows = slot->rows_valid && (int)(slot->cached_ref.size() / 7) == n_ref &&
                         (trusted || n_ref == 0 || std::memcmp(slot->cached_ref.data(), ref7, (size_t)n_ref * 7 * sizeof(double)) == 0);
  if (!same_rows) {
    if (!trusted) slot->cached_ref.assign(ref7, ref7 + (size_t)n_ref * 7);
    slot->rows_valid = true;
    slot->ref_index_valid = false; slot->ranks_pending = false;
    slot->join_valid = false;
    slot->ref7_uploaded = false;
  }
  if (!slot->ref7_uploaded) {
    if ((rc = upload(h, slot->d_ref7, slot->cached_ref, st))) return rc;
    slot->ref7_uploaded = true;
  }
  h->qry_rows.assign(qry7, qry7 + (size_t)n_qry * 7);
  if ((rc = upload(h, h->d_qry7, h->qry_rows, st))) return rc;
  SPR_CUDA(h, h->d_work.ensure(4096 * sizeof(unsigned long long)));
  SPR_CUDA(h, h->d_best.ensure(sizeof(unsigned long long)));
  SPR_CUDA(h, h->d_stats.ensure(4 * sizeof(unsigned long long)));
  SPR_CUDA(h, h->d_match.ensure(std::max<size_t>(n_qry, 1) * sizeof(int32_t)));
  g_trace.mark("rows_upload");
  rc = join_is_default(h) ? join_prepare(h) : lattice_prepare(h);
  if (rc != SLIDE_PR_OK) return rc;
  // no synchronisation here: every upload reads page-locked vectors owned by the handle (the
  // caller's rows were copied), which stay untouched until the next prepare
  SPR_CUDA(h, cudaEventRecord(h->ev_prep, st));  // a search on another stream waits for these uploads
  if (g_pageable_uploads.load() > 0) SPR_CUDA(h, cudaStreamSynchronize(st));  // page-locking failed somewhere: do not rely on it
  h->prepared = true;
  h->prepare_ms = now_ms() - t0;
  return SLIDE_PR_OK;
}

// Structures of the pair-join scorer for the prepared problem: reference landmarks by label and coarse cell
// (kept per reference map), query groups, ring geometry and lattice blocks (kept while the search ranges repeat).
static int join_prepare(slide_pr_handle *h) {
  const double t0 = now_ms();
  cudaStream_t st = h->stream;
  RefSide *rs = h->rs;
  int rc;
  const bool same_ref = rs->join_valid && rs->join_p.match_xy_step_size == h->p.match_xy_step_size &&
                        rs->join_p.match_threshold == h->p.match_threshold &&
                        rs->join_p.match_threshold_dimension == h->p.match_threshold_dimension;
  // ring geometry and lattice samples (no chunks): a function of the scalar search parameters only
  const bool same_lattice = h->lattice_valid && h->lat_hx == h->half_x && h->lat_hy == h->half_y && h->lat_yaw_half == h->yaw_half &&
                            h->lat_p.match_xy_step_size == h->p.match_xy_step_size &&
                            h->lat_p.match_yaw_angle_step_size == h->p.match_yaw_angle_step_size &&
                            h->lat_p.disable_yaw_search == h->p.disable_yaw_search && !h->L.ring_major;
  // the three host builds are independent: lattice + blocks and the query groups on the helper threads (a sleeping
  // thread takes a while to wake up), the reference-side bins -- the longest of the three -- on this one
  std::string lattice_err, query_err;
  bool lattice_started = false, query_started = false;
  struct Joiner {
    slide_pr_handle *h; bool *l, *q;
    ~Joiner() { if (*l) h->worker_lattice.wait(); if (*q) h->worker_query.wait(); }
  } joiner{h, &lattice_started, &query_started};
  if (!same_lattice) {
    h->lattice_valid = false; h->j_blocks_valid = false;
  } else {
    h->reuse_flags |= 1;
  }
  if (!h->j_blocks_valid) {
    const bool build_rings = !same_lattice;
    h->worker_lattice.submit([h, build_rings, &lattice_err]() {
      cudaSetDevice(h->device);  // page-locked buffers grown by this job belong to the handle's device
      int r = SLIDE_PR_OK;
      if (build_rings) r = spr::build_lattice(h->p, h->half_x, h->half_y, h->yaw_half, 0, -1, false, h->L, lattice_err, true);
      if (r == SLIDE_PR_OK && h->L.status == 0) r = spr::build_join_blocks(h->L, h->p.match_xy_step_size, h->j_blocks, &h->j_drift, lattice_err);
      return r;
    });
    lattice_started = true;
  }
  // query groups (label-major, Morton order inside a label): they need the reference map's labels only, which the job
  // collects itself while the bins (and with them rs->J.labels) are being rebuilt
  h->worker_query.submit([h, rs, same_ref, &query_err]() {
    cudaSetDevice(h->device);
    std::vector<double> labels;
    if (!same_ref) spr::unique_labels(rs->cached_ref.data(), h->n_ref, labels);
    return spr::build_query_set(same_ref ? rs->J.labels : labels, h->qry_rows.data(), h->n_qry, h->JQ, query_err);
  });
  query_started = true;
  if (!same_ref) {
    rs->join_valid = false;
    if ((rc = spr::build_join_ref(h->p, rs->cached_ref.data(), h->n_ref, rs->J, h->err)) != SLIDE_PR_OK) return rc;
    g_trace.mark("join_ref_build");
    if ((rc = upload(h, rs->dj_rec0, rs->J.rec[0], st))) return rc;
    if ((rc = upload(h, rs->dj_rec1, rs->J.rec[1], st))) return rc;
    if ((rc = upload(h, rs->dj_xy0, rs->J.xy[0], st))) return rc;
    if ((rc = upload(h, rs->dj_xy1, rs->J.xy[1], st))) return rc;
    if ((rc = upload(h, rs->dj_cs0, rs->J.cell_start[0], st))) return rc;
    if ((rc = upload(h, rs->dj_cs1, rs->J.cell_start[1], st))) return rc;
    if ((rc = upload(h, rs->dj_nbr, rs->J.nbr, st))) return rc;
    if ((rc = upload(h, rs->dj_labelbox, rs->J.labelbox, st))) return rc;
    rs->join_valid = true;
    rs->join_p = h->p;
    g_trace.mark("join_ref_upload");
  } else {
    h->reuse_flags |= 2;
  }
  rc = h->worker_query.wait();
  query_started = false;
  g_trace.mark("join_queries_wait");
  if (rc != SLIDE_PR_OK) { h->err = query_err; return rc; }
  const int n_groups = h->JQ.nqp / SPR_QGROUP;
  h->j_glabel.assign((size_t)std::max(n_groups, 1), 0);
  for (int g = 0; g < n_groups; g++) h->j_glabel[g] = h->JQ.qlabel[(size_t)g * SPR_QGROUP];  // a group's first entry is never padding
  if ((rc = upload(h, h->dj_qxy, h->JQ.qxy, st))) return rc;
  if ((rc = upload(h, h->dj_qdims, h->JQ.qdims, st))) return rc;
  if ((rc = upload(h, h->dj_qlabel, h->JQ.qlabel, st))) return rc;
  if ((rc = upload(h, h->dj_glabel, h->j_glabel, st))) return rc;
  g_trace.mark("join_queries");
  if (lattice_started) {
    rc = h->worker_lattice.wait();
    lattice_started = false;
    g_trace.mark("join_blocks_wait");
    if (rc != SLIDE_PR_OK) { h->err = lattice_err; return rc; }
    if (!same_lattice) {
      h->lat_hx = h->half_x; h->lat_hy = h->half_y; h->lat_yaw_half = h->yaw_half; h->lat_p = h->p;
      h->lattice_valid = true;
    }
    if ((rc = upload(h, h->dj_lat, h->L.lat, st))) return rc;
    if ((rc = upload(h, h->dj_cs, h->L.cs, st))) return rc;
    if ((rc = upload(h, h->dj_blocks, h->j_blocks, st))) return rc;
    h->j_blocks_valid = true;
    g_trace.mark("join_blocks");
  }
  const size_t n_yaw = h->L.yaw.size();
  SPR_CUDA(h, h->dj_qrot.ensure(std::max<size_t>(n_yaw * (size_t)h->JQ.nqp, 1) * 2 * sizeof(double)));
  SPR_CUDA(h, h->dj_gbox.ensure(std::max<size_t>(n_yaw * (size_t)n_groups, 1) * sizeof(SprJoinBox)));
  SprJoinView &V = h->JV;
  V.lat = h->dj_lat.as<double>();
  V.qrot = h->dj_qrot.as<double>();
  V.gbox = h->dj_gbox.as<SprJoinBox>();
  V.qdims = h->dj_qdims.as<double>();
  V.glabel = h->dj_glabel.as<int32_t>();
  V.qxy = h->dj_qxy.as<double>();
  V.qlabel = h->dj_qlabel.as<int32_t>();
  V.cs = h->dj_cs.as<double>();
  V.nqp = h->JQ.nqp; V.n_groups = n_groups; V.n_yaw = (int32_t)n_yaw; V.n_labels = (int32_t)rs->J.labels.size();
  V.rec[0] = rs->dj_rec0.as<SprJoinRef>(); V.rec[1] = rs->dj_rec1.as<SprJoinRef>();
  V.xy[0] = rs->dj_xy0.as<double>(); V.xy[1] = rs->dj_xy1.as<double>();
  V.cell_start[0] = rs->dj_cs0.as<uint32_t>(); V.cell_start[1] = rs->dj_cs1.as<uint32_t>();
  V.nbr = rs->dj_nbr.as<SprJoinNbr>();
  V.labelbox = rs->dj_labelbox.as<double>();
  V.gx0 = rs->J.gx0; V.gy0 = rs->J.gy0; V.inv_w = rs->J.inv_w;
  V.ncx = rs->J.ncx; V.ncy = rs->J.ncy;
  V.Tstar = rs->J.Tstar; V.Sstar = rs->J.Sstar; V.thr_dim = h->p.match_threshold_dimension;
  V.ignore_dim = h->p.ignore_dimension;
  // the enumeration margins also cover the rounding of the reference's test at the magnitude of the coordinates involved
  // (negligible against the 1e-9 slack of J.reach unless the maps sit millions of metres from the origin)
  double qmag = 0.0;
  for (int j = 0; j < h->n_qry; j++) qmag = std::max(qmag, std::hypot(h->qry_rows[7 * (size_t)j + 1], h->qry_rows[7 * (size_t)j + 2]));
  const double round_slack = 64.0 * DBL_EPSILON * (rs->J.max_abs + qmag + std::max(std::fabs(h->half_x), std::fabs(h->half_y)));
  V.reach = rs->J.reach + round_slack;
  V.ireach = V.reach + 2.0 * h->j_drift + 1e-9;
  V.inv_step = 1.0 / h->p.match_xy_step_size;
  V.blocks = h->dj_blocks.as<SprJoinBlock>();
  V.n_blocks = (uint32_t)h->j_blocks.size();
  h->join_ready = true;
  h->prepare_ms += now_ms() - t0;
  return SLIDE_PR_OK;
}

// Index structures of the lattice kernels (occupancy bitmaps, rank tables, chunked lattice, fixed-point query
// groups) for the prepared problem, from the handle's copies of the maps.
static int lattice_prepare(slide_pr_handle *h) {
  const double t0 = now_ms();
  cudaStream_t st = h->stream;
  const double *ref7 = h->rs->cached_ref.data(), *qry7 = h->qry_rows.data();
  const int32_t n_ref = h->n_ref, n_qry = h->n_qry;
  const double half_x = h->half_x, half_y = h->half_y;
  const bool ring_major = h->ring_major;
  const bool trusted = true;   // prepare_impl has already compared / copied the rows
  int rc;
  // lattice: a function of the scalar search parameters only
  const bool same_lattice = h->lattice_valid && h->lat_hx == half_x && h->lat_hy == half_y && h->lat_yaw_half == h->yaw_half &&
                            h->lat_p.match_xy_step_size == h->p.match_xy_step_size &&
                            h->lat_p.match_yaw_angle_step_size == h->p.match_yaw_angle_step_size &&
                            h->lat_p.disable_yaw_search == h->p.disable_yaw_search &&
                            h->L.ring_major == ring_major && h->L.has_chunks;
  std::string lattice_err, query_err;
  bool lattice_started = false, query_started = false;
  // every exit path waits for the helper threads (they write into the handle)
  struct Joiner {
    slide_pr_handle *h; bool *l, *q;
    ~Joiner() { if (*l) h->worker_lattice.wait(); if (*q) h->worker_query.wait(); }
  } joiner{h, &lattice_started, &query_started};
  if (!same_lattice) {  // built on a helper thread while this one builds the bitmaps
    h->lattice_valid = false;
    h->j_blocks_valid = false;   // the blocks of the pair-join scorer index the sample arrays that are rebuilt here
    h->worker_lattice.submit([h, half_x, half_y, ring_major, &lattice_err]() {
      cudaSetDevice(h->device);  // the page-locked buffers grown by this job belong to the handle's device (one process may drive several GPUs)
      return spr::build_lattice(h->p, half_x, half_y, h->yaw_half, 0, -1, ring_major, h->L, lattice_err);
    });
    lattice_started = true;
  } else {
    h->reuse_flags |= 1;
  }
  auto join_lattice = [&]() -> int {
    if (!lattice_started) return SLIDE_PR_OK;
    const int lrc = h->worker_lattice.wait();
    lattice_started = false;
    if (lrc != SLIDE_PR_OK) { h->err = lattice_err; return lrc; }
    h->lat_hx = half_x; h->lat_hy = half_y; h->lat_yaw_half = h->yaw_half; h->lat_p = h->p;
    h->lattice_valid = true;
    g_trace.mark("lattice_join");
    const int urc = upload_lattice(h, st);
    g_trace.mark("lattice_upload");
    return urc;
  };
  double qrad = 0;
  for (int j = 0; j < n_qry; j++) {
    const double r = std::hypot(qry7[7 * (size_t)j + 1], qry7[7 * (size_t)j + 2]);
    if (!std::isfinite(r)) { h->err = "non-finite query coordinate"; return SLIDE_PR_ERR_NONFINITE; }
    qrad = std::max(qrad, r);
  }
  const double reach = qrad + std::max(std::fabs(half_x), std::fabs(half_y)) + h->p.match_xy_step_size;
  // reference index: a function of the reference rows, the cell size / thresholds and the reach
  const bool same_ref = h->rs->ref_index_valid && (int)(h->rs->cached_ref.size() / 7) == n_ref && reach <= h->rs->cached_reach &&
                        h->rs->cached_ref_p.match_xy_step_size == h->p.match_xy_step_size &&
                        h->rs->cached_ref_p.match_threshold == h->p.match_threshold &&
                        h->rs->cached_ref_p.match_threshold_dimension == h->p.match_threshold_dimension &&
                        (trusted || n_ref == 0 || std::memcmp(h->rs->cached_ref.data(), ref7, (size_t)n_ref * 7 * sizeof(double)) == 0);
  auto start_query_job = [&]() {
    h->worker_query.submit([h, qry7, n_qry, &query_err]() {
      cudaSetDevice(h->device);
      return spr::build_query_set(h->rs->R, qry7, n_qry, h->Q, query_err);
    });
    query_started = true;
  };
  if (!same_ref) {
    // stage 1 of the reference index (bitmaps, landmark tables): all the bound phase needs.  The
    // rank tables (stage 2) are built and uploaded by the search while the bound phase runs.
    h->rs->ref_index_valid = false;
    h->rs->ranks_pending = false;
    const double reach_cap = reach * 1.25;  // head-room so that slightly larger queries reuse the index
    if ((rc = spr::build_ref_grid(h->p, ref7, n_ref, reach_cap, h->rs->R, h->err)) != SLIDE_PR_OK) return rc;
    start_query_job();  // the query set only needs the labels and the grid
    if ((rc = spr::build_ref_marks(h->p, ref7, n_ref, h->rs->R, h->err)) != SLIDE_PR_OK) return rc;
    g_trace.mark("ref_bitmaps_build");
    h->rs->cached_reach = h->rs->R.reach_limit; h->rs->cached_ref_p = h->p;
    h->rs->ref_index_valid = true;
    h->rs->ranks_pending = true;
    if ((rc = upload(h, h->rs->d_labelbox, h->rs->R.labelbox, st))) return rc;
    // 4 zero words in front of (and slack behind) the planes: the bound kernel reads the word before
    // and the word after a probe's base word
    SPR_CUDA(h, h->rs->d_bitmap.ensure(h->rs->R.bitmap.size() * sizeof(uint32_t) + 64));
    SPR_CUDA(h, cudaMemsetAsync(h->rs->d_bitmap.p, 0, 16, st));
    SPR_CUDA(h, cudaMemsetAsync(static_cast<char *>(h->rs->d_bitmap.p) + 16 + h->rs->R.bitmap.size() * sizeof(uint32_t), 0, 32, st));
    if (!h->rs->R.bitmap.empty())
      SPR_CUDA(h, cudaMemcpyAsync(static_cast<char *>(h->rs->d_bitmap.p) + 16, h->rs->R.bitmap.data(), h->rs->R.bitmap.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    h->h2d_bytes += (int64_t)(h->rs->R.bitmap.size() * sizeof(uint32_t));
    if ((rc = upload(h, h->rs->d_reftab, h->rs->R.reftab, st))) return rc;
    if ((rc = upload(h, h->rs->d_refbase, h->rs->R.ref_base, st))) return rc;
    if ((rc = upload(h, h->rs->d_labof, h->rs->R.lab_of, st))) return rc;
    g_trace.mark("ref_bitmaps_upload");
  } else {
    h->reuse_flags |= 2;
  }
  if (!query_started) start_query_job();
  rc = h->worker_query.wait();
  query_started = false;
  if (rc != SLIDE_PR_OK) { h->err = query_err; return rc; }
  g_trace.mark("query_set_build");
  if ((rc = upload(h, h->d_qxy, h->Q.qxy, st))) return rc;
  if ((rc = upload(h, h->d_qdims, h->Q.qdims, st))) return rc;
  if ((rc = upload(h, h->d_labelseg, h->Q.label_gseg, st))) return rc;
  if ((rc = upload(h, h->d_qlabel, h->Q.qlabel, st))) return rc;
  if ((rc = join_lattice()) != SLIDE_PR_OK) return rc;
  const size_t nrot = (size_t)std::max<size_t>((size_t)h->L.yaw.size() * (size_t)h->Q.nqp, 1);
  const size_t ngb = (size_t)std::max<size_t>((size_t)h->L.yaw.size() * (size_t)(h->Q.nqp / SPR_QGROUP), 1);
  SPR_CUDA(h, h->d_qrot.ensure(nrot * 2 * sizeof(double)));
  SPR_CUDA(h, h->d_qrotq.ensure(nrot * 2 * sizeof(int32_t)));
  SPR_CUDA(h, h->d_qrotq_yx.ensure(nrot * 2 * sizeof(int32_t)));
  SPR_CUDA(h, h->d_gbox.ensure(ngb * sizeof(SprBox)));

  SprView &V = h->V;
  V.nqp = h->Q.nqp;
  V.n_groups = h->Q.nqp / SPR_QGROUP;
  V.qrotq_xy = h->d_qrotq.as<int32_t>();
  V.qrotq_yx = h->d_qrotq_yx.as<int32_t>();
  V.gbox = h->d_gbox.as<SprBox>();
  V.qrot = h->d_qrot.as<double>();
  V.qxy = h->d_qxy.as<double>();
  V.qdims = h->d_qdims.as<double>();
  V.label_gseg = h->d_labelseg.as<int32_t>();
  V.qlabel = h->d_qlabel.as<int32_t>();
  V.n_labels = (int32_t)h->rs->R.labels.size();
  V.n_ref = n_ref;
  V.labelbox = h->rs->d_labelbox.as<SprBox>();
  V.bitmap = h->rs->d_bitmap.as<uint32_t>() + 4;  // behind the zero words in front
  V.rank16[0] = h->rs->d_rank16.as<uint16_t>();
  V.rank16[1] = h->rs->d_rank16b.as<uint16_t>();
  V.row_rank[0] = h->rs->d_rowrank.as<uint32_t>();
  V.row_rank[1] = h->rs->d_rowrankb.as<uint32_t>();
  V.cellref[0] = h->rs->d_cellref.as<uint16_t>();
  V.cellref[1] = h->rs->d_cellrefb.as<uint16_t>();
  V.cell_base[0] = h->rs->d_cellbase.as<uint32_t>();
  V.cell_base[1] = h->rs->d_cellbaseb.as<uint32_t>();
  V.reftab = h->rs->d_reftab.as<double>();
  V.ref_base = h->rs->d_refbase.as<uint32_t>();
  V.cand[0] = h->rs->d_cand.as<SprCand>();
  V.cand[1] = h->rs->d_cand1.as<SprCand>();
  V.grid = h->rs->R.grid;
  V.Tstar = h->rs->R.Tstar;
  V.Sstar = h->rs->R.Sstar;
  V.thr_dim = h->p.match_threshold_dimension;
  V.ignore_dim = h->p.ignore_dimension;
  h->lattice_ready = true;
  h->prepare_ms += now_ms() - t0;
  g_trace.mark("query_upload");
  return SLIDE_PR_OK;
}

// The hypothesis-list scorer runs on the default engine's structures: the landmark bins of the pair-join scorer, or the
// occupancy bitmaps of the lattice kernels.  Returns through *join which kernel to launch.
static int ensure_list_scorer(slide_pr_handle *h, cudaStream_t st, bool *join);

// The lattice kernels' structures, built the first time a call needs them for the prepared problem.
static int ensure_lattice(slide_pr_handle *h, cudaStream_t st) {
  if (h->lattice_ready) return SLIDE_PR_OK;
  const int rc = lattice_prepare(h);
  if (rc != SLIDE_PR_OK) return rc;
  SPR_CUDA(h, cudaEventRecord(h->ev_prep, h->stream));
  if (st != h->stream) SPR_CUDA(h, cudaStreamWaitEvent(st, h->ev_prep, 0));
  if (g_pageable_uploads.load() > 0) SPR_CUDA(h, cudaStreamSynchronize(h->stream));
  return SLIDE_PR_OK;
}

int slide_pr_prepare(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7, int32_t n_qry,
                     double half_x, double half_y) {
  if (!h) return SLIDE_PR_ERR_INVALID;
  return prepare_impl(h, &h->anon, false, ref7, n_ref, qry7, n_qry, half_x, half_y);
}

static int ensure_list_scorer(slide_pr_handle *h, cudaStream_t st, bool *join) {
  *join = join_is_default(h);
  if (!*join) {
    int rc = ensure_lattice(h, st);
    if (rc == SLIDE_PR_OK && h->rs->ranks_pending) rc = finish_ranks(h, st);
    return rc;
  }
  if (h->join_ready) return SLIDE_PR_OK;
  const int rc = join_prepare(h);
  if (rc != SLIDE_PR_OK) return rc;
  SPR_CUDA(h, cudaEventRecord(h->ev_prep, h->stream));
  if (st != h->stream) SPR_CUDA(h, cudaStreamWaitEvent(st, h->ev_prep, 0));
  return SLIDE_PR_OK;
}

static void fill_result_header(slide_pr_handle *h, slide_pr_match_result *out) {
  std::memset(out, 0, sizeof(*out));
  out->status = h->L.status;
  out->best_num_inliers = -10000;  // PR.cpp:125
  out->R_t[0] = out->R_t[4] = out->R_t[8] = 1.0;  // PR.cpp:127
  out->best_hyp_index = -1;
  out->n_rings = h->L.rings;
  out->n_yaw = (int32_t)h->L.yaw.size();
  out->n_translations = (int64_t)h->L.n_translations;
  out->prepare_ms = (float)h->prepare_ms;
  out->h2d_bytes = h->h2d_bytes;
  out->reuse = h->reuse_flags;
}

// Exact passes over plane (label l, direction d): row bands whose tables fit in shared memory.  Fills the band
// fields of K for band b; band_rows == 0 means "tables read in place" (one pass).  The rank tables must have
// been built (finish_ranks).
struct PlanePlan { uint32_t band_rows; int stage_reftab; uint32_t n_bands; };
static void set_band(slide_pr_handle *h, SprLaunch &K, uint32_t d, int l, const PlanePlan &P, uint32_t b) {
  const spr::RefIndex &R = h->rs->R;
  const uint32_t Rr = (uint32_t)R.grid.R[d];
  K.dir = d; K.label = l;
  const uint32_t label_cells = l >= 0 ? R.cell_base[d][l + 1] - R.cell_base[d][l] : 0u;
  K.tab_refs = l >= 0 ? R.ref_base[l + 1] - R.ref_base[l] : 0u;
  K.tab_cell_base = l >= 0 ? R.cell_base[d][l] : 0u;
  K.tab_ref_base = l >= 0 ? R.ref_base[l] : 0u;
  K.stage_reftab = P.stage_reftab;
  if (!P.band_rows || l < 0) { K.row_begin = K.row_end = 0; K.tab_rank_lo = 0; K.tab_cells = label_cells; return; }
  const uint32_t *rr = R.row_rank[d].data() + (size_t)l * Rr;
  K.row_begin = b * P.band_rows;
  K.row_end = std::min(Rr, (b + 1) * P.band_rows);
  K.tab_rank_lo = rr[K.row_begin];
  K.tab_cells = (K.row_end < Rr ? rr[K.row_end] : label_cells) - K.tab_rank_lo;
}
static PlanePlan plan_plane(slide_pr_handle *h, const SprLaunch &K0, uint32_t d, int l) {
  PlanePlan P{0u, 0, 1u};
  if (l < 0 || h->tables_mode != SPR_TABLES_AUTO) return P;
  const spr::RefIndex &R = h->rs->R;
  spr_score_plan(h->V, d, R.cell_base[d][l + 1] - R.cell_base[d][l], R.ref_base[l + 1] - R.ref_base[l], &P.band_rows, &P.stage_reftab);
  if (!P.band_rows) return P;
  P.n_bands = ((uint32_t)R.grid.R[d] + P.band_rows - 1) / P.band_rows;
  SprLaunch T = K0;
  for (uint32_t b = 0; b < P.n_bands; b++) {   // every band must really fit (the slot table is not spread evenly)
    set_band(h, T, d, l, P, b);
    if (spr_score_smem_warps(h->V, T, h->tables_mode) < 8) return PlanePlan{0u, 0, 1u};
  }
  return P;
}

// The pair-join scorer over the prepared problem: the exact inlier count of every hypothesis of the slice /
// shard (spr_join.cu), arg-max on the device.  Two launches: rotate + score.
static int join_search(slide_pr_handle *h, const slide_pr_search_opts &o, cudaStream_t st, slide_pr_match_result *out) {
  int rc;
  if (!h->join_ready) {
    if ((rc = join_prepare(h)) != SLIDE_PR_OK) return rc;
    SPR_CUDA(h, cudaEventRecord(h->ev_prep, h->stream));
    if (st != h->stream) SPR_CUDA(h, cudaStreamWaitEvent(st, h->ev_prep, 0));
  }
  const int n_yaw = (int)h->L.yaw.size();
  const int64_t n_trans = (int64_t)h->L.n_translations;
  const int64_t tb = std::min<int64_t>(o.trans_begin < 0 ? 0 : o.trans_begin, n_trans);
  const int64_t te = o.trans_end < 0 ? n_trans : std::max<int64_t>(std::min<int64_t>(o.trans_end, n_trans), tb);
  int64_t n_counts = 0;
  if (o.counts_out) {
    if (o.trans_end < 0) { h->err = "counts_out needs trans_end >= 0"; return SLIDE_PR_ERR_INVALID; }
    n_counts = (te - tb) * n_yaw;
    if (n_counts > o.counts_cap) { h->err = "counts_cap too small"; return SLIDE_PR_ERR_INVALID; }
    SPR_CUDA(h, h->d_counts.ensure(std::max<size_t>((size_t)n_counts, 1) * sizeof(int32_t)));
    SPR_CUDA(h, cudaMemsetAsync(h->d_counts.p, 0xff, std::max<size_t>((size_t)n_counts, 1) * sizeof(int32_t), st));
  }
  SPR_CUDA(h, cudaMemsetAsync(h->d_best.p, 0, sizeof(unsigned long long), st));
  // an inlier count reached elsewhere (another shard): see slide_pr_search
  const unsigned long long incumbent_key = o.incumbent_inliers > 0 ? ((unsigned long long)(o.incumbent_inliers + 1) << SPR_KEY_IDX_BITS) : 0ull;
  if (h->h_scalars.size() < 8) h->h_scalars.assign(8, 0ull);
  if (incumbent_key) {
    h->h_scalars[7] = incumbent_key;
    SPR_CUDA(h, cudaMemcpyAsync(h->d_best.p, h->h_scalars.data() + 7, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  }
  SPR_CUDA(h, cudaMemsetAsync(h->d_work.p, 0, sizeof(unsigned long long), st));
  SprJoinLaunch K{};
  K.work_counter = h->d_work.as<unsigned long long>();
  K.best_key = h->d_best.as<unsigned long long>();
  K.counts_out = o.counts_out ? h->d_counts.as<int32_t>() : nullptr;
  K.ord_begin = (unsigned long long)tb; K.ord_end = (unsigned long long)te;
  K.shard_index = o.shard_index; K.shard_count = o.shard_count;
  int launches = 0;
  SPR_CUDA(h, cudaEventRecord(h->ev0, st));
  if (h->JV.nqp > 0 && n_yaw > 0) {
    SPR_CUDA(h, spr_launch_join_rotate(h->JV, h->dj_qrot.as<double>(), h->dj_gbox.as<SprJoinBox>(), st));
    launches++;
  }
  if (n_yaw > 0 && te > tb && !h->j_blocks.empty()) {
    SPR_CUDA(h, spr_launch_join_score(h->JV, K, h->sm_count, st));
    launches++;
  }
  SPR_CUDA(h, cudaEventRecord(h->ev1, st));
  g_trace.mark("search_launch");
  unsigned long long *hs = h->h_scalars.data();
  hs[0] = 0ull;
  SPR_CUDA(h, cudaMemcpyAsync(hs, h->d_best.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (o.counts_out && n_counts > 0)
    SPR_CUDA(h, cudaMemcpyAsync(o.counts_out, h->d_counts.p, (size_t)n_counts * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SPR_CUDA(h, cudaStreamSynchronize(st));
  g_trace.mark("search_sync");
  float ms = 0;
  SPR_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  out->kernel_ms = ms;
  out->gpu_launches = launches;
  out->search_mode = 2;
  out->rings_scored = h->L.rings;
  out->h2d_bytes = h->h2d_bytes;
  out->d2h_bytes = (int64_t)sizeof(unsigned long long) + n_counts * (int64_t)sizeof(int32_t);
  // hypotheses scored by this shard: lattice samples of its blocks inside the slice x yaw candidates
  {
    const uint32_t sc = o.shard_count > 1 ? (uint32_t)o.shard_count : 1u, si = o.shard_count > 1 ? (uint32_t)o.shard_index : 0u;
    uint64_t samples = 0;
    for (size_t b = si; b < h->j_blocks.size(); b += sc) {
      const SprJoinBlock &B = h->j_blocks[b];
      if (tb == 0 && te == n_trans) { samples += (uint64_t)B.nx * B.ny; continue; }
      for (uint32_t i = 0; i < B.nx; i++) {
        const int64_t lo = std::max<int64_t>((int64_t)B.ord0 + (int64_t)i * B.row_stride, tb);
        const int64_t hi = std::min<int64_t>((int64_t)B.ord0 + (int64_t)i * B.row_stride + B.ny, te);
        if (hi > lo) samples += (uint64_t)(hi - lo);
      }
    }
    out->hypotheses_scored = (int64_t)(samples * (uint64_t)n_yaw);
  }
  const unsigned long long key = hs[0];
  if (key != 0ull && key != incumbent_key) {
    out->best_num_inliers = spr_key_count(key);
    out->best_hyp_index = spr_key_index(key);
  }
  return SLIDE_PR_OK;
}

int slide_pr_search(slide_pr_handle *h, const slide_pr_search_opts *opts, slide_pr_match_result *out) {
  if (!h || !out) return SLIDE_PR_ERR_INVALID;
  if (!h->prepared) { h->err = "slide_pr_search before slide_pr_prepare"; return SLIDE_PR_ERR_INVALID; }
  SPR_CUDA(h, cudaSetDevice(h->device));
  slide_pr_search_opts o{};
  o.trans_end = -1;
  if (opts) o = *opts;
  cudaStream_t st = o.stream ? (cudaStream_t)o.stream : h->stream;
  if (st != h->stream) SPR_CUDA(h, cudaStreamWaitEvent(st, h->ev_prep, 0));  // the uploads are asynchronous (page-locked sources)
  fill_result_header(h, out);
  if (h->L.status == SLIDE_PR_SANITY_RETURN) return SLIDE_PR_OK;
  int rc;
  // engine: the pair-join scorer (exhaustive = 4, or the default unless the handle says otherwise), or the
  // lattice kernels (1: every hypothesis verified; 2: bounds only; 3: bound-and-verify)
  {
    const bool want_join = o.exhaustive == 4 || (o.exhaustive == 0 && join_is_default(h));
    const bool can_join = join_supported(h) && !o.collect_stats && !o.reuse_bounds;
    if (o.exhaustive == 4 && !can_join) { h->err = "the pair-join scorer does not serve this request (statistics, reused bounds, a binding compute budget or more than 65535 query landmarks)"; return SLIDE_PR_ERR_UNSUPPORTED; }
    if (want_join && can_join) return join_search(h, o, st, out);
    if (o.exhaustive == 3 || o.exhaustive == 4) o.exhaustive = 0;
    if ((rc = ensure_lattice(h, st)) != SLIDE_PR_OK) return rc;
    out->prepare_ms = (float)h->prepare_ms;
    out->reuse = h->reuse_flags;
  }
  const int64_t tb = o.trans_begin < 0 ? 0 : o.trans_begin, te = o.trans_end;
  if (tb != h->lat_tb || te != h->lat_te) {  // re-chunk the lattice for the requested slice
    if ((rc = spr::build_lattice(h->p, h->half_x, h->half_y, h->yaw_half, tb, te, h->ring_major, h->L, h->err)) != SLIDE_PR_OK) return rc;
    h->lat_tb = tb; h->lat_te = te;
    h->lattice_valid = tb == 0 && te < 0;  // a sliced lattice is not the one prepare may reuse
    if ((rc = upload_lattice(h, st))) return rc;
  }
  const int n_yaw = (int)h->L.yaw.size();
  int64_t n_counts = 0;
  if (o.counts_out) {
    if (te < 0) { h->err = "counts_out needs trans_end >= 0"; return SLIDE_PR_ERR_INVALID; }
    const int64_t te_c = std::min<int64_t>(te, (int64_t)h->L.n_translations);
    n_counts = std::max<int64_t>(te_c - tb, 0) * n_yaw;
    if (n_counts > o.counts_cap) { h->err = "counts_cap too small"; return SLIDE_PR_ERR_INVALID; }
    SPR_CUDA(h, h->d_counts.ensure(std::max<size_t>((size_t)n_counts, 1) * sizeof(int32_t)));
    SPR_CUDA(h, cudaMemsetAsync(h->d_counts.p, 0xff, std::max<size_t>((size_t)n_counts, 1) * sizeof(int32_t), st));
  }
  SPR_CUDA(h, cudaMemsetAsync(h->d_best.p, 0, sizeof(unsigned long long), st));
  // an inlier count reached elsewhere (another shard): a synthetic best with that count and an
  // impossible index, so that an equal count found here still wins and anything below is pruned
  const unsigned long long incumbent_key = o.incumbent_inliers > 0 ? ((unsigned long long)(o.incumbent_inliers + 1) << SPR_KEY_IDX_BITS) : 0ull;
  if (incumbent_key) {
    if (h->h_scalars.size() < 8) h->h_scalars.assign(8, 0ull);
    h->h_scalars[7] = incumbent_key;
    SPR_CUDA(h, cudaMemcpyAsync(h->d_best.p, h->h_scalars.data() + 7, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  }
  if (o.collect_stats) SPR_CUDA(h, cudaMemsetAsync(h->d_stats.p, 0, 4 * sizeof(unsigned long long), st));

  SprLaunch K{};
  K.shard_index = o.shard_index;
  K.shard_count = o.shard_count;
  K.best_key = h->d_best.as<unsigned long long>();
  K.counts_out = o.counts_out ? h->d_counts.as<int32_t>() : nullptr;
  K.counts_cap = n_counts;
  K.ord_begin = (unsigned long long)tb;
  K.stats = o.collect_stats ? h->d_stats.as<unsigned long long>() : nullptr;
  K.work_counter = h->d_work.as<unsigned long long>();  // one work-item counter per pass
  size_t passes_left = 0;

  int launches = 0;
  SPR_CUDA(h, cudaEventRecord(h->ev0, st));
  if (h->V.nqp > 0 && n_yaw > 0) {
    SPR_CUDA(h, spr_launch_rotate(h->V, h->d_qrotq.as<int32_t>(), h->d_qrotq_yx.as<int32_t>(), h->d_qrot.as<double>(),
                                  h->d_gbox.as<SprBox>(), st));
    launches++;
  }
  // one pass per (chunk range, bitmap direction, label with queries); the per-hypothesis counters
  // travel between the label passes of a range in d_gcnt
  std::vector<int> active;
  for (int l = 0; l < h->V.n_labels; l++)
    if (h->Q.label_gseg[l + 1] > h->Q.label_gseg[l]) active.push_back(l);
  if (active.empty()) active.push_back(-1);  // no query can match: every hypothesis scores 0
  K.n_chunks_total = (uint32_t)h->L.chunks.size();
  auto next_counter = [&]() -> int {  // (re)arm a batch of work-item counters with a single memset
    if (passes_left == 0) {
      passes_left = 4096;
      K.work_counter = h->d_work.as<unsigned long long>();
      SPR_CUDA(h, cudaMemsetAsync(h->d_work.p, 0, 4096 * sizeof(unsigned long long), st));
    }
    return SLIDE_PR_OK;
  };
  // bound-and-verify (the lattice kernels' default): upper bounds of all hypotheses first, then exact verification of
  // those whose bound reaches the running best.  Exhaustive verification of every hypothesis
  // when per-hypothesis counts / statistics are requested, with a compute budget (ring by ring),
  // or on request (opts.exhaustive).
  const bool can_bound = !h->ring_major && active[0] >= 0 && h->V.nqp > 0 && h->V.nqp < 65536;
  // exhaustive = 2: bound phase only (first half of a sharded search; test hook when counts_out is given: it
  // receives the upper bounds).  Problems without a bound phase (no query label occurs in the reference map,
  // >= 65536 query landmarks, a compute budget that may bind) are searched exhaustively at once instead:
  // the result says so (search_mode = 0) and the caller skips the verification half.
  if (o.exhaustive == 2 && !can_bound && o.counts_out) { h->err = "upper bounds are not available for this problem"; return SLIDE_PR_ERR_UNSUPPORTED; }
  const bool bounds_only = o.exhaustive == 2 && can_bound;
  const bool prune = bounds_only || (!o.exhaustive && !o.counts_out && !o.collect_stats && can_bound && !h->force_exhaustive && h->p.exhaustive_search != 1);
  const int n_planes = spr_bound_planes(h->V.nqp);
  size_t cand_off[2] = {0, 0};
  if (prune) {
    const size_t n_wg_total = K.n_chunks_total / SPR_WARP_CHUNKS;
    SPR_CUDA(h, h->d_ubplanes.ensure((size_t)n_yaw * n_wg_total * (size_t)n_planes * 32 * sizeof(uint32_t) + 64));
    SPR_CUDA(h, h->d_itemub.ensure((size_t)n_yaw * n_wg_total * sizeof(uint32_t) + 64));
    SPR_CUDA(h, h->d_seed.ensure((size_t)n_yaw * SPR_SEED_SLOTS * sizeof(unsigned long long)));
    SPR_CUDA(h, cudaMemsetAsync(h->d_seed.p, 0, (size_t)n_yaw * SPR_SEED_SLOTS * sizeof(unsigned long long), st));
    SprBoundLaunch B{};
    B.n_chunks_total = K.n_chunks_total;
    B.shard_index = o.shard_index; B.shard_count = o.shard_count;
    B.planes = h->d_ubplanes.as<uint32_t>();
    B.item_ub = h->d_itemub.as<uint32_t>();
    B.seed_key = h->d_seed.as<unsigned long long>();
    // second half of a two-phase (sharded) search: the bounds of the preceding bound-only call are reused
    const bool reuse = o.reuse_bounds && !bounds_only && h->bounds_valid && tb == 0 && te < 0 &&
                       h->bounds_shard_index == o.shard_index && h->bounds_shard_count == o.shard_count;
    h->bounds_valid = false;
    for (uint32_t d = 0; d < 2 && !reuse; d++) {
      if (h->L.dir_end[d] <= h->L.dir_begin[d]) continue;
      int per = 1;
      uint32_t band_rows = 0;
      spr_bound_plan(h->V, d, (int)active.size(), h->bound_smem, &per, &band_rows);
      const uint32_t R = (uint32_t)h->V.grid.R[d];
      const uint32_t n_bands = band_rows ? (R + band_rows - 1) / band_rows : 1;
      for (uint32_t band = 0; band < n_bands; band++) {
        for (size_t i = 0; i < active.size(); i += (size_t)per) {
          B.chunk_begin = h->L.dir_begin[d]; B.chunk_end = h->L.dir_end[d]; B.dir = d;
          B.row_begin = band_rows ? band * band_rows : 0u;
          B.row_end = band_rows ? std::min(R, (band + 1) * band_rows) : 0u;
          B.n_labels = (int32_t)std::min<size_t>((size_t)per, active.size() - i);
          for (int k = 0; k < B.n_labels; k++) B.labels[k] = active[i + (size_t)k];
          B.first = band == 0 && i == 0;
          B.last = band + 1 == n_bands && i + (size_t)per >= active.size();
          if ((rc = next_counter()) != SLIDE_PR_OK) return rc;
          B.work_counter = K.work_counter;
          SPR_CUDA(h, spr_launch_bound_lattice(h->V, B, n_planes, h->sm_count, st, &launches));
          K.work_counter++;
          passes_left--;
        }
      }
    }
    if (h->rs->ranks_pending) {
      // the bound launches above keep the GPU busy: build the rank tables now and upload them on
      // the copy stream; the seed / verification kernels wait for the copies
      if ((rc = finish_ranks(h, h->copy_stream)) != SLIDE_PR_OK) return rc;
      SPR_CUDA(h, cudaEventRecord(h->ev_copy, h->copy_stream));
      SPR_CUDA(h, cudaStreamWaitEvent(st, h->ev_copy, 0));
    }
    if (!reuse) {
      SPR_CUDA(h, spr_launch_seed(h->V, h->d_seed.as<unsigned long long>(), K.best_key, st));
      launches++;
    }
    if (bounds_only && tb == 0 && te < 0) {
      h->bounds_valid = true;
      h->bounds_shard_index = o.shard_index; h->bounds_shard_count = o.shard_count;
    }
    // refinement: when many double groups stay candidates (no sharp peak, dense maps), their bounds
    // are recomputed against half-cell variants of the bitmaps (built on the device on demand); all
    // three kernels return at once when there are fewer than refine_min candidates
    // Worth its cost (about one more bound phase over the candidates) only when the exact
    // verification is expensive, i.e. when some pass has to read its tables in place (planes too
    // large for shared memory, e.g. 20 000-landmark maps); SLIDE_PR_REFINE_MIN=0 forces it.
    // (expensive = some plane does not fit in shared memory as a whole: it is verified in row bands or in place)
    bool slow_verify = h->refine_forced;
    for (uint32_t d = 0; d < 2 && !slow_verify; d++)
      for (size_t i = 0; i < active.size() && !slow_verify; i++) {
        const PlanePlan P = plan_plane(h, K, d, active[i]);
        slow_verify = P.band_rows == 0 || P.n_bands > 1;
      }
    // (not in the bound-only first half of a sharded search: the second half refines against the shared
    // incumbent, which is at least as good as this shard's own seeds)
    if (h->refine_min >= 0 && h->rs->R.mark_rc2 > 0 && slow_verify && !(bounds_only && !o.counts_out)) {
      size_t dcap[2];
      for (int d = 0; d < 2; d++) dcap[d] = (size_t)((h->L.dir_end[d] - h->L.dir_begin[d]) / (2 * SPR_WARP_CHUNKS)) * (size_t)n_yaw;
      const size_t vwords = 4 * (size_t)h->V.grid.label_stride * (size_t)std::max(h->V.n_labels, 1) + 16;
      SPR_CUDA(h, h->d_dgitems.ensure((dcap[0] + dcap[1]) * sizeof(uint32_t) + 64));
      SPR_CUDA(h, h->d_dgcount.ensure(2 * sizeof(uint32_t)));
      SPR_CUDA(h, h->rs->d_vbitmap.ensure(vwords * sizeof(uint32_t)));
      SPR_CUDA(h, cudaMemsetAsync(h->d_dgcount.p, 0, 2 * sizeof(uint32_t), st));
      h->V.vbitmap = h->rs->d_vbitmap.as<uint32_t>() + 4;
      for (uint32_t d = 0; d < 2; d++) {
        if (h->L.dir_end[d] <= h->L.dir_begin[d]) continue;
        B.chunk_begin = h->L.dir_begin[d]; B.chunk_end = h->L.dir_end[d]; B.dir = d;
        SPR_CUDA(h, spr_launch_select_dgroups(h->V, B, K.best_key, h->d_dgitems.as<uint32_t>() + (d ? dcap[0] : 0),
                                              h->d_dgcount.as<uint32_t>() + d, h->sm_count, st));
        launches++;
      }
      SPR_CUDA(h, spr_launch_variant_planes(h->V, h->rs->d_vbitmap.as<uint32_t>(), vwords, h->rs->d_ref7.as<double>(), h->rs->d_labof.as<int32_t>(),
                                            h->n_ref, h->p.match_xy_step_size, h->rs->R.mark_rc, h->rs->R.mark_rc2,
                                            h->d_dgcount.as<uint32_t>(), (uint32_t)h->refine_min, h->sm_count, st));
      launches += 2;
      for (uint32_t d = 0; d < 2; d++) {
        if (h->L.dir_end[d] <= h->L.dir_begin[d]) continue;
        // one label per launch with its four variant planes staged in row bands, or -- small query
        // maps, planes too wide -- up to 8 labels per launch reading the variants in place
        const uint32_t band_rows = spr_refine_band_rows(h->V, d, h->bound_smem);
        const uint32_t R = (uint32_t)h->V.grid.R[d];
        const uint32_t n_bands = band_rows ? (R + band_rows - 1) / band_rows : 1;
        const size_t per = band_rows ? 1 : SPR_BOUND_MAX_LABELS;
        for (size_t i = 0; i < active.size(); i += per) {
          for (uint32_t band = 0; band < n_bands; band++) {
            B.chunk_begin = h->L.dir_begin[d]; B.chunk_end = h->L.dir_end[d]; B.dir = d;
            B.row_begin = band_rows ? band * band_rows : 0u;
            B.row_end = band_rows ? std::min(R, (band + 1) * band_rows) : 0u;
            B.n_labels = (int32_t)std::min<size_t>(per, active.size() - i);
            for (int k = 0; k < B.n_labels; k++) B.labels[k] = active[i + (size_t)k];
            B.first = i == 0 && band == 0;
            B.last = i + per >= active.size() && band + 1 == n_bands;
            B.cand_items = h->d_dgitems.as<uint32_t>() + (d ? dcap[0] : 0);
            B.cand_count = h->d_dgcount.as<uint32_t>() + d;
            B.refine_min = (uint32_t)h->refine_min;
            if ((rc = next_counter()) != SLIDE_PR_OK) return rc;
            B.work_counter = K.work_counter;
            SPR_CUDA(h, spr_launch_bound_lattice(h->V, B, n_planes, h->sm_count, st, &launches));
            K.work_counter++;
            passes_left--;
          }
        }
      }
      B.cand_items = nullptr; B.cand_count = nullptr;
    }
    // candidate work items of each direction (largest bound >= seeded best)
    size_t cap[2];
    for (int d = 0; d < 2; d++) cap[d] = (size_t)((h->L.dir_end[d] - h->L.dir_begin[d]) / SPR_WARP_CHUNKS) * (size_t)n_yaw;
    SPR_CUDA(h, h->d_canditems.ensure((cap[0] + cap[1]) * sizeof(uint32_t) + 64));
    SPR_CUDA(h, h->d_candcount.ensure(2 * sizeof(uint32_t)));
    SPR_CUDA(h, cudaMemsetAsync(h->d_candcount.p, 0, 2 * sizeof(uint32_t), st));
    for (uint32_t d = 0; d < 2; d++) {
      if (h->L.dir_end[d] <= h->L.dir_begin[d]) continue;
      B.chunk_begin = h->L.dir_begin[d]; B.chunk_end = h->L.dir_end[d]; B.dir = d;
      SPR_CUDA(h, spr_launch_select_items(h->V, B, K.best_key, h->d_canditems.as<uint32_t>() + (d ? cap[0] : 0),
                                          h->d_candcount.as<uint32_t>() + d, h->sm_count, st));
      launches++;
    }
    cand_off[1] = cap[0];
    K.ub_planes = h->d_ubplanes.as<uint32_t>();
    K.item_ub = h->d_itemub.as<uint32_t>();
    K.ub_nplanes = n_planes;
  }
  if (h->rs->ranks_pending && !bounds_only && (rc = finish_ranks(h, st)) != SLIDE_PR_OK) return rc;  // exhaustive path
  auto run_range = [&](const uint32_t begin[2], const uint32_t end[2]) -> int {
    // verification phase of the pruned search: few candidate items per pass, so the label passes
    // of the two bitmap directions (disjoint counters) run concurrently on two streams
    const bool fork = prune && end[0] > begin[0] && end[1] > begin[1];
    if (fork) {
      if (passes_left < 1024) passes_left = 0;  // re-arm the work counters before the fork (one per pass: label x row band)
      if ((rc = next_counter()) != SLIDE_PR_OK) return rc;
      SPR_CUDA(h, cudaEventRecord(h->ev_fork, st));
      SPR_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    }
    for (uint32_t d = 0; d < 2; d++) {
      if (end[d] <= begin[d]) continue;
      cudaStream_t sd = fork && d == 1 ? h->side_stream : st;
      for (size_t i = 0; i < active.size(); i++) {
        const PlanePlan P = plan_plane(h, K, d, active[i]);
        for (uint32_t b = 0; b < P.n_bands; b++) {
          K.chunk_begin = begin[d]; K.chunk_end = end[d];
          set_band(h, K, d, active[i], P, b);
          if (prune) {
            K.cand_items = h->d_canditems.as<uint32_t>() + cand_off[d];
            K.cand_count = h->d_candcount.as<uint32_t>() + d;
          }
          K.first = i == 0 && b == 0; K.last = i + 1 == active.size() && b + 1 == P.n_bands;
          if ((rc = next_counter()) != SLIDE_PR_OK) return rc;
          SPR_CUDA(h, spr_launch_score_lattice(h->V, K, h->tables_mode, h->sm_count, sd, &launches));
          K.work_counter++;
          passes_left--;
        }
      }
    }
    if (fork) {
      SPR_CUDA(h, cudaEventRecord(h->ev_join, h->side_stream));
      SPR_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    }
    return SLIDE_PR_OK;
  };
  if (!bounds_only) {
    // per-hypothesis counters carried between the passes of a chunk range (labels x row bands) in HBM
    bool carry = active.size() > 1;
    for (uint32_t d = 0; d < 2 && !carry; d++) carry = plan_plane(h, K, d, active[0]).n_bands > 1;
    if (carry) {
      const size_t per = h->V.nqp > 65535 ? 4 : 2;
      SPR_CUDA(h, h->d_gcnt.ensure((size_t)n_yaw * (size_t)K.n_chunks_total * 32 * per + 64));
    }
    K.gcnt = h->d_gcnt.p;
  }
  int rings_scored = 0;
  if (bounds_only) {
    rings_scored = h->L.rings;
  } else if (h->ring_major) {
    // anytime behaviour of PR.cpp:181-191: whole seconds, checked before every ring
    const auto start = std::chrono::high_resolution_clock::now();
    for (size_t k = 0; k < h->L.ring.size(); k++) {
      const double duration = (double)std::chrono::duration_cast<std::chrono::seconds>(
                                  std::chrono::high_resolution_clock::now() - start).count();
      if (duration > h->p.compute_budget_sec) break;
      if ((rc = run_range(h->L.ring[k].dbegin, h->L.ring[k].dend)) != SLIDE_PR_OK) return rc;
      SPR_CUDA(h, cudaStreamSynchronize(st));
      rings_scored++;
    }
  } else {
    if ((rc = run_range(h->L.dir_begin, h->L.dir_end)) != SLIDE_PR_OK) return rc;
    rings_scored = h->L.rings;
  }
  SPR_CUDA(h, cudaEventRecord(h->ev1, st));
  g_trace.mark("search_launch");
  if (h->h_scalars.size() < 8) h->h_scalars.assign(8, 0ull);
  unsigned long long *hs = h->h_scalars.data();
  for (int i = 0; i < 5; i++) hs[i] = 0ull;
  SPR_CUDA(h, cudaMemcpyAsync(hs, h->d_best.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (o.collect_stats) SPR_CUDA(h, cudaMemcpyAsync(hs + 1, h->d_stats.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (o.counts_out && n_counts > 0 && !bounds_only)
    SPR_CUDA(h, cudaMemcpyAsync(o.counts_out, h->d_counts.p, (size_t)n_counts * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SPR_CUDA(h, cudaStreamSynchronize(st));
  if (bounds_only && o.counts_out && n_counts > 0) {  // decode the bit planes of the slice on the host
    const size_t n_wg_total = K.n_chunks_total / SPR_WARP_CHUNKS;
    std::vector<uint32_t> planes((size_t)n_yaw * n_wg_total * (size_t)n_planes * 32);
    SPR_CUDA(h, cudaMemcpy(planes.data(), h->d_ubplanes.p, planes.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n_counts; i++) o.counts_out[i] = -1;
    for (uint32_t c = 0; c < K.n_chunks_total; c++) {
      const SprChunk &ch = h->L.chunks[c];
      for (int b = 0; b < 32; b++) {
        if (!((ch.valid >> b) & 1u)) continue;
        const int64_t ord = (int64_t)ch.ord_base + (int64_t)b * ch.ord_stride;
        for (int a = 0; a < n_yaw; a++) {
          const int64_t slot = (ord - tb) * n_yaw + a;
          if (slot < 0 || slot >= n_counts) continue;
          const uint32_t *pp = planes.data() + (((size_t)a * n_wg_total + c / SPR_WARP_CHUNKS) * (size_t)n_planes) * 32 + (c % SPR_WARP_CHUNKS);
          int32_t v = 0;
          for (int i = 0; i < n_planes; i++) v |= (int32_t)((pp[(size_t)i * 32] >> b) & 1u) << i;
          o.counts_out[slot] = v;
        }
      }
    }
  }
  g_trace.mark("search_sync");
  const unsigned long long key = hs[0], stats[4] = {hs[1], hs[2], hs[3], hs[4]};
  float ms = 0;
  SPR_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  out->kernel_ms = ms;
  out->gpu_launches = launches;
  out->search_mode = prune ? 1 : 0;
  out->rings_scored = rings_scored;
  out->filter_hits = (int64_t)stats[0];
  out->groups_probed = (int64_t)stats[2];
  out->groups_skipped = (int64_t)stats[3];
  out->h2d_bytes = h->h2d_bytes;
  out->d2h_bytes = (int64_t)sizeof(key) + (o.collect_stats ? (int64_t)sizeof(stats) : 0) + n_counts * (int64_t)sizeof(int32_t);
  // hypotheses scored by this shard = valid bits of its double groups (64 chunks) x yaw candidates
  {
    const int sc = o.shard_count > 1 ? o.shard_count : 1, si = o.shard_count > 1 ? o.shard_index : 0;
    uint64_t bits = 0;
    auto count_range = [&](uint32_t cb, uint32_t ce) {
      for (uint32_t g = cb / 64 + (uint32_t)si; g < ce / 64; g += (uint32_t)sc) bits += h->L.dg_bits[g];
    };
    if (h->ring_major) {
      for (int k = 0; k < rings_scored; k++)
        for (int d = 0; d < 2; d++) count_range(h->L.ring[k].dbegin[d], h->L.ring[k].dend[d]);
    } else {
      for (int d = 0; d < 2; d++) count_range(h->L.dir_begin[d], h->L.dir_end[d]);
    }
    out->hypotheses_scored = (int64_t)(bits * (uint64_t)n_yaw);
  }
  if (key != 0ull && key != incumbent_key) {  // the synthetic incumbent itself: nothing at least as good here
    out->best_num_inliers = spr_key_count(key);
    out->best_hyp_index = spr_key_index(key);
  }
  return SLIDE_PR_OK;
}

int slide_pr_lattice_info(const slide_pr_handle *h, int64_t *n_translations, int32_t *n_yaw, int32_t *n_rings) {
  if (!h || !h->prepared) return SLIDE_PR_ERR_INVALID;
  if (n_translations) *n_translations = (int64_t)h->L.n_translations;
  if (n_yaw) *n_yaw = (int32_t)h->L.yaw.size();
  if (n_rings) *n_rings = h->L.rings;
  return h->L.status;
}

int slide_pr_extract(slide_pr_handle *h, int64_t hyp_index, int32_t *ref_idx_out, int32_t *qry_idx_out,
                     slide_pr_match_result *io) {
  if (!h || !io) return SLIDE_PR_ERR_INVALID;
  if (!h->prepared) { h->err = "slide_pr_extract before slide_pr_prepare"; return SLIDE_PR_ERR_INVALID; }
  SPR_CUDA(h, cudaSetDevice(h->device));
  const int n_yaw = (int)h->L.yaw.size();
  io->n_matched = 0;
  if (hyp_index < 0 || n_yaw <= 0) return SLIDE_PR_OK;
  // the lattice may have been re-chunked for a slice: ordinals are global, translation_of uses rings only
  double tx, ty;
  int ring;
  if (!spr::translation_of(h->L, (uint64_t)(hyp_index / n_yaw), &tx, &ty, &ring)) { h->err = "hypothesis index out of range"; return SLIDE_PR_ERR_INVALID; }
  const int a = (int)(hyp_index % n_yaw);
  const double c = h->L.cs[2 * a], s = h->L.cs[2 * a + 1];
  io->R_t[0] = c; io->R_t[1] = -s; io->R_t[2] = tx;  // PR.cpp:246-251
  io->R_t[3] = s; io->R_t[4] = c;  io->R_t[5] = ty;
  io->R_t[6] = 0; io->R_t[7] = 0;  io->R_t[8] = 1;
  cudaStream_t st = h->stream;
  SPR_CUDA(h, spr_launch_extract(h->rs->d_ref7.as<double>(), h->n_ref, h->d_qry7.as<double>(), h->n_qry, c, s, tx, ty,
                                 h->Tstar, h->Sstar, h->p.match_threshold_dimension, h->p.ignore_dimension,
                                 h->d_match.as<int32_t>(), st));
  io->gpu_launches += 1;
  if ((int)h->h_match.size() < std::max(h->n_qry, 1)) h->h_match.resize(std::max(h->n_qry, 1));
  SPR_CUDA(h, cudaMemcpyAsync(h->h_match.data(), h->d_match.p, (size_t)h->n_qry * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SPR_CUDA(h, cudaStreamSynchronize(st));
  g_trace.mark("extract");
  int k = 0;
  for (int j = 0; j < h->n_qry; j++)
    if (h->h_match[j] >= 0) {
      if (ref_idx_out) ref_idx_out[k] = h->h_match[j];
      if (qry_idx_out) qry_idx_out[k] = j;
      k++;
    }
  io->n_matched = k;
  io->d2h_bytes += (int64_t)h->n_qry * (int64_t)sizeof(int32_t);
  return SLIDE_PR_OK;
}

static int match_maps_impl(slide_pr_handle *h, RefSide *slot, bool trusted, const double *ref7, int32_t n_ref, const double *qry7,
                           int32_t n_qry, double half_x, double half_y, int32_t *ref_idx_out, int32_t *qry_idx_out,
                           slide_pr_match_result *out) {
  if (!h || !out) return SLIDE_PR_ERR_INVALID;
  int rc = prepare_impl(h, slot, trusted, ref7, n_ref, qry7, n_qry, half_x, half_y);
  if (rc != SLIDE_PR_OK) return rc;
  if ((rc = slide_pr_search(h, nullptr, out)) != SLIDE_PR_OK) return rc;
  if (out->status == SLIDE_PR_SANITY_RETURN || out->best_hyp_index < 0) return SLIDE_PR_OK;
  if ((rc = slide_pr_extract(h, out->best_hyp_index, ref_idx_out, qry_idx_out, out)) != SLIDE_PR_OK) return rc;
  if (out->n_matched != out->best_num_inliers) {  // brute-force recount of the winner must agree with the index path
    char buf[160];
    std::snprintf(buf, sizeof(buf), "self-check failed: indexed count %d != brute-force count %d for hypothesis %lld",
                  out->best_num_inliers, out->n_matched, (long long)out->best_hyp_index);
    h->err = buf;
    return SLIDE_PR_ERR_INTERNAL;
  }
  return SLIDE_PR_OK;
}

int slide_pr_match_maps(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7, int32_t n_qry,
                        double half_x, double half_y, int32_t *ref_idx_out, int32_t *qry_idx_out,
                        slide_pr_match_result *out) {
  if (!h || !out) return SLIDE_PR_ERR_INVALID;
  return match_maps_impl(h, &h->anon, false, ref7, n_ref, qry7, n_qry, half_x, half_y, ref_idx_out, qry_idx_out, out);
}

void slide_pr_get_xyz_yaw_from_tf(const double *tf16, double *xyz_yaw4) { spr::xyz_yaw_from_tf(tf16, xyz_yaw4); }

int slide_pr_solve_lsq(const double *tgt3, const double *src3, int32_t k, double *xyz_yaw4, double *transform16) {
  if (!tgt3 || !src3 || !xyz_yaw4 || !transform16 || k < 0) return SLIDE_PR_ERR_INVALID;
  spr::solve_lsq(tgt3, src3, k, xyz_yaw4, transform16);
  return SLIDE_PR_OK;
}

// findTransformation after MatchMaps (PR.cpp:845-944): inlier gate, then the raw lattice transform with the
// centroid shift reverted, or the closed-form refinement on the matched pairs.  ref / qry: the rows MatchMaps
// saw (centroid-shifted in inter-robot mode).
static int finish_transformation(const slide_pr_params &p, const double *ref, const double *qry, const double *cref, const double *cqry,
                                 const int32_t *ri, const int32_t *qi, slide_pr_tf_result *out) {
  const slide_pr_match_result &m = out->match;
  std::memcpy(out->R_t, m.R_t, sizeof(m.R_t));
  // findTransformation starts from best_num_inliers_out = 0 and MatchMaps leaves it untouched on
  // its sanity-check return (PR.cpp:819, 169-175)
  out->best_num_inliers = m.status == SLIDE_PR_SANITY_RETURN ? 0 : m.best_num_inliers;
  out->n_matched = m.status == SLIDE_PR_SANITY_RETURN ? 0 : m.n_matched;
  if (out->best_num_inliers < p.min_num_inliers) { out->found = 0; return SLIDE_PR_NOT_FOUND; }  // PR.cpp:849
  out->found = 1;
  if (!p.use_lsq) {                                                    // PR.cpp:882-905
    double raw[16] = {0};
    raw[0] = m.R_t[0]; raw[1] = m.R_t[1]; raw[4] = m.R_t[3]; raw[5] = m.R_t[4];
    raw[10] = 1; raw[15] = 1;
    raw[3] = m.R_t[2]; raw[7] = m.R_t[5]; raw[11] = 0;
    if (p.inter_loop_closure) {                                        // PR.cpp:947-967
      double H1[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, H2[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      H1[3] = cref[0]; H1[7] = cref[1];
      H2[3] = -cqry[0]; H2[7] = -cqry[1];
      double T1[16];
      spr::mat4_mul(H1, raw, T1);
      spr::mat4_mul(T1, H2, out->transform);
    } else {
      std::memcpy(out->transform, raw, sizeof(raw));
    }
    spr::xyz_yaw_from_tf(out->transform, out->xyz_yaw);
  } else {                                                             // PR.cpp:906-944
    const int k = out->n_matched;
    std::vector<double> tgt(3 * (size_t)std::max(k, 1)), src(3 * (size_t)std::max(k, 1));
    for (int i = 0; i < k; i++) {
      const double *r = ref + 7 * (size_t)ri[i], *q = qry + 7 * (size_t)qi[i];
      tgt[3 * i] = r[1]; tgt[3 * i + 1] = r[2]; tgt[3 * i + 2] = r[3];
      src[3 * i] = q[1]; src[3 * i + 1] = q[2]; src[3 * i + 2] = q[3];
      if (p.inter_loop_closure) {                                      // PR.cpp:925-937
        tgt[3 * i] += cref[0]; tgt[3 * i + 1] += cref[1];
        src[3 * i] += cqry[0]; src[3 * i + 1] += cqry[1];
      }
    }
    spr::solve_lsq(tgt.data(), src.data(), k, out->xyz_yaw, out->transform);
  }
  return SLIDE_PR_OK;
}

int slide_pr_find_transformation(slide_pr_handle *h, const double *ref7_in, int32_t n_ref, const double *qry7_in,
                                 int32_t n_qry, int32_t *ref_idx_out, int32_t *qry_idx_out, slide_pr_tf_result *out) {
  if (!h || !out) return SLIDE_PR_ERR_INVALID;
  if (n_ref < 0 || n_qry < 0 || (n_ref > 0 && !ref7_in) || (n_qry > 0 && !qry7_in)) { h->err = "bad map arguments"; return SLIDE_PR_ERR_INVALID; }
  std::memset(out, 0, sizeof(*out));
  g_trace.start();
  std::vector<double> ref(ref7_in, ref7_in + (size_t)n_ref * 7), qry(qry7_in, qry7_in + (size_t)n_qry * 7);
  double cref[2] = {0, 0}, cqry[2] = {0, 0};
  double half_x, half_y;
  const slide_pr_params &p = h->p;
  if (p.inter_loop_closure) {
    for (int i = 0; i < n_ref; i++) { cref[0] += ref[7 * (size_t)i + 1]; cref[1] += ref[7 * (size_t)i + 2]; }  // PR.cpp:713-722
    cref[0] /= (double)n_ref; cref[1] /= (double)n_ref;
    for (int i = 0; i < n_qry; i++) { cqry[0] += qry[7 * (size_t)i + 1]; cqry[1] += qry[7 * (size_t)i + 2]; }
    cqry[0] /= (double)n_qry; cqry[1] /= (double)n_qry;
    double bxr = 0, byr = 0, bxq = 0, byq = 0;
    for (int i = 0; i < n_ref; i++) {                                  // PR.cpp:755-759, 724-734
      ref[7 * (size_t)i + 1] -= cref[0]; ref[7 * (size_t)i + 2] -= cref[1];
      bxr = std::max(bxr, std::abs(ref[7 * (size_t)i + 1])); byr = std::max(byr, std::abs(ref[7 * (size_t)i + 2]));
    }
    for (int i = 0; i < n_qry; i++) {
      qry[7 * (size_t)i + 1] -= cqry[0]; qry[7 * (size_t)i + 2] -= cqry[1];
      bxq = std::max(bxq, std::abs(qry[7 * (size_t)i + 1])); byq = std::max(byq, std::abs(qry[7 * (size_t)i + 2]));
    }
    double max_x = std::max(bxr, bxq), max_y = std::max(byr, byq);     // PR.cpp:771-774
    if (!p.disable_yaw_search) { const double m = std::max(max_x, max_y); max_x = m; max_y = m; }  // :777-782
    half_x = max_x * p.dilation_factor;                                // :786-787
    half_y = max_y * p.dilation_factor;
    out->yaw_half = p.match_yaw_half_range;
  } else {
    half_x = p.match_x_half_range_intra;                               // :808-810
    half_y = p.match_y_half_range_intra;
    out->yaw_half = p.match_yaw_half_range_intra;
  }
  out->half_x = half_x; out->half_y = half_y;
  out->centroid_ref[0] = cref[0]; out->centroid_ref[1] = cref[1];
  out->centroid_qry[0] = cqry[0]; out->centroid_qry[1] = cqry[1];

  std::vector<int32_t> ri_own, qi_own;
  int32_t *ri = ref_idx_out, *qi = qry_idx_out;
  if (!ri) { ri_own.resize(std::max(n_qry, 1)); ri = ri_own.data(); }
  if (!qi) { qi_own.resize(std::max(n_qry, 1)); qi = qi_own.data(); }
  g_trace.mark("centroid_shift");
  int rc = slide_pr_match_maps(h, ref.data(), n_ref, qry.data(), n_qry, half_x, half_y, ri, qi, &out->match);
  g_trace.flush("find_transformation");
  if (rc != SLIDE_PR_OK) return rc;
  return finish_transformation(p, ref.data(), qry.data(), cref, cqry, ri, qi, out);
}

int slide_pr_find_inter_loop_closure(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7,
                                     int32_t n_qry, double *tf16, slide_pr_tf_result *out_opt) {
  if (!h || !tf16) return SLIDE_PR_ERR_INVALID;
  slide_pr_tf_result local, *out = out_opt ? out_opt : &local;
  std::memset(out, 0, sizeof(*out));
  if (n_ref < h->p.min_num_map_objects_to_start || n_qry < h->p.min_num_map_objects_to_start)  // PR.cpp:508-510
    return SLIDE_PR_NOT_FOUND;
  const int rc = slide_pr_find_transformation(h, ref7, n_ref, qry7, n_qry, nullptr, nullptr, out);
  if (rc != SLIDE_PR_OK) return rc;
  const double x = out->xyz_yaw[0], y = out->xyz_yaw[1], z = out->xyz_yaw[2], yaw = out->xyz_yaw[3];
  for (int i = 0; i < 16; i++) tf16[i] = 0;                            // PR.cpp:523-536
  tf16[0] = std::cos(yaw); tf16[1] = -std::sin(yaw); tf16[4] = std::sin(yaw); tf16[5] = std::cos(yaw);
  tf16[10] = 1; tf16[15] = 1;
  tf16[3] = x; tf16[7] = y; tf16[11] = z;
  return SLIDE_PR_OK;
}

// measurements (query's local frame) into the map frame, PR.cpp:421-439
static void intra_move_measurements(const double *meas7, int32_t n_meas, const double *P, std::vector<double> &moved) {
  moved.resize((size_t)n_meas * 7);
  for (int i = 0; i < n_meas; i++) {
    const double *m = meas7 + 7 * (size_t)i;
    double v[4];
    for (int r = 0; r < 4; r++) v[r] = ((P[r * 4] * m[1] + P[r * 4 + 1] * m[2]) + P[r * 4 + 2] * m[3]) + P[r * 4 + 3] * 1.0;
    double *o = moved.data() + 7 * (size_t)i;
    o[0] = m[0]; o[1] = v[0] / v[3]; o[2] = v[1] / v[3]; o[3] = v[2] / v[3];
    o[4] = m[4]; o[5] = m[5]; o[6] = m[6];
  }
}

// loop-closure transform from the lattice / LSQ result and the two poses, PR.cpp:455-494
static void intra_compose_tf(const slide_pr_tf_result *out, const double *query_pose16, const double *candidate_pose16, double *tf16) {
  const double yaw = out->xyz_yaw[3];
  double lc[16] = {0};
  lc[0] = std::cos(yaw); lc[1] = -std::sin(yaw); lc[4] = std::sin(yaw); lc[5] = std::cos(yaw);
  lc[10] = 1; lc[15] = 1;
  lc[3] = out->xyz_yaw[0]; lc[7] = out->xyz_yaw[1]; lc[11] = 0.0;
  double cinv[16], drift[16];
  spr::mat4_rigid_inverse(candidate_pose16, cinv);
  spr::mat4_mul(cinv, query_pose16, drift);                            // PR.cpp:478
  spr::mat4_mul(drift, lc, tf16);                                      // PR.cpp:483-494
}

int slide_pr_find_intra_loop_closure(slide_pr_handle *h, const double *meas7, int32_t n_meas, const double *submap7,
                                     int32_t n_sub, const double *query_pose16, const double *candidate_pose16,
                                     double *tf16, slide_pr_tf_result *out_opt) {
  if (!h || !tf16 || !query_pose16 || !candidate_pose16) return SLIDE_PR_ERR_INVALID;
  slide_pr_tf_result local, *out = out_opt ? out_opt : &local;
  std::memset(out, 0, sizeof(*out));
  if (n_meas == 0 || n_sub == 0) return SLIDE_PR_NOT_FOUND;            // PR.cpp:395-398
  if (n_meas < 4) return SLIDE_PR_NOT_FOUND;                           // PR.cpp:400-403
  std::vector<double> moved;
  intra_move_measurements(meas7, n_meas, query_pose16, moved);
  const int32_t saved = h->p.inter_loop_closure;
  h->p.inter_loop_closure = 0;  // the caller's intra instance has inter_loop_closure = false (sloamNode.cpp:23)
  const int rc = slide_pr_find_transformation(h, submap7, n_sub, moved.data(), n_meas, nullptr, nullptr, out);
  h->p.inter_loop_closure = saved;
  if (rc != SLIDE_PR_OK) return rc;
  intra_compose_tf(out, query_pose16, candidate_pose16, tf16);
  return SLIDE_PR_OK;
}

// Several candidate key poses for ONE set of measurements (the reference tries one
// candidate per attempt, sloamNode.cpp:355-486).  The candidates' searches are enqueued back to back on the
// handle's stream -- the host prepares candidate k + 1 while the GPU scores candidate k -- and the host waits
// once for all of them: one synchronisation for the searches, one for the correspondences.
int slide_pr_find_intra_loop_closure_batch(slide_pr_handle *h, const double *meas7, int32_t n_meas, const double *const *submaps7,
                                           const int32_t *n_subs, const double *query_pose16, const double *candidate_poses16,
                                           int32_t n_cand, double *tf16_out, slide_pr_tf_result *out) {
  if (!h || !out || !tf16_out || !query_pose16 || !candidate_poses16 || n_cand < 0 || (n_cand > 0 && (!submaps7 || !n_subs))) return SLIDE_PR_ERR_INVALID;
  if (n_meas < 0 || (n_meas > 0 && !meas7)) { h->err = "bad measurement arguments"; return SLIDE_PR_ERR_INVALID; }
  for (int k = 0; k < n_cand; k++) {
    std::memset(&out[k], 0, sizeof(out[k]));
    if (n_subs[k] < 0 || (n_subs[k] > 0 && !submaps7[k])) { h->err = "bad submap arguments"; return SLIDE_PR_ERR_INVALID; }
  }
  if (n_cand == 0 || n_meas < 4) return SLIDE_PR_OK;                   // PR.cpp:395-403: no closure for any candidate
  const int32_t saved = h->p.inter_loop_closure;
  struct Restore { slide_pr_handle *h; int32_t v; ~Restore() { h->p.inter_loop_closure = v; } } restore{h, saved};
  h->p.inter_loop_closure = 0;  // the caller's intra instance has inter_loop_closure = false (sloamNode.cpp:23)
  const double hx = h->p.match_x_half_range_intra, hy = h->p.match_y_half_range_intra;   // PR.cpp:808-810
  const bool pipelined = h->p.exhaustive_search == 0 && !h->force_exhaustive && !h->env_lattice && n_meas <= 65535 && n_cand <= 4096 &&
                         !budget_may_bind(h->p, hx, hy, h->p.match_yaw_half_range_intra, n_meas);
  if (!pipelined) {   // the lattice kernels keep per-search state on the handle: one candidate at a time
    h->p.inter_loop_closure = saved;
    for (int k = 0; k < n_cand; k++) {
      const int rc = slide_pr_find_intra_loop_closure(h, meas7, n_meas, submaps7[k], n_subs[k], query_pose16, candidate_poses16 + 16 * (size_t)k,
                                                      tf16_out + 16 * (size_t)k, &out[k]);
      if (rc < 0) return rc;
    }
    return SLIDE_PR_OK;
  }
  SPR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  std::vector<double> moved;
  intra_move_measurements(meas7, n_meas, query_pose16, moved);
  DevBuf &d_keys = h->d_batch_keys, &d_matches = h->d_batch_match;
  SPR_CUDA(h, d_keys.ensure((size_t)n_cand * sizeof(unsigned long long)));
  SPR_CUDA(h, d_matches.ensure((size_t)n_cand * (size_t)n_meas * sizeof(int32_t)));
  SPR_CUDA(h, cudaMemsetAsync(d_keys.p, 0, (size_t)n_cand * sizeof(unsigned long long), st));
  if (h->h_batch_keys.size() < (size_t)n_cand) h->h_batch_keys.resize((size_t)n_cand);
  if (h->h_batch_match.size() < (size_t)n_cand * (size_t)n_meas) h->h_batch_match.resize((size_t)n_cand * (size_t)n_meas);
  std::vector<std::unique_ptr<RefSide>> slots((size_t)n_cand);
  auto give_back = [&]() {
    cudaStreamSynchronize(st);
    h->rs = &h->anon; h->prepared = false;
    for (auto &sl : slots) if (sl) retire_slot(h, std::move(sl));
  };
  int rc = SLIDE_PR_OK;
  bool work_zeroed = false;
  int n_yaw = 0;
  bool sanity = false;
  for (int k = 0; k < n_cand && rc == SLIDE_PR_OK && !sanity; k++) {
    if (n_subs[k] == 0) continue;                                      // PR.cpp:395-398
    if (!h->free_slots.empty()) { slots[k] = std::move(h->free_slots.back()); h->free_slots.pop_back(); }
    else slots[k].reset(new RefSide());
    RefSide &E = *slots[k];
    E.cached_ref.assign(submaps7[k], submaps7[k] + (size_t)n_subs[k] * 7);
    E.rows_valid = true; E.ref_index_valid = false; E.ranks_pending = false; E.join_valid = false; E.ref7_uploaded = false;
    E.n_rows = n_subs[k]; E.shifted = false;
    if ((rc = prepare_impl(h, &E, true, E.cached_ref.data(), n_subs[k], moved.data(), n_meas, hx, hy)) != SLIDE_PR_OK) break;
    if (h->L.status == SLIDE_PR_SANITY_RETURN) { sanity = true; break; }   // the lattice is the same for every candidate
    if (!h->join_ready) { h->err = "internal: pair-join structures missing in the batched intra search"; rc = SLIDE_PR_ERR_INTERNAL; break; }
    n_yaw = (int)h->L.yaw.size();
    if (!work_zeroed) {   // after the first prepare: d_work exists
      if (cudaMemsetAsync(h->d_work.p, 0, (size_t)n_cand * sizeof(unsigned long long), st) != cudaSuccess) { h->err = "cudaMemsetAsync"; rc = SLIDE_PR_ERR_CUDA; break; }
      work_zeroed = true;
    }
    SprJoinLaunch K{};
    K.work_counter = h->d_work.as<unsigned long long>() + k;
    K.best_key = d_keys.as<unsigned long long>() + k;
    K.ord_begin = 0ull; K.ord_end = (unsigned long long)h->L.n_translations;
    cudaError_t e = cudaSuccess;
    if (h->JV.nqp > 0 && n_yaw > 0) e = spr_launch_join_rotate(h->JV, h->dj_qrot.as<double>(), h->dj_gbox.as<SprJoinBox>(), st);
    if (e == cudaSuccess && n_yaw > 0 && !h->j_blocks.empty()) e = spr_launch_join_score(h->JV, K, h->sm_count, st);
    if (e != cudaSuccess) { h->err = std::string("batched intra search: ") + cudaGetErrorString(e); rc = SLIDE_PR_ERR_CUDA; break; }
    fill_result_header(h, &out[k].match);
    out[k].match.search_mode = 2;
    out[k].match.gpu_launches = 2;
    out[k].match.rings_scored = h->L.rings;
    out[k].match.hypotheses_scored = (int64_t)h->L.n_translations * n_yaw;
  }
  if (rc != SLIDE_PR_OK) { give_back(); return rc; }
  if (sanity) {   // MatchMaps' early return for every candidate: nothing found (PR.cpp:169-175, 819, 849)
    for (int k = 0; k < n_cand; k++) { out[k].match.status = SLIDE_PR_SANITY_RETURN; out[k].half_x = hx; out[k].half_y = hy; }
    give_back();
    return SLIDE_PR_OK;
  }
  if (cudaMemcpyAsync(h->h_batch_keys.data(), d_keys.p, (size_t)n_cand * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) { h->err = "batched intra search: copy of the results failed"; give_back(); return SLIDE_PR_ERR_CUDA; }
  // correspondences of every candidate's winner: all extraction kernels, then one wait
  for (int k = 0; k < n_cand; k++) {
    if (!slots[k]) continue;
    const unsigned long long key = h->h_batch_keys[k];
    slide_pr_match_result &m = out[k].match;
    if (key == 0ull || n_yaw <= 0) continue;
    m.best_num_inliers = spr_key_count(key);
    m.best_hyp_index = spr_key_index(key);
    double tx, ty;
    int ring;
    if (!spr::translation_of(h->L, (uint64_t)(m.best_hyp_index / n_yaw), &tx, &ty, &ring)) { h->err = "hypothesis index out of range"; give_back(); return SLIDE_PR_ERR_INTERNAL; }
    const int a = (int)(m.best_hyp_index % n_yaw);
    const double c = h->L.cs[2 * a], s = h->L.cs[2 * a + 1];
    m.R_t[0] = c; m.R_t[1] = -s; m.R_t[2] = tx;  // PR.cpp:246-251
    m.R_t[3] = s; m.R_t[4] = c;  m.R_t[5] = ty;
    m.R_t[6] = 0; m.R_t[7] = 0;  m.R_t[8] = 1;
    const cudaError_t e = spr_launch_extract(slots[k]->d_ref7.as<double>(), n_subs[k], h->d_qry7.as<double>(), n_meas, c, s, tx, ty, h->Tstar, h->Sstar,
                                             h->p.match_threshold_dimension, h->p.ignore_dimension, d_matches.as<int32_t>() + (size_t)k * n_meas, st);
    if (e != cudaSuccess) { h->err = std::string("batched intra search: ") + cudaGetErrorString(e); give_back(); return SLIDE_PR_ERR_CUDA; }
    m.gpu_launches += 1;
  }
  if (cudaMemcpyAsync(h->h_batch_match.data(), d_matches.p, (size_t)n_cand * (size_t)n_meas * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) { h->err = "batched intra search: copy of the correspondences failed"; give_back(); return SLIDE_PR_ERR_CUDA; }
  std::vector<int32_t> ri((size_t)std::max(n_meas, 1)), qi((size_t)std::max(n_meas, 1));
  const double zero2[2] = {0, 0};
  for (int k = 0; k < n_cand; k++) {
    out[k].half_x = hx; out[k].half_y = hy; out[k].yaw_half = h->p.match_yaw_half_range_intra;
    if (!slots[k]) continue;
    slide_pr_match_result &m = out[k].match;
    int n = 0;
    if (m.best_hyp_index >= 0) {
      const int32_t *mt = h->h_batch_match.data() + (size_t)k * n_meas;
      for (int j = 0; j < n_meas; j++)
        if (mt[j] >= 0) { ri[n] = mt[j]; qi[n] = j; n++; }
      if (n != m.best_num_inliers) {   // brute-force recount of the winner must agree with the scorer
        char buf[160];
        std::snprintf(buf, sizeof(buf), "self-check failed: candidate %d, counted %d != brute-force %d for hypothesis %lld", k, m.best_num_inliers, n, (long long)m.best_hyp_index);
        h->err = buf;
        give_back();
        return SLIDE_PR_ERR_INTERNAL;
      }
    }
    m.n_matched = n;
    if (finish_transformation(h->p, slots[k]->cached_ref.data(), moved.data(), zero2, zero2, ri.data(), qi.data(), &out[k]) == SLIDE_PR_OK)
      intra_compose_tf(&out[k], query_pose16, candidate_poses16 + 16 * (size_t)k, tf16_out + 16 * (size_t)k);
  }
  give_back();
  return SLIDE_PR_OK;
}

// ---- device-resident map cache ---------------------------------------------
// The reference keeps one object map per robot (databaseManager::robotMapDict_, databaseManager.h:99-102)
// and deep-copies both maps of a pair for every attempt (sloamNode.cpp:603-614).  Here a map is handed
// over once per version: its centroid-shifted rows stay page-locked on the host and on the device, and
// the reference-side index (occupancy bitmaps, rank tables) is built the first time the map is searched
// AS A REFERENCE and reused by every later pair until the version changes.
int slide_pr_map_cache_put(slide_pr_handle *h, int64_t robot_id, uint64_t version, const double *rows7, int32_t n) {
  if (!h || n < 0 || (n > 0 && !rows7)) return SLIDE_PR_ERR_INVALID;
  SPR_CUDA(h, cudaSetDevice(h->device));
  auto it = h->cache.find(robot_id);
  if (it != h->cache.end() && it->second->version == version && it->second->n_rows == n &&
      it->second->shifted == (h->p.inter_loop_closure != 0)) {
    it->second->last_use = ++h->use_clock;
    return SLIDE_PR_OK;
  }
  for (int i = 0; i < n; i++)
    if (!std::isfinite(rows7[7 * (size_t)i + 1]) || !std::isfinite(rows7[7 * (size_t)i + 2])) { h->err = "non-finite map coordinate"; return SLIDE_PR_ERR_NONFINITE; }
  // uploads of the slot being replaced may still be in flight on the handle's stream
  SPR_CUDA(h, cudaStreamSynchronize(h->stream));
  if (it == h->cache.end()) {
    while (h->cache.size() >= h->cache_capacity) {   // drop the least recently used slot
      auto lru = h->cache.begin();
      for (auto k = h->cache.begin(); k != h->cache.end(); ++k)
        if (k->second->last_use < lru->second->last_use) lru = k;
      if (h->rs == lru->second.get()) { h->rs = &h->anon; h->prepared = false; }
      retire_slot(h, std::move(lru->second));
      h->cache.erase(lru);
    }
    std::unique_ptr<RefSide> fresh;
    if (!h->free_slots.empty()) { fresh = std::move(h->free_slots.back()); h->free_slots.pop_back(); }
    else fresh.reset(new RefSide());
    it = h->cache.emplace(robot_id, std::move(fresh)).first;
  }
  RefSide &E = *it->second;
  if (h->rs == &E) h->prepared = false;
  E.robot_id = robot_id; E.version = version; E.n_rows = n;
  E.ref_index_valid = false; E.ranks_pending = false;
  E.join_valid = false; E.ref7_uploaded = false; E.rows_valid = true;
  E.shifted = h->p.inter_loop_closure != 0;
  E.cached_ref.assign(rows7, rows7 + (size_t)n * 7);
  double c[2] = {0, 0}, b[2] = {0, 0};
  if (E.shifted) {                                                     // getCentroid / getMapBoundaries, PR.cpp:713-734, 752-765
    for (int i = 0; i < n; i++) { c[0] += rows7[7 * (size_t)i + 1]; c[1] += rows7[7 * (size_t)i + 2]; }
    c[0] /= (double)n; c[1] /= (double)n;
    for (int i = 0; i < n; i++) {
      double *r = E.cached_ref.data() + 7 * (size_t)i;
      r[1] -= c[0]; r[2] -= c[1];
      b[0] = std::max(b[0], std::abs(r[1])); b[1] = std::max(b[1], std::abs(r[2]));
    }
  }
  E.centroid[0] = c[0]; E.centroid[1] = c[1]; E.max_abs[0] = b[0]; E.max_abs[1] = b[1];
  E.last_use = ++h->use_clock;
  return SLIDE_PR_OK;
}

int slide_pr_map_cache_drop(slide_pr_handle *h, int64_t robot_id) {
  if (!h) return SLIDE_PR_ERR_INVALID;
  auto it = h->cache.find(robot_id);
  if (it == h->cache.end()) return SLIDE_PR_NOT_FOUND;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->rs == it->second.get()) { h->rs = &h->anon; h->prepared = false; }
  retire_slot(h, std::move(it->second));
  h->cache.erase(it);
  return SLIDE_PR_OK;
}

int32_t slide_pr_map_cache_size(const slide_pr_handle *h) { return h ? (int32_t)h->cache.size() : 0; }

// findTransformation (PR.cpp:736-945) on two cached maps; inter-robot mode (the cache holds centroid-shifted rows)
int slide_pr_find_transformation_cached(slide_pr_handle *h, int64_t ref_robot_id, int64_t qry_robot_id, int32_t *ref_idx_out,
                                        int32_t *qry_idx_out, slide_pr_tf_result *out) {
  if (!h || !out) return SLIDE_PR_ERR_INVALID;
  std::memset(out, 0, sizeof(*out));
  const slide_pr_params &p = h->p;
  if (!p.inter_loop_closure) { h->err = "the map cache serves the inter-robot search (inter_loop_closure = 1)"; return SLIDE_PR_ERR_INVALID; }
  auto ir = h->cache.find(ref_robot_id), iq = h->cache.find(qry_robot_id);
  if (ir == h->cache.end() || iq == h->cache.end()) { h->err = "robot id not in the map cache"; return SLIDE_PR_ERR_INVALID; }
  RefSide &Rf = *ir->second, &Qr = *iq->second;
  if (!Rf.shifted || !Qr.shifted) { h->err = "cached map was stored in intra-robot mode: put it again"; return SLIDE_PR_ERR_INVALID; }
  g_trace.start();
  Rf.last_use = ++h->use_clock; Qr.last_use = ++h->use_clock;
  double max_x = std::max(Rf.max_abs[0], Qr.max_abs[0]), max_y = std::max(Rf.max_abs[1], Qr.max_abs[1]);   // PR.cpp:771-774
  if (!p.disable_yaw_search) { const double m = std::max(max_x, max_y); max_x = m; max_y = m; }           // :777-782
  out->half_x = max_x * p.dilation_factor; out->half_y = max_y * p.dilation_factor;                       // :786-787
  out->yaw_half = p.match_yaw_half_range;
  out->centroid_ref[0] = Rf.centroid[0]; out->centroid_ref[1] = Rf.centroid[1];
  out->centroid_qry[0] = Qr.centroid[0]; out->centroid_qry[1] = Qr.centroid[1];
  std::vector<int32_t> ri_own, qi_own;
  int32_t *ri = ref_idx_out, *qi = qry_idx_out;
  if (!ri) { ri_own.resize(std::max(Qr.n_rows, 1)); ri = ri_own.data(); }
  if (!qi) { qi_own.resize(std::max(Qr.n_rows, 1)); qi = qi_own.data(); }
  const int rc = match_maps_impl(h, &Rf, true, Rf.cached_ref.data(), Rf.n_rows, Qr.cached_ref.data(), Qr.n_rows, out->half_x, out->half_y,
                                 ri, qi, &out->match);
  g_trace.flush("find_transformation_cached");
  if (rc != SLIDE_PR_OK) return rc;
  return finish_transformation(p, Rf.cached_ref.data(), Qr.cached_ref.data(), out->centroid_ref, out->centroid_qry, ri, qi, out);
}

// n_pairs findTransformation calls over n_maps maps (all-pairs multi-robot matching, BASELINE config 4): every
// map goes through the cache once, so a map's reference-side index is built once however many pairs use it.
int slide_pr_find_transformation_batch(slide_pr_handle *h, const double *const *maps, const int32_t *map_sizes,
                                       int32_t n_maps, const int32_t *ref_of, const int32_t *qry_of, int32_t n_pairs,
                                       slide_pr_tf_result *out) {
  if (!h || !maps || !map_sizes || !ref_of || !qry_of || !out || n_maps < 0 || n_pairs < 0) return SLIDE_PR_ERR_INVALID;
  for (int p = 0; p < n_pairs; p++)
    if (ref_of[p] < 0 || ref_of[p] >= n_maps || qry_of[p] < 0 || qry_of[p] >= n_maps) { h->err = "pair index out of range"; return SLIDE_PR_ERR_INVALID; }
  if (!h->p.inter_loop_closure) {   // intra-robot mode keeps raw frames: no shared index, plain calls
    for (int p = 0; p < n_pairs; p++) {
      const int rc = slide_pr_find_transformation(h, maps[ref_of[p]], map_sizes[ref_of[p]], maps[qry_of[p]], map_sizes[qry_of[p]], nullptr, nullptr, &out[p]);
      if (rc < 0) return rc;
    }
    return SLIDE_PR_OK;
  }
  const int64_t id0 = INT64_MIN / 2;   // private id range of this call
  const size_t keep = h->cache_capacity;
  h->cache_capacity = std::max<size_t>(keep, h->cache.size() + (size_t)n_maps);
  int rc = SLIDE_PR_OK;
  for (int m = 0; m < n_maps && rc >= 0; m++) rc = slide_pr_map_cache_put(h, id0 + m, 1, maps[m], map_sizes[m]);
  // pairs in the caller's order; consecutive pairs with the same reference map reuse its index at once,
  // later ones find it in the cache
  for (int p = 0; p < n_pairs && rc >= 0; p++) {
    rc = slide_pr_find_transformation_cached(h, id0 + ref_of[p], id0 + qry_of[p], nullptr, nullptr, &out[p]);
    if (rc > 0) rc = SLIDE_PR_OK;  // "not found" is a per-pair result
  }
  for (int m = 0; m < n_maps; m++) slide_pr_map_cache_drop(h, id0 + m);
  h->cache_capacity = keep;
  return rc < 0 ? rc : SLIDE_PR_OK;
}

void slide_pr_pack_record(const slide_pr_match_result *r, int32_t rank, slide_pr_topk_record *rec) {
  rec->hyp_index = r->best_hyp_index;
  rec->inliers = r->best_hyp_index >= 0 ? r->best_num_inliers : -10000;
  rec->rank = rank;
}

int slide_pr_merge_records(const slide_pr_topk_record *recs, int32_t n) {
  int best = -1;
  for (int i = 0; i < n; i++) {
    if (recs[i].hyp_index < 0) continue;
    if (best < 0 || recs[i].inliers > recs[best].inliers ||
        (recs[i].inliers == recs[best].inliers && recs[i].hyp_index < recs[best].hyp_index))
      best = i;
  }
  return best;
}

// Device-side generator front (spr_generate.cu): descriptors, binning of the data triangles by their
// first descriptor component, windowed matching, radix sort back into the reference's order.  Leaves on
// the device (h->d_tri arena): both triangle lists, descriptors, vertex permutations, class signatures and
// the sorted match list as (model_idx, data_idx) arrays.
struct GenDevice {
  const double *tris_m, *tris_d;
  const int32_t *perm_m, *perm_d;
  int32_t *model_idx, *data_idx;
  long long n_matches;
  float match_ms;
};

static int gen_match_on_device(slide_pr_handle *h, const double *tris_model6, const double *labels_model3, int32_t t_model,
                               const double *tris_data6, const double *labels_data3, int32_t t_data, double threshold,
                               GenDevice *G) {
  std::memset(G, 0, sizeof(*G));
  const bool labeled = labels_model3 != nullptr;
  cudaStream_t st = h->stream;
  const size_t tm = (size_t)t_model, td = (size_t)t_data;
  // bins: width >= the matching window, at most 2^20 of them over the data descriptors' possible range
  // (a vertex is never farther from its triangle's centroid than the bounding box' diagonal)
  double lo[2] = {HUGE_VAL, HUGE_VAL}, hi[2] = {-HUGE_VAL, -HUGE_VAL};
  for (size_t k = 0; k < td * 3; k++)
    for (int c = 0; c < 2; c++) {
      const double v = tris_data6[2 * k + c];
      if (v < lo[c]) lo[c] = v;
      if (v > hi[c]) hi[c] = v;
    }
  double diag = std::hypot(hi[0] - lo[0], hi[1] - lo[1]);
  if (!std::isfinite(diag)) diag = 0.0;  // non-finite coordinates give NaN descriptors, which never match
  const bool can_match = threshold > 0 && threshold == threshold;
  const double window = can_match && std::isfinite(threshold) ? threshold * (1.0 + 1e-9) + 1e-300 : HUGE_VAL;
  const double Tstar = std::isinf(threshold) && threshold > 0 ? HUGE_VAL : spr::sqrt_threshold(threshold);
  double width = std::isfinite(window) ? window : diag + 1.0;
  if (diag / width > 1048000.0) width = diag / 1048000.0;
  if (!(width > 0)) width = 1.0;
  const uint32_t n_bins = (uint32_t)std::min(1048576.0, std::floor(diag / width) + 2.0);
  const double inv_w = 1.0 / width;
  // arena
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_tm = take(tm * 6 * 8), o_td = take(td * 6 * 8), o_lm = take(tm * 3 * 8), o_ld = take(td * 3 * 8);
  const size_t o_dm = take(tm * 3 * 8), o_dd = take(td * 3 * 8), o_sm = take(tm * 3 * 8), o_sd = take(td * 3 * 8);
  const size_t o_pm = take(tm * 3 * 4), o_pd = take(td * 3 * 4);
  const size_t o_bs = take(((size_t)n_bins + 2) * 4), o_bf = take(((size_t)n_bins + 1) * 4), o_bn = take(spr_gen_binned_bytes(t_data));
  const size_t o_tot = take(8);
  SPR_CUDA(h, h->d_tri.ensure(off));
  char *base = h->d_tri.as<char>();
  SPR_CUDA(h, cudaMemcpyAsync(base + o_tm, tris_model6, tm * 6 * 8, cudaMemcpyHostToDevice, st));
  SPR_CUDA(h, cudaMemcpyAsync(base + o_td, tris_data6, td * 6 * 8, cudaMemcpyHostToDevice, st));
  if (labeled) {
    SPR_CUDA(h, cudaMemcpyAsync(base + o_lm, labels_model3, tm * 3 * 8, cudaMemcpyHostToDevice, st));
    SPR_CUDA(h, cudaMemcpyAsync(base + o_ld, labels_data3, td * 3 * 8, cudaMemcpyHostToDevice, st));
  }
  double *dm = (double *)(base + o_dm), *dd = (double *)(base + o_dd);
  double *sm = labeled ? (double *)(base + o_sm) : nullptr, *sd = labeled ? (double *)(base + o_sd) : nullptr;
  int32_t *pm = (int32_t *)(base + o_pm), *pd = (int32_t *)(base + o_pd);
  unsigned long long *tot = (unsigned long long *)(base + o_tot);
  SPR_CUDA(h, cudaEventRecord(h->ev0, st));
  SPR_CUDA(h, spr_launch_tri_desc((const double *)(base + o_tm), labeled ? (const double *)(base + o_lm) : nullptr, t_model, dm, pm, sm, st));
  SPR_CUDA(h, spr_launch_tri_desc((const double *)(base + o_td), labeled ? (const double *)(base + o_ld) : nullptr, t_data, dd, pd, sd, st));
  SPR_CUDA(h, spr_launch_gen_bin(dd, t_data, inv_w, n_bins, (uint32_t *)(base + o_bs), (uint32_t *)(base + o_bf), base + o_bn, st));
  // match keys: capacity grows on overflow (one retry with the exact count)
  unsigned long long cap = std::max<unsigned long long>(h->gen_cap, 1ull << 16);
  unsigned long long total = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    SPR_CUDA(h, h->d_tri_out.ensure((size_t)cap * 8 * 3 + spr_radix_sort_hist_words((long long)cap) * 4 + 1024));
    SPR_CUDA(h, cudaMemsetAsync(tot, 0, 8, st));
    SPR_CUDA(h, spr_launch_gen_match(dm, t_model, base + o_bn, (const uint32_t *)(base + o_bs), n_bins, inv_w, window, Tstar, sm, sd,
                                     t_data, h->d_tri_out.as<unsigned long long>(), cap, tot, h->sm_count, st));
    SPR_CUDA(h, cudaMemcpyAsync(&total, tot, 8, cudaMemcpyDeviceToHost, st));
    SPR_CUDA(h, cudaStreamSynchronize(st));
    if (total <= cap) break;
    cap = total + total / 8 + 1024;
  }
  h->gen_cap = cap;
  unsigned long long *keys = h->d_tri_out.as<unsigned long long>(), *tmp = keys + cap;
  int32_t *mi = (int32_t *)(tmp + cap), *di = mi + cap;
  uint32_t *hist = (uint32_t *)(di + cap);
  int bits = 1;
  while (bits < 64 && ((unsigned long long)t_model * (unsigned long long)std::max(t_data, 1)) >> bits) bits++;
  SPR_CUDA(h, spr_radix_sort_u64(keys, tmp, (long long)total, bits, hist, st));
  SPR_CUDA(h, spr_launch_gen_unpack(keys, (long long)total, t_data, mi, di, st));
  SPR_CUDA(h, cudaEventRecord(h->ev1, st));
  G->tris_m = (const double *)(base + o_tm); G->tris_d = (const double *)(base + o_td);
  G->perm_m = pm; G->perm_d = pd;
  G->model_idx = mi; G->data_idx = di;
  G->n_matches = (long long)total;
  return SLIDE_PR_OK;
}

int slide_pr_match_triangles_labeled(slide_pr_handle *h, const double *tris_model6, const double *labels_model3,
                                     int32_t t_model, const double *tris_data6, const double *labels_data3,
                                     int32_t t_data, double threshold, int32_t *model_idx_out, int32_t *data_idx_out,
                                     int32_t *perm_model_out, int32_t *perm_data_out, int64_t cap, int64_t *n_matches) {
  if (!h || !n_matches || t_model < 0 || t_data < 0 || cap < 0) return SLIDE_PR_ERR_INVALID;
  if ((t_model > 0 && !tris_model6) || (t_data > 0 && !tris_data6)) { h->err = "null triangle array"; return SLIDE_PR_ERR_INVALID; }
  if ((labels_model3 == nullptr) != (labels_data3 == nullptr)) { h->err = "labels must be given for both maps or for none"; return SLIDE_PR_ERR_INVALID; }
  *n_matches = 0;
  if (t_model == 0 || t_data == 0) return SLIDE_PR_OK;
  SPR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  GenDevice G;
  const int rc = gen_match_on_device(h, tris_model6, labels_model3, t_model, tris_data6, labels_data3, t_data, threshold, &G);
  if (rc != SLIDE_PR_OK) return rc;
  *n_matches = (int64_t)G.n_matches;
  const int64_t n_out = std::min<int64_t>((int64_t)G.n_matches, cap);
  if (n_out > 0 && model_idx_out && data_idx_out) {
    const size_t tm = (size_t)t_model, td = (size_t)t_data;
    SPR_CUDA(h, cudaMemcpyAsync(model_idx_out, G.model_idx, (size_t)n_out * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SPR_CUDA(h, cudaMemcpyAsync(data_idx_out, G.data_idx, (size_t)n_out * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    std::vector<int32_t> hpm, hpd;
    if (perm_model_out || perm_data_out) {
      hpm.resize(tm * 3); hpd.resize(td * 3);
      SPR_CUDA(h, cudaMemcpyAsync(hpm.data(), G.perm_m, tm * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      SPR_CUDA(h, cudaMemcpyAsync(hpd.data(), G.perm_d, td * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    SPR_CUDA(h, cudaStreamSynchronize(st));
    for (int64_t k = 0; k < n_out; k++)
      for (int v = 0; v < 3; v++) {  // vertices in sorted-distance order, SC.cpp:102-105
        if (perm_model_out) perm_model_out[3 * k + v] = hpm[3 * (size_t)model_idx_out[k] + v];
        if (perm_data_out) perm_data_out[3 * k + v] = hpd[3 * (size_t)data_idx_out[k] + v];
      }
  }
  return SLIDE_PR_OK;
}

// The whole generator half on the device: matched triangles -> one 2-D Kabsch hypothesis per match ->
// MatchMaps predicate on every hypothesis (spr_score_list_kernel) -> best hypothesis.  Needs a
// slide_pr_prepare'd map pair (the maps the triangles were built from).
int slide_pr_generate_and_score(slide_pr_handle *h, const double *tris_model6, const double *labels_model3, int32_t t_model,
                                const double *tris_data6, const double *labels_data3, int32_t t_data, double threshold,
                                slide_pr_match_result *out, slide_pr_generate_info *info, int32_t *model_idx_out,
                                int32_t *data_idx_out, double *hyps4_out, int32_t *counts_out, int64_t cap) {
  if (!h || !out || t_model < 0 || t_data < 0 || cap < 0) return SLIDE_PR_ERR_INVALID;
  if (!h->prepared) { h->err = "slide_pr_generate_and_score before slide_pr_prepare"; return SLIDE_PR_ERR_INVALID; }
  if ((t_model > 0 && !tris_model6) || (t_data > 0 && !tris_data6)) { h->err = "null triangle array"; return SLIDE_PR_ERR_INVALID; }
  if ((labels_model3 == nullptr) != (labels_data3 == nullptr)) { h->err = "labels must be given for both maps or for none"; return SLIDE_PR_ERR_INVALID; }
  SPR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  fill_result_header(h, out);
  out->status = SLIDE_PR_OK;
  slide_pr_generate_info local{}, *I = info ? info : &local;
  std::memset(I, 0, sizeof(*I));
  if (t_model == 0 || t_data == 0) return SLIDE_PR_OK;
  bool list_join = false;
  { const int lrc = ensure_list_scorer(h, st, &list_join); if (lrc != SLIDE_PR_OK) return lrc; }
  GenDevice G;
  int rc = gen_match_on_device(h, tris_model6, labels_model3, t_model, tris_data6, labels_data3, t_data, threshold, &G);
  if (rc != SLIDE_PR_OK) return rc;
  const long long n = G.n_matches;
  I->n_matches = n;
  I->n_triangles_model = t_model; I->n_triangles_data = t_data;
  if (n == 0) { SPR_CUDA(h, cudaStreamSynchronize(st)); SPR_CUDA(h, cudaEventElapsedTime(&I->match_ms, h->ev0, h->ev1)); return SLIDE_PR_OK; }
  SPR_CUDA(h, h->d_hyps.ensure((size_t)n * 4 * sizeof(double)));
  SPR_CUDA(h, h->d_counts.ensure((size_t)n * sizeof(int32_t)));
  SPR_CUDA(h, cudaMemsetAsync(h->d_best.p, 0, sizeof(unsigned long long), st));
  cudaEvent_t e2, e3;
  SPR_CUDA(h, cudaEventCreate(&e2)); SPR_CUDA(h, cudaEventCreate(&e3));
  SPR_CUDA(h, spr_launch_gen_kabsch(G.tris_m, G.tris_d, G.perm_m, G.perm_d, G.model_idx, G.data_idx, n, h->d_hyps.as<double>(),
                                    nullptr, nullptr, st));
  SPR_CUDA(h, cudaEventRecord(e2, st));
  if (list_join) {
    SPR_CUDA(h, spr_launch_join_score_list(h->JV, h->d_hyps.as<double>(), n, h->d_counts.as<int32_t>(), h->d_best.as<unsigned long long>(),
                                           h->sm_count, st));
  } else {
    SPR_CUDA(h, spr_launch_score_list(h->V, h->d_hyps.as<double>(), n, h->d_counts.as<int32_t>(), h->d_best.as<unsigned long long>(),
                                      h->sm_count, st));
  }
  SPR_CUDA(h, cudaEventRecord(e3, st));
  if (h->h_scalars.size() < 8) h->h_scalars.assign(8, 0ull);
  SPR_CUDA(h, cudaMemcpyAsync(h->h_scalars.data(), h->d_best.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  const int64_t n_out = std::min<int64_t>(n, cap);
  if (n_out > 0) {
    if (model_idx_out) SPR_CUDA(h, cudaMemcpyAsync(model_idx_out, G.model_idx, (size_t)n_out * 4, cudaMemcpyDeviceToHost, st));
    if (data_idx_out) SPR_CUDA(h, cudaMemcpyAsync(data_idx_out, G.data_idx, (size_t)n_out * 4, cudaMemcpyDeviceToHost, st));
    if (hyps4_out) SPR_CUDA(h, cudaMemcpyAsync(hyps4_out, h->d_hyps.p, (size_t)n_out * 32, cudaMemcpyDeviceToHost, st));
    if (counts_out) SPR_CUDA(h, cudaMemcpyAsync(counts_out, h->d_counts.p, (size_t)n_out * 4, cudaMemcpyDeviceToHost, st));
  }
  SPR_CUDA(h, cudaStreamSynchronize(st));
  SPR_CUDA(h, cudaEventElapsedTime(&I->match_ms, h->ev0, h->ev1));
  SPR_CUDA(h, cudaEventElapsedTime(&I->kabsch_ms, h->ev1, e2));
  SPR_CUDA(h, cudaEventElapsedTime(&I->score_ms, e2, e3));
  cudaEventDestroy(e2); cudaEventDestroy(e3);
  const unsigned long long key = h->h_scalars[0];
  out->kernel_ms = I->match_ms + I->kabsch_ms + I->score_ms;
  out->gpu_launches = 2 + 3 + 1 + 3 * ((64 + 7) / 8) + 3;  // upper bound; the sort passes depend on the key width
  out->hypotheses_scored = n;
  if (key) {
    out->best_num_inliers = spr_key_count(key);
    out->best_hyp_index = spr_key_index(key);
    double hp[4];
    SPR_CUDA(h, cudaMemcpy(hp, h->d_hyps.as<double>() + 4 * (size_t)out->best_hyp_index, sizeof(hp), cudaMemcpyDeviceToHost));
    out->R_t[0] = hp[0]; out->R_t[1] = -hp[1]; out->R_t[2] = hp[2];
    out->R_t[3] = hp[1]; out->R_t[4] = hp[0];  out->R_t[5] = hp[3];
  }
  return SLIDE_PR_OK;
}

int slide_pr_match_triangles(slide_pr_handle *h, const double *tris_model6, int32_t t_model, const double *tris_data6,
                             int32_t t_data, double threshold, int32_t *model_idx_out, int32_t *data_idx_out,
                             int32_t *perm_model_out, int32_t *perm_data_out, int64_t cap, int64_t *n_matches) {
  return slide_pr_match_triangles_labeled(h, tris_model6, nullptr, t_model, tris_data6, nullptr, t_data, threshold,
                                          model_idx_out, data_idx_out, perm_model_out, perm_data_out, cap, n_matches);
}

int slide_pr_estimate_tf(const double *pts_a2, const double *pts_b2, int32_t k, double *tf9) {
  if (!pts_a2 || !pts_b2 || !tf9 || k <= 0) return SLIDE_PR_ERR_INVALID;
  spr::estimate_tf(pts_a2, pts_b2, k, tf9);
  return SLIDE_PR_OK;
}

int slide_pr_triangle_hypotheses(const double *tris_model6, const double *tris_data6, const int32_t *model_idx,
                                 const int32_t *data_idx, const int32_t *perm_model, const int32_t *perm_data, int64_t n,
                                 double *hyps4_out) {
  if (n < 0 || (n > 0 && (!tris_model6 || !tris_data6 || !model_idx || !data_idx || !perm_model || !perm_data || !hyps4_out)))
    return SLIDE_PR_ERR_INVALID;
  for (int64_t k = 0; k < n; k++) {
    const double *tm = tris_model6 + 6 * (size_t)model_idx[k], *td = tris_data6 + 6 * (size_t)data_idx[k];
    double a[6], b[6], tf[9];
    for (int v = 0; v < 3; v++) {  // data (query) vertex v  ->  model (reference) vertex v, sorted order
      a[2 * v] = td[2 * perm_data[3 * k + v]]; a[2 * v + 1] = td[2 * perm_data[3 * k + v] + 1];
      b[2 * v] = tm[2 * perm_model[3 * k + v]]; b[2 * v + 1] = tm[2 * perm_model[3 * k + v] + 1];
    }
    spr::estimate_tf(a, b, 3, tf);
    hyps4_out[4 * k] = tf[0]; hyps4_out[4 * k + 1] = tf[3]; hyps4_out[4 * k + 2] = tf[2]; hyps4_out[4 * k + 3] = tf[5];
  }
  return SLIDE_PR_OK;
}

int slide_pr_score_hypotheses(slide_pr_handle *h, const double *hyps4, int64_t n, int32_t *counts_out,
                              slide_pr_match_result *out) {
  if (!h || !out || (n > 0 && !hyps4) || n < 0) return SLIDE_PR_ERR_INVALID;
  if (!h->prepared) { h->err = "slide_pr_score_hypotheses before slide_pr_prepare"; return SLIDE_PR_ERR_INVALID; }
  SPR_CUDA(h, cudaSetDevice(h->device));
  bool list_join = false;
  { const int lrc = ensure_list_scorer(h, h->stream, &list_join); if (lrc != SLIDE_PR_OK) return lrc; }
  cudaStream_t st = h->stream;
  fill_result_header(h, out);
  out->status = SLIDE_PR_OK;
  if (n == 0) return SLIDE_PR_OK;
  int rc;
  if ((rc = upload_raw(h, h->d_hyps, hyps4, (size_t)n * 4 * sizeof(double), st))) return rc;
  if (counts_out) SPR_CUDA(h, h->d_counts.ensure((size_t)n * sizeof(int32_t)));
  SPR_CUDA(h, cudaMemsetAsync(h->d_best.p, 0, sizeof(unsigned long long), st));
  SPR_CUDA(h, cudaEventRecord(h->ev0, st));
  if (list_join) {
    SPR_CUDA(h, spr_launch_join_score_list(h->JV, h->d_hyps.as<double>(), n, counts_out ? h->d_counts.as<int32_t>() : nullptr,
                                           h->d_best.as<unsigned long long>(), h->sm_count, st));
  } else {
    SPR_CUDA(h, spr_launch_score_list(h->V, h->d_hyps.as<double>(), n, counts_out ? h->d_counts.as<int32_t>() : nullptr,
                                      h->d_best.as<unsigned long long>(), h->sm_count, st));
  }
  SPR_CUDA(h, cudaEventRecord(h->ev1, st));
  unsigned long long key = 0;
  SPR_CUDA(h, cudaMemcpyAsync(&key, h->d_best.p, sizeof(key), cudaMemcpyDeviceToHost, st));
  if (counts_out) SPR_CUDA(h, cudaMemcpyAsync(counts_out, h->d_counts.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SPR_CUDA(h, cudaStreamSynchronize(st));
  float ms = 0;
  SPR_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  out->kernel_ms = ms;
  out->gpu_launches = 1;
  out->hypotheses_scored = n;
  if (key) {
    out->best_num_inliers = spr_key_count(key);
    out->best_hyp_index = spr_key_index(key);
    const double *hp = hyps4 + 4 * (size_t)out->best_hyp_index;
    out->R_t[0] = hp[0]; out->R_t[1] = -hp[1]; out->R_t[2] = hp[2];
    out->R_t[3] = hp[1]; out->R_t[4] = hp[0];  out->R_t[5] = hp[3];
  }
  return SLIDE_PR_OK;
}

// ---- SlideGraph: CLIPPER ------------------------------------------------------------------------
void slide_clipper_default_params(slide_clipper_params *p) {  // clipper.h:28-60, euclidean_distance.h:27-30
  std::memset(p, 0, sizeof(*p));
  p->sigma = 0.01; p->epsilon = 0.06; p->mindist = 0;
  p->tol_u = 1e-8; p->tol_F = 1e-9; p->tol_Fop = 1e-10;
  p->maxiniters = 200; p->maxoliters = 1000;
  p->beta = 0.25; p->maxlsiters = 99;
  p->eps = 1e-9; p->affinityeps = 1e-4;
  p->rescale_u0 = 1;
  p->rounding = SLIDE_CLIPPER_ROUND_DSD_HEU;
}

static SprClipper *clipper_of(slide_pr_handle *h) {
  if (!h->clipper) h->clipper = spr_clipper_create();
  return h->clipper;
}

int slide_pr_clipper_score_pairwise_consistency(slide_pr_handle *h, const slide_clipper_params *p, const double *D1, int32_t n1,
                                                const double *D2, int32_t n2, int32_t dim, const int32_t *A, int32_t m,
                                                int64_t *nnz_upper) {
  if (!h || !p || (n1 > 0 && !D1) || (n2 > 0 && !D2)) return SLIDE_PR_ERR_INVALID;
  SPR_CUDA(h, cudaSetDevice(h->device));
  long long nnz = 0;
  const int rc = spr_clipper_score(clipper_of(h), *p, D1, n1, D2, n2, dim, A, m, nullptr, h->sm_count, h->stream, &nnz, nullptr, h->err);
  if (nnz_upper) *nnz_upper = nnz;
  return rc;
}

int32_t slide_pr_clipper_get_initial_associations(slide_pr_handle *h, int32_t *A_out, int32_t cap) {
  if (!h || !h->clipper) return 0;
  int m = 0;
  spr_clipper_size(h->clipper, &m, nullptr);
  if (A_out && cap > 0) std::memcpy(A_out, spr_clipper_associations(h->clipper), sizeof(int32_t) * 2 * (size_t)std::min(m, cap));
  return m;
}

int slide_pr_clipper_get_affinity_csr(slide_pr_handle *h, int64_t *row_ptr, int32_t *col, double *val, int64_t cap) {
  if (!h || !row_ptr) return SLIDE_PR_ERR_INVALID;
  if (!h->clipper) { h->err = "no pairwise consistency has been scored on this handle"; return SLIDE_PR_ERR_INVALID; }
  SPR_CUDA(h, cudaSetDevice(h->device));
  return spr_clipper_get_csr(h->clipper, row_ptr, col, val, cap, h->stream, h->err);
}

int slide_pr_clipper_get_affinity_matrix(slide_pr_handle *h, double *M_out, int64_t cap) {
  if (!h || !M_out) return SLIDE_PR_ERR_INVALID;
  if (!h->clipper) { h->err = "no pairwise consistency has been scored on this handle"; return SLIDE_PR_ERR_INVALID; }
  SPR_CUDA(h, cudaSetDevice(h->device));
  int m = 0;
  long long nnz = 0;
  spr_clipper_size(h->clipper, &m, &nnz);
  if ((int64_t)m * m > cap) { h->err = "dense affinity matrix does not fit the given capacity"; return SLIDE_PR_ERR_INVALID; }
  std::vector<int64_t> rp((size_t)m + 1);
  std::vector<int32_t> col((size_t)std::max<long long>(nnz, 1));
  std::vector<double> val((size_t)std::max<long long>(nnz, 1));
  const int rc = spr_clipper_get_csr(h->clipper, rp.data(), col.data(), val.data(), nnz, h->stream, h->err);
  if (rc != SLIDE_PR_OK) return rc;
  std::fill(M_out, M_out + (size_t)m * m, 0.0);
  for (int i = 0; i < m; i++) {
    M_out[(size_t)i * m + i] = 1.0;  // + Identity (clipper.cpp:124)
    for (int64_t k = rp[i]; k < rp[i + 1]; k++) M_out[(size_t)i * m + col[k]] = val[k];
  }
  return SLIDE_PR_OK;
}

int slide_pr_clipper_solve(slide_pr_handle *h, const slide_clipper_params *p, const double *u0, uint64_t seed, int32_t *nodes_out,
                           int32_t cap, slide_clipper_solution *sol, double *u_out) {
  if (!h || !p || !sol) return SLIDE_PR_ERR_INVALID;
  if (!h->clipper) { h->err = "no pairwise consistency has been scored on this handle"; return SLIDE_PR_ERR_INVALID; }
  SPR_CUDA(h, cudaSetDevice(h->device));
  int m = 0;
  spr_clipper_size(h->clipper, &m, nullptr);
  std::vector<double> own;
  if (!u0) {  // stand-in for utils::randvec (utils.cpp:22-29): U[0, 1) from a splitmix64 stream
    own.resize((size_t)std::max(m, 1));
    uint64_t x = seed;
    for (int i = 0; i < m; i++) {
      x += 0x9e3779b97f4a7c15ull;
      uint64_t z = x;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      z ^= z >> 31;
      own[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
    u0 = own.data();
  }
  return spr_clipper_solve(h->clipper, *p, u0, h->sm_count, h->stream, nodes_out, cap, sol, u_out, h->err);
}

int slide_pr_measure_issue_peaks(slide_pr_handle *h, double *alu_winst_per_s, double *fma_winst_per_s, double *mixed_winst_per_s) {
  if (!h || !alu_winst_per_s || !fma_winst_per_s || !mixed_winst_per_s) return SLIDE_PR_ERR_INVALID;
  SPR_CUDA(h, cudaSetDevice(h->device));
  SPR_CUDA(h, spr_measure_issue_peaks(h->sm_count, alu_winst_per_s, fma_winst_per_s, mixed_winst_per_s, h->stream));
  return SLIDE_PR_OK;
}

// ---- SlideGraph entry points -----------------------------------------------------------------------
int slide_pr_delaunay(const double *xy2, int32_t n, int32_t *tri3_out, int64_t cap, int64_t *n_tri) {
  if (n < 0 || (n > 0 && !xy2) || !n_tri || cap < 0) return SLIDE_PR_ERR_INVALID;
  std::vector<int32_t> tri;
  const int k = spr::delaunay_triangulate(xy2, n, tri);
  if (k < 0) return SLIDE_PR_ERR_NONFINITE;
  *n_tri = k;
  if (tri3_out) std::memcpy(tri3_out, tri.data(), sizeof(int32_t) * 3 * (size_t)std::min<int64_t>(k, cap));
  return SLIDE_PR_OK;
}

void slide_pr_slidegraph_default_params(slide_pr_slidegraph_params *p) {  // PR.cpp:64-75
  std::memset(p, 0, sizeof(*p));
  p->sigma = 0.1; p->epsilon = 0.1; p->matching_threshold = 0.1;
  p->num_inliers_threshold = 10; p->min_num_map_objects_to_start = 20;
}

static void triangles_of_map(const double *rows7, int n, std::vector<double> &tris6, std::vector<double> &labels3, bool want_labels) {
  std::vector<double> xy(2 * (size_t)std::max(n, 1));
  for (int i = 0; i < n; i++) { xy[2 * (size_t)i] = rows7[7 * (size_t)i + 1]; xy[2 * (size_t)i + 1] = rows7[7 * (size_t)i + 2]; }  // SC.cpp:150-166
  std::vector<int32_t> tri;
  const int k = spr::delaunay_triangulate(xy.data(), n, tri);
  tris6.resize(6 * (size_t)std::max(k, 0));
  labels3.resize(want_labels ? 3 * (size_t)std::max(k, 0) : 0);
  for (int t = 0; t < k; t++)
    for (int v = 0; v < 3; v++) {
      const int id = tri[3 * (size_t)t + v];
      tris6[6 * (size_t)t + 2 * v] = xy[2 * (size_t)id];
      tris6[6 * (size_t)t + 2 * v + 1] = xy[2 * (size_t)id + 1];
      if (want_labels) labels3[3 * (size_t)t + v] = rows7[7 * (size_t)id];
    }
}

int slide_pr_run_semantic_clipper(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7, int32_t n_qry,
                                  const slide_pr_slidegraph_params *sp, const double *tris_model6, int32_t t_model,
                                  const double *tris_data6, int32_t t_data, const double *u0, int64_t u0_len, double *tf16,
                                  slide_pr_sc_info *info_opt) {
  if (!h || !sp || !tf16 || n_ref < 0 || n_qry < 0 || (n_ref > 0 && !ref7) || (n_qry > 0 && !qry7)) return SLIDE_PR_ERR_INVALID;
  slide_pr_sc_info local, *I = info_opt ? info_opt : &local;
  std::memset(I, 0, sizeof(*I));
  SPR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // triangles: the caller's (e.g. qhull facets in the reference's order) or our own Delaunay triangulation
  const bool own_tris = !tris_model6 && !tris_data6;
  if ((tris_model6 == nullptr) != (tris_data6 == nullptr)) { h->err = "give the triangles of both maps or of none"; return SLIDE_PR_ERR_INVALID; }
  std::vector<double> tm_own, td_own, lm, ld;
  const double t0 = now_ms();
  if (own_tris) {
    for (int i = 0; i < n_ref; i++) if (!std::isfinite(ref7[7 * (size_t)i + 1]) || !std::isfinite(ref7[7 * (size_t)i + 2])) { h->err = "non-finite reference coordinate"; return SLIDE_PR_ERR_NONFINITE; }
    for (int i = 0; i < n_qry; i++) if (!std::isfinite(qry7[7 * (size_t)i + 1]) || !std::isfinite(qry7[7 * (size_t)i + 2])) { h->err = "non-finite query coordinate"; return SLIDE_PR_ERR_NONFINITE; }
    triangles_of_map(ref7, n_ref, tm_own, lm, sp->use_class_signature != 0);
    triangles_of_map(qry7, n_qry, td_own, ld, sp->use_class_signature != 0);
    tris_model6 = tm_own.data(); t_model = (int32_t)(tm_own.size() / 6);
    tris_data6 = td_own.data(); t_data = (int32_t)(td_own.size() / 6);
  } else if (sp->use_class_signature) {
    h->err = "the class signature needs the internal triangulation (vertex labels are not part of caller-supplied triangles)";
    return SLIDE_PR_ERR_UNSUPPORTED;
  }
  I->delaunay_ms = (float)(now_ms() - t0);
  I->n_triangles_model = t_model; I->n_triangles_data = t_data;
  if (t_model <= 0 || t_data <= 0) return SLIDE_PR_NOT_FOUND;
  // match_triangles (SC.cpp:111-118) on the device, reference order
  GenDevice G;
  int rc = gen_match_on_device(h, tris_model6, lm.empty() ? nullptr : lm.data(), t_model, tris_data6, ld.empty() ? nullptr : ld.data(),
                               t_data, sp->matching_threshold, &G);
  if (rc != SLIDE_PR_OK) return rc;
  I->n_triangle_matches = G.n_matches;
  const long long m = 3 * G.n_matches;     // 3 matched points per triangle pair (SC.cpp:102-105)
  I->n_associations = m;
  if (m == 0) { SPR_CUDA(h, cudaStreamSynchronize(st)); return SLIDE_PR_NOT_FOUND; }
  if (m > 0x3fffffffLL) { h->err = "more than 2^30 putative associations"; return SLIDE_PR_ERR_UNSUPPORTED; }
  // matched point lists (model / data, m x 2 each) straight from the match list, on the device
  SPR_CUDA(h, h->d_hyps.ensure((size_t)m * 2 * 2 * sizeof(double)));
  double *pm = h->d_hyps.as<double>(), *pd = pm + 2 * (size_t)m;
  SPR_CUDA(h, spr_launch_gen_kabsch(G.tris_m, G.tris_d, G.perm_m, G.perm_d, G.model_idx, G.data_idx, G.n_matches, nullptr, pm, pd, st));
  SPR_CUDA(h, cudaStreamSynchronize(st));
  SPR_CUDA(h, cudaEventElapsedTime(&I->match_ms, h->ev0, h->ev1));
  // CLIPPER: associations (i, i) (SC.cpp:203-207), EuclideanDistance with sigma / epsilon (SC.cpp:210-212)
  slide_clipper_params cp;
  slide_clipper_default_params(&cp);
  cp.sigma = sp->sigma; cp.epsilon = sp->epsilon;
  std::vector<int32_t> A(2 * (size_t)m);
  for (long long i = 0; i < m; i++) { A[2 * (size_t)i] = (int32_t)i; A[2 * (size_t)i + 1] = (int32_t)i; }
  // centre + extent of each cloud for the fp32 prefilter: the triangles' bounding boxes
  double hint[7] = {0, 0, 0, 0, 0, 0, 0};
  {
    double lo[2][2] = {{HUGE_VAL, HUGE_VAL}, {HUGE_VAL, HUGE_VAL}}, hi[2][2] = {{-HUGE_VAL, -HUGE_VAL}, {-HUGE_VAL, -HUGE_VAL}};
    const double *src[2] = {tris_model6, tris_data6};
    const size_t cnt[2] = {3 * (size_t)t_model, 3 * (size_t)t_data};
    for (int s = 0; s < 2; s++)
      for (size_t k = 0; k < cnt[s]; k++)
        for (int c = 0; c < 2; c++) { const double v = src[s][2 * k + c]; lo[s][c] = std::min(lo[s][c], v); hi[s][c] = std::max(hi[s][c], v); }
    for (int s = 0; s < 2; s++)
      for (int c = 0; c < 2; c++) { hint[3 * s + c] = 0.5 * (lo[s][c] + hi[s][c]); hint[6] = std::max(hint[6], 0.5 * (hi[s][c] - lo[s][c])); }
  }
  long long nnz = 0;
  rc = spr_clipper_score(clipper_of(h), cp, pm, (int)m, pd, (int)m, 2, A.data(), (int)m, hint, h->sm_count, st, &nnz, &I->affinity_ms, h->err);
  if (rc != SLIDE_PR_OK) return rc;
  I->nnz_upper = nnz;
  std::vector<double> u0_own;
  if (!u0 || u0_len < m) {  // stand-in for utils::randvec (std::random_device, utils.cpp:22-29)
    u0_own.resize((size_t)m);
    uint64_t x = sp->seed;
    for (long long i = 0; i < m; i++) {
      x += 0x9e3779b97f4a7c15ull;
      uint64_t z = x;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      z ^= z >> 31;
      u0_own[(size_t)i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
    u0 = u0_own.data();
  }
  std::vector<int32_t> nodes((size_t)m);
  slide_clipper_solution sol;
  rc = spr_clipper_solve(h->clipper, cp, u0, h->sm_count, st, nodes.data(), (int32_t)m, &sol, nullptr, h->err);
  if (rc != SLIDE_PR_OK) return rc;
  I->solve_ms = sol.kernel_ms;
  I->n_inliers = sol.n_nodes;
  I->score = sol.score;
  if (sol.n_nodes < sp->num_inliers_threshold) return SLIDE_PR_NOT_FOUND;           // SC.cpp:249-255
  // estimate_tf on the selected pairs (SC.cpp:238-258): model -> data
  std::vector<double> hp((size_t)m * 4);
  SPR_CUDA(h, cudaMemcpy(hp.data(), pm, (size_t)m * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<double> a(2 * (size_t)sol.n_nodes), b(2 * (size_t)sol.n_nodes);
  for (int k = 0; k < sol.n_nodes; k++) {
    const size_t nd = (size_t)nodes[k];
    a[2 * k] = hp[2 * nd]; a[2 * k + 1] = hp[2 * nd + 1];
    b[2 * k] = hp[2 * (size_t)m + 2 * nd]; b[2 * k + 1] = hp[2 * (size_t)m + 2 * nd + 1];
  }
  double tf9[9];
  spr::estimate_tf(a.data(), b.data(), sol.n_nodes, tf9);
  const double yaw = std::atan2(tf9[3], tf9[0]);                                    // SC.cpp:261-268
  for (int i = 0; i < 16; i++) tf16[i] = (i % 5 == 0) ? 1.0 : 0.0;
  tf16[3] = tf9[2]; tf16[7] = tf9[5];
  tf16[0] = std::cos(yaw); tf16[1] = -std::sin(yaw); tf16[4] = std::sin(yaw); tf16[5] = std::cos(yaw);
  I->found = 1;
  return SLIDE_PR_OK;
}

int slide_pr_find_inter_loop_closure_with_clipper(slide_pr_handle *h, const double *ref7, int32_t n_ref, const double *qry7,
                                                  int32_t n_qry, const slide_pr_slidegraph_params *sp, double *tf16,
                                                  slide_pr_sc_info *info_opt) {
  if (!h || !sp || !tf16 || n_ref < 0 || n_qry < 0 || (n_ref > 0 && !ref7) || (n_qry > 0 && !qry7)) return SLIDE_PR_ERR_INVALID;
  if (info_opt) std::memset(info_opt, 0, sizeof(*info_opt));
  // objects at exactly (0, 0) are invalid (PR.cpp:576-612)
  std::vector<double> ref, qry;
  ref.reserve(7 * (size_t)n_ref); qry.reserve(7 * (size_t)n_qry);
  for (int i = 0; i < n_ref; i++)
    if (!(ref7[7 * (size_t)i + 1] == 0.0 && ref7[7 * (size_t)i + 2] == 0.0)) ref.insert(ref.end(), ref7 + 7 * (size_t)i, ref7 + 7 * (size_t)i + 7);
  for (int i = 0; i < n_qry; i++)
    if (!(qry7[7 * (size_t)i + 1] == 0.0 && qry7[7 * (size_t)i + 2] == 0.0)) qry.insert(qry.end(), qry7 + 7 * (size_t)i, qry7 + 7 * (size_t)i + 7);
  const int nr = (int)(ref.size() / 7), nq = (int)(qry.size() / 7);
  if (nr < sp->min_num_map_objects_to_start || nq < sp->min_num_map_objects_to_start) return SLIDE_PR_NOT_FOUND;  // PR.cpp:618
  double fwd[16];
  const int rc = slide_pr_run_semantic_clipper(h, ref.data(), nr, qry.data(), nq, sp, nullptr, 0, nullptr, 0, nullptr, 0, fwd, info_opt);
  if (rc != SLIDE_PR_OK) return rc;
  spr::mat4_rigid_inverse(fwd, tf16);                                                // PR.cpp:622-623 (a yaw + translation matrix)
  return SLIDE_PR_OK;
}

#pragma GCC visibility pop
}  // extern "C"
