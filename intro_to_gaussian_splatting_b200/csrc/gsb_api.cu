// gsb_api.cu -- the C ABI (include/gsb.h): context, scratch management, kernel-chain orchestration.
//
// Frame pipeline (one stream, no CPU work between kernels except ONE read-back of the tile-instance
// count K, needed to size the key buffers and the sort grid):
//
//   FULL :  project -> scan(index order) -> [K] -> emit -> histogram + p x onesweep(K, 64-bit keys)
//           -> ranges -> composite                          p = ceil((32 + tile_bits) / 8)
//   SPLIT:  project -> histogram + 4 x onesweep(N, 32-bit depth keys) -> scan(depth order) -> [K]
//           -> emit(depth order) -> histogram + ceil(tile_bits/8) x onesweep(K, tile digits only)
//           -> ranges -> composite
//
// Both leave bit-identical sorted (key, payload) arrays: an LSD radix sort orders by the low (depth)
// digits first, and every tile instance of a Gaussian shares those digits, so they can be sorted once
// per Gaussian BEFORE the expansion instead of once per instance after it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "gsb_internal.cuh"

using namespace gsb;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return GSB_OK;
    size_t want = bytes + bytes / 4 + 256;  // geometric growth
    if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return (int)e; }
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); return GSB_E_ALLOC; }
    cap = want;
    return GSB_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

bool is_device_pointer(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int tile_grid_dim(int extent, int T, int full_cover) {
  if (full_cover) return (extent + T - 1) / T;
  int span = extent - T;  // len(range(0, extent - T, T)): splat/gaussian_scene.py:208,:214
  return span <= 0 ? 0 : (span + T - 1) / T;
}

int ceil_log2(int64_t v) {
  int b = 0;
  while (((int64_t)1 << b) < v) ++b;
  return b < 1 ? 1 : b;
}

// control block, zeroed once per frame:
//   [hdr 16: 0 = M, 1 = K mod 2^32 (scan), 2-3 = K (tile grid, 64-bit)] [hist 8 x 256: rows 0-3 depth digits, 4-7 tile digits]
//   [difference grid (tiles_x+1)*(tiles_y+1)] [tile cursors (BINNED)] [emit scan status] [depth-sort tickets + status]
constexpr int kCtlHeaderWords = 16;
constexpr int kCtlHistWords = kMaxPasses * kRadix;

struct CtlLayout {
  size_t hist, grid, cursor, scan, dsort, total;  // word offsets
};
CtlLayout ctl_layout(FrameGeom g, int64_t n_rows, size_t dsort_words) {
  auto up4 = [](size_t w) { return (w + 3) & ~(size_t)3; };  // sections start 16-byte aligned (64-bit status words)
  CtlLayout L;
  L.hist = kCtlHeaderWords;
  L.grid = L.hist + kCtlHistWords;
  L.cursor = up4(L.grid + (size_t)(g.tiles_x + 1) * (size_t)(g.tiles_y + 1));  // BINNED: one fill cursor per tile
  L.scan = up4(L.cursor + (size_t)g.tiles_x * (size_t)g.tiles_y);
  L.dsort = up4(L.scan + scan_status_words(n_rows));
  L.total = up4(L.dsort + dsort_words);
  return L;
}

}  // namespace

struct GsbContext {
  int device = 0;
  int64_t n = 0, n_pad = 0;
  DevBuf planes, staging;
  // per-Gaussian frame data
  DevBuf depth_key, rec, rect, count, offsets, bbox;
  DevBuf dbg_cov2d, dbg_conic, dbg_bbox;
  DevBuf ord_keys_a, ord_keys_b, ord_vals_a, ord_vals_b, rank;
  // per-instance
  DevBuf keys_a, keys_b, vals_a, vals_b;
  DevBuf ranges, control, control2;
  DevBuf image, image2, scratch;
  // save_for_backward: per-pixel blended count / final transmittance of the last frame; gradient scratch
  DevBuf aux_t, aux_n, grad2d, grad_stage;
  bool allow_keys32 = true;
  bool frame_projected = false;  // last frame came from render_device: last_cam / last_prm describe its projection
  GsbCamera last_cam{};
  GsbParams last_prm{};
  bool have_saved = false;
  GsbCamera saved_cam{};
  GsbParams saved_prm{};
  uint64_t scene_gen = 0, saved_gen = 0;
  uint32_t* pinned = nullptr;  // mailbox written by tile_stats_kernel: [0]=M, [2..3]=K, [4]=frame sequence number
  uint32_t* pinned_dev = nullptr;  // device alias of the mailbox
  uint32_t seq = 0;
  cudaStream_t aux = nullptr;   // side stream for work that is off the critical path (tile stats)
  cudaEvent_t ev_fork = nullptr, ev_stats = nullptr;
  // asynchronous image egress: two device staging images, a copy stream, one event per staging image
  cudaStream_t copy = nullptr;
  DevBuf host_stage[2];
  cudaEvent_t ev_rendered = nullptr, ev_copied[2] = {nullptr, nullptr};
  bool copy_pending[2] = {false, false};
  int stage_next = 0;
  // state of the last frame
  bool have_frame = false;
  bool sorted_in_a = true;
  bool emitted_valid = false;
  bool keys_materialized = true;  // false in SPLIT mode: sorted 64-bit keys are rebuilt on demand (debug)
  bool order_in_a = true;
  bool have_order = false;
  int64_t frame_rows = 0;  // rows of the per-Gaussian arrays of the last frame (N, or M for gsb_render_image)
  GsbFrameInfo info{};
  cudaEvent_t ev[GSB_NUM_STAGES + 5]{};
  int mark_stage[GSB_NUM_STAGES + 5]{};
  int n_marks = 0;
  float stage_ms[GSB_NUM_STAGES]{};
  bool have_times = false;
};

namespace {

// chronological list of (stage, event): a stage's time is its event minus the previous one in the list
struct StageTimer {
  GsbContext* c;
  cudaStream_t st;
  bool on;
  void start() {
    c->n_marks = 0;
    if (on) cudaEventRecord(c->ev[0], st);
  }
  void mark(int stage) {
    if (!on || c->n_marks >= GSB_NUM_STAGES + 4) return;
    ++c->n_marks;
    c->mark_stage[c->n_marks] = stage;
    cudaEventRecord(c->ev[c->n_marks], st);
  }
};

int check_params(const GsbCamera* cam, const GsbParams* prm) {
  if (!cam || !prm) return GSB_E_INVALID_ARG;
  if (prm->tile_size != kTile) return GSB_E_UNSUPPORTED;
  if (cam->width <= 0 || cam->height <= 0) return GSB_E_INVALID_ARG;
  if (cam->width > 65535 * kTile || cam->height > 65535 * kTile) return GSB_E_UNSUPPORTED;
  if (prm->semantics != GSB_SEM_REF_CPU && prm->semantics != GSB_SEM_REF_CU) return GSB_E_INVALID_ARG;
  if (prm->sort_mode < GSB_SORT_AUTO || prm->sort_mode > GSB_SORT_BINNED) return GSB_E_INVALID_ARG;
  return GSB_OK;
}

// tile_stats_kernel needs only the projection's difference grid, not the depth sort: run it on the context's
// auxiliary stream so that it overlaps the (latency-bound) depth-sort passes; the main stream joins on ev_stats.
int launch_stats_async(GsbContext* c, FrameGeom geom, uint32_t* ctl, const CtlLayout& L, cudaStream_t st, int* launches) {
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  GSB_TRY(c->ranges.ensure((size_t)(tiles > 0 ? tiles : 1) * 8));
  GSB_CUDA_TRY(cudaEventRecord(c->ev_fork, st));
  GSB_CUDA_TRY(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
  ++c->seq;
  GSB_CUDA_TRY((cudaError_t)launch_tile_stats(reinterpret_cast<int32_t*>(ctl + L.grid), geom, ctl + L.hist + 4 * kRadix,
                                              c->ranges.as<uint2>(), ctl + 2, ctl, c->pinned_dev, c->seq, c->aux));
  if (tiles > 0) ++*launches;
  GSB_CUDA_TRY(cudaEventRecord(c->ev_stats, c->aux));
  return GSB_OK;
}

// Wait for tile_stats_kernel's mailbox (M, K) without draining the main stream: the depth-sort passes queued
// behind the projection keep running while the host sizes the key buffers and queues emit / sort / composite.
int wait_counts(GsbContext* c, int64_t tiles, const uint32_t* ctl, cudaStream_t st, int64_t* m, int64_t* k,
                int64_t* v = nullptr) {
  if (v) *v = 0;
  if (tiles <= 0) {  // no tile grid, no tile_stats launch: plain read-back of M, K = 0
    GSB_CUDA_TRY(cudaMemcpyAsync(c->pinned, ctl, 4, cudaMemcpyDeviceToHost, st));
    GSB_CUDA_TRY(cudaStreamSynchronize(st));
    *m = c->pinned[0]; *k = 0;
    return GSB_OK;
  }
  volatile uint32_t* box = c->pinned;
  const auto t0 = std::chrono::steady_clock::now();
  for (uint64_t spins = 0; box[4] != c->seq; ++spins) {
    if ((spins & 0x3FFF) == 0x3FFF) {
      cudaError_t q = cudaStreamQuery(c->aux);
      if (q != cudaSuccess && q != cudaErrorNotReady) return (int)q;
      if (q == cudaSuccess && box[4] != c->seq) return GSB_E_INTERNAL;  // kernel finished, mailbox never written
      // a stream that never runs (e.g. waiting on an event nobody records) must not hang the caller for ever
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return GSB_E_INTERNAL;
    }
  }
  *m = box[0];
  *k = (int64_t)(((uint64_t)box[3] << 32) | box[2]);
  if (v) *v = box[1];
  return GSB_OK;
}

// depth sort of the per-Gaussian keys (N items): leaves the order in ord_vals_{a|b}.  The first pass reads
// depth_key directly and synthesises payload = index; `hist` = the 4 depth-digit histograms (unweighted).
int depth_sort(GsbContext* c, int64_t n, const uint32_t* hist, uint32_t* control_words, cudaStream_t st, int* launches,
               int* passes) {
  GSB_TRY(c->ord_keys_a.ensure((size_t)n * 4));
  GSB_TRY(c->ord_keys_b.ensure((size_t)n * 4));
  GSB_TRY(c->ord_vals_a.ensure((size_t)n * 4));
  GSB_TRY(c->ord_vals_b.ensure((size_t)n * 4));
  SortPlan plan = make_sort_plan<uint32_t>(n, 0, 32);
  bool in_a = true;
  GSB_CUDA_TRY((cudaError_t)launch_sort<uint32_t>(plan, c->depth_key.as<uint32_t>(), nullptr, c->ord_keys_a.as<uint32_t>(),
                                                  c->ord_vals_a.as<uint32_t>(), c->ord_keys_b.as<uint32_t>(),
                                                  c->ord_vals_b.as<uint32_t>(), hist, control_words, &in_a, launches, st));
  c->order_in_a = in_a;
  c->have_order = true;
  *passes = plan.passes;
  return GSB_OK;
}

// everything after the per-Gaussian records exist: tile stats -> scan -> K -> emit -> sort.
// `n_rows` per-Gaussian rows; `perm` optional emission order; `low_bits_sorted`: emission order already
// sorts the low key word (SPLIT mode / pre-sorted rows), so only the tile digits need radix passes.
// `rows_with_tiles`: upper bound of the emission positions that emit anything when the HOST knows one
// (pre-sorted rows: M), -1 when the projection counted it (V, read from the mailbox).
// With low_bits_sorted the keys shrink to 32 bits -- tile << rank_bits | emission position -- whenever
// ceil(log2 tiles) + ceil(log2 V) <= 32 (config 3: 13 + 19); the last radix pass then writes perm[position].
int bin_and_sort(GsbContext* c, int64_t n_rows, const uint32_t* perm, bool low_bits_sorted, int64_t rows_with_tiles,
                 FrameGeom geom, uint32_t* ctl, const CtlLayout& L, cudaStream_t st, StageTimer& tm, int* launches) {
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  GSB_TRY(c->offsets.ensure((size_t)n_rows * 4 + 4));
  GSB_CUDA_TRY((cudaError_t)launch_scan(c->count.as<uint32_t>(), perm, n_rows, c->offsets.as<uint32_t>(), ctl + L.scan, st));
  if (n_rows > 0) ++*launches;
  tm.mark(GSB_STAGE_SCAN);
  // tile stats were launched on the auxiliary stream right after the projection (they do not depend on the
  // depth sort) and post M and K to the host mailbox; pick them up without draining the main stream
  int64_t m = 0, k = 0, v = 0;
  GSB_TRY(wait_counts(c, tiles, ctl, st, &m, &k, &v));
  GSB_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_stats, 0));  // ranges / tile histograms are inputs of what follows
  if (k >= ((int64_t)1 << 32) - 1) return GSB_E_UNSUPPORTED;  // payload positions and ranges are u32
  c->info.m_in_view = m;
  c->info.k_instances = k;
  if (rows_with_tiles >= 0) v = rows_with_tiles;

  GSB_TRY(c->keys_a.ensure((size_t)k * 8 + 8));
  GSB_TRY(c->keys_b.ensure((size_t)k * 8 + 8));
  GSB_TRY(c->vals_a.ensure((size_t)k * 4 + 4));
  GSB_TRY(c->vals_b.ensure((size_t)k * 4 + 4));

  const int tile_bits = ceil_log2(tiles);
  const int rank_bits = ceil_log2(v);
  const bool keys32 = low_bits_sorted && c->allow_keys32 && tile_bits + rank_bits <= 32;
  const uint32_t* hist = ctl + L.hist + (low_bits_sorted ? 4 * kRadix : 0);
  bool in_a = true;
  int passes = 0;
  if (keys32) {
    SortPlan plan = make_sort_plan<uint32_t>(k, rank_bits, rank_bits + tile_bits);
    plan.keys_only = 1;
    plan.low_bits = rank_bits;
    plan.gather_table = perm;  // nullptr (pre-sorted rows): the position IS the row
    GSB_TRY(c->control2.ensure(plan.control_words * 4));
    GSB_CUDA_TRY(cudaMemsetAsync(c->control2.p, 0, plan.control_words * 4, st));
    GSB_CUDA_TRY((cudaError_t)launch_emit(c->offsets.as<uint32_t>(), perm, ctl + 2, n_rows, k, c->depth_key.as<uint32_t>(),
                                          c->rect.as<ushort4>(), geom.tiles_x, true, rank_bits,
                                          c->keys_a.as<uint64_t>(), c->vals_a.as<uint32_t>(), st));
    if (n_rows > 0 && k > 0) ++*launches;
    tm.mark(GSB_STAGE_EMIT);
    GSB_CUDA_TRY((cudaError_t)launch_sort<uint32_t>(plan, c->keys_a.as<uint32_t>(), nullptr, c->keys_a.as<uint32_t>(),
                                                    c->vals_a.as<uint32_t>(), c->keys_b.as<uint32_t>(),
                                                    c->vals_b.as<uint32_t>(), hist, c->control2.as<uint32_t>(), &in_a,
                                                    launches, st));
    passes = plan.passes;
  } else {
    SortPlan plan = low_bits_sorted ? make_sort_plan<uint64_t>(k, 32, 32 + tile_bits)
                                    : make_sort_plan<uint64_t>(k, 0, 32 + tile_bits);
    GSB_TRY(c->control2.ensure(plan.control_words * 4));
    GSB_CUDA_TRY(cudaMemsetAsync(c->control2.p, 0, plan.control_words * 4, st));
    plan.keys_only = low_bits_sorted ? 1 : 0;
    GSB_CUDA_TRY((cudaError_t)launch_emit(c->offsets.as<uint32_t>(), perm, ctl + 2, n_rows, k, c->depth_key.as<uint32_t>(),
                                          c->rect.as<ushort4>(), geom.tiles_x, /*combined=*/low_bits_sorted, 0,
                                          c->keys_a.as<uint64_t>(), c->vals_a.as<uint32_t>(), st));
    if (n_rows > 0 && k > 0) ++*launches;
    tm.mark(GSB_STAGE_EMIT);
    GSB_CUDA_TRY((cudaError_t)launch_sort<uint64_t>(plan, c->keys_a.as<uint64_t>(), c->vals_a.as<uint32_t>(),
                                                    c->keys_a.as<uint64_t>(), c->vals_a.as<uint32_t>(),
                                                    c->keys_b.as<uint64_t>(), c->vals_b.as<uint32_t>(), hist,
                                                    c->control2.as<uint32_t>(), &in_a, launches, st));
    passes = plan.passes;
  }
  c->keys_materialized = !low_bits_sorted;
  c->sorted_in_a = in_a;
  c->emitted_valid = (passes <= 1) && !low_bits_sorted;  // one pass: the a-buffers still hold the emitted order
  c->info.sort_passes = (k > 0) ? passes : 0;
  c->info.key_bits = keys32 ? 32 : 64;
  tm.mark(GSB_STAGE_SORT);
  return GSB_OK;
}

// BINNED mode: tile stats -> K -> unordered per-tile segments (atomic cursors) -> per-tile sort by depth rank.
int bin_by_tile(GsbContext* c, int64_t n_rows, const uint32_t* order, FrameGeom geom, uint32_t* ctl, const CtlLayout& L,
                cudaStream_t st, StageTimer& tm, int* launches) {
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  int64_t m = 0, k = 0;
  GSB_TRY(wait_counts(c, tiles, ctl, st, &m, &k));
  GSB_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_stats, 0));  // tile stats ran on the auxiliary stream
  if (k >= ((int64_t)1 << 32) - 1) return GSB_E_UNSUPPORTED;  // payload positions and ranges are u32
  c->info.m_in_view = m;
  c->info.k_instances = k;
  GSB_TRY(c->vals_a.ensure((size_t)k * 4 + 4));
  GSB_TRY(c->keys_a.ensure((size_t)k * 4 + 4));  // scratch for tile segments longer than shared memory
  GSB_TRY(c->keys_b.ensure((size_t)k * 4 + 4));
  if (k > 0 && n_rows > 0) {
    GSB_CUDA_TRY((cudaError_t)launch_emit_binned(n_rows, c->rect.as<ushort4>(), c->count.as<uint32_t>(), geom.tiles_x,
                                                 c->ranges.as<uint2>(), ctl + L.cursor, c->vals_a.as<uint32_t>(), st));
    ++*launches;
    tm.mark(GSB_STAGE_EMIT);
    const int rank_bits = ceil_log2(n_rows);
    GSB_CUDA_TRY((cudaError_t)launch_tile_sort(c->ranges.as<uint2>(), (int)tiles, c->rank.as<uint32_t>(), order,
                                               c->vals_a.as<uint32_t>(), rank_bits, c->keys_a.as<uint32_t>(),
                                               c->keys_b.as<uint32_t>(), st));
    ++*launches;
    c->info.sort_passes = (rank_bits + 7) / 8;
    tm.mark(GSB_STAGE_SORT);
  }
  c->sorted_in_a = true;
  c->emitted_valid = false;
  c->keys_materialized = false;
  return GSB_OK;
}

void finish_times(GsbContext* c, cudaStream_t st, bool on) {
  c->have_times = false;
  if (!on) return;
  if (cudaStreamSynchronize(st) != cudaSuccess) return;
  for (float& v : c->stage_ms) v = 0.f;
  for (int i = 1; i <= c->n_marks; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[i - 1], c->ev[i]) == cudaSuccess) c->stage_ms[c->mark_stage[i]] += ms;
  }
  c->have_times = true;
}

// projection + binning + sort + compositing into a DEVICE image buffer
int render_device(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, float* dev_image, cudaStream_t st) {
  GSB_TRY(check_params(cam, prm));
  if (c->n_pad == 0 && c->n != 0) return GSB_E_NO_SCENE;
  if (!c->planes.p && c->n > 0) return GSB_E_NO_SCENE;
  if (prm->semantics != GSB_SEM_REF_CPU) return GSB_E_UNSUPPORTED;  // REF_CU is served by gsb_render_image
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const int64_t n = c->n;
  FrameGeom geom{cam->width, cam->height, tile_grid_dim(cam->width, kTile, prm->full_cover),
                 tile_grid_dim(cam->height, kTile, prm->full_cover)};
  // AUTO = SPLIT, the fastest measured (config 3: tile sort 183 us vs 562 us FULL; BINNED spends 140 us in the
  // atomic emit and 370 us in the per-tile sort, profiles/r1_summary.md)
  const int mode = prm->sort_mode == GSB_SORT_AUTO ? GSB_SORT_SPLIT : prm->sort_mode;
  const bool split = mode != GSB_SORT_FULL;  // SPLIT and BINNED both sort the depth keys per Gaussian first
  int launches = 0;
  c->have_frame = false;
  c->frame_projected = false;
  c->have_order = false;
  c->have_saved = false;
  std::memset(&c->info, 0, sizeof(c->info));
  c->info.n = n; c->info.tiles_x = geom.tiles_x; c->info.tiles_y = geom.tiles_y;
  StageTimer tm{c, st, prm->collect_stage_times != 0};

  const size_t rows = (size_t)(n > 0 ? n : 1);
  GSB_TRY(c->depth_key.ensure(rows * 4));
  GSB_TRY(c->rec.ensure(rows * 48));
  GSB_TRY(c->rect.ensure(rows * 8));
  GSB_TRY(c->count.ensure(rows * 4));
  SortPlan dplan = make_sort_plan<uint32_t>(n, 0, 32);
  const CtlLayout L = ctl_layout(geom, n, split ? dplan.control_words : 0);
  GSB_TRY(c->control.ensure(L.total * 4));
  GSB_CUDA_TRY(cudaMemsetAsync(c->control.p, 0, L.total * 4, st));
  uint32_t* ctl = c->control.as<uint32_t>();

  tm.start();
  GSB_CUDA_TRY((cudaError_t)launch_project(c->planes.as<float>(), n, c->n_pad, *cam, *prm, geom, c->depth_key.as<uint32_t>(),
                                           c->rec.as<float4>(), c->rect.as<ushort4>(), c->count.as<uint32_t>(), ctl,
                                           ctl + L.hist, /*hist_weighted=*/split ? 0 : 1,
                                           reinterpret_cast<int32_t*>(ctl + L.grid), nullptr, st));
  if (n > 0) ++launches;
  tm.mark(GSB_STAGE_PROJECT);
  GSB_TRY(launch_stats_async(c, geom, ctl, L, st, &launches));
  const uint32_t* perm = nullptr;
  if (split && n > 0) {
    int dp = 0;
    GSB_TRY(depth_sort(c, n, ctl + L.hist, ctl + L.dsort, st, &launches, &dp));
    c->info.depth_passes = dp;
    perm = c->order_in_a ? c->ord_vals_a.as<uint32_t>() : c->ord_vals_b.as<uint32_t>();
    if (mode == GSB_SORT_BINNED) {
      GSB_TRY(c->rank.ensure(rows * 4));
      GSB_CUDA_TRY((cudaError_t)launch_invert_perm(perm, n, c->rank.as<uint32_t>(), st));
      ++launches;
    }
    tm.mark(GSB_STAGE_DEPTH_SORT);
  }
  if (mode == GSB_SORT_BINNED) {
    GSB_TRY(bin_by_tile(c, n, perm, geom, ctl, L, st, tm, &launches));
  } else {
    GSB_TRY(bin_and_sort(c, n, perm, split, /*rows_with_tiles=*/-1, geom, ctl, L, st, tm, &launches));
  }

  if (!prm->full_cover)  // pixels outside the reference tile grid stay 0 (splat/gaussian_scene.py:206)
    GSB_CUDA_TRY(cudaMemsetAsync(dev_image, 0, (size_t)cam->width * cam->height * 3 * sizeof(float), st));
  const uint32_t* sv = c->sorted_in_a ? c->vals_a.as<uint32_t>() : c->vals_b.as<uint32_t>();
  float* aux_t = nullptr;
  uint32_t* aux_n = nullptr;
  if (prm->save_for_backward) {
    const size_t px = (size_t)cam->width * cam->height;
    GSB_TRY(c->aux_t.ensure(px * 4));
    GSB_TRY(c->aux_n.ensure(px * 4));
    aux_t = c->aux_t.as<float>();
    aux_n = c->aux_n.as<uint32_t>();
    if (!prm->full_cover) GSB_CUDA_TRY(cudaMemsetAsync(aux_n, 0, px * 4, st));  // pixels outside the grid: nothing blended
  }
  GSB_CUDA_TRY((cudaError_t)launch_composite(c->ranges.as<uint2>(), sv, c->rec.as<float4>(), dev_image, geom, *prm,
                                             aux_t, aux_n, st));
  if (geom.tiles_x * geom.tiles_y > 0) ++launches;
  tm.mark(GSB_STAGE_COMPOSITE);
  c->info.kernel_launches = launches;
  c->frame_rows = n;
  c->have_frame = true;
  c->frame_projected = true;
  c->last_cam = *cam;
  c->last_prm = *prm;
  if (prm->save_for_backward) {
    c->have_saved = true;
    c->saved_cam = *cam;
    c->saved_prm = *prm;
    c->saved_gen = c->scene_gen;
  }
  finish_times(c, st, tm.on);
  return GSB_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int gsb_version(void) { return GSB_API_VERSION; }

const char* gsb_error_string(int s) {
  switch (s) {
    case GSB_OK: return "ok";
    case GSB_E_INVALID_ARG: return "invalid argument";
    case GSB_E_NO_SCENE: return "no Gaussians uploaded (call gsb_upload first)";
    case GSB_E_NO_FRAME: return "no frame rendered yet";
    case GSB_E_UNSUPPORTED: return "unsupported configuration (tile_size must be 16; fewer than 2^32-1 tile instances)";
    case GSB_E_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case GSB_E_ALLOC: return "device memory allocation failed";
    case GSB_E_INTERNAL: return "internal error (device-side consistency check failed, or the stream made no progress)";
    case GSB_E_NO_SAVED: return "no saved frame for the backward pass (render with save_for_backward = 1 and the same camera / params first)";
    default: return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown error";
  }
}

void gsb_default_params(GsbParams* p) {
  if (!p) return;
  p->tile_size = 16;
  p->minimum_z = 0.2f;
  p->fov_clamp = 1.3f;
  p->det_min = 1e-3f;
  p->lambda_floor = 0.1f;
  p->sigma_extent = 3.0f;
  p->min_weight = 1e-6f;
  p->alpha_max = 0.99f;
  p->semantics = GSB_SEM_REF_CPU;
  p->full_cover = 0;
  p->sort_mode = GSB_SORT_AUTO;
  p->collect_stage_times = 0;
  p->async_host_copy = 0;
  p->save_for_backward = 0;
}

int gsb_create(GsbContext** out, int device) {
  if (!out) return GSB_E_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return GSB_E_NO_DEVICE; }
  if (device < 0 || device >= count) return GSB_E_INVALID_ARG;
  GSB_CUDA_TRY(cudaSetDevice(device));
  GsbContext* c = new (std::nothrow) GsbContext();
  if (!c) return GSB_E_ALLOC;
  c->device = device;
  {  // process-wide tuning / test knobs, re-read whenever a context is created
    const char* e = std::getenv("GSB_SORT_ITEMS");        // onesweep keys per thread: 16 (default) or 8
    set_sort_items(e ? std::atoi(e) : 16);
    e = std::getenv("GSB_FORCE_WIDE_STATUS");             // 1: 64-bit look-back words even below 2^30 keys
    set_force_wide_status(e ? std::atoi(e) : 0);
    e = std::getenv("GSB_KEYS32");                        // 0: never use the 32-bit tile keys of SPLIT mode
    c->allow_keys32 = !e || std::atoi(e) != 0;
  }
  if (cudaHostAlloc((void**)&c->pinned, 64, cudaHostAllocMapped) != cudaSuccess) { delete c; return GSB_E_ALLOC; }
  std::memset(c->pinned, 0, 64);
  if (cudaHostGetDevicePointer((void**)&c->pinned_dev, c->pinned, 0) != cudaSuccess) { gsb_destroy(c); return GSB_E_ALLOC; }
  if (cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_stats, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_rendered, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_copied[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_copied[1], cudaEventDisableTiming) != cudaSuccess) {
    gsb_destroy(c);
    return GSB_E_ALLOC;
  }
  for (auto& e : c->ev)
    if (cudaEventCreate(&e) != cudaSuccess) { gsb_destroy(c); return GSB_E_ALLOC; }
  *out = c;
  return GSB_OK;
}

void gsb_destroy(GsbContext* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&c->planes, &c->staging, &c->depth_key, &c->rec, &c->rect, &c->count, &c->offsets, &c->bbox,
                    &c->dbg_cov2d, &c->dbg_conic, &c->dbg_bbox, &c->ord_keys_a, &c->ord_keys_b, &c->ord_vals_a,
                    &c->ord_vals_b, &c->rank, &c->keys_a, &c->keys_b, &c->vals_a, &c->vals_b, &c->ranges, &c->control,
                    &c->control2, &c->image, &c->image2, &c->scratch};
  for (DevBuf* b : bufs) b->release();
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->copy) cudaStreamDestroy(c->copy);
  if (c->ev_rendered) cudaEventDestroy(c->ev_rendered);
  for (auto& e : c->ev_copied) if (e) cudaEventDestroy(e);
  c->host_stage[0].release(); c->host_stage[1].release();
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_stats) cudaEventDestroy(c->ev_stats);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  delete c;
}

int gsb_upload(GsbContext* c, int64_t n, const float* xyz, const float* scales, const float* quats,
               const float* colors, const float* opacity_logit, void* stream) {
  if (!c || n < 0) return GSB_E_INVALID_ARG;
  if (n > 0 && (!xyz || !scales || !quats || !colors || !opacity_logit)) return GSB_E_INVALID_ARG;
  if (n >= ((int64_t)1 << 31)) return GSB_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  c->have_frame = false;
  c->frame_projected = false;
  c->have_saved = false;
  ++c->scene_gen;
  c->n = n;
  c->n_pad = (n + 3) & ~(int64_t)3;
  if (n == 0) return GSB_OK;
  GSB_TRY(c->planes.ensure((size_t)c->n_pad * kNumPlanes * sizeof(float)));
  const float* src[5] = {xyz, scales, quats, colors, opacity_logit};
  const int width[5] = {3, 3, 4, 3, 1};
  const float* dev[5];
  size_t need = 0;
  for (int i = 0; i < 5; ++i)
    if (!is_device_pointer(src[i])) need += (size_t)n * width[i] * sizeof(float);
  if (need) GSB_TRY(c->staging.ensure(need));
  size_t off = 0;
  for (int i = 0; i < 5; ++i) {
    if (is_device_pointer(src[i])) { dev[i] = src[i]; continue; }
    float* d = reinterpret_cast<float*>(c->staging.as<char>() + off);
    size_t bytes = (size_t)n * width[i] * sizeof(float);
    GSB_CUDA_TRY(cudaMemcpyAsync(d, src[i], bytes, cudaMemcpyHostToDevice, st));
    dev[i] = d;
    off += bytes;
  }
  GSB_CUDA_TRY((cudaError_t)launch_repack(dev[0], dev[1], dev[2], dev[3], dev[4], c->planes.as<float>(), n, c->n_pad, st));
  if (need) {  // host sources may be reused by the caller; staging is ours, so just make the copy complete
    GSB_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return GSB_OK;
}

int gsb_join_host_copies(GsbContext* c, void* stream) {
  if (!c) return GSB_E_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  for (int b = 0; b < 2; ++b)
    if (c->copy_pending[b]) {
      GSB_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copied[b], 0));
      c->copy_pending[b] = false;
    }
  return GSB_OK;
}

int gsb_render(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, float* out_image, void* stream) {
  if (!c || !out_image) return GSB_E_INVALID_ARG;
  GSB_TRY(check_params(cam, prm));
  cudaStream_t st = (cudaStream_t)stream;
  if (is_device_pointer(out_image)) return render_device(c, cam, prm, out_image, st);
  const size_t bytes = (size_t)cam->width * cam->height * 3 * sizeof(float);
  if (prm->async_host_copy) {
    // double-buffered staging: frame i is copied to the host on the copy stream while frame i+1 renders
    const int b = c->stage_next;
    c->stage_next ^= 1;
    GSB_TRY(c->host_stage[b].ensure(bytes));
    if (c->copy_pending[b]) GSB_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copied[b], 0));  // staging image still in flight
    GSB_TRY(render_device(c, cam, prm, c->host_stage[b].as<float>(), st));
    GSB_CUDA_TRY(cudaEventRecord(c->ev_rendered, st));
    GSB_CUDA_TRY(cudaStreamWaitEvent(c->copy, c->ev_rendered, 0));
    GSB_CUDA_TRY(cudaMemcpyAsync(out_image, c->host_stage[b].p, bytes, cudaMemcpyDeviceToHost, c->copy));
    GSB_CUDA_TRY(cudaEventRecord(c->ev_copied[b], c->copy));
    c->copy_pending[b] = true;
    return GSB_OK;
  }
  GSB_TRY(c->image.ensure(bytes));
  GSB_TRY(render_device(c, cam, prm, c->image.as<float>(), st));
  GSB_CUDA_TRY(cudaMemcpyAsync(out_image, c->image.p, bytes, cudaMemcpyDeviceToHost, st));
  return GSB_OK;
}

int gsb_render_backward(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, const float* grad_image,
                        float* grad_points, float* grad_scales, float* grad_quats, float* grad_colors,
                        float* grad_opacity, void* stream) {
  if (!c || !grad_image) return GSB_E_INVALID_ARG;
  GSB_TRY(check_params(cam, prm));
  if (!c->have_frame || !c->have_saved || c->saved_gen != c->scene_gen ||
      std::memcmp(&c->saved_cam, cam, sizeof(GsbCamera)) != 0 || std::memcmp(&c->saved_prm, prm, sizeof(GsbParams)) != 0)
    return GSB_E_NO_SAVED;
  cudaStream_t st = (cudaStream_t)stream;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const int64_t n = c->n;
  if (n == 0) return GSB_OK;
  FrameGeom geom{cam->width, cam->height, c->info.tiles_x, c->info.tiles_y};
  const size_t img_bytes = (size_t)cam->width * cam->height * 3 * sizeof(float);
  const float* g_img = grad_image;
  if (!is_device_pointer(grad_image)) {
    GSB_TRY(c->image2.ensure(img_bytes));
    GSB_CUDA_TRY(cudaMemcpyAsync(c->image2.p, grad_image, img_bytes, cudaMemcpyHostToDevice, st));
    g_img = c->image2.as<float>();
  }
  GSB_TRY(c->grad2d.ensure((size_t)n * 12 * sizeof(float)));
  GSB_CUDA_TRY(cudaMemsetAsync(c->grad2d.p, 0, (size_t)n * 12 * sizeof(float), st));
  const uint32_t* sv = c->sorted_in_a ? c->vals_a.as<uint32_t>() : c->vals_b.as<uint32_t>();
  GSB_CUDA_TRY((cudaError_t)launch_composite_backward(c->ranges.as<uint2>(), sv, c->rec.as<float4>(), g_img,
                                                      c->aux_t.as<float>(), c->aux_n.as<uint32_t>(),
                                                      c->grad2d.as<float>(), geom, *prm, st));
  // outputs: device pointers are written directly; host / NULL ones go through one staging block
  float* out[5] = {grad_points, grad_scales, grad_quats, grad_colors, grad_opacity};
  const int width[5] = {3, 3, 4, 3, 1};
  float* dev[5];
  size_t need = 0;
  for (int i = 0; i < 5; ++i)
    if (!out[i] || !is_device_pointer(out[i])) need += (size_t)n * width[i] * sizeof(float);
  if (need) GSB_TRY(c->grad_stage.ensure(need));
  size_t off = 0;
  for (int i = 0; i < 5; ++i) {
    if (out[i] && is_device_pointer(out[i])) { dev[i] = out[i]; continue; }
    dev[i] = reinterpret_cast<float*>(c->grad_stage.as<char>() + off);
    off += (size_t)n * width[i] * sizeof(float);
  }
  GSB_CUDA_TRY((cudaError_t)launch_project_backward(c->planes.as<float>(), n, c->n_pad, *cam, *prm,
                                                    c->depth_key.as<uint32_t>(), c->count.as<uint32_t>(),
                                                    c->grad2d.as<float>(), dev[0], dev[1], dev[2], dev[3], dev[4], st));
  for (int i = 0; i < 5; ++i)
    if (out[i] && dev[i] != out[i])
      GSB_CUDA_TRY(cudaMemcpyAsync(out[i], dev[i], (size_t)n * width[i] * sizeof(float), cudaMemcpyDeviceToHost, st));
  return GSB_OK;
}

int gsb_render_wh(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, float* out_wh, void* stream) {
  if (!c || !out_wh) return GSB_E_INVALID_ARG;
  GSB_TRY(check_params(cam, prm));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)cam->width * cam->height * 3 * sizeof(float);
  GSB_TRY(c->image.ensure(bytes));
  GSB_TRY(render_device(c, cam, prm, c->image.as<float>(), st));
  if (is_device_pointer(out_wh)) {
    GSB_CUDA_TRY((cudaError_t)launch_hwc_to_whc(c->image.as<float>(), out_wh, cam->width, cam->height, st));
    return GSB_OK;
  }
  GSB_TRY(c->image2.ensure(bytes));
  GSB_CUDA_TRY((cudaError_t)launch_hwc_to_whc(c->image.as<float>(), c->image2.as<float>(), cam->width, cam->height, st));
  GSB_CUDA_TRY(cudaMemcpyAsync(out_wh, c->image2.p, bytes, cudaMemcpyDeviceToHost, st));
  return GSB_OK;
}

int gsb_render_u8(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, uint8_t* out, void* stream) {
  if (!c || !out) return GSB_E_INVALID_ARG;
  GSB_TRY(check_params(cam, prm));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t elems = (int64_t)cam->width * cam->height * 3;
  GSB_TRY(c->image.ensure((size_t)elems * sizeof(float)));
  GSB_TRY(render_device(c, cam, prm, c->image.as<float>(), st));
  if (is_device_pointer(out)) return launch_to_u8(c->image.as<float>(), out, elems, st);
  GSB_TRY(c->image2.ensure((size_t)elems));
  GSB_CUDA_TRY((cudaError_t)launch_to_u8(c->image.as<float>(), c->image2.as<uint8_t>(), elems, st));
  GSB_CUDA_TRY(cudaMemcpyAsync(out, c->image2.p, (size_t)elems, cudaMemcpyDeviceToHost, st));
  return GSB_OK;
}

int gsb_preprocess(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, int64_t* m_out, float* points_xy,
                   float* colors, float* covariance_2d, float* depths, float* inverse_covariance_2d, float* radius,
                   float* min_x, float* min_y, float* max_x, float* max_y, float* sigmoid_opacity,
                   int32_t* source_index, void* stream) {
  if (!c) return GSB_E_INVALID_ARG;
  GSB_TRY(check_params(cam, prm));
  if (!c->planes.p && c->n > 0) return GSB_E_NO_SCENE;
  cudaStream_t st = (cudaStream_t)stream;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const int64_t n = c->n;
  c->have_frame = false;
  c->frame_projected = false;
  if (m_out) *m_out = 0;
  if (n == 0) return GSB_OK;
  FrameGeom geom{cam->width, cam->height, tile_grid_dim(cam->width, kTile, prm->full_cover),
                 tile_grid_dim(cam->height, kTile, prm->full_cover)};
  GSB_TRY(c->depth_key.ensure((size_t)n * 4));
  GSB_TRY(c->rec.ensure((size_t)n * 48));
  GSB_TRY(c->rect.ensure((size_t)n * 8));
  GSB_TRY(c->count.ensure((size_t)n * 4));
  GSB_TRY(c->dbg_cov2d.ensure((size_t)n * 16));
  GSB_TRY(c->dbg_conic.ensure((size_t)n * 16));
  GSB_TRY(c->dbg_bbox.ensure((size_t)n * 16));
  SortPlan dplan = make_sort_plan<uint32_t>(n, 0, 32);
  const CtlLayout L = ctl_layout(geom, n, dplan.control_words);
  GSB_TRY(c->control.ensure(L.total * 4));
  GSB_CUDA_TRY(cudaMemsetAsync(c->control.p, 0, L.total * 4, st));
  uint32_t* hdr = c->control.as<uint32_t>();
  DebugOut dbg{c->dbg_cov2d.as<float>(), c->dbg_conic.as<float>(), c->dbg_bbox.as<float>()};
  GSB_CUDA_TRY((cudaError_t)launch_project(c->planes.as<float>(), n, c->n_pad, *cam, *prm, geom, c->depth_key.as<uint32_t>(),
                                           c->rec.as<float4>(), c->rect.as<ushort4>(), c->count.as<uint32_t>(), hdr,
                                           hdr + L.hist, /*hist_weighted=*/0, reinterpret_cast<int32_t*>(hdr + L.grid),
                                           &dbg, st));
  int launches = 1, dp = 0;
  GSB_TRY(depth_sort(c, n, hdr + L.hist, hdr + L.dsort, st, &launches, &dp));
  GSB_CUDA_TRY(cudaMemcpyAsync(c->pinned, hdr, 4, cudaMemcpyDeviceToHost, st));
  GSB_CUDA_TRY(cudaStreamSynchronize(st));
  const int64_t m = c->pinned[0];
  if (m_out) *m_out = m;
  if (m == 0) return GSB_OK;
  // gather into one staging block, then copy each requested field out (device or host destination)
  const size_t words_per_row = 2 + 3 + 4 + 1 + 4 + 1 + 4 + 1 + 1;
  GSB_TRY(c->scratch.ensure((size_t)m * words_per_row * 4));
  float* base = c->scratch.as<float>();
  float* s_xy = base;            float* s_col = s_xy + 2 * m;   float* s_cov = s_col + 3 * m;
  float* s_dep = s_cov + 4 * m;  float* s_con = s_dep + m;      float* s_rad = s_con + 4 * m;
  float* s_mnx = s_rad + m;      float* s_mny = s_mnx + m;      float* s_mxx = s_mny + m;
  float* s_mxy = s_mxx + m;      float* s_sig = s_mxy + m;      int32_t* s_idx = reinterpret_cast<int32_t*>(s_sig + m);
  const uint32_t* order = c->order_in_a ? c->ord_vals_a.as<uint32_t>() : c->ord_vals_b.as<uint32_t>();
  GSB_CUDA_TRY((cudaError_t)launch_gather_preprocess(order, m, c->rec.as<float4>(), c->planes.as<float>(), c->n_pad,
                                                     c->depth_key.as<uint32_t>(), dbg, s_xy, s_col, s_cov, s_dep, s_con,
                                                     s_rad, s_mnx, s_mny, s_mxx, s_mxy, s_sig, s_idx, st));
  struct { void* dst; const void* src; size_t words; } cp[] = {
      {points_xy, s_xy, 2}, {colors, s_col, 3}, {covariance_2d, s_cov, 4}, {depths, s_dep, 1},
      {inverse_covariance_2d, s_con, 4}, {radius, s_rad, 1}, {min_x, s_mnx, 1}, {min_y, s_mny, 1},
      {max_x, s_mxx, 1}, {max_y, s_mxy, 1}, {sigmoid_opacity, s_sig, 1}, {source_index, s_idx, 1}};
  for (auto& e : cp)
    if (e.dst) GSB_CUDA_TRY(cudaMemcpyAsync(e.dst, e.src, (size_t)m * e.words * 4, cudaMemcpyDefault, st));
  GSB_CUDA_TRY(cudaStreamSynchronize(st));
  return GSB_OK;
}

int gsb_render_image(GsbContext* c, int32_t H, int32_t W, int32_t tile_size, int64_t m, const float* point_means,
                     const float* point_colors, const float* inverse_covariance_2d, const float* min_x,
                     const float* max_x, const float* min_y, const float* max_y, const float* opacity,
                     const GsbParams* prm_in, float* out_image, void* stream) {
  if (!c || !prm_in || !out_image || m < 0) return GSB_E_INVALID_ARG;
  if (m > 0 && (!point_means || !point_colors || !inverse_covariance_2d || !min_x || !max_x || !min_y || !max_y || !opacity))
    return GSB_E_INVALID_ARG;
  GsbParams prm = *prm_in;
  prm.tile_size = tile_size;
  GsbCamera cam{};
  cam.width = W; cam.height = H;
  GSB_TRY(check_params(&cam, &prm));
  if (m >= ((int64_t)1 << 31)) return GSB_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const bool cu = prm.semantics == GSB_SEM_REF_CU;
  const int cover = cu ? 1 : prm.full_cover;  // render.cu covers every pixel (:119-124)
  FrameGeom geom{W, H, tile_grid_dim(W, kTile, cover), tile_grid_dim(H, kTile, cover)};
  c->have_frame = false;
  c->frame_projected = false;
  c->have_order = false;
  std::memset(&c->info, 0, sizeof(c->info));
  c->info.n = m; c->info.tiles_x = geom.tiles_x; c->info.tiles_y = geom.tiles_y;
  StageTimer tm{c, st, prm.collect_stage_times != 0};
  int launches = 0;

  // stage host inputs
  const float* src[8] = {point_means, point_colors, inverse_covariance_2d, min_x, max_x, min_y, max_y, opacity};
  const int width[8] = {2, 3, 4, 1, 1, 1, 1, 1};
  const float* dev[8];
  size_t need = 0;
  for (int i = 0; i < 8; ++i)
    if (m > 0 && !is_device_pointer(src[i])) need += (size_t)m * width[i] * 4;
  if (need) GSB_TRY(c->staging.ensure(need));
  size_t off = 0;
  for (int i = 0; i < 8; ++i) {
    if (m == 0 || is_device_pointer(src[i])) { dev[i] = src[i]; continue; }
    float* d = reinterpret_cast<float*>(c->staging.as<char>() + off);
    GSB_CUDA_TRY(cudaMemcpyAsync(d, src[i], (size_t)m * width[i] * 4, cudaMemcpyHostToDevice, st));
    dev[i] = d;
    off += (size_t)m * width[i] * 4;
  }
  const size_t rows = (size_t)(m > 0 ? m : 1);
  GSB_TRY(c->depth_key.ensure(rows * 4));
  GSB_TRY(c->rec.ensure(rows * 48));
  GSB_TRY(c->rect.ensure(rows * 8));
  GSB_TRY(c->count.ensure(rows * 4));
  GSB_TRY(c->bbox.ensure(rows * 16));
  const CtlLayout L = ctl_layout(geom, m, 0);
  GSB_TRY(c->control.ensure(L.total * 4));
  GSB_CUDA_TRY(cudaMemsetAsync(c->control.p, 0, L.total * 4, st));
  uint32_t* hdr = c->control.as<uint32_t>();
  tm.start();
  GSB_CUDA_TRY((cudaError_t)launch_ingest_preprocessed(m, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7], geom,
                                                       prm, c->depth_key.as<uint32_t>(), c->rec.as<float4>(),
                                                       c->bbox.as<float4>(), c->rect.as<ushort4>(), c->count.as<uint32_t>(),
                                                       reinterpret_cast<int32_t*>(hdr + L.grid), st));
  if (m > 0) ++launches;
  tm.mark(GSB_STAGE_PROJECT);
  GSB_TRY(launch_stats_async(c, geom, hdr, L, st, &launches));
  GSB_TRY(bin_and_sort(c, m, nullptr, /*low_bits_sorted=*/true, /*rows_with_tiles=*/m, geom, hdr, L, st, tm, &launches));
  c->info.m_in_view = m;

  float* dev_image = out_image;
  const size_t bytes = (size_t)W * H * 3 * sizeof(float);
  const bool host_out = !is_device_pointer(out_image);
  if (host_out) { GSB_TRY(c->image.ensure(bytes)); dev_image = c->image.as<float>(); }
  if (!cover) GSB_CUDA_TRY(cudaMemsetAsync(dev_image, 0, bytes, st));
  const uint32_t* sv = c->sorted_in_a ? c->vals_a.as<uint32_t>() : c->vals_b.as<uint32_t>();
  if (cu)
    GSB_CUDA_TRY((cudaError_t)launch_composite_cu(c->ranges.as<uint2>(), sv, c->rec.as<float4>(), c->bbox.as<float4>(),
                                                  dev_image, geom, prm, st));
  else
    GSB_CUDA_TRY((cudaError_t)launch_composite(c->ranges.as<uint2>(), sv, c->rec.as<float4>(), dev_image, geom, prm, nullptr, nullptr, st));
  if (geom.tiles_x * geom.tiles_y > 0) ++launches;
  tm.mark(GSB_STAGE_COMPOSITE);
  if (host_out) GSB_CUDA_TRY(cudaMemcpyAsync(out_image, dev_image, bytes, cudaMemcpyDeviceToHost, st));
  c->info.kernel_launches = launches;
  c->frame_rows = m;
  c->have_frame = true;
  finish_times(c, st, tm.on);
  return GSB_OK;
}

int gsb_frame_info(GsbContext* c, GsbFrameInfo* info) {
  if (!c || !info) return GSB_E_INVALID_ARG;
  if (!c->have_frame) return GSB_E_NO_FRAME;
  *info = c->info;
  return GSB_OK;
}

}  // extern "C"

// ---- debug getters ---------------------------------------------------------------------------------
namespace {
__global__ void unpack_projection_kernel(int64_t n, const uint32_t* __restrict__ depth_key, const float4* __restrict__ rec,
                                         const ushort4* __restrict__ rect, const uint32_t* __restrict__ count,
                                         uint8_t* in_view, float* depth, float* pxy, float* radius, int32_t* trect) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool keep = depth_key[i] != 0xFFFFFFFFu;
  in_view[i] = keep ? 1 : 0;
  float4 r0 = make_float4(0, 0, 0, 0), r2 = r0;
  if (keep) { r0 = rec[3 * i]; r2 = rec[3 * i + 2]; }
  depth[i] = keep ? __uint_as_float(depth_key[i]) : 0.f;
  pxy[2 * i] = r0.x; pxy[2 * i + 1] = r0.y;
  radius[i] = r2.z;
  const ushort4 r = rect[i];
  const bool any = keep && count[i] > 0;
  trect[4 * i + 0] = any ? r.x : 0; trect[4 * i + 1] = any ? r.y : -1;
  trect[4 * i + 2] = any ? r.z : 0; trect[4 * i + 3] = any ? r.w : -1;
}
}  // namespace

extern "C" {

int gsb_debug_projection(GsbContext* c, uint8_t* in_view, float* depth, float* points_xy, float* radius,
                         int32_t* tile_rect, uint32_t* tile_count) {
  if (!c) return GSB_E_INVALID_ARG;
  if (!c->have_frame) return GSB_E_NO_FRAME;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const int64_t n = c->frame_rows;
  if (n == 0) return GSB_OK;
  if (c->frame_projected) {
    // The frame variant of the projection writes no record for Gaussians without tiles and keys them like culled
    // ones; the getter reports every in-view row, so run the debug variant once more for the frame's camera
    // (identical values for the rows the frame did write; the control words it accumulates into are dead by now).
    FrameGeom geom{c->last_cam.width, c->last_cam.height, c->info.tiles_x, c->info.tiles_y};
    GSB_TRY(c->dbg_cov2d.ensure((size_t)n * 16));
    GSB_TRY(c->dbg_conic.ensure((size_t)n * 16));
    GSB_TRY(c->dbg_bbox.ensure((size_t)n * 16));
    DebugOut dbg{c->dbg_cov2d.as<float>(), c->dbg_conic.as<float>(), c->dbg_bbox.as<float>()};
    const CtlLayout L = ctl_layout(geom, n, 0);
    GSB_TRY(c->control.ensure(L.total * 4));
    uint32_t* ctl = c->control.as<uint32_t>();
    GSB_CUDA_TRY((cudaError_t)launch_project(c->planes.as<float>(), n, c->n_pad, c->last_cam, c->last_prm, geom,
                                             c->depth_key.as<uint32_t>(), c->rec.as<float4>(), c->rect.as<ushort4>(),
                                             c->count.as<uint32_t>(), ctl, ctl + L.hist, 0,
                                             reinterpret_cast<int32_t*>(ctl + L.grid), &dbg, 0));
  }
  // staging: u8[n] (padded to 4) | depth | pxy | radius | rect
  const size_t n4 = ((size_t)n + 3) & ~(size_t)3;
  GSB_TRY(c->scratch.ensure(n4 + (size_t)n * 4 * (1 + 2 + 1 + 4)));
  uint8_t* s_v = c->scratch.as<uint8_t>();
  float* s_d = reinterpret_cast<float*>(s_v + n4);
  float* s_xy = s_d + n; float* s_r = s_xy + 2 * n; int32_t* s_t = reinterpret_cast<int32_t*>(s_r + n);
  unpack_projection_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, c->depth_key.as<uint32_t>(), c->rec.as<float4>(),
                                                                c->rect.as<ushort4>(), c->count.as<uint32_t>(), s_v, s_d,
                                                                s_xy, s_r, s_t);
  GSB_CUDA_TRY(cudaGetLastError());
  if (in_view) GSB_CUDA_TRY(cudaMemcpy(in_view, s_v, (size_t)n, cudaMemcpyDefault));
  if (depth) GSB_CUDA_TRY(cudaMemcpy(depth, s_d, (size_t)n * 4, cudaMemcpyDefault));
  if (points_xy) GSB_CUDA_TRY(cudaMemcpy(points_xy, s_xy, (size_t)n * 8, cudaMemcpyDefault));
  if (radius) GSB_CUDA_TRY(cudaMemcpy(radius, s_r, (size_t)n * 4, cudaMemcpyDefault));
  if (tile_rect) GSB_CUDA_TRY(cudaMemcpy(tile_rect, s_t, (size_t)n * 16, cudaMemcpyDefault));
  if (tile_count) GSB_CUDA_TRY(cudaMemcpy(tile_count, c->count.p, (size_t)n * 4, cudaMemcpyDefault));
  return GSB_OK;
}

int gsb_debug_sorted_keys(GsbContext* c, uint64_t* keys, uint32_t* payload) {
  if (!c) return GSB_E_INVALID_ARG;
  if (!c->have_frame) return GSB_E_NO_FRAME;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const size_t k = (size_t)c->info.k_instances;
  if (k == 0) return GSB_OK;
  const uint32_t* sv = c->sorted_in_a ? c->vals_a.as<uint32_t>() : c->vals_b.as<uint32_t>();
  if (keys) {
    if (c->keys_materialized) {
      GSB_CUDA_TRY(cudaMemcpy(keys, c->sorted_in_a ? c->keys_a.p : c->keys_b.p, k * 8, cudaMemcpyDefault));
    } else {
      // SPLIT mode moved only (tile | index) through the tile passes: rebuild tile<<32 | depth bits
      GSB_TRY(c->scratch.ensure(k * 8));
      GSB_CUDA_TRY((cudaError_t)launch_rebuild_keys(c->ranges.as<uint2>(), c->info.tiles_x * c->info.tiles_y, sv,
                                                    c->depth_key.as<uint32_t>(), c->scratch.as<uint64_t>(), 0));
      GSB_CUDA_TRY(cudaMemcpy(keys, c->scratch.p, k * 8, cudaMemcpyDefault));
    }
  }
  if (payload) GSB_CUDA_TRY(cudaMemcpy(payload, sv, k * 4, cudaMemcpyDefault));
  return GSB_OK;
}

int gsb_debug_emitted_keys(GsbContext* c, uint64_t* keys, uint32_t* payload) {
  if (!c) return GSB_E_INVALID_ARG;
  if (!c->have_frame || !c->emitted_valid) return GSB_E_NO_FRAME;  // multi-pass sorts recycle the emit buffer
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const size_t k = (size_t)c->info.k_instances;
  if (k == 0) return GSB_OK;
  if (keys) GSB_CUDA_TRY(cudaMemcpy(keys, c->keys_a.p, k * 8, cudaMemcpyDefault));
  if (payload) GSB_CUDA_TRY(cudaMemcpy(payload, c->vals_a.p, k * 4, cudaMemcpyDefault));
  return GSB_OK;
}

int gsb_debug_tile_ranges(GsbContext* c, uint32_t* ranges) {
  if (!c) return GSB_E_INVALID_ARG;
  if (!c->have_frame) return GSB_E_NO_FRAME;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  const size_t tiles = (size_t)c->info.tiles_x * c->info.tiles_y;
  if (tiles == 0) return GSB_OK;
  if (!ranges) return GSB_E_INVALID_ARG;
  GSB_CUDA_TRY(cudaMemcpy(ranges, c->ranges.p, tiles * 8, cudaMemcpyDefault));
  return GSB_OK;
}

int gsb_stage_times(GsbContext* c, float ms[GSB_NUM_STAGES]) {
  if (!c || !ms) return GSB_E_INVALID_ARG;
  if (!c->have_frame || !c->have_times) return GSB_E_NO_FRAME;
  for (int i = 0; i < GSB_NUM_STAGES; ++i) ms[i] = c->stage_ms[i];
  return GSB_OK;
}

int gsb_sort_pairs_u64(GsbContext* c, int64_t n, uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out,
                       uint32_t* vals_out, int32_t begin_bit, int32_t end_bit, void* stream) {
  if (!c || n < 0 || begin_bit < 0 || end_bit > 64 || end_bit < begin_bit) return GSB_E_INVALID_ARG;
  if (n >= ((int64_t)1 << 32) - 1) return GSB_E_UNSUPPORTED;
  if (n > 0 && (!keys_in || !vals_in || !keys_out || !vals_out)) return GSB_E_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  if (n == 0) return GSB_OK;
  SortPlan plan = make_sort_plan<uint64_t>(n, begin_bit, end_bit);
  const size_t words = (size_t)kCtlHistWords + plan.control_words;
  GSB_TRY(c->control2.ensure(words * 4));
  GSB_CUDA_TRY(cudaMemsetAsync(c->control2.p, 0, words * 4, st));
  uint32_t* hist = c->control2.as<uint32_t>();
  GSB_CUDA_TRY((cudaError_t)launch_key_histogram<uint64_t>(plan, keys_in, hist, st));
  bool in_a = true;
  int launches = 0;
  GSB_CUDA_TRY((cudaError_t)launch_sort<uint64_t>(plan, keys_in, vals_in, keys_in, vals_in, keys_out, vals_out, hist,
                                                  hist + kCtlHistWords, &in_a, &launches, st));
  if (in_a) {  // even number of passes (or none): result sits in the input buffers
    GSB_CUDA_TRY(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    GSB_CUDA_TRY(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  }
  return GSB_OK;
}

}  // extern "C"
