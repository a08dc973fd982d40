// gsb_api.cu -- the C ABI (include/gsb.h): context, scratch management, kernel-chain orchestration.
//
// Frame pipeline (one stream; tile statistics on an auxiliary stream):
//
//   FULL :  project -> scan(index order) -> [K -> host] -> emit -> p x onesweep(K, 64-bit keys) -> composite
//           p = ceil((32 + tile_bits) / 8); the literal pipeline of the north star, kept as the cross-check
//   SPLIT:  project -> 4 x onesweep(N, 32-bit depth keys) -> scan(depth order, super-tile counts)
//           -> emit(one key per SUPER-TILE instance) -> onesweep over the super-tile ids -> expand(per-tile lists)
//           -> composite
//
// Both leave bit-identical per-tile lists: an LSD radix sort orders by the low (depth) digits first, and every tile
// instance of a Gaussian shares those digits, so they can be sorted once per Gaussian BEFORE the expansion; what is
// left is a stable partition by tile, done in two levels (binning.cu).
//
// SPLIT queues the WHOLE frame before the host reads anything data-dependent.  Grids and buffers are sized from
// the context's capacities; tile_stats_kernel, which learns the counts, decides on the device whether they fit
// (ctl[kCtlAbort]) and posts them to a mailbox in mapped pinned host memory.  The host picks them up after queueing
// -- the GPU never waits for it -- and only when a count outgrew its buffer does it grow the buffer and queue the
// tail of the frame again.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "gsb_internal.cuh"

using namespace gsb;

namespace gsb {
int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 148; }
  if (dev >= 0 && dev < 64) cache[dev] = v;
  return v;
}
}  // namespace gsb

namespace {

// device scratch owned by a context; released when the context is deleted (no list of members to keep in sync)
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
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
//   [hdr 16: kCtl* words] [hist 8 x 256: rows 0-3 depth digits, 4-7 tile (FULL) / super-tile (SPLIT) digits]
//   [difference grid (tiles_x+1)*(tiles_y+1)] [super-tile difference grid] [scan status] [depth-sort tickets + status]
constexpr int kCtlHeaderWords = 16;
constexpr int kCtlHistWords = kMaxPasses * kRadix;

struct CtlLayout {
  size_t hist, grid, grid_s, scan, dsort, total;  // word offsets
};
CtlLayout ctl_layout(FrameGeom g, SuperGeom sg, int64_t n_rows, size_t dsort_words) {
  auto up4 = [](size_t w) { return (w + 3) & ~(size_t)3; };  // sections start 16-byte aligned (64-bit status words)
  CtlLayout L;
  L.hist = kCtlHeaderWords;
  L.grid = L.hist + kCtlHistWords;
  L.grid_s = up4(L.grid + (size_t)(g.tiles_x + 1) * (size_t)(g.tiles_y + 1));
  L.scan = up4(L.grid_s + (size_t)(sg.nx + 1) * (size_t)(sg.ny + 1));
  L.dsort = (L.scan + scan_status_words(n_rows) + 31) & ~(size_t)31;  // 128-byte aligned: see make_sort_plan
  L.total = up4(L.dsort + dsort_words);
  return L;
}

// Super-tile shape of SPLIT mode: 2^lw x 2^lh tiles with lw + lh <= 5, so that "which tiles of the super-tile does
// this rect cover" is ONE 32-bit mask (8 x 4 by default: 255 super-tiles at 1080p -> one radix pass over their
// ids; 1 020 at 4K -> two).  lw = lh = 0 switches the second level off (one key per tile instance).
SuperGeom choose_super(FrameGeom g, int lw, int lh) {
  SuperGeom s{lw, lh, 0, 0};
  s.nx = (g.tiles_x + (1 << s.lw) - 1) >> s.lw;
  s.ny = (g.tiles_y + (1 << s.lh) - 1) >> s.lh;
  return s;
}

}  // namespace

struct GsbContext {
  int device = 0;
  int64_t n = 0, n_pad = 0;
  DevBuf planes, staging;
  // per-Gaussian frame data
  DevBuf depth_key, rec, rect, count, offsets, bbox;
  DevBuf dbg_cov2d, dbg_conic, dbg_bbox;
  DevBuf ord_keys_a, ord_keys_b, ord_vals_a, ord_vals_b;
  // per tile instance: FULL mode key/payload ping-pong; vals_a is also SPLIT's per-tile payload
  DevBuf keys_a, keys_b, vals_a, vals_b;
  // per super-tile instance (SPLIT): key ping-pong and the sorted Gaussian indices
  DevBuf ckeys_a, ckeys_b, cvals;
  DevBuf ranges, ranges_s, control, control2;
  DevBuf image, image2, scratch;
  // save_for_backward: per-pixel blended count / final transmittance of the last frame; gradient scratch
  DevBuf aux_t, aux_n, grad2d, grad_stage;
  DevBuf host_stage[2];
  bool allow_keys32 = true;
  int super_lw = 3, super_lh = 2;
  bool frame_projected = false;  // last frame came from render_device: last_cam / last_prm describe its projection
  GsbCamera last_cam{};
  GsbParams last_prm{};
  bool have_saved = false;
  GsbCamera saved_cam{};
  GsbParams saved_prm{};
  uint64_t scene_gen = 0, saved_gen = 0;
  int64_t frame_id = 0, saved_frame_id = 0;
  uint32_t* pinned = nullptr;      // mailbox written by tile_stats_kernel: {M, V, K lo, K hi, seq, abort, Ks lo, Ks hi}
  uint32_t* pinned_dev = nullptr;  // device alias of the mailbox
  uint32_t seq = 0;
  // the frame as ONE CUDA graph (see FrameCapture): instantiated once, updated in place every frame
  bool allow_graph = true;
  cudaGraphExec_t frame_exec = nullptr;
  bool stats_in_graph = false;  // this frame's tile_stats launches run inside the graph on the caller's stream
  cudaStream_t aux = nullptr;   // side stream for work that is off the critical path (tile stats)
  cudaEvent_t ev_fork = nullptr, ev_stats = nullptr;
  // asynchronous image egress: two device staging images, a copy stream, one event per staging image
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_rendered = nullptr, ev_copied[2] = {nullptr, nullptr};
  bool copy_pending[2] = {false, false};
  int stage_next = 0;
  // state of the last frame
  bool have_frame = false;
  bool sorted_in_a = true;
  bool emitted_valid = false;
  bool keys_materialized = true;  // false in SPLIT mode: sorted 64-bit keys are rebuilt on demand (debug)
  bool lists_materialized = true; // false after a two-level SPLIT frame: per-tile lists exist only as super-tile lists
  SuperGeom last_sg{0, 0, 0, 0};
  bool order_in_a = true;
  bool have_order = false;
  int64_t frame_rows = 0;  // rows of the per-Gaussian arrays of the last frame (N, or M for gsb_render_image)
  int tail_reruns = 0;     // frames whose tail had to be queued twice (a count outgrew its buffer)
  int64_t ks_hint = 0;     // super-tile instances of the last frame of this scene (0: none yet)
  GsbFrameInfo info{};
  cudaEvent_t ev[GSB_NUM_STAGES + 9]{};
  int mark_stage[GSB_NUM_STAGES + 9]{};
  int n_marks = 0;
  float stage_ms[GSB_NUM_STAGES]{};
  bool have_times = false;
};

namespace {

constexpr int kMaxMarks = GSB_NUM_STAGES + 8;

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
    if (!on || c->n_marks >= kMaxMarks) return;
    ++c->n_marks;
    c->mark_stage[c->n_marks] = stage;
    cudaEventRecord(c->ev[c->n_marks], st);
  }
};

// The frame as one CUDA graph.  A frame is 11 kernels + 2 memsets + an event fork / join, all queued before the host
// reads anything (section "queue first" of DESIGN.md): captured from the very same launch code (relaxed stream
// capture, the auxiliary stream joins the capture through its events), then cudaGraphExecUpdate writes this frame's
// kernel arguments (camera, mailbox sequence number, image pointer, buffer addresses) into the graph instantiated by
// the first frame, and ONE cudaGraphLaunch submits it.  Why: each stream command is fetched by the GPU from host
// memory, and while image copies saturate PCIe every such fetch waits ~35 us behind them (measured,
// tools/e2e_probe.py) -- eleven of them per frame cost the end-to-end figure 14 %.  Not used on the legacy default
// stream (capture is not allowed there), inside a caller's own capture, or with stage timing on.
struct FrameCapture {
  GsbContext* c;
  cudaStream_t st;
  bool active = false;
  int begin() {
    GSB_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    active = true;
    c->stats_in_graph = true;
    return GSB_OK;
  }
  // end of the queued frame: instantiate / update, launch.  No-op when nothing is being captured.
  int submit() {
    if (!active) return GSB_OK;
    active = false;
    cudaGraph_t g = nullptr;
    GSB_CUDA_TRY(cudaStreamEndCapture(st, &g));
    if (c->frame_exec) {
      cudaGraphExecUpdateResultInfo why;
      if (cudaGraphExecUpdate(c->frame_exec, g, &why) != cudaSuccess) {  // another topology (pass count, key width, ...)
        cudaGetLastError();
        cudaGraphExecDestroy(c->frame_exec);
        c->frame_exec = nullptr;
      }
    }
    cudaError_t e = cudaSuccess;
    if (!c->frame_exec) e = cudaGraphInstantiate(&c->frame_exec, g, 0);
    cudaGraphDestroy(g);
    GSB_CUDA_TRY(e);
    GSB_CUDA_TRY(cudaGraphLaunch(c->frame_exec, st));
    c->info.graph_launch = 1;
    return GSB_OK;
  }
  ~FrameCapture() {  // an error path left the stream capturing: drop what was captured, nothing of it ran
    if (!active) return;
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(st, &g);
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    c->stats_in_graph = false;
  }
};

// A caller's own capture cannot contain a frame: the host waits for the frame's counts, which a captured frame never
// produces.  Refuse instead of spinning on the mailbox.
bool caller_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
}

// may this frame go out as a graph on `st`?
bool graph_allowed(GsbContext* c, cudaStream_t st, const GsbParams* prm) {
  if (!c->allow_graph || prm->collect_stage_times) return false;
  if (st == nullptr || st == cudaStreamLegacy) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs == cudaStreamCaptureStatusNone;
}

int check_params(const GsbCamera* cam, const GsbParams* prm) {
  if (!cam || !prm) return GSB_E_INVALID_ARG;
  if (prm->tile_size != kTile) return GSB_E_UNSUPPORTED;
  if (cam->width <= 0 || cam->height <= 0) return GSB_E_INVALID_ARG;
  if (cam->width > 65535 * kTile || cam->height > 65535 * kTile) return GSB_E_UNSUPPORTED;
  if (prm->semantics != GSB_SEM_REF_CPU && prm->semantics != GSB_SEM_REF_CU) return GSB_E_INVALID_ARG;
  if (prm->sort_mode < GSB_SORT_AUTO || prm->sort_mode > GSB_SORT_SPLIT) return GSB_E_INVALID_ARG;
  if (prm->cull_alpha != prm->cull_alpha) return GSB_E_INVALID_ARG;
  return GSB_OK;
}

// every entry point that overwrites per-frame state starts a new frame: a frame saved for the backward pass is gone
void begin_frame(GsbContext* c) {
  c->have_frame = false;
  c->frame_projected = false;
  c->have_order = false;
  c->have_saved = false;
  c->have_times = false;
  c->stats_in_graph = false;
  ++c->frame_id;
  std::memset(&c->info, 0, sizeof(c->info));
  c->info.frame_id = c->frame_id;
}

struct Counts {
  int64_t m = 0, v = 0, k = 0, ks = 0;
  bool abort = false;
};

// tile_stats_kernel needs only the projection's difference grids, not the depth sort: run it on the context's
// auxiliary stream so that it overlaps the (latency-bound) depth-sort passes; the main stream joins on ev_stats.
// `sg` with lw + lh > 0 adds the launch over the super-tile grid; the last launch decides abort and posts the mailbox.
int launch_stats_async(GsbContext* c, FrameGeom geom, SuperGeom sg, bool split, uint32_t* ctl, const CtlLayout& L,
                       uint64_t cap_k, uint64_t cap_ks, cudaStream_t st, int* launches) {
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  GSB_TRY(c->ranges.ensure((size_t)(tiles > 0 ? tiles : 1) * 8));
  if (tiles <= 0) return GSB_OK;
  const bool two_level = split && sg.lw + sg.lh > 0;
  if (two_level) GSB_TRY(c->ranges_s.ensure((size_t)sg.nx * sg.ny * 8));
  GSB_CUDA_TRY(cudaEventRecord(c->ev_fork, st));
  GSB_CUDA_TRY(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
  ++c->seq;
  uint32_t* tile_hist = ctl + L.hist + 4 * kRadix;
  StatsPost post{0, two_level ? 0 : 1, cap_k, cap_ks, c->pinned_dev, c->seq};
  GSB_CUDA_TRY((cudaError_t)launch_tile_stats(reinterpret_cast<int32_t*>(ctl + L.grid), geom.tiles_x, geom.tiles_y,
                                              two_level ? nullptr : tile_hist, c->ranges.as<uint2>(), ctl, post, c->aux));
  ++*launches;
  if (two_level) {
    post.level = 1;
    post.enabled = 1;
    GSB_CUDA_TRY((cudaError_t)launch_tile_stats(reinterpret_cast<int32_t*>(ctl + L.grid_s), sg.nx, sg.ny, tile_hist,
                                                c->ranges_s.as<uint2>(), ctl, post, c->aux));
    ++*launches;
  }
  GSB_CUDA_TRY(cudaEventRecord(c->ev_stats, c->aux));
  return GSB_OK;
}

// Pick up the counts tile_stats_kernel posted, without draining any stream.
int wait_counts(GsbContext* c, int64_t tiles, const uint32_t* ctl, cudaStream_t st, Counts* out) {
  *out = Counts{};
  if (tiles <= 0) {  // no tile grid, no tile_stats launch: plain read-back of M
    GSB_CUDA_TRY(cudaMemcpyAsync(c->pinned, ctl, 4, cudaMemcpyDeviceToHost, st));
    GSB_CUDA_TRY(cudaStreamSynchronize(st));
    out->m = c->pinned[0];
    return GSB_OK;
  }
  const uint32_t* box = c->pinned;
  const auto t0 = std::chrono::steady_clock::now();
  // acquire: the counts below must not be read before the sequence number (weakly ordered hosts)
  for (uint64_t spins = 0; __atomic_load_n(&box[4], __ATOMIC_ACQUIRE) != c->seq; ++spins) {
    if ((spins & 0x3FFF) == 0x3FFF) {
      cudaError_t q = cudaStreamQuery(c->stats_in_graph ? st : c->aux);  // where this frame's tile_stats run
      if (q != cudaSuccess && q != cudaErrorNotReady) return (int)q;
      if (q == cudaSuccess && __atomic_load_n(&box[4], __ATOMIC_ACQUIRE) != c->seq) return GSB_E_INTERNAL;  // kernel finished, mailbox never written
      // a stream that never runs (e.g. waiting on an event nobody records) must not hang the caller for ever
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return GSB_E_INTERNAL;
    }
  }
  auto ld = [&](int i) { return (uint64_t)__atomic_load_n(&box[i], __ATOMIC_RELAXED); };
  out->m = (int64_t)ld(0);
  out->v = (int64_t)ld(1);
  out->k = (int64_t)((ld(3) << 32) | ld(2));
  out->abort = ld(5) != 0;
  out->ks = (int64_t)((ld(7) << 32) | ld(6));
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

// ---- FULL mode: scan (index order) -> K to the host -> emit -> 64-bit sort.  The host needs K to size the grids,
// so this path (the literal north-star pipeline, kept as the cross-check of SPLIT) waits for the mailbox mid-frame.
int bin_full(GsbContext* c, int64_t n_rows, FrameGeom geom, uint32_t* ctl, const CtlLayout& L, cudaStream_t st,
             StageTimer& tm, int* launches) {
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  GSB_TRY(c->offsets.ensure((size_t)n_rows * 4 + 4));
  GSB_CUDA_TRY((cudaError_t)launch_scan(c->count.as<uint32_t>(), nullptr, n_rows, c->offsets.as<uint32_t>(), ctl + L.scan, st));
  if (n_rows > 0) ++*launches;
  tm.mark(GSB_STAGE_SCAN);
  Counts cn;
  GSB_TRY(wait_counts(c, tiles, ctl, st, &cn));
  if (tiles > 0) GSB_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_stats, 0));  // tile histograms are inputs of the sort
  if (cn.k >= ((int64_t)1 << 32) - 1) return GSB_E_UNSUPPORTED;  // payload positions and ranges are u32
  c->info.m_in_view = cn.m;
  c->info.v_with_tiles = cn.v;
  c->info.k_instances = cn.k;
  c->info.k_sorted = cn.k;
  const int64_t k = cn.k;
  GSB_TRY(c->keys_a.ensure((size_t)k * 8 + 8));
  GSB_TRY(c->keys_b.ensure((size_t)k * 8 + 8));
  GSB_TRY(c->vals_a.ensure((size_t)k * 4 + 4));
  GSB_TRY(c->vals_b.ensure((size_t)k * 4 + 4));
  const int tile_bits = ceil_log2(tiles);
  SortPlan plan = make_sort_plan<uint64_t>(k, 0, 32 + tile_bits);
  GSB_TRY(c->control2.ensure(plan.control_words * 4));
  GSB_CUDA_TRY(cudaMemsetAsync(c->control2.p, 0, plan.control_words * 4, st));
  GSB_CUDA_TRY((cudaError_t)launch_emit(c->offsets.as<uint32_t>(), nullptr, ctl + kCtlK, n_rows, k,
                                        c->depth_key.as<uint32_t>(), c->rect.as<ushort4>(), geom.tiles_x,
                                        c->keys_a.as<uint64_t>(), c->vals_a.as<uint32_t>(), st));
  if (n_rows > 0 && k > 0) ++*launches;
  tm.mark(GSB_STAGE_EMIT);
  bool in_a = true;
  GSB_CUDA_TRY((cudaError_t)launch_sort<uint64_t>(plan, c->keys_a.as<uint64_t>(), c->vals_a.as<uint32_t>(),
                                                  c->keys_a.as<uint64_t>(), c->vals_a.as<uint32_t>(),
                                                  c->keys_b.as<uint64_t>(), c->vals_b.as<uint32_t>(), ctl + L.hist,
                                                  c->control2.as<uint32_t>(), &in_a, launches, st));
  c->keys_materialized = true;
  c->sorted_in_a = in_a;
  c->emitted_valid = plan.passes <= 1;  // one pass: the a-buffers still hold the emitted order
  c->info.sort_passes = (k > 0) ? plan.passes : 0;
  c->info.key_bits = 64;
  c->info.super_w = c->info.super_h = 1;
  tm.mark(GSB_STAGE_SORT);
  return GSB_OK;
}

// ---- SPLIT mode, everything between the scan and the compositing kernel, queued from capacities.
// Rows are in emission order (`perm`, or identity) with their depth digits already sorted; `v_limit` (optional):
// device word holding the number of leading emission positions that touch tiles.
struct SplitPlan {
  SuperGeom sg;
  bool two_level, keys32;
  // two_level frames store per-tile lists only when somebody reads them as arrays: the gradient pass (the
  // compositing kernel then writes what it consumes: payload_side) or the REF_CU kernel (expand_in_stream)
  bool payload_side, expand_in_stream;
  int rank_bits, sbits;
  int64_t cap_k, cap_ks;
  bool need_payload() const { return !two_level || payload_side || expand_in_stream; }
};

int split_capacities(GsbContext* c, int64_t n_rows, const SplitPlan& sp, int64_t need_k, int64_t need_ks) {
  // payload: 4 B per tile instance (if stored at all); super-tile level: one key ping-pong + 8 B list entries
  const size_t kb = sp.keys32 ? 4 : 8;
  if (need_k < 16 * n_rows) need_k = 16 * n_rows;  // first guess of a fresh context
  if (!sp.two_level) need_ks = need_k;
  else if (need_ks < 4 * n_rows) need_ks = 4 * n_rows;
  if (sp.need_payload()) GSB_TRY(c->vals_a.ensure((size_t)need_k * 4 + 16));
  GSB_TRY(c->ckeys_a.ensure((size_t)need_ks * kb + 16));
  GSB_TRY(c->ckeys_b.ensure((size_t)need_ks * kb + 16));
  if (sp.two_level) GSB_TRY(c->cvals.ensure((size_t)need_ks * 8 + 16));
  return GSB_OK;
}

SplitPlan make_split_plan(GsbContext* c, int64_t n_rows, SuperGeom sg, bool payload_side, bool expand_in_stream) {
  SplitPlan sp;
  sp.sg = sg;
  sp.two_level = sg.lw + sg.lh > 0;
  sp.payload_side = sp.two_level && payload_side;
  sp.expand_in_stream = sp.two_level && expand_in_stream;
  sp.sbits = ceil_log2((int64_t)sg.nx * sg.ny);
  sp.rank_bits = ceil_log2(n_rows);
  sp.keys32 = c->allow_keys32 && sp.sbits + sp.rank_bits <= 32;
  sp.cap_k = sp.cap_ks = 0;
  return sp;
}

void read_capacities(GsbContext* c, SplitPlan& sp) {
  const size_t kb = sp.keys32 ? 4 : 8;
  const int64_t lim = ((int64_t)1 << 32) - 2;  // positions are u32
  int64_t ck = sp.need_payload() ? (int64_t)((c->vals_a.cap - 16) / 4) : lim;
  int64_t cks = (int64_t)((c->ckeys_a.cap - 16) / kb);
  const int64_t cks_b = (int64_t)((c->ckeys_b.cap - 16) / kb);
  if (cks_b < cks) cks = cks_b;
  if (sp.two_level) {
    const int64_t cv = (int64_t)((c->cvals.cap - 16) / 8);
    if (cv < cks) cks = cv;
  } else {
    if (ck < cks) cks = ck; else ck = cks;
  }
  sp.cap_k = ck < lim ? ck : lim;
  sp.cap_ks = cks < lim ? cks : lim;
}

// Keys per thread of the super-tile radix passes.  Around a million keys (config 3) the pass is a few hundred tiles
// whose last pass gathers two records per key: 8 keys per thread (twice the CTAs) run it in 22.5 us instead of 28-32;
// from a few million keys on 16 is as good or better (config 4: 45 vs 45 us) and halves the status words.  The count
// is not known when the frame is queued: the last frame's stands in for it.
int tile_sort_items(const GsbContext* c) { return (c->ks_hint > 0 && c->ks_hint <= (int64_t)2 << 20) ? 8 : 16; }

int queue_split_tail(GsbContext* c, int64_t n_rows, const uint32_t* perm, const uint32_t* v_limit, FrameGeom geom,
                     const SplitPlan& sp, uint32_t* ctl, const CtlLayout& L, cudaStream_t st, StageTimer& tm,
                     int* launches) {
  const uint32_t* abort = ctl + kCtlAbort;
  const uint32_t* hist = ctl + L.hist + 4 * kRadix;
  uint32_t* out_vals = sp.two_level ? c->cvals.as<uint32_t>() : c->vals_a.as<uint32_t>();
  bool in_a = true;
  int passes = 0;
  if (sp.keys32) {
    SortPlan plan = make_sort_plan<uint32_t>(sp.cap_ks, sp.rank_bits, sp.rank_bits + sp.sbits, tile_sort_items(c));
    plan.keys_only = 1;
    plan.low_bits = sp.rank_bits;
    plan.gather_table = perm;  // nullptr (pre-sorted rows): the position IS the row
    plan.n_dev = ctl + kCtlKs;
    plan.abort = abort;
    if (sp.two_level) { plan.entry_rect = c->rect.as<ushort4>(); plan.entry_lw = sp.sg.lw; plan.entry_lh = sp.sg.lh; plan.entry_snx = sp.sg.nx; }
    GSB_TRY(c->control2.ensure(plan.control_words * 4));
    GSB_CUDA_TRY(cudaMemsetAsync(c->control2.p, 0, plan.control_words * 4, st));
    GSB_CUDA_TRY((cudaError_t)launch_emit_coarse(c->offsets.as<uint32_t>(), perm, ctl + kCtlKs, n_rows, v_limit, abort,
                                                 sp.cap_ks, c->rect.as<ushort4>(), sp.sg, sp.rank_bits, c->ckeys_a.p, st));
    ++*launches;
    tm.mark(GSB_STAGE_EMIT);
    GSB_CUDA_TRY((cudaError_t)launch_sort<uint32_t>(plan, c->ckeys_a.as<uint32_t>(), nullptr, c->ckeys_a.as<uint32_t>(),
                                                    out_vals, c->ckeys_b.as<uint32_t>(), out_vals, hist,
                                                    c->control2.as<uint32_t>(), &in_a, launches, st));
    passes = plan.passes;
  } else {
    SortPlan plan = make_sort_plan<uint64_t>(sp.cap_ks, 32, 32 + sp.sbits, tile_sort_items(c));
    plan.keys_only = 1;
    plan.n_dev = ctl + kCtlKs;
    plan.abort = abort;
    if (sp.two_level) { plan.entry_rect = c->rect.as<ushort4>(); plan.entry_lw = sp.sg.lw; plan.entry_lh = sp.sg.lh; plan.entry_snx = sp.sg.nx; }
    GSB_TRY(c->control2.ensure(plan.control_words * 4));
    GSB_CUDA_TRY(cudaMemsetAsync(c->control2.p, 0, plan.control_words * 4, st));
    GSB_CUDA_TRY((cudaError_t)launch_emit_coarse(c->offsets.as<uint32_t>(), perm, ctl + kCtlKs, n_rows, v_limit, abort,
                                                 sp.cap_ks, c->rect.as<ushort4>(), sp.sg, 0, c->ckeys_a.p, st));
    ++*launches;
    tm.mark(GSB_STAGE_EMIT);
    GSB_CUDA_TRY((cudaError_t)launch_sort<uint64_t>(plan, c->ckeys_a.as<uint64_t>(), nullptr, c->ckeys_a.as<uint64_t>(),
                                                    out_vals, c->ckeys_b.as<uint64_t>(), out_vals, hist,
                                                    c->control2.as<uint32_t>(), &in_a, launches, st));
    passes = plan.passes;
  }
  tm.mark(GSB_STAGE_SORT);
  if (sp.expand_in_stream) {
    GSB_CUDA_TRY((cudaError_t)launch_expand(c->ranges_s.as<uint2>(), c->cvals.as<uint2>(), c->ranges.as<uint2>(),
                                            c->vals_a.as<uint32_t>(), geom, sp.sg, abort, st));
    ++*launches;
    tm.mark(GSB_STAGE_EXPAND);
  }
  c->keys_materialized = false;
  c->lists_materialized = !sp.two_level || sp.expand_in_stream;
  c->last_sg = sp.sg;
  c->sorted_in_a = true;
  c->emitted_valid = false;
  c->info.sort_passes = passes;
  c->info.key_bits = sp.keys32 ? 32 : 64;
  c->info.super_w = 1 << sp.sg.lw;
  c->info.super_h = 1 << sp.sg.lh;
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

// What follows the per-Gaussian records in SPLIT mode, shared by gsb_render (projected rows, depth order in `perm`)
// and gsb_render_image (the caller's pre-sorted rows): scan -> [emit, sort, expand, composite] queued from
// capacities -> counts from the mailbox -> on overflow grow and queue the bracketed tail once more.
TileSource tile_source(GsbContext* c, const SplitPlan* sp) {
  TileSource src{c->ranges.as<uint2>(), c->vals_a.as<uint32_t>(), nullptr, nullptr, nullptr, 1, 0, 0};
  if (sp && sp->two_level && !sp->expand_in_stream) {
    src.payload = nullptr;
    src.ranges_s = c->ranges_s.as<uint2>();
    src.clist = c->cvals.as<uint2>();
    src.payload_out = sp->payload_side ? c->vals_a.as<uint32_t>() : nullptr;
    src.snx = sp->sg.nx; src.lw = sp->sg.lw; src.lh = sp->sg.lh;
  }
  return src;
}

template <typename CompositeFn>
int run_split(GsbContext* c, int64_t n_rows, const uint32_t* perm, bool rows_sorted_by_visibility, FrameGeom geom,
              SuperGeom sg, bool payload_side, bool expand_in_stream, uint32_t* ctl, const CtlLayout& L, cudaStream_t st,
              StageTimer& tm, int* launches, FrameCapture* cap, CompositeFn&& queue_composite) {
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  if (tiles <= 0 || n_rows <= 0) {  // nothing to bin; the compositing launcher handles an empty grid itself
    Counts cn;
    if (n_rows > 0) GSB_TRY(wait_counts(c, 0, ctl, st, &cn));
    c->info.m_in_view = cn.m;
    c->info.super_w = c->info.super_h = 1;
    if (tiles > 0) {  // no rows: every range is (0,0)
      GSB_TRY(c->ranges.ensure((size_t)tiles * 8));
      GSB_CUDA_TRY(cudaMemsetAsync(c->ranges.p, 0, (size_t)tiles * 8, st));
      GSB_TRY(c->vals_a.ensure(16));
    }
    c->keys_materialized = false;
    c->lists_materialized = true;
    c->sorted_in_a = true;
    c->emitted_valid = false;
    return queue_composite(tile_source(c, nullptr), nullptr);
  }
  const uint32_t* v_limit = rows_sorted_by_visibility ? ctl + kCtlVisible : nullptr;
  GSB_TRY(c->offsets.ensure((size_t)n_rows * 4 + 4));
  GSB_CUDA_TRY((cudaError_t)launch_scan_coarse(c->rect.as<ushort4>(), sg, perm, n_rows, v_limit, c->offsets.as<uint32_t>(),
                                               ctl + L.scan, st));
  ++*launches;
  tm.mark(GSB_STAGE_SCAN);
  GSB_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_stats, 0));  // ranges / super-tile histograms are inputs of what follows
  SplitPlan sp = make_split_plan(c, n_rows, sg, payload_side, expand_in_stream);
  read_capacities(c, sp);  // the capacities tile_stats_kernel was given (launch_stats_async ran with the same values)
  GSB_TRY(queue_split_tail(c, n_rows, perm, v_limit, geom, sp, ctl, L, st, tm, launches));
  GSB_TRY(queue_composite(tile_source(c, &sp), ctl + kCtlAbort));
  if (cap) GSB_TRY(cap->submit());  // the whole frame is queued: as one graph launch when it was captured
  Counts cn;
  GSB_TRY(wait_counts(c, tiles, ctl, st, &cn));
  if (cn.k >= ((int64_t)1 << 32) - 1 || cn.ks >= ((int64_t)1 << 32) - 1) return GSB_E_UNSUPPORTED;  // positions are u32
  c->info.m_in_view = cn.m;
  c->info.v_with_tiles = rows_sorted_by_visibility ? cn.v : 0;
  c->info.k_instances = cn.k;
  c->info.k_sorted = cn.ks;
  c->ks_hint = cn.ks;
  if (cn.abort) {
    c->info.tail_requeued = 1;
    // a count outgrew its buffer: none of the tail kernels touched anything.  Grow (cudaFree drains the device),
    // clear the flag and queue the tail once more; the projection, depth order, ranges and offsets stand.
    ++c->tail_reruns;
    GSB_TRY(split_capacities(c, n_rows, sp, cn.k, cn.ks));
    read_capacities(c, sp);
    if (cn.k > sp.cap_k || cn.ks > sp.cap_ks) return GSB_E_INTERNAL;
    GSB_CUDA_TRY(cudaMemsetAsync(ctl + kCtlAbort, 0, 4, st));
    GSB_TRY(queue_split_tail(c, n_rows, perm, v_limit, geom, sp, ctl, L, st, tm, launches));
    GSB_TRY(queue_composite(tile_source(c, &sp), ctl + kCtlAbort));
  }
  return GSB_OK;
}

// capacities the stats kernel is told for this frame (must equal what run_split reads back: both derive them from
// the same buffers, and nothing reallocates in between)
int prepare_split(GsbContext* c, int64_t n_rows, SuperGeom sg, bool payload_side, bool expand_in_stream, uint64_t* cap_k,
                  uint64_t* cap_ks) {
  SplitPlan sp = make_split_plan(c, n_rows, sg, payload_side, expand_in_stream);
  GSB_TRY(split_capacities(c, n_rows, sp, 0, 0));  // no-op once the buffers exist
  read_capacities(c, sp);
  *cap_k = (uint64_t)sp.cap_k;
  *cap_ks = (uint64_t)sp.cap_ks;
  return GSB_OK;
}

// projection + binning + sort + compositing into a DEVICE image buffer
int render_device(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, float* dev_image, cudaStream_t st) {
  GSB_TRY(check_params(cam, prm));
  if (c->n_pad == 0 && c->n != 0) return GSB_E_NO_SCENE;
  if (!c->planes.p && c->n > 0) return GSB_E_NO_SCENE;
  if (prm->semantics != GSB_SEM_REF_CPU) return GSB_E_UNSUPPORTED;  // REF_CU is served by gsb_render_image
  GSB_CUDA_TRY(cudaSetDevice(c->device));
  if (caller_is_capturing(st)) return GSB_E_UNSUPPORTED;
  const int64_t n = c->n;
  FrameGeom geom{cam->width, cam->height, tile_grid_dim(cam->width, kTile, prm->full_cover),
                 tile_grid_dim(cam->height, kTile, prm->full_cover)};
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  const bool split = prm->sort_mode != GSB_SORT_FULL;  // AUTO = SPLIT
  const SuperGeom sg = split ? choose_super(geom, c->super_lw, c->super_lh) : SuperGeom{0, 0, geom.tiles_x, geom.tiles_y};
  const bool two_level = split && sg.lw + sg.lh > 0;
  int launches = 0;
  begin_frame(c);
  c->info.n = n; c->info.tiles_x = geom.tiles_x; c->info.tiles_y = geom.tiles_y;
  StageTimer tm{c, st, prm->collect_stage_times != 0};

  const size_t rows = (size_t)(n > 0 ? n : 1);
  GSB_TRY(c->depth_key.ensure(rows * 4));
  GSB_TRY(c->rec.ensure(rows * 48));
  GSB_TRY(c->rect.ensure(rows * 8));
  GSB_TRY(c->count.ensure(rows * 4));
  SortPlan dplan = make_sort_plan<uint32_t>(n, 0, 32);
  const CtlLayout L = ctl_layout(geom, sg, n, split ? dplan.control_words : 0);
  GSB_TRY(c->control.ensure(L.total * 4));
  uint64_t cap_k = ~0ull, cap_ks = ~0ull;  // FULL: the host sizes everything after it has seen K
  const bool payload_side = prm->save_for_backward != 0;  // the gradient pass walks per-tile lists as arrays
  if (split && n > 0 && tiles > 0) GSB_TRY(prepare_split(c, n, sg, payload_side, false, &cap_k, &cap_ks));
  FrameCapture cap{c, st};
  if (split && n > 0 && tiles > 0 && graph_allowed(c, st, prm)) {
    // buffers the captured section would otherwise allocate on its way (an allocation is legal in a relaxed
    // capture, a cudaFree of a grown buffer drains the device: keep both out of it)
    GSB_TRY(c->ord_keys_a.ensure((size_t)n * 4)); GSB_TRY(c->ord_keys_b.ensure((size_t)n * 4));
    GSB_TRY(c->ord_vals_a.ensure((size_t)n * 4)); GSB_TRY(c->ord_vals_b.ensure((size_t)n * 4));
    GSB_TRY(c->offsets.ensure((size_t)n * 4 + 4));
    GSB_TRY(c->ranges.ensure((size_t)tiles * 8));
    if (two_level) GSB_TRY(c->ranges_s.ensure((size_t)sg.nx * sg.ny * 8));
    if (prm->save_for_backward) {
      GSB_TRY(c->aux_t.ensure((size_t)cam->width * cam->height * 4));
      GSB_TRY(c->aux_n.ensure((size_t)cam->width * cam->height * 4));
    }
    GSB_TRY(cap.begin());
  }
  GSB_CUDA_TRY(cudaMemsetAsync(c->control.p, 0, L.total * 4, st));
  uint32_t* ctl = c->control.as<uint32_t>();

  tm.start();
  GSB_CUDA_TRY((cudaError_t)launch_project(c->planes.as<float>(), n, c->n_pad, *cam, *prm, geom, c->depth_key.as<uint32_t>(),
                                           c->rec.as<float4>(), c->rect.as<ushort4>(), c->count.as<uint32_t>(), ctl,
                                           ctl + L.hist, /*hist_weighted=*/split ? 0 : 1,
                                           reinterpret_cast<int32_t*>(ctl + L.grid),
                                           two_level ? reinterpret_cast<int32_t*>(ctl + L.grid_s) : nullptr, sg, nullptr, st));
  if (n > 0) ++launches;
  tm.mark(GSB_STAGE_PROJECT);
  if (n > 0) GSB_TRY(launch_stats_async(c, geom, sg, split, ctl, L, cap_k, cap_ks, st, &launches));

  float* aux_t = nullptr;
  uint32_t* aux_n = nullptr;
  if (prm->save_for_backward) {
    const size_t px = (size_t)cam->width * cam->height;
    GSB_TRY(c->aux_t.ensure(px * 4));
    GSB_TRY(c->aux_n.ensure(px * 4));
    aux_t = c->aux_t.as<float>();
    aux_n = c->aux_n.as<uint32_t>();
  }
  auto queue_composite = [&](const TileSource& src, const uint32_t* abort) -> int {
    if (!prm->full_cover) {  // pixels outside the reference tile grid stay 0 (splat/gaussian_scene.py:206)
      GSB_CUDA_TRY(cudaMemsetAsync(dev_image, 0, (size_t)cam->width * cam->height * 3 * sizeof(float), st));
      if (aux_n) GSB_CUDA_TRY(cudaMemsetAsync(aux_n, 0, (size_t)cam->width * cam->height * 4, st));  // nothing blended there
    }
    GSB_CUDA_TRY((cudaError_t)launch_composite(src, c->rec.as<float4>(), dev_image, geom, *prm, aux_t, aux_n, abort, st));
    if (tiles > 0) ++launches;
    tm.mark(GSB_STAGE_COMPOSITE);
    return GSB_OK;
  };

  if (split) {
    const uint32_t* perm = nullptr;
    if (n > 0) {
      int dp = 0;
      GSB_TRY(depth_sort(c, n, ctl + L.hist, ctl + L.dsort, st, &launches, &dp));
      c->info.depth_passes = dp;
      perm = c->order_in_a ? c->ord_vals_a.as<uint32_t>() : c->ord_vals_b.as<uint32_t>();
      tm.mark(GSB_STAGE_DEPTH_SORT);
    }
    GSB_TRY(run_split(c, n, perm, /*rows_sorted_by_visibility=*/true, geom, sg, payload_side, false, ctl, L, st, tm,
                      &launches, &cap, queue_composite));
  } else {
    if (n > 0 && tiles > 0) {
      GSB_TRY(bin_full(c, n, geom, ctl, L, st, tm, &launches));
    } else {
      Counts cn;
      if (n > 0) GSB_TRY(wait_counts(c, 0, ctl, st, &cn));
      c->info.m_in_view = cn.m;
      if (tiles > 0) {
        GSB_TRY(c->ranges.ensure((size_t)tiles * 8));
        GSB_CUDA_TRY(cudaMemsetAsync(c->ranges.p, 0, (size_t)tiles * 8, st));
        GSB_TRY(c->vals_a.ensure(16));
      }
      c->sorted_in_a = true;
      c->keys_materialized = false;
      c->emitted_valid = false;
    }
    c->lists_materialized = true;
    if (!c->sorted_in_a) {  // the compositing launcher above reads vals_a: an odd pass count left the order in vals_b
      std::swap(c->vals_a.p, c->vals_b.p); std::swap(c->vals_a.cap, c->vals_b.cap);
      std::swap(c->keys_a.p, c->keys_b.p); std::swap(c->keys_a.cap, c->keys_b.cap);
      c->sorted_in_a = true;
      c->emitted_valid = false;
    }
    GSB_TRY(queue_composite(tile_source(c, nullptr), nullptr));
  }
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
    c->saved_frame_id = c->frame_id;
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
    case GSB_E_UNSUPPORTED: return "unsupported configuration (tile_size must be 16; fewer than 2^32-1 tile instances; no render calls inside a caller's stream capture)";
    case GSB_E_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case GSB_E_ALLOC: return "device memory allocation failed";
    case GSB_E_INTERNAL: return "internal error (device-side consistency check failed, or the stream made no progress)";
    case GSB_E_NO_SAVED: return "no saved frame for the backward pass (render with save_for_backward = 1 and the same camera / params first; any other render, preprocess or upload on the context discards it)";
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
  p->cull_alpha = 9.31322574615478515625e-10f;  // 2^-30
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
    const char* e = std::getenv("GSB_SORT_ITEMS");        // onesweep keys per thread, all sorts: 16 or 8 (unset: per sort)
    set_sort_items(e ? std::atoi(e) : 0);
    e = std::getenv("GSB_FORCE_WIDE_STATUS");             // 1: 64-bit look-back words even below 2^30 keys
    set_force_wide_status(e ? std::atoi(e) : 0);
    e = std::getenv("GSB_KEYS32");                        // 0: never use 32-bit keys for the super-tile passes
    c->allow_keys32 = !e || std::atoi(e) != 0;
    e = std::getenv("GSB_GRAPH");                         // 0: queue every frame launch by launch
    c->allow_graph = !e || std::atoi(e) != 0;
    e = std::getenv("GSB_SUPER");                         // "lw,lh": log2 tiles per super-tile; "0,0": one level
    if (e) {
      int lw = 3, lh = 2;
      if (std::sscanf(e, "%d,%d", &lw, &lh) == 2 && lw >= 0 && lh >= 0 && lw + lh <= 5) {
        c->super_lw = lw; c->super_lh = lh;
      }
    }
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
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->frame_exec) cudaGraphExecDestroy(c->frame_exec);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->copy) cudaStreamDestroy(c->copy);
  if (c->ev_rendered) cudaEventDestroy(c->ev_rendered);
  for (auto& e : c->ev_copied) if (e) cudaEventDestroy(e);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_stats) cudaEventDestroy(c->ev_stats);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  delete c;  // every DevBuf member frees its device memory in its destructor
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
  c->ks_hint = 0;
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

int gsb_render_backward(GsbContext* c, const GsbCamera* cam, const GsbParams* prm, int64_t frame_id,
                        const float* grad_image, float* grad_points, float* grad_scales, float* grad_quats,
                        float* grad_colors, float* grad_opacity, void* stream) {
  if (!c || !grad_image) return GSB_E_INVALID_ARG;
  GSB_TRY(check_params(cam, prm));
  // the saved state IS the context's per-frame scratch: valid only while the saved frame is still the last thing
  // that wrote it (begin_frame / gsb_upload / the debug projection clear have_saved)
  if (!c->have_frame || !c->have_saved || c->saved_gen != c->scene_gen || c->saved_frame_id != c->frame_id ||
      (frame_id != 0 && frame_id != c->saved_frame_id) ||
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
  if (is_device_pointer(out)) {
    GSB_TRY(render_device(c, cam, prm, c->image.as<float>(), st));
    return launch_to_u8(c->image.as<float>(), out, elems, st);
  }
  if (prm->async_host_copy) {
    // same double-buffered egress as gsb_render, a quarter of the bytes: the conversion runs on the render stream,
    // the copy on the copy stream
    const int b = c->stage_next;
    c->stage_next ^= 1;
    GSB_TRY(c->host_stage[b].ensure((size_t)elems));
    if (c->copy_pending[b]) GSB_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_copied[b], 0));
    GSB_TRY(render_device(c, cam, prm, c->image.as<float>(), st));
    GSB_CUDA_TRY((cudaError_t)launch_to_u8(c->image.as<float>(), c->host_stage[b].as<uint8_t>(), elems, st));
    GSB_CUDA_TRY(cudaEventRecord(c->ev_rendered, st));
    GSB_CUDA_TRY(cudaStreamWaitEvent(c->copy, c->ev_rendered, 0));
    GSB_CUDA_TRY(cudaMemcpyAsync(out, c->host_stage[b].p, (size_t)elems, cudaMemcpyDeviceToHost, c->copy));
    GSB_CUDA_TRY(cudaEventRecord(c->ev_copied[b], c->copy));
    c->copy_pending[b] = true;
    return GSB_OK;
  }
  GSB_TRY(render_device(c, cam, prm, c->image.as<float>(), st));
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
  if (caller_is_capturing(st)) return GSB_E_UNSUPPORTED;
  const int64_t n = c->n;
  begin_frame(c);  // overwrites the per-frame records: a frame saved for the backward pass is gone
  if (m_out) *m_out = 0;
  if (n == 0) return GSB_OK;
  FrameGeom geom{cam->width, cam->height, tile_grid_dim(cam->width, kTile, prm->full_cover),
                 tile_grid_dim(cam->height, kTile, prm->full_cover)};
  const SuperGeom sg{0, 0, geom.tiles_x, geom.tiles_y};
  GSB_TRY(c->depth_key.ensure((size_t)n * 4));
  GSB_TRY(c->rec.ensure((size_t)n * 48));
  GSB_TRY(c->rect.ensure((size_t)n * 8));
  GSB_TRY(c->count.ensure((size_t)n * 4));
  GSB_TRY(c->dbg_cov2d.ensure((size_t)n * 16));
  GSB_TRY(c->dbg_conic.ensure((size_t)n * 16));
  GSB_TRY(c->dbg_bbox.ensure((size_t)n * 16));
  SortPlan dplan = make_sort_plan<uint32_t>(n, 0, 32);
  const CtlLayout L = ctl_layout(geom, sg, n, dplan.control_words);
  GSB_TRY(c->control.ensure(L.total * 4));
  GSB_CUDA_TRY(cudaMemsetAsync(c->control.p, 0, L.total * 4, st));
  uint32_t* hdr = c->control.as<uint32_t>();
  DebugOut dbg{c->dbg_cov2d.as<float>(), c->dbg_conic.as<float>(), c->dbg_bbox.as<float>()};
  GSB_CUDA_TRY((cudaError_t)launch_project(c->planes.as<float>(), n, c->n_pad, *cam, *prm, geom, c->depth_key.as<uint32_t>(),
                                           c->rec.as<float4>(), c->rect.as<ushort4>(), c->count.as<uint32_t>(), hdr,
                                           hdr + L.hist, /*hist_weighted=*/0, reinterpret_cast<int32_t*>(hdr + L.grid),
                                           nullptr, sg, &dbg, st));
  int launches = 1, dp = 0;
  GSB_TRY(depth_sort(c, n, hdr + L.hist, hdr + L.dsort, st, &launches, &dp));
  GSB_CUDA_TRY(cudaMemcpyAsync(c->pinned + 8, hdr, 4, cudaMemcpyDeviceToHost, st));
  GSB_CUDA_TRY(cudaStreamSynchronize(st));
  const int64_t m = c->pinned[8];
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
  if (caller_is_capturing(st)) return GSB_E_UNSUPPORTED;
  const bool cu = prm.semantics == GSB_SEM_REF_CU;
  const int cover = cu ? 1 : prm.full_cover;  // render.cu covers every pixel (:119-124)
  FrameGeom geom{W, H, tile_grid_dim(W, kTile, cover), tile_grid_dim(H, kTile, cover)};
  const int64_t tiles = (int64_t)geom.tiles_x * geom.tiles_y;
  const SuperGeom sg = choose_super(geom, c->super_lw, c->super_lh);
  const bool two_level = sg.lw + sg.lh > 0;
  begin_frame(c);  // overwrites rec / ranges / payload with M-row data: a frame saved for the backward pass is gone
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
  const CtlLayout L = ctl_layout(geom, sg, m, 0);
  GSB_TRY(c->control.ensure(L.total * 4));
  uint64_t cap_k = ~0ull, cap_ks = ~0ull;
  if (m > 0 && tiles > 0) GSB_TRY(prepare_split(c, m, sg, false, /*expand_in_stream=*/cu, &cap_k, &cap_ks));
  GSB_CUDA_TRY(cudaMemsetAsync(c->control.p, 0, L.total * 4, st));
  uint32_t* hdr = c->control.as<uint32_t>();
  tm.start();
  GSB_CUDA_TRY((cudaError_t)launch_ingest_preprocessed(m, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7], geom,
                                                       prm, c->depth_key.as<uint32_t>(), c->rec.as<float4>(),
                                                       c->bbox.as<float4>(), c->rect.as<ushort4>(), c->count.as<uint32_t>(),
                                                       reinterpret_cast<int32_t*>(hdr + L.grid),
                                                       two_level ? reinterpret_cast<int32_t*>(hdr + L.grid_s) : nullptr, sg, st));
  if (m > 0) ++launches;
  tm.mark(GSB_STAGE_PROJECT);
  if (m > 0) GSB_TRY(launch_stats_async(c, geom, sg, /*split=*/true, hdr, L, cap_k, cap_ks, st, &launches));

  float* dev_image = out_image;
  const size_t bytes = (size_t)W * H * 3 * sizeof(float);
  const bool host_out = !is_device_pointer(out_image);
  if (host_out) { GSB_TRY(c->image.ensure(bytes)); dev_image = c->image.as<float>(); }
  auto queue_composite = [&](const TileSource& src, const uint32_t* abort) -> int {
    if (!cover) GSB_CUDA_TRY(cudaMemsetAsync(dev_image, 0, bytes, st));
    if (cu)  // the REF_CU kernel walks materialised per-tile lists (expand_in_stream)
      GSB_CUDA_TRY((cudaError_t)launch_composite_cu(c->ranges.as<uint2>(), c->vals_a.as<uint32_t>(), c->rec.as<float4>(),
                                                    c->bbox.as<float4>(), dev_image, geom, prm, abort, st));
    else
      GSB_CUDA_TRY((cudaError_t)launch_composite(src, c->rec.as<float4>(), dev_image, geom, prm, nullptr, nullptr, abort, st));
    if (tiles > 0) ++launches;
    tm.mark(GSB_STAGE_COMPOSITE);
    return GSB_OK;
  };
  GSB_TRY(run_split(c, m, nullptr, /*rows_sorted_by_visibility=*/false, geom, sg, false, /*expand_in_stream=*/cu, hdr, L, st,
                    tm, &launches, /*cap=*/nullptr, queue_composite));
  c->info.m_in_view = m;
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
  GSB_CUDA_TRY(cudaDeviceSynchronize());  // the getters run on the legacy stream: order them after the frame's stream
  if (c->frame_projected) {
    // The frame variant of the projection writes no record for Gaussians without tiles and keys them like culled
    // ones; the getter reports every in-view row, so run the debug variant once more for the frame's camera
    // (identical values for the rows the frame did write; the control words it accumulates into are dead by now).
    // It rewrites depth_key / count / rec: a frame saved for the backward pass is gone.
    c->have_saved = false;
    FrameGeom geom{c->last_cam.width, c->last_cam.height, c->info.tiles_x, c->info.tiles_y};
    const SuperGeom sg{0, 0, geom.tiles_x, geom.tiles_y};
    GSB_TRY(c->dbg_cov2d.ensure((size_t)n * 16));
    GSB_TRY(c->dbg_conic.ensure((size_t)n * 16));
    GSB_TRY(c->dbg_bbox.ensure((size_t)n * 16));
    DebugOut dbg{c->dbg_cov2d.as<float>(), c->dbg_conic.as<float>(), c->dbg_bbox.as<float>()};
    const CtlLayout L = ctl_layout(geom, sg, n, 0);
    GSB_TRY(c->control.ensure(L.total * 4));
    uint32_t* ctl = c->control.as<uint32_t>();
    GSB_CUDA_TRY((cudaError_t)launch_project(c->planes.as<float>(), n, c->n_pad, c->last_cam, c->last_prm, geom,
                                             c->depth_key.as<uint32_t>(), c->rec.as<float4>(), c->rect.as<ushort4>(),
                                             c->count.as<uint32_t>(), ctl, ctl + L.hist, 0,
                                             reinterpret_cast<int32_t*>(ctl + L.grid), nullptr, sg, &dbg, 0));
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
  GSB_CUDA_TRY(cudaDeviceSynchronize());  // legacy-stream getter: order it after the frame's stream
  const size_t k = (size_t)c->info.k_instances;
  if (k == 0) return GSB_OK;
  if (!c->lists_materialized) {
    // two-level SPLIT frames keep the per-tile lists implicit (super-tile lists + tile masks; the compositing kernel
    // filters them on the fly): write them out now, with the same filter, as one array per tile
    GSB_TRY(c->vals_a.ensure(k * 4 + 16));
    FrameGeom geom{0, 0, c->info.tiles_x, c->info.tiles_y};
    GSB_CUDA_TRY((cudaError_t)launch_expand(c->ranges_s.as<uint2>(), c->cvals.as<uint2>(), c->ranges.as<uint2>(),
                                            c->vals_a.as<uint32_t>(), geom, c->last_sg, nullptr, 0));
    GSB_CUDA_TRY(cudaDeviceSynchronize());
    c->lists_materialized = true;
    c->sorted_in_a = true;
  }
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
  GSB_CUDA_TRY(cudaDeviceSynchronize());
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
  GSB_CUDA_TRY(cudaDeviceSynchronize());
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
