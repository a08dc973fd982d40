// gsb_internal.cuh -- shared declarations of libgsb_b200.so (sm_100a only).
//
// Data layout in HBM (one GsbContext = one device):
//   scene (gsb_upload):  planar SoA, 14 fp32 planes of length n_pad (n rounded up to 4):
//                        x y z | sx sy sz | qw qx qy qz | r g b | opacity_logit      (56 B/Gaussian)
//   per frame (project): depth_key u32[N]  float_as_uint(z_view), 0xFFFFFFFF when culled
//                        rec float4[3N]    AoS compositing record, 48 B/Gaussian:
//                                          {mx,my,a,b} {c,d,log2(op2),r} {g,b,radius,sig_op}
//                                          a,b,c,d = -0.5 * inverse covariance (exact scaling)
//                        rect ushort4[N]   tile rect tx0,tx1,ty0,ty1 (count==0 => unused)
//                        count u32[N]      tiles touched
//                        keys u64[K] x2, payload u32[K] x2   ping-pong buffers of the radix sort
//   ranges:              uint2[tiles]      [start,end) into the sorted arrays
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/gsb.h"

#define GSB_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

#define GSB_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != GSB_OK) return _s; \
  } while (0)

namespace gsb {

constexpr int kTile = 16;              // pixels per tile edge (the only size the kernels are built for)
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kTicketWords = 32;  // head of a radix sort's control block: 8 pass tickets padded to a 128-byte line
constexpr int kMaxPasses = 8;
// words of the per-frame control header (zeroed once per frame)
constexpr int kCtlM = 0;        // Gaussians in view
constexpr int kCtlK = 2;        // tile instances, 64-bit (tile_stats_kernel, fine level)
constexpr int kCtlVisible = 4;  // Gaussians WITH tiles (V)
constexpr int kCtlAbort = 5;    // 1: a count exceeded the capacity the frame's tail was queued with -> those kernels return at once
constexpr int kCtlKs = 6;       // super-tile instances, 64-bit (= K when the frame is binned in one level)

// ---- scene planes ----
enum Plane { PX = 0, PY, PZ, PSX, PSY, PSZ, PQW, PQX, PQY, PQZ, PR, PG, PB, POP, kNumPlanes };

struct FrameGeom {
  int width, height;
  int tiles_x, tiles_y;
};

// SPLIT mode bins in two levels: super-tiles of 2^lw x 2^lh tiles (nx x ny of them), then tiles.  lw = lh = 0: one level.
struct SuperGeom {
  int lw, lh;
  int nx, ny;
};

// what the (last) tile_stats launch of a frame reports besides its own total
struct StatsPost {
  int level;               // 0: fine grid (total -> ctl[kCtlK]), 1: super-tile grid (total -> ctl[kCtlKs])
  int enabled;             // this launch decides ctl[kCtlAbort] and posts the mailbox
  unsigned long long cap_k, cap_ks;  // capacities (in instances) the frame's tail was queued with
  uint32_t* mailbox;       // mapped pinned host memory (device alias) or nullptr: {M, V, K lo, K hi, seq, abort, Ks lo, Ks hi}
  uint32_t seq;
};

// extra per-Gaussian outputs written only by the debug/preprocess variant of the projection kernel
struct DebugOut {
  float* cov2d;   // (N,4)
  float* conic;   // (N,4) unscaled inverse covariance
  float* bbox;    // (N,4) min_x,min_y,max_x,max_y
};

// ---- kernel launchers (each returns a cudaError_t as int; all async on `st`) ----
int launch_repack(const float* xyz, const float* scales, const float* quats, const float* colors,
                  const float* opacity, float* planes, int64_t n, int64_t n_pad, cudaStream_t st);

// m_counter: the control header -- [0] += M (in view), [kCtlVisible] += V (in view and touching a tile).
// Without `dbg` (the frame variant) rows that touch no tile get depth_key 0xFFFFFFFF and no record / rect.
// depth_hist: 4*256 zeroed words (digit histograms of the depth keys; weighted by tile count when
// hist_weighted).  diff_grid: (tiles_x+1)*(tiles_y+1) zeroed ints (2-D difference grid of the tile rects).
// super_grid (optional, with sg): (sg.nx+1)*(sg.ny+1) zeroed ints, the same for the super-tile rects.
// Rows that touch no tile get rect = (1,0,1,0) (tx1 < tx0).
int launch_project(const float* planes, int64_t n, int64_t n_pad, const GsbCamera& cam, const GsbParams& prm,
                   FrameGeom geom, uint32_t* depth_key, float4* rec, ushort4* rect, uint32_t* count,
                   uint32_t* m_counter, uint32_t* depth_hist, int hist_weighted, int32_t* diff_grid,
                   int32_t* super_grid, SuperGeom sg, const DebugOut* dbg, cudaStream_t st);

// 2-D prefix sum of a difference grid (in place) -> instances per tile -> per-tile [start,end) ranges
// (empty tiles (0,0)), histograms of the tile-id digits (tile_hist: 4*256 words, or nullptr), total -> ctl.
int launch_tile_stats(int32_t* diff_grid, int tiles_x, int tiles_y, uint32_t* tile_hist, uint2* ranges, uint32_t* ctl,
                      const StatsPost& post, cudaStream_t st);

// exclusive scan of count[perm ? perm[i] : i] for i < n -> offsets[i] (u32, wraps if K >= 2^32: the host rejects
// that from tile_stats' 64-bit total).  `status` needs scan_status_words(n) zeroed u32 words, 8-byte aligned.
size_t scan_status_words(int64_t n);
int launch_scan(const uint32_t* count, const uint32_t* perm, int64_t n, uint32_t* offsets, uint32_t* status,
                cudaStream_t st);
// the same over the number of super-tiles each rect touches; rows at positions >= *v_limit (optional) are skipped
int launch_scan_coarse(const ushort4* rect, SuperGeom sg, const uint32_t* perm, int64_t n, const uint32_t* v_limit,
                       uint32_t* offsets, uint32_t* status, cudaStream_t st);
// FULL mode: keys[o] = tile<<32 | depth bits, payload[o] = Gaussian index, in index order.  `total`: device pointer
// to K (low word); k: the host's copy of K (sizes the grid: one block per 8 192 output slots)
int launch_emit(const uint32_t* offsets, const uint32_t* perm, const uint32_t* total, int64_t n, int64_t k,
                const uint32_t* depth_key, const ushort4* rect, int tiles_x, uint64_t* keys, uint32_t* payload,
                cudaStream_t st);
// SPLIT mode: one key per super-tile of the rect, in emission (depth) order: rank_bits > 0: u32 keys
// super-tile << rank_bits | emission position; rank_bits == 0: u64 keys super-tile << 32 | Gaussian index.
// The grid covers `capacity` keys; blocks past *total return, and nothing runs when *abort is set.
int launch_emit_coarse(const uint32_t* offsets, const uint32_t* perm, const uint32_t* total, int64_t n,
                       const uint32_t* v_limit, const uint32_t* abort, int64_t capacity, const ushort4* rect,
                       SuperGeom sg, int rank_bits, void* keys, cudaStream_t st);
// per-tile lists from per-super-tile lists of {Gaussian index, tile mask} entries (stable compaction; the tile
// starts come from `ranges`).  On demand only: the compositing kernel filters the super-tile lists itself.
int launch_expand(const uint2* ranges_s, const uint2* clist, const uint2* ranges, uint32_t* payload, FrameGeom geom,
                  SuperGeom sg, const uint32_t* abort, cudaStream_t st);
// debug only: sorted keys tile<<32 | depth bits from ranges + sorted payload (SPLIT mode never stores them)
int launch_rebuild_keys(const uint2* ranges, int tiles, const uint32_t* payload, const uint32_t* depth_key,
                        uint64_t* keys, cudaStream_t st);

// ---- onesweep radix sort ----
struct SortPlan {
  int begin_bit, end_bit, passes;
  int items;              // keys per thread (8 or 16)
  int keys_only;          // 1: payload packed in unsorted key bits; the last pass writes, per key, the low
                          // `low_bits` bits (all 32 when low_bits == 0) -- or gather_table[those bits] -- to vals
  int low_bits;
  const uint32_t* gather_table;
  int wide_status;        // 1: 64-bit look-back words (2^30 keys and more)
  const uint32_t* n_dev;  // optional: the key count lives on the device (u32); `n` is then the CAPACITY the grid and
                          // the status words are sized for, tiles past *n_dev return at once
  const uint32_t* abort;  // optional: nothing runs when *abort != 0
  // keys_only, optional: the last pass writes super-tile list entries {Gaussian index, tile mask} (uint2, to the vals
  // buffer of its destination side) instead of bare indices; the mask comes from entry_rect[index] and the
  // super-tile id in the key's sorted bits (super-tiles of 2^entry_lw x 2^entry_lh tiles, entry_snx per row)
  const ushort4* entry_rect;
  int entry_lw, entry_lh, entry_snx;
  int64_t n;
  int64_t tiles;          // onesweep tiles per pass
  size_t control_words;   // u32 words of control memory (tickets + look-back status), zeroed by the caller
};
void set_sort_items(int items);
void set_force_wide_status(int on);
template <typename KeyT>
SortPlan make_sort_plan(int64_t n, int begin_bit, int end_bit, int items = 16);
// digit histograms computed from the keys (stand-alone sort only); hist = kMaxPasses*256 zeroed words
template <typename KeyT>
int launch_key_histogram(const SortPlan& plan, const KeyT* keys, uint32_t* hist, cudaStream_t st);
// Pass 0 reads (keys_src, vals_src) -- vals_src == nullptr means "payload = index" -- and writes the
// b-buffers; later passes ping-pong b -> a -> b.  keys_src may be keys_a itself.  hist: [passes][256]
// digit histograms of the keys.  *result_in_a: sorted data ended in the a-buffers (even pass count).
template <typename KeyT>
int launch_sort(const SortPlan& plan, const KeyT* keys_src, const uint32_t* vals_src, KeyT* keys_a, uint32_t* vals_a,
                KeyT* keys_b, uint32_t* vals_b, const uint32_t* hist, uint32_t* control, bool* result_in_a,
                int* launches, cudaStream_t st);

// Where the compositing kernel gets a tile's list from.  clist == nullptr: a materialised per-tile payload array
// (`payload` sliced by `ranges`; FULL mode, one-level SPLIT).  clist != nullptr: the tile filters its super-tile's
// list of {Gaussian index, tile mask} entries (`clist` sliced by `ranges_s`) on the fly; with save_for_backward it
// also writes the list it consumes to payload_out[ranges[tile].x ..] for the gradient pass.
struct TileSource {
  const uint2* ranges;
  const uint32_t* payload;
  const uint2* ranges_s;
  const uint2* clist;
  uint32_t* payload_out;
  int snx, lw, lh;
};

// aux_t / aux_n (both or neither; save_for_backward): per pixel, transmittance after the last blended Gaussian
// and the length of the list prefix that reached the pixel.  abort (optional): device flag; the kernel returns at
// once when it is set (see kCtlAbort).
int launch_composite(const TileSource& src, const float4* rec, float* image, FrameGeom geom, const GsbParams& prm,
                     float* aux_t, uint32_t* aux_n, const uint32_t* abort, cudaStream_t st);

// ---- warp-level culling, shared by the compositing kernel and its gradient pass (both must skip the same records) ----
// log2 of GsbParams.cull_alpha as the kernels use it
inline float cull_threshold_log2(const GsbParams& prm) {
  // cull_alpha < 0: no warp-level skipping; 0: skip only what is exactly zero in fp32 (ex2.approx.ftz flushes below
  // 2^-126; one binade of slack for its own rounding); > 0: skip when every alpha of the warp is below it
  if (prm.cull_alpha < 0.f) return -INFINITY;
  if (prm.cull_alpha == 0.f) return -127.f;
  const float l = log2f(prm.cull_alpha);
  return l < -127.f ? -127.f : l;
}
#ifdef __CUDACC__
// max over t in [lo, hi] of  q t^2 + s t + c0  for q < 0;  rq = 1 / (2 q)
__device__ __forceinline__ float edge_max(float q, float s, float lo, float hi, float rq, float c0) {
  const float t = fminf(fmaxf(-s * rq, lo), hi);
  return fmaf(fmaf(q, t, s), t, c0);
}

// May any pixel of the rectangle dx in [dxl, dxh], dy in [dyl, dyh] (offsets mean - pixel) see
// exp2(power * log2e + l2op) >= 2^thr ?  (A B; C D) = -0.5 * inverse covariance.  Conservative: answers true
// whenever it cannot prove otherwise (indefinite or non-finite conic, NaNs, large cancellation).
__device__ __forceinline__ bool may_contribute(float A, float B, float C, float D, float l2op, float dxl, float dxh,
                                               float dyl, float dyh, float thr) {
  const float S = B + C;
  const bool ok = A < 0.f && A > -1e30f && D < 0.f && D > -1e30f && fmaf(4.f * A, D, -S * S) > 0.f &&
                  fabsf(dxl) < 1e30f && fabsf(dyl) < 1e30f;
  const bool inside = dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f;
  const float rA = __frcp_rn(2.f * A), rD = __frcp_rn(2.f * D);
  const float g1 = edge_max(D, S * dxl, dyl, dyh, rD, A * dxl * dxl);
  const float g2 = edge_max(D, S * dxh, dyl, dyh, rD, A * dxh * dxh);
  const float h1 = edge_max(A, S * dyl, dxl, dxh, rA, D * dyl * dyl);
  const float h2 = edge_max(A, S * dyh, dxl, dxh, rA, D * dyh * dyh);
  const float ub = inside ? 0.f : fmaxf(fmaxf(g1, g2), fmaxf(h1, h2));
  const float ax = fmaxf(fabsf(dxl), fabsf(dxh)), ay = fmaxf(fabsf(dyl), fabsf(dyh));
  // |rounding error| of the fp32 evaluation in the blend loop (and of this bound) <= 2^-21 * sum of |terms|
  const float mag = fmaf(fabsf(A) * ax, ax, fmaf((fabsf(B) + fabsf(C)) * ax, ay, fabsf(D) * ay * ay));
  const float arg = fmaf(ub + fmaf(mag, 4.76837158e-7f, 1e-3f), 1.4426950408889634f, l2op);
  return !ok || !(arg < thr);
}
#endif  // __CUDACC__

// ---- backward pass (backward.cu) ----
// grad2d: 12 zeroed floats per Gaussian row: d mean x,y | d a, d (b+c), d d (a b; c d = -0.5 inverse covariance) |
// sum of dL/dalpha * alpha | d colour r,g,b | 3 unused.  Accumulated with float atomics.
int launch_composite_backward(const uint2* ranges, const uint32_t* payload, const float4* rec,
                              const float* grad_image, const float* aux_t, const uint32_t* aux_n, float* grad2d,
                              FrameGeom geom, const GsbParams& prm, cudaStream_t st);
// chain rule through the projection (project.cu's forward, recomputed): grad2d -> the five attribute gradients in
// the reference layouts (N,3) (N,3) (N,4) (N,3) (N,1); rows without tile instances get zeros.
int launch_project_backward(const float* planes, int64_t n, int64_t n_pad, const GsbCamera& cam, const GsbParams& prm,
                            const uint32_t* depth_key, const uint32_t* count, const float* grad2d, float* g_points,
                            float* g_scales, float* g_quats, float* g_colors, float* g_opacity, cudaStream_t st);

// egress helpers
int launch_hwc_to_whc(const float* src_hw3, float* dst_wh3, int width, int height, cudaStream_t st);
int launch_to_u8(const float* src, uint8_t* dst, int64_t n, cudaStream_t st);
int launch_gather_preprocess(const uint32_t* order, int64_t m, const float4* rec, const float* planes, int64_t n_pad,
                             const uint32_t* depth_key, const DebugOut& dbg, float* points_xy, float* colors,
                             float* cov2d, float* depths, float* conic, float* radius, float* min_x, float* min_y,
                             float* max_x, float* max_y, float* sig_op, int32_t* src_index, cudaStream_t st);
int launch_iota(uint32_t* p, int64_t n, cudaStream_t st);

// pre-projected entry (gsb_render_image): build rec/rect/count/depth_key from the reference op's arguments
int launch_ingest_preprocessed(int64_t m, const float* means, const float* colors, const float* conic,
                               const float* min_x, const float* max_x, const float* min_y, const float* max_y,
                               const float* opacity, FrameGeom geom, const GsbParams& prm, uint32_t* depth_key,
                               float4* rec, float4* bbox, ushort4* rect, uint32_t* count, int32_t* diff_grid,
                               int32_t* super_grid, SuperGeom sg, cudaStream_t st);
int launch_composite_cu(const uint2* ranges, const uint32_t* payload, const float4* rec, const float4* bbox,
                        float* image, FrameGeom geom, const GsbParams& prm, const uint32_t* abort, cudaStream_t st);

// number of SMs of the current device (cached per device)
int sm_count();

}  // namespace gsb
