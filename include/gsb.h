/*
 * gsb.h -- C ABI of libgsb_b200.so, the B200 (sm_100a) forward Gaussian-splat rasterizer.
 *
 * This is the drop-in boundary for the native render op of
 * dcaustin33/intro_to_gaussian_splatting (reference file:line below are relative to the
 * reference checkout):
 *
 *   reference native op      splat/c/render.cu:90-101   torch::Tensor render_image(int H, int W,
 *                                                       int tile, Tensor means, colors, inv_cov2d,
 *                                                       min_x, max_x, min_y, max_y, opacity)
 *   its Python binding       splat/gaussian_scene.py:240-285 (compile_cuda_ext / render_image_cuda),
 *                            JIT-built by splat/utils.py:426-434 (load_inline)
 *   torch preprocessing it   splat/gaussian_scene.py:70-144 (GaussianScene.preprocess) and the
 *   depends on               math in splat/utils.py:132-155, :293-423, splat/gaussians.py:54-69
 *
 * Everything crosses the boundary as plain pointers and sizes.  No torch / ATen / pybind types.
 * Conventions:
 *   - every function returns an int status: 0 = ok, <0 = GSB_E_* argument/state error,
 *     >0 = a cudaError_t value.  Nothing throws across the ABI.  gsb_error_string() explains.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - "dev-or-host" pointers may be device memory or (pinned or pageable) host memory; the library
 *     asks the driver (cudaPointerGetAttributes) and stages through device memory when needed.
 *   - a context belongs to one device; it is not thread-safe; different contexts are independent.
 *   - the caller owns all inputs and outputs; the context owns only its scratch (projection
 *     records, key/payload double buffers, histograms, tile ranges), grown geometrically.
 */
#ifndef GSB_H_
#define GSB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_API_VERSION 2

/* ---- status codes (negative: library; positive: cudaError_t) ---- */
#define GSB_OK 0
#define GSB_E_INVALID_ARG (-1)
#define GSB_E_NO_SCENE (-2)      /* render before gsb_upload */
#define GSB_E_NO_FRAME (-3)      /* debug getter before a render */
#define GSB_E_UNSUPPORTED (-4)   /* e.g. tile_size != 16, image too large for the key layout, or a render call on a
                                    stream the caller is capturing (the host needs the frame's counts) */
#define GSB_E_NO_DEVICE (-5)     /* no usable CUDA device: the library has NO CPU fallback */
#define GSB_E_ALLOC (-6)
#define GSB_E_INTERNAL (-7)   /* a device-side consistency check failed */
#define GSB_E_NO_SAVED (-8)   /* gsb_render_backward: the last render of this context did not save_for_backward,
                                 or camera / params / scene changed since */

/* ---- compositing semantics ---- */
#define GSB_SEM_REF_CPU 0 /* splat/gaussian_scene.py:146-238: the parity target */
#define GSB_SEM_REF_CU 1  /* splat/c/render.cu:21-87: informational */

/* ---- sort organisation; both give bit-identical sorted (key,payload) arrays ---- */
#define GSB_SORT_AUTO 0  /* = SPLIT */
#define GSB_SORT_FULL 1  /* expand in index order, 64-bit LSD onesweep over all K (tile<<32 | depth) keys */
#define GSB_SORT_SPLIT 2 /* depth digits sorted per Gaussian (N items) BEFORE the expansion; the expansion then runs
                            in two levels: one key per SUPER-TILE (8x4 tiles at 1080p) of the rect, one radix pass
                            over the super-tile ids, and a stable per-tile compaction of each super-tile's list */

/* Per-view camera constants, exactly the tensors GaussianImage holds (splat/image.py:19-70).
 * Matrices are row-major with the reference's row-vector convention: row = [x y z 1] @ M.
 * f_x,f_y,tan_fovx,tan_fovy are the fp32 values of the GaussianImage attributes, NOT recomputed. */
typedef struct GsbCamera {
  float world2view[16];  /* image.world2view           splat/image.py:51-53 */
  float full_proj[16];   /* image.full_proj_transform  splat/image.py:61-65 */
  float f_x, f_y;        /* image.f_x, image.f_y       splat/image.py:28-29 */
  float tan_fovx, tan_fovy; /* image.tan_fovX/Y        splat/image.py:42-43 */
  int32_t width, height; /* image.width/height (1-element float tensors there) splat/image.py:37-38 */
} GsbCamera;

/* The literals scattered through the reference; defaults = reference values. */
typedef struct GsbParams {
  int32_t tile_size;   /* 16   splat/gaussian_scene.py:181,:200,:263 */
  float minimum_z;     /* 0.2  splat/utils.py:294 */
  float fov_clamp;     /* 1.3  splat/utils.py:336-337 */
  float det_min;       /* 1e-3 splat/utils.py:387 */
  float lambda_floor;  /* 0.1  splat/utils.py:414 */
  float sigma_extent;  /* 3.0  splat/utils.py:421 */
  float min_weight;    /* 1e-6 splat/gaussian_scene.py:153 (REF_CPU); 1e-3 for REF_CU render.cu:73 */
  float alpha_max;     /* REF_CU only: 0.99 render.cu:71 */
  int32_t semantics;   /* GSB_SEM_* */
  int32_t full_cover;  /* 0: reference tile grid range(0, W-T, T) (last row/col never rendered,
                          splat/gaussian_scene.py:208,:214); 1: ceil(W/T) x ceil(H/T) tiles */
  int32_t sort_mode;   /* GSB_SORT_* */
  int32_t collect_stage_times; /* 1: record CUDA events per stage (adds event overhead) */
  int32_t async_host_copy; /* 1: when the output pointer is HOST memory, copy the image on the context's copy
                              stream so that it overlaps the next frame; the host buffer is valid only after
                              gsb_join_host_copies(ctx, stream) + a synchronisation of that stream */
  int32_t save_for_backward; /* 1: gsb_render also keeps, per pixel, the number of blended Gaussians and the final
                                transmittance, so that gsb_render_backward can follow (REF_CPU semantics only) */
  float cull_alpha;    /* REF_CPU compositing: a warp skips a Gaussian when a conservative upper bound of its alpha
                          over the warp's 16x8 pixels is below this value.  The reference evaluates every Gaussian
                          of a tile at every pixel (no per-pixel bbox test, splat/gaussian_scene.py:209-226), most of
                          them with alpha that fp32 cannot see.  0: skip only alpha that is EXACTLY zero in fp32
                          (frames bit-identical to no skipping); t > 0: a skipped step would have changed a pixel by
                          < t, a frame differs by < t x list length; < 0: never skip.  Default 2^-30. */
} GsbParams;

/* Stage indices for gsb_stage_times (CUDA events on the caller's stream) */
#define GSB_STAGE_PROJECT 0
#define GSB_STAGE_DEPTH_SORT 1 /* per-Gaussian depth sort (SPLIT / BINNED) */
#define GSB_STAGE_SCAN 2       /* prefix sum of the tile counts */
#define GSB_STAGE_EMIT 3       /* key emission (FULL: one key per tile instance; SPLIT: one per super-tile instance) */
#define GSB_STAGE_SORT 4       /* radix passes over those keys */
#define GSB_STAGE_RANGES 5     /* tile statistics; runs on the auxiliary stream and reads 0 */
#define GSB_STAGE_COMPOSITE 6
#define GSB_STAGE_EXPAND 7     /* SPLIT: per-tile lists from the super-tile lists */
#define GSB_NUM_STAGES 8

/* Per-frame counts (valid after a render / preprocess call has completed on its stream). */
typedef struct GsbFrameInfo {
  int64_t n;           /* Gaussians uploaded */
  int64_t m_in_view;   /* z_view >= minimum_z (splat/utils.py:293-310) */
  int64_t k_instances; /* tile instances = sort keys */
  int32_t tiles_x, tiles_y;
  int32_t sort_passes; /* onesweep passes executed over the K keys */
  int32_t depth_passes;/* onesweep passes executed over the M depth keys (split mode) */
  int32_t kernel_launches; /* kernels launched by the last gsb_render */
  int32_t key_bits;    /* width of the keys of the last frame's tile-level radix passes: 32 or 64 */
  int64_t k_sorted;    /* keys those passes moved: K (FULL) or the number of super-tile instances (SPLIT) */
  int64_t frame_id;    /* increases with every frame this context renders; gsb_render_backward checks it */
  int32_t super_w, super_h; /* SPLIT: tiles per super-tile (1 x 1: single-level binning) */
  int64_t v_with_tiles; /* in-view Gaussians whose tile rect is not empty (the rows the binning stages touch) */
  int32_t tail_requeued; /* 1: a count outgrew the capacities the frame was queued with and its tail was queued twice */
  int32_t graph_launch;     /* 1: the frame went out as ONE CUDA graph launch (non-default stream, no stage timing) */
} GsbFrameInfo;

typedef struct GsbContext GsbContext;

int gsb_version(void);
const char* gsb_error_string(int status);
void gsb_default_params(GsbParams* p);

/* Create a context on CUDA device `device`.  Fails with GSB_E_NO_DEVICE when there is none. */
int gsb_create(GsbContext** out, int device);
void gsb_destroy(GsbContext* ctx);

/* Upload (or replace) the Gaussian set: the attributes of `Gaussians` (splat/gaussians.py:9-33),
 * in the reference's own layouts: xyz (N,3), scales (N,3) linear, quats (N,4) wxyz unnormalised,
 * colors (N,3) = rgb/256, opacity_logit (N,1).  dev-or-host pointers.  The library repacks them
 * once into its planar SoA (one device pass); the caller's buffers are not referenced afterwards. */
int gsb_upload(GsbContext* ctx, int64_t n, const float* xyz, const float* scales, const float* quats,
               const float* colors, const float* opacity_logit, void* stream);

/* The forward render: projection -> tile binning -> radix sort -> tile ranges -> compositing.
 * Replaces GaussianScene.preprocess + ext.render_image (splat/gaussian_scene.py:263-285).
 * out_image: (H,W,3) fp32, image[y][x][c] like render.cu:83-85; dev-or-host.  `cam`/`params`
 * are host structs, copied before return.  The WHOLE frame is queued before the host looks at anything
 * data-dependent: grids and buffers are sized from the context's capacities, the device decides whether
 * the frame's counts fit, and the call returns once those counts (M, K) have reached the host through a
 * mailbox in mapped pinned memory -- the stream is never drained and never waits for the host.  Only when
 * a count outgrew its buffer (first frames, or a view with many more tile instances) does the host grow
 * the buffer and queue the tail of the frame again.  `stream` must be able to make progress without further
 * action of the calling thread (no wait on an event that is only recorded later): the call waits, on the host, for
 * the frame's projection to have run, and gives up with GSB_E_INTERNAL after two minutes. */
int gsb_render(GsbContext* ctx, const GsbCamera* cam, const GsbParams* params, float* out_image,
               void* stream);

/* Makes `stream` wait for every device->host image copy still in flight on the context's copy stream (see
 * GsbParams.async_host_copy).  After this call, work queued on `stream` -- or a synchronisation of it -- is
 * ordered after those copies. */
int gsb_join_host_copies(GsbContext* ctx, void* stream);

/* Backward pass of the LAST gsb_render of this context (the reference announces training,
 * README.md:3, and marks splat/gaussians.py:19-21 requires_grad, but never wrote it).  That render must have been
 * made with params.save_for_backward = 1 and the same camera / params (compared bytewise), and nothing else may
 * have been rendered, preprocessed or uploaded on the context since (GSB_E_NO_SAVED otherwise).  frame_id: the
 * GsbFrameInfo.frame_id of the render being differentiated, or 0 for "whatever the last saved frame is".
 * grad_image: dL/d image, (H,W,3) fp32, device or host.  Outputs (device or host, any may be NULL), in the
 * reference's attribute layouts: grad_points (N,3), grad_scales (N,3), grad_quats (N,4), grad_colors (N,3),
 * grad_opacity (N,1) -- the gradient wrt the opacity LOGIT.  Tile membership, depth order, early termination and
 * the clamps are treated as piecewise constant.  Sums use float atomics: results are reproducible to rounding,
 * not bit for bit. */
int gsb_render_backward(GsbContext* ctx, const GsbCamera* cam, const GsbParams* params, int64_t frame_id,
                        const float* grad_image, float* grad_points, float* grad_scales, float* grad_quats, float* grad_colors,
                        float* grad_opacity, void* stream);

/* Same frame, egress variants: (W,H,3) layout of the reference CPU
 * path (image[x][y][c], splat/gaussian_scene.py:206,:227), or 8-bit (H,W,3) clamp(v,0,1)*255. */
int gsb_render_wh(GsbContext* ctx, const GsbCamera* cam, const GsbParams* params, float* out_image_wh,
                  void* stream);
int gsb_render_u8(GsbContext* ctx, const GsbCamera* cam, const GsbParams* params, uint8_t* out_image,
                  void* stream);

/* GaussianScene.preprocess (splat/gaussian_scene.py:70-144): runs projection + the stable depth
 * sort and writes the 12 PreprocessedScene fields (splat/schema.py:13-25) depth-sorted, ties in
 * Gaussian-index order.  Each output is dev-or-host with room for N rows (M <= N are written);
 * any may be NULL.  *m_out receives M.  Synchronous. */
int gsb_preprocess(GsbContext* ctx, const GsbCamera* cam, const GsbParams* params, int64_t* m_out,
                   float* points_xy /*M,2*/, float* colors /*M,3*/, float* covariance_2d /*M,2,2*/,
                   float* depths /*M*/, float* inverse_covariance_2d /*M,2,2*/, float* radius /*M*/,
                   float* min_x, float* min_y, float* max_x, float* max_y, float* sigmoid_opacity /*M,1*/,
                   int32_t* source_index /*M: original Gaussian index of each row*/, void* stream);

/* Drop-in for the reference op itself, same argument list as render.cu:90-101 (rows already
 * depth-sorted by the caller, as GaussianScene.preprocess leaves them): binning + sort + compositing
 * only.  `params->semantics` selects REF_CU (what render.cu computes) or REF_CPU.  dev-or-host. */
int gsb_render_image(GsbContext* ctx, int32_t image_height, int32_t image_width, int32_t tile_size,
                     int64_t m, const float* point_means /*M,2*/, const float* point_colors /*M,3*/,
                     const float* inverse_covariance_2d /*M,2,2*/, const float* min_x, const float* max_x,
                     const float* min_y, const float* max_y, const float* opacity /*M,1*/,
                     const GsbParams* params, float* out_image /*H,W,3*/, void* stream);

/* ---- parity/debug surface: state of the LAST frame rendered by this context ---- */
int gsb_frame_info(GsbContext* ctx, GsbFrameInfo* info);
/* per-Gaussian projection records in Gaussian-index order (N rows each; any pointer may be NULL):
 * in_view (u8), depth = z_view, pixel centre, tile rect [tx0,tx1]x[ty0,ty1] (int32 x4, tx1<tx0 = empty),
 * tile count.  dev-or-host. */
int gsb_debug_projection(GsbContext* ctx, uint8_t* in_view, float* depth, float* points_xy,
                         float* radius, int32_t* tile_rect, uint32_t* tile_count);
/* sorted keys (tile_id<<32 | float_as_uint(z_view)) and payload (Gaussian index), K each. */
int gsb_debug_sorted_keys(GsbContext* ctx, uint64_t* keys, uint32_t* payload);
/* unsorted keys/payload as emitted (order depends on sort_mode), K each. */
int gsb_debug_emitted_keys(GsbContext* ctx, uint64_t* keys, uint32_t* payload);
/* per-tile [start,end) into the sorted arrays: uint32 pairs, tiles_x*tiles_y of them (0,0 = empty). */
int gsb_debug_tile_ranges(GsbContext* ctx, uint32_t* ranges);
/* per-stage milliseconds of the last frame rendered with params->collect_stage_times = 1. */
int gsb_stage_times(GsbContext* ctx, float ms[GSB_NUM_STAGES]);

/* ---- the radix sort on its own (unit tests / micro-benchmarks) ----
 * Stable LSD onesweep sort of (u64 key, u32 payload) pairs on bits [begin_bit, end_bit).
 * All pointers are DEVICE pointers; keys_in/vals_in are clobbered (used as the ping-pong buffer).
 * The result is left in keys_out/vals_out. */
int gsb_sort_pairs_u64(GsbContext* ctx, int64_t n, uint64_t* keys_in, uint32_t* vals_in,
                       uint64_t* keys_out, uint32_t* vals_out, int32_t begin_bit, int32_t end_bit,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSB_H_ */
