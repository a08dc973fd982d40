// project.cu -- scene repack + the fused projection kernel.            (compile with --fmad=false)
//
// Replaces the torch preprocessing of the reference, GaussianScene.preprocess
// (splat/gaussian_scene.py:70-111) and everything it calls:
//   in_view_frustum            splat/utils.py:293-310
//   get_3d_covariance_matrix   splat/gaussians.py:54-69  (+ build_rotation splat/utils.py:132-155)
//   ndc2Pix                    splat/utils.py:313-317
//   compute_2d_covariance      splat/utils.py:320-354
//   compute_inverted_covariance splat/utils.py:368-393
//   compute_extent_and_radius  splat/utils.py:409-423
// plus the closed form of the tile masks of render_image (splat/gaussian_scene.py:208-220).
//
// Parity contract: depth, pixel centre, radius, bbox, tile rect and tile count must be BIT-EXACT
// with the reference's fp32 results, because they decide the sort keys.  torch's CPU kernels fuse
// multiply-add in some products and not in others (probed), so this TU is
// compiled with --fmad=false and every fused operation is written as __fmaf_rn explicitly; divisions
// and square roots are the IEEE-rounded intrinsics.
//
// Roofline: HBM.  Algorithmic bytes: 56 B read + 64 B written per in-view Gaussian
// (depth_key 4 + record 48 + rect 8 + count 4).
#include <cmath>

#include "gsb_internal.cuh"

namespace gsb {

// ------------------------------------------------------------------------------------------------
// repack: reference layouts (N,3)/(N,4)/(N,1) -> 14 planes.  One-time per scene.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) repack_kernel(const float* __restrict__ xyz, const float* __restrict__ scales,
                                                     const float* __restrict__ quats, const float* __restrict__ colors,
                                                     const float* __restrict__ opacity, float* __restrict__ planes,
                                                     int64_t n, int64_t n_pad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  bool ok = i < n;
  // padding rows: a Gaussian behind the camera for any view is impossible, so the kernels bound-check n.
  planes[PX * n_pad + i] = ok ? xyz[3 * i + 0] : 0.f;
  planes[PY * n_pad + i] = ok ? xyz[3 * i + 1] : 0.f;
  planes[PZ * n_pad + i] = ok ? xyz[3 * i + 2] : 0.f;
  planes[PSX * n_pad + i] = ok ? scales[3 * i + 0] : 0.f;
  planes[PSY * n_pad + i] = ok ? scales[3 * i + 1] : 0.f;
  planes[PSZ * n_pad + i] = ok ? scales[3 * i + 2] : 0.f;
  planes[PQW * n_pad + i] = ok ? quats[4 * i + 0] : 1.f;
  planes[PQX * n_pad + i] = ok ? quats[4 * i + 1] : 0.f;
  planes[PQY * n_pad + i] = ok ? quats[4 * i + 2] : 0.f;
  planes[PQZ * n_pad + i] = ok ? quats[4 * i + 3] : 0.f;
  planes[PR * n_pad + i] = ok ? colors[3 * i + 0] : 0.f;
  planes[PG * n_pad + i] = ok ? colors[3 * i + 1] : 0.f;
  planes[PB * n_pad + i] = ok ? colors[3 * i + 2] : 0.f;
  planes[POP * n_pad + i] = ok ? opacity[i] : 0.f;
}

int launch_repack(const float* xyz, const float* scales, const float* quats, const float* colors,
                  const float* opacity, float* planes, int64_t n, int64_t n_pad, cudaStream_t st) {
  if (n_pad == 0) return 0;
  unsigned blocks = (unsigned)((n_pad + 255) / 256);
  repack_kernel<<<blocks, 256, 0, st>>>(xyz, scales, quats, colors, opacity, planes, n, n_pad);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// projection
// ------------------------------------------------------------------------------------------------
struct ProjectArgs {
  GsbCamera cam;
  float minimum_z, fov_clamp, det_min, lambda_floor, sigma_extent;
  int tile_size, tiles_x, tiles_y;
  int super_lw, super_lh, super_nx, super_cells;  // super-tile grid of SPLIT mode (super_cells == 0: none)
};

constexpr int kMaxSuperCells = 1280;  // (nx+1)*(ny+1) of the 8 x 4-tile super-tile grid: 288 at 1080p, 1 085 at 4K

// [x y z 1] @ M[:, j] as torch's (N,4)@(4,4) evaluates it: one rounded product, then an FMA chain.
__device__ __forceinline__ float rowvec_col(float x, float y, float z, const float* M, int j) {
  float t = __fmul_rn(x, M[0 * 4 + j]);
  t = __fmaf_rn(y, M[1 * 4 + j], t);
  t = __fmaf_rn(z, M[2 * 4 + j], t);
  t = __fmaf_rn(1.0f, M[3 * 4 + j], t);
  return t;
}

__device__ __forceinline__ float dot3_unfused(float a0, float b0, float a1, float b1, float a2, float b2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

__device__ __forceinline__ float dot3_fmachain(float a0, float b0, float a1, float b1, float a2, float b2) {
  float t = __fmul_rn(a0, b0);
  t = __fmaf_rn(a1, b1, t);
  return __fmaf_rn(a2, b2, t);
}

__device__ __forceinline__ float clamp_torch(float v, float lo, float hi) {
  if (v != v) return v;
  float t = v < lo ? lo : v;
  return t > hi ? hi : t;
}

__device__ __forceinline__ int floor_div(int a, int b) {  // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

// Closed form of the reference tile mask along one axis (splat/gaussian_scene.py:209-217):
//   tile t (t_min = t*T) is hit  iff  mn <= t_min + T  and  mx >= t_min.
// mn, mx are integer-valued floats (floor/ceil results) or non-finite.
__device__ __forceinline__ void tile_interval(float mn, float mx, int T, int ntiles, int& lo, int& hi) {
  if (!(mn == mn) || !(mx == mx)) { lo = 0; hi = -1; return; }  // NaN never passes a comparison
  const float big = 1073741824.0f;                               // 2^30, exact; keeps the int math in range
  int imn = (int)fminf(fmaxf(mn, -big), big);
  int imx = (int)fminf(fmaxf(mx, -big), big);
  int l = floor_div(imn - T + (T - 1), T);  // ceil((mn - T)/T)
  int h = floor_div(imx, T);                // floor(mx/T)
  lo = l < 0 ? 0 : l;
  hi = h > ntiles - 1 ? ntiles - 1 : h;
}

// Besides the per-Gaussian records the kernel accumulates everything the radix sort needs to know
// about the keys WITHOUT ever reading them back:
//   * histograms of the four depth-key digits (block-private in shared memory, flushed with one RED
//     per non-empty bin per block): unweighted over all N rows for the per-Gaussian depth sort of SPLIT
//     mode, or weighted by the tile count for the 64-bit key sort of FULL mode;
//   * a 2-D difference grid of the tile rects (+1,-1,-1,+1 at the rect corners, 4 REDs per Gaussian);
//     its 2-D prefix sum (tile_stats_kernel) is the exact number of instances per tile, which yields
//     the tile-digit histograms, the per-tile [start,end) ranges and K;
//   * the same difference grid one level up, for the super-tile rects of SPLIT mode's two-level binning: a few
//     hundred cells that every Gaussian would hammer, so it is accumulated block-privately in shared memory and
//     flushed once per block like the histograms.
// The grid is persistent (a multiple of the SM count) so that the flushes stay small.
struct ProjectAccum {
  uint32_t (*hist)[kRadix];  // shared: [4][256]
  int* super_cells;          // shared super-tile difference grid, or the global one
  int kept, with_tiles;
  uint32_t n_ff;
};

// Everything the reference computes for ONE in-view Gaussian (view z already known to pass the cull), in its rounding
// order, and everything the frame keeps of it.  x, y, z: the position; i: the row.
template <bool kDebug>
__device__ __forceinline__ void project_one(const float* __restrict__ planes, int64_t n_pad, int64_t i, float x, float y,
                                            float z, float vz, const ProjectArgs& a, uint32_t* __restrict__ depth_key,
                                            float4* __restrict__ rec, ushort4* __restrict__ rect,
                                            uint32_t* __restrict__ count, int hist_weighted,
                                            int32_t* __restrict__ diff_grid, const DebugOut& dbg, ProjectAccum& acc) {
  uint32_t cnt = 0;
  uint32_t dkey = 0xFFFFFFFFu;
  int tx0 = 0, tx1 = -1, ty0 = 0, ty1 = -1;
  const float* V = a.cam.world2view;
  const float* F = a.cam.full_proj;
  {
    const float sx = planes[PSX * n_pad + i], sy = planes[PSY * n_pad + i], sz = planes[PSZ * n_pad + i];
    float q0 = planes[PQW * n_pad + i], q1 = planes[PQX * n_pad + i], q2 = planes[PQY * n_pad + i],
          q3 = planes[PQZ * n_pad + i];

    const float vx = rowvec_col(x, y, z, V, 0);
    const float vy = rowvec_col(x, y, z, V, 1);
    // pixel centre: clip space -> NDC -> ndc2Pix (principal point is not used on this path)
    const float cx = rowvec_col(x, y, z, F, 0);
    const float cy = rowvec_col(x, y, z, F, 1);
    const float cw = rowvec_col(x, y, z, F, 3);
    const float ndx = __fdiv_rn(cx, cw), ndy = __fdiv_rn(cy, cw);
    const float Wf = (float)a.cam.width, Hf = (float)a.cam.height;
    const float px = __fmul_rn(__fmul_rn(__fadd_rn(ndx, 1.0f), __fsub_rn(Wf, 1.0f)), 0.5f);
    const float py = __fmul_rn(__fmul_rn(__fadd_rn(ndy, 1.0f), __fsub_rn(Hf, 1.0f)), 0.5f);

    // ---- 3-D covariance: F.normalize, then build_rotation normalises again ----
    float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2)),
                                     __fmul_rn(q3, q3)));
    float dn = nrm > 1e-12f ? nrm : 1e-12f;
    q0 = __fdiv_rn(q0, dn); q1 = __fdiv_rn(q1, dn); q2 = __fdiv_rn(q2, dn); q3 = __fdiv_rn(q3, dn);
    float n2 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2)),
                                    __fmul_rn(q3, q3)));
    const float r = __fdiv_rn(q0, n2), qx = __fdiv_rn(q1, n2), qy = __fdiv_rn(q2, n2), qz = __fdiv_rn(q3, n2);
    float R[9];
    R[0] = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(__fmul_rn(qy, qy), __fmul_rn(qz, qz))));
    R[1] = __fmul_rn(2.0f, __fsub_rn(__fmul_rn(qx, qy), __fmul_rn(r, qz)));
    R[2] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(qx, qz), __fmul_rn(r, qy)));
    R[3] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(qx, qy), __fmul_rn(r, qz)));
    R[4] = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qz, qz))));
    R[5] = __fmul_rn(2.0f, __fsub_rn(__fmul_rn(qy, qz), __fmul_rn(r, qx)));
    R[6] = __fmul_rn(2.0f, __fsub_rn(__fmul_rn(qx, qz), __fmul_rn(r, qy)));
    R[7] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(qy, qz), __fmul_rn(r, qx)));
    R[8] = __fsub_rn(1.0f, __fmul_rn(2.0f, __fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy))));
    const float s[3] = {sx, sy, sz};
    float Ms[9];
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) Ms[ii * 3 + jj] = __fmul_rn(R[ii * 3 + jj], s[jj]);  // R @ diag(s)
    float S3[9];
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
      for (int jj = 0; jj < 3; ++jj)  // M @ M^T, batched bmm: unfused
        S3[ii * 3 + jj] = dot3_unfused(Ms[ii * 3 + 0], Ms[jj * 3 + 0], Ms[ii * 3 + 1], Ms[jj * 3 + 1], Ms[ii * 3 + 2],
                                       Ms[jj * 3 + 2]);

    // ---- EWA 2-D covariance ----
    const float limx = __fmul_rn(a.fov_clamp, a.cam.tan_fovx);
    const float limy = __fmul_rn(a.fov_clamp, a.cam.tan_fovy);
    const float tx = __fmul_rn(clamp_torch(__fdiv_rn(vx, vz), -limx, limx), vz);
    const float ty = __fmul_rn(clamp_torch(__fdiv_rn(vy, vz), -limy, limy), vz);
    const float z2 = __fmul_rn(vz, vz);
    float J[6];  // rows 0,1 of J (row 2 is zero and only feeds discarded outputs)
    J[0] = __fdiv_rn(a.cam.f_x, vz); J[1] = 0.f; J[2] = __fdiv_rn(-__fmul_rn(a.cam.f_x, tx), z2);
    J[3] = 0.f; J[4] = __fdiv_rn(a.cam.f_y, vz); J[5] = __fdiv_rn(-__fmul_rn(a.cam.f_y, ty), z2);
    // W = world2view[:3,:3].T  =>  W[l][k] = V[k][l];  (W.T)[l][k] = V[l][k]
    float T1[6], T2[6], T3[6];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
      for (int k = 0; k < 3; ++k)  // J @ W (broadcast 3x3): FMA chain
        T1[ii * 3 + k] = dot3_fmachain(J[ii * 3 + 0], V[k * 4 + 0], J[ii * 3 + 1], V[k * 4 + 1], J[ii * 3 + 2], V[k * 4 + 2]);
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
      for (int k = 0; k < 3; ++k)  // @ Sigma (batched): unfused
        T2[ii * 3 + k] = dot3_unfused(T1[ii * 3 + 0], S3[0 * 3 + k], T1[ii * 3 + 1], S3[1 * 3 + k], T1[ii * 3 + 2], S3[2 * 3 + k]);
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
      for (int k = 0; k < 3; ++k)  // @ W.T (broadcast): FMA chain
        T3[ii * 3 + k] = dot3_fmachain(T2[ii * 3 + 0], V[0 * 4 + k], T2[ii * 3 + 1], V[1 * 4 + k], T2[ii * 3 + 2], V[2 * 4 + k]);
    float c2[4];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)  // @ J^T (batched): unfused
        c2[ii * 2 + jj] = dot3_unfused(T3[ii * 3 + 0], J[jj * 3 + 0], T3[ii * 3 + 1], J[jj * 3 + 1], T3[ii * 3 + 2], J[jj * 3 + 2]);
    const float ca = c2[0], cbb = c2[1], cc = c2[2], cd = c2[3];

    // ---- conic (det clamp 1e-3), radius (lambda floor 0.1, 3 sigma), bbox ----
    float det = __fsub_rn(__fmul_rn(ca, cd), __fmul_rn(cbb, cc));
    det = (det != det) ? det : (det < a.det_min ? a.det_min : det);
    const float i00 = __fdiv_rn(cd, det), i11 = __fdiv_rn(ca, det);
    const float i01 = __fdiv_rn(-cbb, det), i10 = __fdiv_rn(-cc, det);
    const float mid = __fmul_rn(0.5f, __fadd_rn(ca, cd));
    const float det2 = __fsub_rn(__fmul_rn(ca, cd), __fmul_rn(cbb, cbb));
    const float im = __fsub_rn(__fmul_rn(mid, mid), det2);
    const float mv = (im != im) ? im : (im > a.lambda_floor ? im : a.lambda_floor);
    const float sq = __fsqrt_rn(mv);
    const float l1 = __fadd_rn(mid, sq), l2 = __fsub_rn(mid, sq);
    const float lm = (l1 != l1 || l2 != l2) ? __int_as_float(0x7fc00000) : (l1 > l2 ? l1 : l2);
    const float rad = ceilf(__fmul_rn(a.sigma_extent, __fsqrt_rn(lm)));
    const float mnx = floorf(__fsub_rn(px, rad)), mny = floorf(__fsub_rn(py, rad));
    const float mxx = ceilf(__fadd_rn(px, rad)), mxy = ceilf(__fadd_rn(py, rad));

    tile_interval(mnx, mxx, a.tile_size, a.tiles_x, tx0, tx1);
    tile_interval(mny, mxy, a.tile_size, a.tiles_y, ty0, ty1);
    if (tx1 >= tx0 && ty1 >= ty0) cnt = (uint32_t)(tx1 - tx0 + 1) * (uint32_t)(ty1 - ty0 + 1);

    // A Gaussian that touches no tile is never read again on the render path: the frame variant skips its
    // record and keys it like a culled one (0xFFFFFFFF), so the depth sort leaves the V Gaussians WITH tiles
    // first, in depth order.  The debug variant (gsb_preprocess, gsb_debug_projection) keeps every in-view row.
    if (kDebug || cnt) {
      // colour and opacity are only read (and the two sigmoids only evaluated) for Gaussians that are drawn
      const float cr = planes[PR * n_pad + i], cg = planes[PG * n_pad + i], cb = planes[PB * n_pad + i];
      const float logit = planes[POP * n_pad + i];
      // opacity as the CPU path uses it: sigmoid(sigmoid(logit)) (splat/gaussian_scene.py:143 then :164)
      const float sig1 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-logit)));
      const float op2 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-sig1)));
      // conic pre-scaled by -0.5: exact (power of two), so (-0.5*d) @ inv rounds identically (composite.cu)
      rec[3 * i + 0] = make_float4(px, py, -0.5f * i00, -0.5f * i01);
      // the blend loop evaluates alpha = op2 * exp(power) as exp2(power*log2e + log2(op2)): one FFMA + MUFU.EX2
      rec[3 * i + 1] = make_float4(-0.5f * i10, -0.5f * i11, log2f(op2), cr);
      rec[3 * i + 2] = make_float4(cg, cb, rad, sig1);
      dkey = __float_as_uint(vz);
    }
    if (kDebug) {
      reinterpret_cast<float4*>(dbg.cov2d)[i] = make_float4(ca, cbb, cc, cd);
      reinterpret_cast<float4*>(dbg.conic)[i] = make_float4(i00, i01, i10, i11);
      reinterpret_cast<float4*>(dbg.bbox)[i] = make_float4(mnx, mny, mxx, mxy);
    }
  }
  // tx1 < tx0 marks "touches no tile" for everything downstream that reads rects without the count
  rect[i] = cnt ? make_ushort4((unsigned short)tx0, (unsigned short)tx1, (unsigned short)ty0, (unsigned short)ty1)
                : make_ushort4(1, 0, 1, 0);
  count[i] = cnt;
  depth_key[i] = dkey;  // 0xFFFFFFFF when it touches no tile: sorts behind every real depth (z >= 0.2 > 0, finite)
  acc.with_tiles += cnt ? 1 : 0;
  const uint32_t w = hist_weighted ? cnt : 1u;
  if (dkey == 0xFFFFFFFFu) {
    acc.n_ff += w;  // all four digits are 255: counted per thread, added once per warp (no same-address atomics)
  } else if (w) {
    atomicAdd(&acc.hist[0][dkey & 255u], w);
    atomicAdd(&acc.hist[1][(dkey >> 8) & 255u], w);
    atomicAdd(&acc.hist[2][(dkey >> 16) & 255u], w);
    atomicAdd(&acc.hist[3][dkey >> 24], w);
  }
  if (cnt) {
    const int gw = a.tiles_x + 1;  // width of the difference grid
    atomicAdd(&diff_grid[ty0 * gw + tx0], 1);
    atomicAdd(&diff_grid[ty0 * gw + tx1 + 1], -1);
    atomicAdd(&diff_grid[(ty1 + 1) * gw + tx0], -1);
    atomicAdd(&diff_grid[(ty1 + 1) * gw + tx1 + 1], 1);
    if (a.super_cells) {
      const int sx0 = tx0 >> a.super_lw, sx1 = (tx1 >> a.super_lw) + 1;
      const int sy0 = ty0 >> a.super_lh, sy1 = (ty1 >> a.super_lh) + 1;
      const int gs = a.super_nx + 1;
      atomicAdd(&acc.super_cells[sy0 * gs + sx0], 1);
      atomicAdd(&acc.super_cells[sy0 * gs + sx1], -1);
      atomicAdd(&acc.super_cells[sy1 * gs + sx0], -1);
      atomicAdd(&acc.super_cells[sy1 * gs + sx1], 1);
    }
  }
}

// A row that is culled, or proven to touch no tile: what the frame keeps of it.
__device__ __forceinline__ void project_none(int64_t i, uint32_t* __restrict__ depth_key, ushort4* __restrict__ rect,
                                             uint32_t* __restrict__ count, int hist_weighted, ProjectAccum& acc) {
  rect[i] = make_ushort4(1, 0, 1, 0);
  count[i] = 0u;
  depth_key[i] = 0xFFFFFFFFu;
  if (!hist_weighted) acc.n_ff += 1u;
}

// One grid-stride loop, one row per thread and trip.  63 % of the in-view Gaussians of config 3 end up touching no
// tile after ~900 instructions each; a two-phase variant (cheap conservative radius bound from the largest scale
// first -- it rejects 97 % of those rows and never a drawn one --, survivors queued in shared memory until 256 of them
// run the full arithmetic with dense lanes) executed 35 % fewer instructions (20.9 M vs 31.9 M warp instructions) and
// was SLOWER, 55-60 us against 51-53: the arithmetic is long chains of IEEE divisions, latency bound, and the barriers
// of the queue cost more than the idle lanes (profiles/r2_summary.md).  Dropped.
template <bool kDebug>
__global__ void __launch_bounds__(256, 4)
project_kernel(const float* __restrict__ planes, int64_t n, int64_t n_pad, const __grid_constant__ ProjectArgs a,
               uint32_t* __restrict__ depth_key, float4* __restrict__ rec, ushort4* __restrict__ rect,
               uint32_t* __restrict__ count, uint32_t* __restrict__ m_counter, uint32_t* __restrict__ depth_hist,
               int hist_weighted, int32_t* __restrict__ diff_grid, int32_t* __restrict__ super_grid, DebugOut dbg) {
  __shared__ uint32_t s_hist[4][kRadix];
  __shared__ int s_cg[kMaxSuperCells];
  __shared__ int s_cnt, s_vis;
  for (int t = threadIdx.x; t < 4 * kRadix; t += blockDim.x) (&s_hist[0][0])[t] = 0;
  const bool cg_smem = a.super_cells <= kMaxSuperCells;  // (larger grids: global atomics)
  if (cg_smem)
    for (int t = threadIdx.x; t < a.super_cells; t += blockDim.x) s_cg[t] = 0;
  if (threadIdx.x == 0) { s_cnt = 0; s_vis = 0; }
  __syncthreads();
  ProjectAccum acc{s_hist, cg_smem ? s_cg : super_grid, 0, 0, 0u};
  const float* V = a.cam.world2view;
  {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float x = planes[PX * n_pad + i], y = planes[PY * n_pad + i], z = planes[PZ * n_pad + i];
      const float vz = rowvec_col(x, y, z, V, 2);
      if (vz >= a.minimum_z) {  // in_view_frustum; the ONLY cull (no x/y frustum test in the reference)
        acc.kept += 1;
        project_one<kDebug>(planes, n_pad, i, x, y, z, vz, a, depth_key, rec, rect, count, hist_weighted, diff_grid, dbg, acc);
      } else {
        project_none(i, depth_key, rect, count, hist_weighted, acc);
      }
    }
  }
  // M = number of in-view Gaussians, V = number that touch a tile: one atomic each per block
  int kept = acc.kept, with_tiles = acc.with_tiles;
  uint32_t n_ff = acc.n_ff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kept += __shfl_xor_sync(0xffffffffu, kept, o);
    with_tiles += __shfl_xor_sync(0xffffffffu, with_tiles, o);
    n_ff += __shfl_xor_sync(0xffffffffu, n_ff, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (kept) atomicAdd(&s_cnt, kept);
    if (with_tiles) atomicAdd(&s_vis, with_tiles);
    if (n_ff) {
      atomicAdd(&s_hist[0][255], n_ff); atomicAdd(&s_hist[1][255], n_ff);
      atomicAdd(&s_hist[2][255], n_ff); atomicAdd(&s_hist[3][255], n_ff);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(m_counter, (uint32_t)s_cnt);
  if (threadIdx.x == 0 && s_vis) atomicAdd(m_counter + kCtlVisible, (uint32_t)s_vis);
  for (int t = threadIdx.x; t < 4 * kRadix; t += blockDim.x) {
    const uint32_t v = (&s_hist[0][0])[t];
    if (v) atomicAdd(&depth_hist[t], v);
  }
  if (cg_smem)
    for (int t = threadIdx.x; t < a.super_cells; t += blockDim.x) {
      const int v = s_cg[t];
      if (v) atomicAdd(&super_grid[t], v);
    }
}

int launch_project(const float* planes, int64_t n, int64_t n_pad, const GsbCamera& cam, const GsbParams& prm,
                   FrameGeom geom, uint32_t* depth_key, float4* rec, ushort4* rect, uint32_t* count,
                   uint32_t* m_counter, uint32_t* depth_hist, int hist_weighted, int32_t* diff_grid,
                   int32_t* super_grid, SuperGeom sg, const DebugOut* dbg, cudaStream_t st) {
  if (n == 0) return 0;
  ProjectArgs a;
  a.cam = cam;
  a.minimum_z = prm.minimum_z; a.fov_clamp = prm.fov_clamp; a.det_min = prm.det_min;
  a.lambda_floor = prm.lambda_floor; a.sigma_extent = prm.sigma_extent;
  a.tile_size = prm.tile_size; a.tiles_x = geom.tiles_x; a.tiles_y = geom.tiles_y;
  a.super_lw = sg.lw; a.super_lh = sg.lh; a.super_nx = sg.nx;
  a.super_cells = super_grid ? (sg.nx + 1) * (sg.ny + 1) : 0;
  int64_t want = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;  // persistent: 8 CTAs per SM
  unsigned blocks = (unsigned)(want < cap ? want : cap);
  if (dbg)
    project_kernel<true><<<blocks, 256, 0, st>>>(planes, n, n_pad, a, depth_key, rec, rect, count, m_counter, depth_hist,
                                                 hist_weighted, diff_grid, super_grid, *dbg);
  else
    project_kernel<false><<<blocks, 256, 0, st>>>(planes, n, n_pad, a, depth_key, rec, rect, count, m_counter, depth_hist,
                                                  hist_weighted, diff_grid, super_grid, DebugOut{});
  return (int)cudaGetLastError();
}

}  // namespace gsb
