// composite.cu -- per-tile front-to-back alpha compositing.
//
// Replaces GaussianScene.render_tile / render_pixel (splat/gaussian_scene.py:146-198) and
// compute_gaussian_weight (splat/utils.py:357-365) [semantics REF_CPU, the parity target], and the
// brute-force render_tile kernel of splat/c/render.cu:21-87 [semantics REF_CU].
//
// REF_CPU per (pixel, Gaussian) step, in list order (front to back):
//     d = mean - pixel;  w = exp(-0.5 * d^T inv d);  alpha = w * sigmoid(sigmoid(logit))
//     test = T * (1 - alpha);  if test < 1e-6: STOP, this Gaussian is NOT added
//     C += T * alpha * c;  T = test
// No per-pixel bbox test, no alpha clamp, no 1/255 skip.
//
// One CTA per 16x16 tile, one thread per pixel; each warp owns an 8x4 pixel block so that early
// termination is spatially coherent.  The tile's instance list is staged through shared memory in
// batches of 256: thread t gathers the 48-byte record of instance batch+t (prefetched into registers
// one batch ahead, so the gather latency hides behind the blend loop), every thread then reads the
// records as shared-memory broadcasts.  A warp leaves the blend loop when all its pixels are done; the
// CTA stops fetching batches when __syncthreads_and says every pixel is done.
//
// Roofline: issue slots (fp32 FMA/ALU + MUFU.EX2), not HBM: ~20 warp-instructions per
// (warp, Gaussian) step; HBM traffic is 4 B payload + 48 B record per instance + 12 B per pixel.
#include <cmath>
#include <cstdlib>

#include "gsb_internal.cuh"

namespace gsb {

namespace {

constexpr int kBatch = 256;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct CompositeArgs {
  int width, height, tiles_x, tiles_y;
  float min_weight, alpha_max;
  float cull_log2;  // warp-level skip threshold as log2(alpha); -inf = never skip
};

template <int kSem>
__global__ void __launch_bounds__(256)
composite_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ payload,
                 const float4* __restrict__ rec, const float4* __restrict__ bbox, float* __restrict__ image,
                 const uint32_t* __restrict__ abort, const __grid_constant__ CompositeArgs a) {
  if (abort && *abort) return;
  __shared__ float4 s0[kBatch];  // mx, my, a, b      (a b; c d) = -0.5 * inverse covariance
  __shared__ float4 s1[kBatch];  // c, d, op, r
  __shared__ float2 s2[kBatch];  // g, b
  __shared__ float4 s3[kSem == GSB_SEM_REF_CU ? kBatch : 1];  // bbox (REF_CU only)

  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int px = tx * kTile + (warp & 1) * 8 + (lane & 7);
  const int py = ty * kTile + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < a.width && py < a.height;
  const float fx = (float)px, fy = (float)py;

  const uint2 rg = ranges[tile];
  const uint32_t len = rg.y - rg.x;

  float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f;
  bool done = !inside;

  // register-staged prefetch of the first batch
  float4 r0 = make_float4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
  if ((uint32_t)tid < len) {
    const uint32_t g = payload[rg.x + tid];
    r0 = rec[3 * (size_t)g]; r1 = rec[3 * (size_t)g + 1]; r2 = rec[3 * (size_t)g + 2];
    if (kSem == GSB_SEM_REF_CU) r3 = bbox[g];
  }
  for (uint32_t b0 = 0; b0 < len; b0 += kBatch) {
    s0[tid] = r0; s1[tid] = r1; s2[tid] = make_float2(r2.x, r2.y);
    if (kSem == GSB_SEM_REF_CU) s3[tid] = r3;
    __syncthreads();
    const uint32_t nxt = b0 + kBatch + tid;
    if (nxt < len) {
      const uint32_t g = payload[rg.x + nxt];
      r0 = rec[3 * (size_t)g]; r1 = rec[3 * (size_t)g + 1]; r2 = rec[3 * (size_t)g + 2];
      if (kSem == GSB_SEM_REF_CU) r3 = bbox[g];
    }
    const int cnt = (int)min((uint32_t)kBatch, len - b0);
    if (!done) {
#pragma unroll 4
      for (int i = 0; i < cnt; ++i) {
        const float4 g0 = s0[i];
        const float4 g1 = s1[i];
        if (kSem == GSB_SEM_REF_CU) {
          const float4 bb = s3[i];
          if (fx < bb.x || fx > bb.z || fy < bb.y || fy > bb.w) continue;
        }
        const float dx = g0.x - fx, dy = g0.y - fy;
        // The reference's own rounding sequence (compute_gaussian_weight, splat/utils.py:363-364), probed
        // bit-exact against torch on ill-conditioned conics: ((-0.5 d) @ inv) is an FMA chain in k order,
        // (.) @ d^T is two rounded products and a rounded sum.  Long thin Gaussians make this sum cancel
        // catastrophically, so any other order drifts from the reference by far more than 1e-4.
        const float u0 = __fmaf_rn(dy, g1.x, __fmul_rn(dx, g0.z));
        const float u1 = __fmaf_rn(dy, g1.y, __fmul_rn(dx, g0.w));
        const float power = __fadd_rn(__fmul_rn(u0, dx), __fmul_rn(u1, dy));
        float alpha = ex2_approx(power * 1.4426950408889634f) * g1.z;
        if (kSem == GSB_SEM_REF_CU) alpha = fminf(a.alpha_max, alpha);
        const float ta = T * alpha;
        const float test = T - ta;
        if (test < a.min_weight) { done = true; break; }
        const float2 gb = s2[i];
        cr = fmaf(ta, g1.w, cr);
        cg = fmaf(ta, gb.x, cg);
        cb = fmaf(ta, gb.y, cb);
        T = test;
      }
    }
    if (__syncthreads_and(done)) break;
  }
  if (inside) {
    float* o = image + ((size_t)py * a.width + px) * 3;
    o[0] = cr; o[1] = cg; o[2] = cb;
  }
}


// ------------------------------------------------------------------------------------------------
// REF_CPU fast path: 64 threads per 16x16 tile, each thread owns a 1x4 pixel column.
//
//   warp w (0/1) -> rows 8w..8w+7;  lane l -> column l&15, rows 8w + 4*(l>>4) + {0,1,2,3}
//
// Why four pixels per thread: the per-Gaussian work that does not depend on the row (three LDS, dx,
// dx*a, dx*b) is paid once per thread-step instead of once per pixel-step, ~15 instead of ~22 issue
// slots per pixel-step.  The price is coarser termination (a warp now spans a 16x8 region); measured
// on config 3 with the oracle's per-pixel step counts that costs 5.5% more lane-steps than 8x4 regions
// (lane efficiency 0.893 vs 0.948) -- termination is spatially very coherent.
//
// Termination without divergence: each pixel carries a predicate `live`; the step computes
// test = T - T*alpha, live &= (test >= min_weight), and the three colour FMAs are predicated on it.  T
// itself is updated unconditionally: once live is false it can never become true again (it is ANDed),
// so a decaying T contributes nothing.  The Gaussian that trips the threshold is therefore not added,
// exactly like `return pixel_color` at splat/gaussian_scene.py:166-167.  Every 4 Gaussians the warp
// votes and leaves when no pixel is live.
//
// Staging: records are copied global->shared with cp.async (LDGSTS, 3 x 16 B per record, no register
// staging) into a double buffer, one batch ahead of the blend loop; the payload index of the batch
// after that is prefetched into a register so the cp.async addresses are ready when the buffer frees.
//
// Per-warp culling.  The reference has no per-pixel bounding-box test: every pixel of a
// tile evaluates every Gaussian of the tile's list, and for a large share of those steps alpha is exactly 0 in
// fp32 (ex2.approx.ftz underflows below 2^-126) or far below anything fp32 can see.  Once a batch has landed, each
// thread computes for its two records a CONSERVATIVE upper bound of  power * log2(e) + log2(opacity)  over each
// warp's 16x8 pixel rectangle: the exponent is a concave quadratic in the pixel offset, so its maximum over a
// rectangle is 0 when the centre is inside and otherwise the best of four 1-D edge maxima; a rounding margin
// proportional to the magnitude of the cancelling terms covers the difference between the real-valued quadratic
// and the kernel's (the reference's) fp32 evaluation order, so ill-conditioned conics are simply never culled.
// One ballot per (record slot, warp footprint) turns the verdicts into a 128-bit survivor mask per warp and
// batch, and the blend loop walks the set bits only.  With cull_alpha = 0 only steps whose alpha is exactly zero
// are skipped (bit-identical frames); with cull_alpha = t > 0 a skipped step would have changed a pixel by less
// than t (T and live are untouched for alpha < 2^-25), so the frame differs by < t * (list length).
// ------------------------------------------------------------------------------------------------
constexpr int kFastThreads = 64;
constexpr int kFastBatch = 128;                        // records per stage
constexpr int kFastPerThread = kFastBatch / kFastThreads;  // records each thread stages
#ifndef GSB_FAST_UNROLL
#define GSB_FAST_UNROLL 4
#endif
constexpr int kFastUnroll = GSB_FAST_UNROLL;                 // Gaussians between two warp votes

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

// kAux (save_for_backward): also records, per pixel, the length of the list prefix that reached the pixel (index of
// the last blended Gaussian + 1) and the transmittance after it -- what the back-to-front gradient pass starts
// from (backward.cu).
template <bool kAux>
__global__ void __launch_bounds__(kFastThreads)
composite_fast_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ payload,
                      const float4* __restrict__ rec, float* __restrict__ image, float* __restrict__ aux_t,
                      uint32_t* __restrict__ aux_n, const uint32_t* __restrict__ abort,
                      const __grid_constant__ CompositeArgs a) {
  __shared__ __align__(16) float4 sm[2][kFastBatch * 3];
  __shared__ __align__(16) uint32_t s_mask[2][2][kFastBatch / 32];  // [buffer][warp footprint][survivor bits]
  if (abort && *abort) return;  // the lists do not exist: the host re-queues the frame's tail (gsb_api.cu)

  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int px = tx * kTile + (lane & 15);
  const int py0 = ty * kTile + warp * 8 + (lane >> 4) * 4;
  const float fx = (float)px;
  const float fy0 = (float)py0, fy1 = (float)(py0 + 1), fy2 = (float)(py0 + 2), fy3 = (float)(py0 + 3);
  const float minw = a.min_weight;
  const float cull = a.cull_log2;
  // pixel rectangles of the two warps (columns shared)
  const float rx0 = (float)(tx * kTile), rx1 = rx0 + (float)(kTile - 1);
  const float ry0 = (float)(ty * kTile), ry1 = ry0 + 7.f, ry2 = ry0 + 8.f, ry3 = ry0 + 15.f;

  const uint2 rg = ranges[tile];
  const uint32_t len = rg.y - rg.x;
  const uint32_t* pl = payload + rg.x;

  bool l0 = px < a.width && py0 < a.height, l1 = px < a.width && py0 + 1 < a.height;
  bool l2 = px < a.width && py0 + 2 < a.height, l3 = px < a.width && py0 + 3 < a.height;
  float T0 = 1.f, T1 = 1.f, T2 = 1.f, T3 = 1.f;
  float r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;
  float r2 = 0.f, g2 = 0.f, b2 = 0.f, r3 = 0.f, g3 = 0.f, b3 = 0.f;
  uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;          // kAux only
  float tf0 = 1.f, tf1 = 1.f, tf2 = 1.f, tf3 = 1.f;  // kAux only

  // stage batch `b` (records b*128 .. b*128+127) into buffer `buf`; idx[] holds this thread's payload indices
  uint32_t idx[kFastPerThread];
  auto load_idx = [&](uint32_t b) {
#pragma unroll
    for (int j = 0; j < kFastPerThread; ++j) {
      const uint32_t slot = b * kFastBatch + j * kFastThreads + tid;
      idx[j] = slot < len ? pl[slot] : 0xFFFFFFFFu;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int j = 0; j < kFastPerThread; ++j) {
      float4* dst = &sm[buf][(j * kFastThreads + tid) * 3];
      if (idx[j] != 0xFFFFFFFFu) {  // slots past the end of the list are never read: their survivor bit is 0
        const float4* src = rec + 3 * (size_t)idx[j];
        cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
      }
    }
    cp_async_commit();
  };
  // survivor masks of batch `b` (resident in buffer `buf`): record slot j*64 + tid is bit `lane` of word 2j + warp
  auto build_masks = [&](int buf, uint32_t b) {
#pragma unroll
    for (int j = 0; j < kFastPerThread; ++j) {
      const int r = j * kFastThreads + tid;
      bool k0 = false, k1 = false;
      if (b * kFastBatch + r < len) {
        const float4 q0 = sm[buf][r * 3], q1 = sm[buf][r * 3 + 1];
        const float dxl = q0.x - rx1, dxh = q0.x - rx0;
        k0 = may_contribute(q0.z, q0.w, q1.x, q1.y, q1.z, dxl, dxh, q0.y - ry1, q0.y - ry0, cull);
        k1 = may_contribute(q0.z, q0.w, q1.x, q1.y, q1.z, dxl, dxh, q0.y - ry3, q0.y - ry2, cull);
      }
      const unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
      if (lane == 0) { s_mask[buf][0][2 * j + warp] = m0; s_mask[buf][1][2 * j + warp] = m1; }
    }
  };

  const uint32_t nb = (len + kFastBatch - 1) / kFastBatch;
  if (nb > 0) { load_idx(0); stage(0); }
  if (nb > 1) load_idx(1);
  bool warp_live = true;
  for (uint32_t b = 0; b < nb; ++b) {
    const int buf = (int)(b & 1);
    cp_async_wait<0>();
    // batch b visible to all; everyone is done reading the other buffer; stop when no pixel of the tile is live
    if (!__syncthreads_or(warp_live)) break;
    if (b + 1 < nb) {
      stage(buf ^ 1);
      if (b + 2 < nb) load_idx(b + 2);
    }
    build_masks(buf, b);
    __syncthreads();
    if (warp_live) {
#pragma unroll 1
      for (int q = 0; q < kFastBatch / 32; ++q) {
        uint32_t m = s_mask[buf][warp][q];
        const float4* pq = sm[buf] + q * 32 * 3;
        const uint32_t nbase = b * kFastBatch + q * 32 + 1;  // kAux: list position + 1 of bit 0
        while (m) {
#pragma unroll
          for (int u = 0; u < kFastUnroll; ++u) {
            if (m == 0u) break;
            const int i = __ffs(m) - 1;
            m &= m - 1u;
            const float4* p = pq + i * 3;
            const float4 q0 = p[0];  // mx, my, a, b
            const float4 q1 = p[1];  // c, d, log2(op), r
            const float2 q2 = *reinterpret_cast<const float2*>(p + 2);  // g, b
            const float dx = q0.x - fx;
            const float ta_ = __fmul_rn(dx, q0.z), tb_ = __fmul_rn(dx, q0.w);
#define GSB_PIXEL_STEP(FY, T, LIVE, R, G, B, NC, TF)                                       \
            {                                                                              \
              const float dy = q0.y - FY;                                                  \
              const float u0 = __fmaf_rn(dy, q1.x, ta_);                                   \
              const float u1 = __fmaf_rn(dy, q1.y, tb_);                                   \
              const float pw = __fadd_rn(__fmul_rn(u0, dx), __fmul_rn(u1, dy));            \
              const float al = ex2_approx(fmaf(pw, 1.4426950408889634f, q1.z)); /* q1.z = log2(op) */ \
              const float ta = T * al;                                                     \
              T = T - ta;                                                                  \
              LIVE = LIVE && (T >= minw);                                                  \
              if (LIVE) { R = fmaf(ta, q1.w, R); G = fmaf(ta, q2.x, G); B = fmaf(ta, q2.y, B); } \
              if (kAux) { if (LIVE) { NC = nbase + (uint32_t)i; TF = T; } }                \
            }
            GSB_PIXEL_STEP(fy0, T0, l0, r0, g0, b0, n0, tf0)
            GSB_PIXEL_STEP(fy1, T1, l1, r1, g1, b1, n1, tf1)
            GSB_PIXEL_STEP(fy2, T2, l2, r2, g2, b2, n2, tf2)
            GSB_PIXEL_STEP(fy3, T3, l3, r3, g3, b3, n3, tf3)
#undef GSB_PIXEL_STEP
          }
          if (!__any_sync(0xffffffffu, l0 || l1 || l2 || l3)) { warp_live = false; m = 0u; }
        }
        if (!warp_live) break;
      }
    }
  }
  cp_async_wait<0>();
  if (px < a.width) {
    float* o = image + ((size_t)py0 * a.width + px) * 3;
    const size_t row = (size_t)a.width * 3;
    if (py0 < a.height) { o[0] = r0; o[1] = g0; o[2] = b0; }
    if (py0 + 1 < a.height) { o[row] = r1; o[row + 1] = g1; o[row + 2] = b1; }
    if (py0 + 2 < a.height) { o[2 * row] = r2; o[2 * row + 1] = g2; o[2 * row + 2] = b2; }
    if (py0 + 3 < a.height) { o[3 * row] = r3; o[3 * row + 1] = g3; o[3 * row + 2] = b3; }
    if (kAux) {
      const size_t p = (size_t)py0 * a.width + px, w = (size_t)a.width;
      if (py0 < a.height) { aux_t[p] = tf0; aux_n[p] = n0; }
      if (py0 + 1 < a.height) { aux_t[p + w] = tf1; aux_n[p + w] = n1; }
      if (py0 + 2 < a.height) { aux_t[p + 2 * w] = tf2; aux_n[p + 2 * w] = n2; }
      if (py0 + 3 < a.height) { aux_t[p + 3 * w] = tf3; aux_n[p + 3 * w] = n3; }
    }
  }
}


}  // namespace

static float cull_threshold_log2(const GsbParams& prm) {
  // cull_alpha < 0: no warp-level skipping; 0: skip only what is exactly zero in fp32 (ex2.approx.ftz flushes below
  // 2^-126; one binade of slack for its own rounding); > 0: skip when every alpha of the warp is below it
  if (prm.cull_alpha < 0.f) return -INFINITY;
  if (prm.cull_alpha == 0.f) return -127.f;
  const float l = log2f(prm.cull_alpha);
  return l < -127.f ? -127.f : l;
}

int launch_composite(const uint2* ranges, const uint32_t* payload, const float4* rec, float* image,
                     FrameGeom geom, const GsbParams& prm, float* aux_t, uint32_t* aux_n, const uint32_t* abort,
                     cudaStream_t st) {
  const int tiles = geom.tiles_x * geom.tiles_y;
  if (tiles <= 0) return 0;
  CompositeArgs a{geom.width, geom.height, geom.tiles_x, geom.tiles_y, prm.min_weight, prm.alpha_max,
                  cull_threshold_log2(prm)};
  if (aux_t && aux_n)
    composite_fast_kernel<true><<<tiles, kFastThreads, 0, st>>>(ranges, payload, rec, image, aux_t, aux_n, abort, a);
  else
    composite_fast_kernel<false><<<tiles, kFastThreads, 0, st>>>(ranges, payload, rec, image, nullptr, nullptr, abort, a);
  return (int)cudaGetLastError();
}

int launch_composite_cu(const uint2* ranges, const uint32_t* payload, const float4* rec, const float4* bbox,
                        float* image, FrameGeom geom, const GsbParams& prm, const uint32_t* abort, cudaStream_t st) {
  const int tiles = geom.tiles_x * geom.tiles_y;
  if (tiles <= 0) return 0;
  CompositeArgs a{geom.width, geom.height, geom.tiles_x, geom.tiles_y, prm.min_weight, prm.alpha_max, -INFINITY};
  composite_kernel<GSB_SEM_REF_CU><<<tiles, 256, 0, st>>>(ranges, payload, rec, bbox, image, abort, a);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// ingest of the reference op's own arguments (gsb_render_image; splat/c/render.cu:90-101): rows are
// already depth-sorted, so the row index is the depth rank and becomes the low key word.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int floor_div_i(int a, int b) {
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

__global__ void __launch_bounds__(256)
ingest_kernel(int64_t m, const float* __restrict__ means, const float* __restrict__ colors,
              const float* __restrict__ conic, const float* __restrict__ min_x, const float* __restrict__ max_x,
              const float* __restrict__ min_y, const float* __restrict__ max_y, const float* __restrict__ opacity,
              FrameGeom geom, int sem, int T, uint32_t* __restrict__ depth_key, float4* __restrict__ rec,
              float4* __restrict__ bbox, ushort4* __restrict__ rect, uint32_t* __restrict__ count,
              int32_t* __restrict__ diff_grid, int32_t* __restrict__ super_grid, SuperGeom sg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float mx = means[2 * i], my = means[2 * i + 1];
  const float i00 = conic[4 * i], i01 = conic[4 * i + 1], i10 = conic[4 * i + 2], i11 = conic[4 * i + 3];
  const float mnx = min_x[i], mxx = max_x[i], mny = min_y[i], mxy = max_y[i];
  const float op = opacity[i];
  int tx0, tx1, ty0, ty1;
  bool nan = !(mnx == mnx) || !(mxx == mxx) || !(mny == mny) || !(mxy == mxy);
  const float big = 1073741824.0f;
  float op_used;
  if (sem == GSB_SEM_REF_CU) {
    // render.cu:55-60: pixel p is a candidate iff min <= p <= max (inclusive, per pixel); mean -> int (:8-9);
    // power = dx*a*dx + 2*dx*dy*b + dy*dy*c with inv[0], inv[1], inv[3] (:17,:66-68)  ==  d^T [[a,b],[b,c]] d
    rec[3 * i + 0] = make_float4(truncf(mx), truncf(my), -0.5f * i00, -0.5f * i01);
    rec[3 * i + 1] = make_float4(-0.5f * i01, -0.5f * i11, op, colors[3 * i]);
    op_used = op;
    int x0 = (int)ceilf(fminf(fmaxf(mnx, -big), big)), x1 = (int)floorf(fminf(fmaxf(mxx, -big), big));
    int y0 = (int)ceilf(fminf(fmaxf(mny, -big), big)), y1 = (int)floorf(fminf(fmaxf(mxy, -big), big));
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, geom.width - 1); y1 = min(y1, geom.height - 1);
    tx0 = x0 / T; tx1 = x1 >= x0 ? x1 / T : -1; ty0 = y0 / T; ty1 = y1 >= y0 ? y1 / T : -1;
    if (x1 < x0) { tx0 = 0; tx1 = -1; }
    if (y1 < y0) { ty0 = 0; ty1 = -1; }
  } else {
    const float op2 = 1.0f / (1.0f + expf(-op));  // the CPU path applies a second sigmoid (:164)
    rec[3 * i + 0] = make_float4(mx, my, -0.5f * i00, -0.5f * i01);
    rec[3 * i + 1] = make_float4(-0.5f * i10, -0.5f * i11, log2f(op2), colors[3 * i]);  // log2: see composite_fast_kernel
    op_used = op2;
    int imn = (int)fminf(fmaxf(mnx, -big), big), imx = (int)fminf(fmaxf(mxx, -big), big);
    tx0 = max(floor_div_i(imn - 1, T), 0); tx1 = min(floor_div_i(imx, T), geom.tiles_x - 1);
    imn = (int)fminf(fmaxf(mny, -big), big); imx = (int)fminf(fmaxf(mxy, -big), big);
    ty0 = max(floor_div_i(imn - 1, T), 0); ty1 = min(floor_div_i(imx, T), geom.tiles_y - 1);
  }
  rec[3 * i + 2] = make_float4(colors[3 * i + 1], colors[3 * i + 2], 0.f, op_used);
  bbox[i] = make_float4(mnx, mny, mxx, mxy);
  uint32_t cnt = 0;
  if (!nan && tx1 >= tx0 && ty1 >= ty0) cnt = (uint32_t)(tx1 - tx0 + 1) * (uint32_t)(ty1 - ty0 + 1);
  rect[i] = cnt ? make_ushort4((unsigned short)tx0, (unsigned short)tx1, (unsigned short)ty0, (unsigned short)ty1)
                : make_ushort4(1, 0, 1, 0);  // tx1 < tx0: touches no tile
  count[i] = cnt;
  depth_key[i] = (uint32_t)i;
  if (cnt) {  // same 2-D difference grids as the projection kernel (tile_stats_kernel turns them into ranges)
    const int gw = geom.tiles_x + 1;
    atomicAdd(&diff_grid[ty0 * gw + tx0], 1);
    atomicAdd(&diff_grid[ty0 * gw + tx1 + 1], -1);
    atomicAdd(&diff_grid[(ty1 + 1) * gw + tx0], -1);
    atomicAdd(&diff_grid[(ty1 + 1) * gw + tx1 + 1], 1);
    if (super_grid) {
      const int sx0 = tx0 >> sg.lw, sx1 = (tx1 >> sg.lw) + 1, sy0 = ty0 >> sg.lh, sy1 = (ty1 >> sg.lh) + 1;
      const int gs = sg.nx + 1;
      atomicAdd(&super_grid[sy0 * gs + sx0], 1);
      atomicAdd(&super_grid[sy0 * gs + sx1], -1);
      atomicAdd(&super_grid[sy1 * gs + sx0], -1);
      atomicAdd(&super_grid[sy1 * gs + sx1], 1);
    }
  }
}

int launch_ingest_preprocessed(int64_t m, const float* means, const float* colors, const float* conic,
                               const float* min_x, const float* max_x, const float* min_y, const float* max_y,
                               const float* opacity, FrameGeom geom, const GsbParams& prm, uint32_t* depth_key,
                               float4* rec, float4* bbox, ushort4* rect, uint32_t* count, int32_t* diff_grid,
                               int32_t* super_grid, SuperGeom sg, cudaStream_t st) {
  if (m == 0) return 0;
  ingest_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, means, colors, conic, min_x, max_x, min_y, max_y,
                                                            opacity, geom, prm.semantics, prm.tile_size, depth_key,
                                                            rec, bbox, rect, count, diff_grid, super_grid, sg);
  return (int)cudaGetLastError();
}

}  // namespace gsb
