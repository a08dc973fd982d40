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
};

template <int kSem>
__global__ void __launch_bounds__(256)
composite_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ payload,
                 const float4* __restrict__ rec, const float4* __restrict__ bbox, float* __restrict__ image,
                 const __grid_constant__ CompositeArgs a) {
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
// Slots past the end of the list are filled with a null record (log2 opacity -inf => alpha = 0 => exact no-op)
// so the blend loop needs no tail handling.
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

// kAux (save_for_backward): also records, per pixel, how many Gaussians were blended (the list prefix [0, n)) and
// the transmittance after the last of them -- what the back-to-front gradient pass starts from (backward.cu).
template <bool kAux>
__global__ void __launch_bounds__(kFastThreads)
composite_fast_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ payload,
                      const float4* __restrict__ rec, float* __restrict__ image, float* __restrict__ aux_t,
                      uint32_t* __restrict__ aux_n, const __grid_constant__ CompositeArgs a) {
  __shared__ __align__(16) float4 sm[2][kFastBatch * 3];

  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int px = tx * kTile + (lane & 15);
  const int py0 = ty * kTile + warp * 8 + (lane >> 4) * 4;
  const float fx = (float)px;
  const float fy0 = (float)py0, fy1 = (float)(py0 + 1), fy2 = (float)(py0 + 2), fy3 = (float)(py0 + 3);
  const float minw = a.min_weight;

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

  const float4 null0 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 null1 = make_float4(0.f, 0.f, -INFINITY, 0.f);  // log2(opacity) = -inf => alpha = 0 => exact no-op
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
      if (idx[j] != 0xFFFFFFFFu) {
        const float4* src = rec + 3 * (size_t)idx[j];
        cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
      } else {
        dst[0] = null0; dst[1] = null1; dst[2] = null0;
      }
    }
    cp_async_commit();
  };

  const uint32_t nb = (len + kFastBatch - 1) / kFastBatch;
  if (nb > 0) { load_idx(0); stage(0); }
  if (nb > 1) load_idx(1);
  bool warp_live = true;
  for (uint32_t b = 0; b < nb; ++b) {
    const int buf = (int)(b & 1);
    cp_async_wait<0>();
    __syncthreads();  // batch b visible to all; everyone is done reading the other buffer
    if (b + 1 < nb) {
      stage(buf ^ 1);
      if (b + 2 < nb) load_idx(b + 2);
    }
    if (warp_live) {
      const float4* p = sm[buf];
#pragma unroll 1
      for (int i = 0; i < kFastBatch; i += kFastUnroll) {
#pragma unroll
        for (int u = 0; u < kFastUnroll; ++u) {
          const float4 q0 = p[0];  // mx, my, a, b
          const float4 q1 = p[1];  // c, d, op, r
          const float2 q2 = *reinterpret_cast<const float2*>(p + 2);  // g, b
          p += 3;
          const float dx = q0.x - fx;
          const float ta_ = __fmul_rn(dx, q0.z), tb_ = __fmul_rn(dx, q0.w);
#define GSB_PIXEL_STEP(FY, T, LIVE, R, G, B, NC, TF)                                       \
          {                                                                                \
            const float dy = q0.y - FY;                                                    \
            const float u0 = __fmaf_rn(dy, q1.x, ta_);                                     \
            const float u1 = __fmaf_rn(dy, q1.y, tb_);                                     \
            const float pw = __fadd_rn(__fmul_rn(u0, dx), __fmul_rn(u1, dy));              \
            const float al = ex2_approx(fmaf(pw, 1.4426950408889634f, q1.z)); /* q1.z = log2(op) */ \
            const float ta = T * al;                                                       \
            T = T - ta;                                                                    \
            LIVE = LIVE && (T >= minw);                                                    \
            if (LIVE) { R = fmaf(ta, q1.w, R); G = fmaf(ta, q2.x, G); B = fmaf(ta, q2.y, B); } \
            if (kAux) { if (LIVE) { NC += 1; TF = T; } }                                   \
          }
          GSB_PIXEL_STEP(fy0, T0, l0, r0, g0, b0, n0, tf0)
          GSB_PIXEL_STEP(fy1, T1, l1, r1, g1, b1, n1, tf1)
          GSB_PIXEL_STEP(fy2, T2, l2, r2, g2, b2, n2, tf2)
          GSB_PIXEL_STEP(fy3, T3, l3, r3, g3, b3, n3, tf3)
#undef GSB_PIXEL_STEP
        }
        if (!__any_sync(0xffffffffu, l0 || l1 || l2 || l3)) { warp_live = false; break; }
      }
    }
    if (!__syncthreads_or(warp_live)) break;  // also orders this batch's reads before the next overwrite
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
      // a null record (tail padding of the last batch) passes the LIVE test with alpha = 0: clamp to the list
      const size_t p = (size_t)py0 * a.width + px, w = (size_t)a.width;
      if (py0 < a.height) { aux_t[p] = tf0; aux_n[p] = min(n0, len); }
      if (py0 + 1 < a.height) { aux_t[p + w] = tf1; aux_n[p + w] = min(n1, len); }
      if (py0 + 2 < a.height) { aux_t[p + 2 * w] = tf2; aux_n[p + 2 * w] = min(n2, len); }
      if (py0 + 3 < a.height) { aux_t[p + 3 * w] = tf3; aux_n[p + 3 * w] = min(n3, len); }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// REF_CPU path, packed-fp32 variant (Blackwell): the same 64-thread / 1x4-pixel-column mapping as
// composite_fast_kernel, but the four pixels are processed as TWO PAIRS with the sm_100 dual-fp32
// instructions (FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2 ...).  Each packed op performs two IEEE-rn
// operations, bit-identical to two scalar ones, for one issue slot -- and issue slots are what bound this
// kernel (ncu: issue active 85 %, FMA pipe 65 %, DRAM 1 %).  Per pixel-step: ~12.5 SASS instructions
// instead of 17.4.
//
// Packed operands must sit in aligned register pairs, so every per-Gaussian scalar that multiplies a pixel
// pair is stored DUPLICATED in shared memory: a record is 5 x float4
//     {mx,mx,my,my} {a,a,b,b} {c,c,d,d} {op,op,r,r} {g,g,bl,bl}            (a b; c d) = -0.5 * inv cov
// read with five broadcast LDS.128 per thread-step.  cp.async cannot duplicate, so staging goes through
// registers (one record per thread per batch of 64, prefetched one batch ahead, double-buffered in smem).
//
// Termination: live &= (T >= min_weight) per pixel as before; the colour FMAs are packed, so instead of
// predicating them the blend weight is zeroed with FSEL when the pixel is no longer live.
// ------------------------------------------------------------------------------------------------
constexpr int kPkThreads = 64;
constexpr int kPkBatch = 64;

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

__global__ void __launch_bounds__(kPkThreads)
composite_packed_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ payload,
                        const float4* __restrict__ rec, float* __restrict__ image,
                        const __grid_constant__ CompositeArgs a) {
  __shared__ __align__(16) float4 sm[2][kPkBatch * 5];

  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int px = tx * kTile + (lane & 15);
  const int py0 = ty * kTile + warp * 8 + (lane >> 4) * 4;
  const float2 nfx = f2(-(float)px, -(float)px);
  const float2 nfyA = f2(-(float)py0, -(float)(py0 + 1));
  const float2 nfyB = f2(-(float)(py0 + 2), -(float)(py0 + 3));
  const float2 kLog2e = f2(1.4426950408889634f, 1.4426950408889634f);
  const float2 kNeg1 = f2(-1.f, -1.f);
  const float minw = a.min_weight;

  const uint2 rg = ranges[tile];
  const uint32_t len = rg.y - rg.x;
  const uint32_t* pl = payload + rg.x;

  bool l0 = px < a.width && py0 < a.height, l1 = px < a.width && py0 + 1 < a.height;
  bool l2 = px < a.width && py0 + 2 < a.height, l3 = px < a.width && py0 + 3 < a.height;
  float2 TA = f2(1.f, 1.f), TB = f2(1.f, 1.f);
  float2 RA = f2(0.f, 0.f), GA = RA, BA = RA, RB = RA, GB = RA, BB = RA;

  // register-staged prefetch: this thread's record of the next batch
  float4 p0 = make_float4(0, 0, 0, 0), p1 = make_float4(0.f, 0.f, -INFINITY, 0.f), p2 = p0;
  auto fetch = [&](uint32_t b) {
    const uint32_t slot = b * kPkBatch + tid;
    p0 = make_float4(0, 0, 0, 0); p2 = p0;
    p1 = make_float4(0.f, 0.f, -INFINITY, 0.f);  // null record: log2(opacity) = -inf => alpha 0 => exact no-op
    if (slot < len) {
      const float4* src = rec + 3 * (size_t)pl[slot];
      p0 = src[0]; p1 = src[1]; p2 = src[2];
    }
  };
  auto stash = [&](int buf) {
    float4* d = &sm[buf][tid * 5];
    d[0] = make_float4(p0.x, p0.x, p0.y, p0.y);  // mx mx my my
    d[1] = make_float4(p0.z, p0.z, p0.w, p0.w);  // a a b b
    d[2] = make_float4(p1.x, p1.x, p1.y, p1.y);  // c c d d
    d[3] = make_float4(p1.z, p1.z, p1.w, p1.w);  // op op r r
    d[4] = make_float4(p2.x, p2.x, p2.y, p2.y);  // g g bl bl
  };

  const uint32_t nb = (len + kPkBatch - 1) / kPkBatch;
  if (nb > 0) { fetch(0); stash(0); }
  if (nb > 1) fetch(1);
  bool warp_live = true;
  for (uint32_t b = 0; b < nb; ++b) {
    const int buf = (int)(b & 1);
    __syncthreads();  // batch b visible; everyone is done reading the other buffer
    if (b + 1 < nb) {
      stash(buf ^ 1);
      if (b + 2 < nb) fetch(b + 2);
    }
    if (warp_live) {
      const float4* p = sm[buf];
#pragma unroll 1
      for (int i = 0; i < kPkBatch; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3], q4 = p[4];
          p += 5;
          const float2 dx = __fadd2_rn(f2(q0.x, q0.y), nfx);            // (dx, dx)
          const float2 ta_ = __fmul2_rn(dx, f2(q1.x, q1.y));            // (dx*a, dx*a)
          const float2 tb_ = __fmul2_rn(dx, f2(q1.z, q1.w));            // (dx*b, dx*b)
          const float2 my = f2(q0.z, q0.w), cc = f2(q2.x, q2.y), dd = f2(q2.z, q2.w);
          const float2 op = f2(q3.x, q3.y), cr = f2(q3.z, q3.w), cg = f2(q4.x, q4.y), cb = f2(q4.z, q4.w);
#define GSB_PAIR_STEP(NFY, T, LA, LB, R, G, B)                                                   \
          {                                                                                      \
            const float2 dy = __fadd2_rn(my, NFY);                                               \
            const float2 u0 = __ffma2_rn(dy, cc, ta_);                                           \
            const float2 u1 = __ffma2_rn(dy, dd, tb_);                                           \
            /* ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (seen in SASS, even with */ \
            /* --fmad=false); the reference sum is UNFUSED, so the add is done with scalar FADDs */ \
            const float2 m0 = __fmul2_rn(u0, dx), m1 = __fmul2_rn(u1, dy);                       \
            const float2 pw = f2(__fadd_rn(m0.x, m1.x), __fadd_rn(m0.y, m1.y));                  \
            const float2 e = __ffma2_rn(pw, kLog2e, op); /* op = (log2(op), log2(op)) */          \
            const float2 al = f2(ex2_approx(e.x), ex2_approx(e.y));                              \
            float2 ta = __fmul2_rn(T, al);                                                       \
            T = __ffma2_rn(ta, kNeg1, T);                                                        \
            LA = LA && (T.x >= minw);                                                            \
            LB = LB && (T.y >= minw);                                                            \
            ta.x = LA ? ta.x : 0.f;                                                              \
            ta.y = LB ? ta.y : 0.f;                                                              \
            R = __ffma2_rn(ta, cr, R); G = __ffma2_rn(ta, cg, G); B = __ffma2_rn(ta, cb, B);     \
          }
          GSB_PAIR_STEP(nfyA, TA, l0, l1, RA, GA, BA)
          GSB_PAIR_STEP(nfyB, TB, l2, l3, RB, GB, BB)
#undef GSB_PAIR_STEP
        }
        if (!__any_sync(0xffffffffu, l0 || l1 || l2 || l3)) { warp_live = false; break; }
      }
    }
    if (!__syncthreads_or(warp_live)) break;
  }
  if (px < a.width) {
    float* o = image + ((size_t)py0 * a.width + px) * 3;
    const size_t row = (size_t)a.width * 3;
    if (py0 < a.height) { o[0] = RA.x; o[1] = GA.x; o[2] = BA.x; }
    if (py0 + 1 < a.height) { o[row] = RA.y; o[row + 1] = GA.y; o[row + 2] = BA.y; }
    if (py0 + 2 < a.height) { o[2 * row] = RB.x; o[2 * row + 1] = GB.x; o[2 * row + 2] = BB.x; }
    if (py0 + 3 < a.height) { o[3 * row] = RB.y; o[3 * row + 1] = GB.y; o[3 * row + 2] = BB.y; }
  }
}


}  // namespace

int launch_composite(const uint2* ranges, const uint32_t* payload, const float4* rec, float* image,
                     FrameGeom geom, const GsbParams& prm, float* aux_t, uint32_t* aux_n, cudaStream_t st) {
  const int tiles = geom.tiles_x * geom.tiles_y;
  if (tiles <= 0) return 0;
  CompositeArgs a{geom.width, geom.height, geom.tiles_x, geom.tiles_y, prm.min_weight, prm.alpha_max};
  // GSB_COMPOSITE=2 selects the packed-fp32x2 variant (A/B knob).  Measured on B200, config 3: 432 us vs 424 us
  // for the scalar kernel -- FFMA2/FMUL2/FADD2 halve the issue slots (289 M vs 387 M warp instructions) but
  // occupy the FMA pipe for two cycles each, so the pipe-bound time is unchanged (profiles/r1_summary.md).
  static const int variant = [] { const char* e = std::getenv("GSB_COMPOSITE"); return e ? std::atoi(e) : 1; }();
  if (aux_t && aux_n)
    composite_fast_kernel<true><<<tiles, kFastThreads, 0, st>>>(ranges, payload, rec, image, aux_t, aux_n, a);
  else if (variant == 2)
    composite_packed_kernel<<<tiles, kPkThreads, 0, st>>>(ranges, payload, rec, image, a);
  else
    composite_fast_kernel<false><<<tiles, kFastThreads, 0, st>>>(ranges, payload, rec, image, nullptr, nullptr, a);
  return (int)cudaGetLastError();
}

int launch_composite_cu(const uint2* ranges, const uint32_t* payload, const float4* rec, const float4* bbox,
                        float* image, FrameGeom geom, const GsbParams& prm, cudaStream_t st) {
  const int tiles = geom.tiles_x * geom.tiles_y;
  if (tiles <= 0) return 0;
  CompositeArgs a{geom.width, geom.height, geom.tiles_x, geom.tiles_y, prm.min_weight, prm.alpha_max};
  composite_kernel<GSB_SEM_REF_CU><<<tiles, 256, 0, st>>>(ranges, payload, rec, bbox, image, a);
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
              int32_t* __restrict__ diff_grid) {
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
                : make_ushort4(0, 0, 0, 0);
  count[i] = cnt;
  depth_key[i] = (uint32_t)i;
  if (cnt) {  // same 2-D difference grid as the projection kernel (tile_stats_kernel turns it into ranges)
    const int gw = geom.tiles_x + 1;
    atomicAdd(&diff_grid[ty0 * gw + tx0], 1);
    atomicAdd(&diff_grid[ty0 * gw + tx1 + 1], -1);
    atomicAdd(&diff_grid[(ty1 + 1) * gw + tx0], -1);
    atomicAdd(&diff_grid[(ty1 + 1) * gw + tx1 + 1], 1);
  }
}

int launch_ingest_preprocessed(int64_t m, const float* means, const float* colors, const float* conic,
                               const float* min_x, const float* max_x, const float* min_y, const float* max_y,
                               const float* opacity, FrameGeom geom, const GsbParams& prm, uint32_t* depth_key,
                               float4* rec, float4* bbox, ushort4* rect, uint32_t* count, int32_t* diff_grid,
                               cudaStream_t st) {
  if (m == 0) return 0;
  ingest_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, means, colors, conic, min_x, max_x, min_y, max_y,
                                                            opacity, geom, prm.semantics, prm.tile_size, depth_key,
                                                            rec, bbox, rect, count, diff_grid);
  return (int)cudaGetLastError();
}

}  // namespace gsb
