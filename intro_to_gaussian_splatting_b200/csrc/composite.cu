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
// Two kernels:
//   composite_fast_kernel   REF_CPU, the frame path: 64 threads per tile, four pixels per thread, list filtered from
//                           the super-tile's list on the fly, per-warp culling (documented at the kernel)
//   composite_kernel<REF_CU> render.cu's arithmetic over materialised per-tile lists: one thread per pixel, 8x4 pixels
//                           per warp, batches of 256 records gathered through registers into shared memory
//
// Roofline: issue slots (fp32 FMA/ALU + MUFU.EX2 out of shared memory), not HBM: 16.4 SASS instructions per
// executed (pixel, Gaussian) step in the fast kernel; DRAM traffic 22 MB per 1080p frame (the records are L2 resident).
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
// Where the tile's list comes from (kMasked).  SPLIT mode never stores per-tile lists: the tile reads the list of
// its SUPER-TILE (8 x 4 tiles; entries {Gaussian index, 32-bit tile mask} in depth order, written by the last radix
// pass) and keeps the entries whose mask has the tile's bit -- 256 entries per round, four ballots per warp, the
// survivors' indices go into a small ring in shared memory.  Early termination therefore also ends the binning
// work: a tile that saturates after 350 Gaussians never looks at the rest of its super-tile's list, where a
// separate expansion pass would have written (and this kernel read back) all of it.  FULL mode and one-level SPLIT
// feed the same ring from a materialised per-tile payload array (every entry a hit).
//
// Staging: the ring's next 128 indices are turned into records with cp.async (LDGSTS, 3 x 16 B per record, no
// register staging) into a double buffer, one batch ahead of the blend loop; the entries of the filter round after
// that are prefetched into registers across the blend loop.
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
// batch; each warp compacts its mask into a byte list of record slots (padded to a multiple of four with the slot
// of a null record) and walks that list four at a time -- straight-line code with all twelve LDS of a group
// issued up front.  (Walking the mask bits directly costs a data-dependent branch per record and serialises the
// loads: measured 491 us against 409 us for the frame's compositing with nothing culled.)  With cull_alpha = 0
// only steps whose alpha is exactly zero are skipped (bit-identical frames); with cull_alpha = t > 0 a skipped step
// would have changed a pixel by less than t (T and live are untouched for alpha < 2^-25), so the frame differs by
// < t * (list length).
// ------------------------------------------------------------------------------------------------
constexpr int kFastThreads = 64;
constexpr int kFastBatch = 128;                        // records per stage
constexpr int kFastPerThread = kFastBatch / kFastThreads;  // records each thread stages
constexpr int kFastUnroll = 4;                         // Gaussians between two warp votes = slots per list word
#ifndef GSB_FILTER_PER
#define GSB_FILTER_PER 4
#endif
constexpr int kFilterPer = GSB_FILTER_PER;             // list entries each thread examines per filter round
constexpr int kFilterRound = kFilterPer * kFastThreads;  // entries per round
constexpr int kRing = 512;                             // survivor ring: < 128 pending + <= 256 from one round

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// -DGSB_STAGE_BULK: stage the records with cp.async.bulk (the TMA unit's 1-D bulk copy, one 48-byte copy per record,
// completion counted in bytes on an mbarrier per buffer) instead of three LDGSTS per record.  Built on request only
// (tools/build_variant.sh bulk -DGSB_STAGE_BULK): measured on B200, profiles/r2_summary.md.
#ifdef GSB_STAGE_BULK
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
  unsigned ok = 0;
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_copy48(void* smem_dst, const void* gmem_src, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 48, [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
#endif

// kAux (save_for_backward): also records, per pixel, the length of the list prefix that reached the pixel (index of
// the last blended Gaussian + 1) and the transmittance after it -- what the back-to-front gradient pass starts
// from (backward.cu) -- and, when the list is filtered on the fly (kMasked), writes the tile's list to
// src.payload_out as it is consumed: the gradient pass walks exactly that prefix.
// One (pixel, Gaussian) step in the reference's own rounding order (compute_gaussian_weight, splat/utils.py:363-364):
// ((-0.5 d) @ inv) is an FMA chain in k order, (.) @ d^T two rounded products and a rounded sum.  Expects q0, q1, q2,
// dx, ta_, tb_, minw, nbase, i in scope.
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
            if (kAux) { if (LIVE && i < (uint32_t)kFastBatch) { NC = nbase + i; TF = T; } }  \
          }

#ifndef GSB_FAST_MINBLOCKS
#define GSB_FAST_MINBLOCKS 9
#endif
template <bool kAux, bool kMasked>
__global__ void __launch_bounds__(kFastThreads, GSB_FAST_MINBLOCKS)
composite_fast_kernel(const __grid_constant__ TileSource src, const float4* __restrict__ rec, float* __restrict__ image,
                      float* __restrict__ aux_t, uint32_t* __restrict__ aux_n, const uint32_t* __restrict__ abort,
                      const __grid_constant__ CompositeArgs a) {
  __shared__ __align__(16) float4 sm[2][(kFastBatch + 1) * 3];      // record double buffer; slot 128 = null record
  __shared__ __align__(16) uint32_t s_mask[2][2][kFastBatch / 32];  // [buffer][warp footprint][survivor bits]
  __shared__ uint32_t s_ring[kRing];                                // Gaussian indices of the tile's list, in order
  __shared__ __align__(8) uint16_t s_list[2][kFastBatch + 8];       // per warp: shared-memory addresses (16 bits) of the
                                                                    // records that survive culling
  __shared__ uint32_t s_cnt[2][2 * kFilterPer];                     // filter round: hits per (entry slice, warp)
  __shared__ uint32_t s_dead[2];                                    // warp w has no live pixel left
#ifdef GSB_STAGE_BULK
  __shared__ __align__(8) uint64_t s_mbar[2];                       // one per record buffer, every thread arrives
#endif
  if (abort && *abort) return;  // the lists do not exist: the host re-queues the frame's tail (gsb_api.cu)

  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int px = tx * kTile + (lane & 15);
  const int py0 = ty * kTile + warp * 8 + (lane >> 4) * 4;
  const float fx = (float)px;
  const float fy0 = (float)py0, fy1 = (float)(py0 + 1), fy2 = (float)(py0 + 2), fy3 = (float)(py0 + 3);
  const float minw = a.min_weight;
  const float cull = a.cull_log2;
  const bool cull_on = cull > -INFINITY;
  // pixel rectangles of the two warps (columns shared)
  const float rx0 = (float)(tx * kTile), rx1 = rx0 + (float)(kTile - 1);
  const float ry0 = (float)(ty * kTile), ry1 = ry0 + 7.f, ry2 = ry0 + 8.f, ry3 = ry0 + 15.f;

  // the list this tile filters: its super-tile's entries (kMasked) or its own payload slice
  uint32_t len, my_bit = 0;
  const uint2* cl = nullptr;
  const uint32_t* pl = nullptr;
  uint32_t* out_list = nullptr;
  if (kMasked) {
    const int s = (ty >> src.lh) * src.snx + (tx >> src.lw);
    const uint2 rs = src.ranges_s[s];
    len = rs.y - rs.x;
    cl = src.clist + rs.x;
    my_bit = (uint32_t)(((ty & ((1 << src.lh) - 1)) << src.lw) | (tx & ((1 << src.lw) - 1)));
    if (kAux) out_list = src.payload_out + src.ranges[tile].x;
  } else {
    const uint2 rg = src.ranges[tile];
    len = rg.y - rg.x;
    pl = src.payload + rg.x;
  }
  if (tid < 2) {  // null record: log2(opacity) = -inf => alpha = 0 => an exact no-op for T, live and the colours
    float4* d = &sm[tid][kFastBatch * 3];
    d[0] = make_float4(0.f, 0.f, 0.f, 0.f); d[1] = make_float4(0.f, 0.f, -INFINITY, 0.f); d[2] = d[0];
    s_dead[tid] = 0u;
#ifdef GSB_STAGE_BULK
    mbar_init(&s_mbar[tid], kFastThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  }

  bool l0 = px < a.width && py0 < a.height, l1 = px < a.width && py0 + 1 < a.height;
  bool l2 = px < a.width && py0 + 2 < a.height, l3 = px < a.width && py0 + 3 < a.height;
  float T0 = 1.f, T1 = 1.f, T2 = 1.f, T3 = 1.f;
  float r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;
  float r2 = 0.f, g2 = 0.f, b2 = 0.f, r3 = 0.f, g3 = 0.f, b3 = 0.f;
  uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;          // kAux only
  float tf0 = 1.f, tf1 = 1.f, tf2 = 1.f, tf3 = 1.f;  // kAux only

  // ---- list filter: entries pos .. pos+127 per round, hits appended to the ring ----
  uint32_t pos = 0, head = 0, tail = 0;  // entries examined; ring indices consumed / produced (uniform over the CTA)
  uint32_t e_g[kFilterPer], e_hit[kFilterPer];
  auto prefetch = [&]() {  // this thread's entries of the next round: pos + 64 j + tid
#pragma unroll
    for (int j = 0; j < kFilterPer; ++j) {
      const uint32_t i = pos + j * kFastThreads + tid;
      e_g[j] = 0u; e_hit[j] = 0u;
      if (i < len) {
        if (kMasked) { const uint2 e = cl[i]; e_g[j] = e.x; e_hit[j] = (e.y >> my_bit) & 1u; }
        else { e_g[j] = pl[i]; e_hit[j] = 1u; }
      }
    }
  };
  int parity = 0;
  auto fill = [&]() {  // until a full batch is pending or the list is exhausted
    while (tail - head < (uint32_t)kFastBatch && pos < len) {
      unsigned h[kFilterPer];
#pragma unroll
      for (int j = 0; j < kFilterPer; ++j) {
        h[j] = __ballot_sync(0xffffffffu, e_hit[j] != 0u);
        if (lane == 0) s_cnt[parity][2 * j + warp] = (uint32_t)__popc(h[j]);  // list order: slice j, then warp
      }
      __syncthreads();
      uint32_t o = tail;
#pragma unroll
      for (int j = 0; j < kFilterPer; ++j) {
        const uint32_t c0 = s_cnt[parity][2 * j], c1 = s_cnt[parity][2 * j + 1];
        if (e_hit[j]) s_ring[(o + (warp ? c0 : 0u) + (uint32_t)__popc(h[j] & lt_mask)) & (kRing - 1)] = e_g[j];
        o += c0 + c1;
      }
      tail = o;
      pos += (uint32_t)kFilterRound;
      parity ^= 1;
      prefetch();
    }
  };
  // stage the ring's next `cnt` indices as records of buffer `buf`; list_pos = position of the batch in the tile's list
  auto stage = [&](int buf, uint32_t cnt, uint32_t list_pos) {
#ifdef GSB_STAGE_BULK
    unsigned mine = 0;
#pragma unroll
    for (int j = 0; j < kFastPerThread; ++j)
      if ((uint32_t)(j * kFastThreads + tid) < cnt) mine += 48u;
    mbar_arrive_expect_tx(&s_mbar[buf], mine);  // every thread arrives once per use of the buffer
#endif
#pragma unroll
    for (int j = 0; j < kFastPerThread; ++j) {
      const uint32_t r = (uint32_t)(j * kFastThreads + tid);
      if (r < cnt) {  // slots past the batch are never read: their survivor bit is 0
        const uint32_t g = s_ring[(head + r) & (kRing - 1)];
        const float4* s3 = rec + 3 * (size_t)g;
        float4* dst = &sm[buf][r * 3];
#ifdef GSB_STAGE_BULK
        bulk_copy48(dst, s3, &s_mbar[buf]);
#else
        cp_async16(dst, s3); cp_async16(dst + 1, s3 + 1); cp_async16(dst + 2, s3 + 2);
#endif
        if (kAux && kMasked) out_list[list_pos + r] = g;
      }
    }
#ifndef GSB_STAGE_BULK
    cp_async_commit();
#endif
  };
  // survivor masks of the batch resident in buffer `buf`: record slot j*64 + tid is bit `lane` of word 2j + warp
  auto build_masks = [&](int buf, uint32_t cnt) {
    const bool dead0 = s_dead[0] != 0u, dead1 = s_dead[1] != 0u;  // written before the barrier at the loop top
#pragma unroll
    for (int j = 0; j < kFastPerThread; ++j) {
      const uint32_t r = (uint32_t)(j * kFastThreads + tid);
      bool k0 = r < cnt && !dead0, k1 = r < cnt && !dead1;  // a warp without live pixels reads no list
      if (cull_on && r < cnt) {
        const float4 q0 = sm[buf][r * 3], q1 = sm[buf][r * 3 + 1];
        const float dxl = q0.x - rx1, dxh = q0.x - rx0;
        if (k0) k0 = may_contribute(q0.z, q0.w, q1.x, q1.y, q1.z, dxl, dxh, q0.y - ry1, q0.y - ry0, cull);
        if (k1) k1 = may_contribute(q0.z, q0.w, q1.x, q1.y, q1.z, dxl, dxh, q0.y - ry3, q0.y - ry2, cull);
      }
      const unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
      if (lane == 0) { s_mask[buf][0][2 * j + warp] = m0; s_mask[buf][1][2 * j + warp] = m1; }
    }
  };

  prefetch();
  fill();
  __syncthreads();  // ring (and the null records) visible
  uint32_t cnt_cur = min(tail - head, (uint32_t)kFastBatch), consumed = 0;
  stage(0, cnt_cur, 0u);
  head += cnt_cur;
  bool warp_live = true;
  for (uint32_t b = 0; cnt_cur > 0u; ++b) {
    const int buf = (int)(b & 1u);
#ifdef GSB_STAGE_BULK
    mbar_wait(&s_mbar[buf], (b >> 1) & 1u);
#else
    cp_async_wait<0>();
#endif
    // batch b visible to all; everyone is done with the other buffer and with the ring slots of batch b;
    // stop when no pixel of the tile is live
    if (!__syncthreads_or(warp_live)) break;
    fill();
    build_masks(buf, cnt_cur);
    __syncthreads();  // masks + ring visible
    const uint32_t cnt_next = min(tail - head, (uint32_t)kFastBatch);
    stage(buf ^ 1, cnt_next, consumed + cnt_cur);
    head += cnt_next;
    if (warp_live) {
      // compact this warp's survivor mask into a list of record slots
      const uint4 mk = *reinterpret_cast<const uint4*>(s_mask[buf][warp]);
      const uint32_t p1 = (uint32_t)__popc(mk.x), p2 = p1 + (uint32_t)__popc(mk.y), p3 = p2 + (uint32_t)__popc(mk.z);
      const uint32_t total = p3 + (uint32_t)__popc(mk.w);
      // list entries are the records' 16-bit shared-memory addresses: the loop below turns an entry into three LDS
      // with one extract, no multiply, no base register
      uint16_t* lst = s_list[warp];
      const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm[buf]);
      if ((mk.x >> lane) & 1u) lst[__popc(mk.x & lt_mask)] = (uint16_t)(sbase + 48u * (uint32_t)lane);
      if ((mk.y >> lane) & 1u) lst[p1 + __popc(mk.y & lt_mask)] = (uint16_t)(sbase + 48u * (uint32_t)(32 + lane));
      if ((mk.z >> lane) & 1u) lst[p2 + __popc(mk.z & lt_mask)] = (uint16_t)(sbase + 48u * (uint32_t)(64 + lane));
      if ((mk.w >> lane) & 1u) lst[p3 + __popc(mk.w & lt_mask)] = (uint16_t)(sbase + 48u * (uint32_t)(96 + lane));
      if (lane < 4) lst[total + lane] = (uint16_t)(sbase + 48u * (uint32_t)kFastBatch);  // pad with the null record
      __syncwarp();
      const uint32_t nbase = consumed + 1u;  // kAux: list position + 1 of record slot 0
#pragma unroll 1
      for (uint32_t j = 0; j < total; j += kFastUnroll) {
        const uint2 four = *reinterpret_cast<const uint2*>(lst + j);
#pragma unroll
        for (int u = 0; u < kFastUnroll; ++u) {
          const uint32_t w2 = (u & 2) ? four.y : four.x;
          const uint32_t addr = (u & 1) ? (w2 >> 16) : (w2 & 0xFFFFu);
          const uint32_t i = (addr - sbase) / 48u;  // record slot (kAux only)
          float4 q0, q1;  // mx, my, a, b | c, d, log2(op), r
          float2 q2;      // g, b
          asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(addr));
          asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "r"(addr));
          asm("ld.shared.v2.f32 {%0,%1}, [%2+32];" : "=f"(q2.x), "=f"(q2.y) : "r"(addr));
          const float dx = q0.x - fx;
          const float ta_ = __fmul_rn(dx, q0.z), tb_ = __fmul_rn(dx, q0.w);
          GSB_PIXEL_STEP(fy0, T0, l0, r0, g0, b0, n0, tf0)
          GSB_PIXEL_STEP(fy1, T1, l1, r1, g1, b1, n1, tf1)
          GSB_PIXEL_STEP(fy2, T2, l2, r2, g2, b2, n2, tf2)
          GSB_PIXEL_STEP(fy3, T3, l3, r3, g3, b3, n3, tf3)
        }
        if (!__any_sync(0xffffffffu, l0 || l1 || l2 || l3)) {
          warp_live = false;
          if (lane == 0) s_dead[warp] = 1u;
          break;
        }
      }
    }
    consumed += cnt_cur;
    cnt_cur = cnt_next;
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


int launch_composite(const TileSource& src, const float4* rec, float* image, FrameGeom geom, const GsbParams& prm,
                     float* aux_t, uint32_t* aux_n, const uint32_t* abort, cudaStream_t st) {
  const int tiles = geom.tiles_x * geom.tiles_y;
  if (tiles <= 0) return 0;
  CompositeArgs a{geom.width, geom.height, geom.tiles_x, geom.tiles_y, prm.min_weight, prm.alpha_max,
                  cull_threshold_log2(prm)};
  const bool aux = aux_t && aux_n, masked = src.clist != nullptr;
  if (aux && masked)
    composite_fast_kernel<true, true><<<tiles, kFastThreads, 0, st>>>(src, rec, image, aux_t, aux_n, abort, a);
  else if (aux)
    composite_fast_kernel<true, false><<<tiles, kFastThreads, 0, st>>>(src, rec, image, aux_t, aux_n, abort, a);
  else if (masked)
    composite_fast_kernel<false, true><<<tiles, kFastThreads, 0, st>>>(src, rec, image, nullptr, nullptr, abort, a);
  else
    composite_fast_kernel<false, false><<<tiles, kFastThreads, 0, st>>>(src, rec, image, nullptr, nullptr, abort, a);
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
