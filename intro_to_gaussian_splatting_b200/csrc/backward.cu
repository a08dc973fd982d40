// backward.cu -- gradient of the rendered image with respect to the Gaussian attributes (sm_100a).
//
// The reference announces a training loop (README.md:3) and marks its parameters requires_grad
// (splat/gaussians.py:19-21) but never wrote the backward pass: compute_gaussian_weight returns a Python float
// (splat/utils.py:365), which cuts the autograd graph.  What is differentiated here is the forward the reference
// DOES define:
//   image = render_pixel over the per-tile depth-ordered lists (splat/gaussian_scene.py:146-171) of
//   preprocess (splat/gaussian_scene.py:70-144) -- double sigmoid, no per-pixel bbox test, terminating Gaussian
//   dropped, det clamp 1e-3, x/z and y/z clamp at 1.3 tan(fov/2).
// Tile membership, depth order, termination and the clamps are piecewise constant: they pass no gradient
// (clamped branches pass exactly what torch.clamp passes, i.e. nothing through the clamped operand).
//
// Two kernels:
//   composite_backward_kernel  per tile, same 64-thread / 1x4-pixel-column mapping and cp.async staging as the
//                              forward kernel, walking the list BACK TO FRONT from the last blended Gaussian
//                              (aux_n) with the transmittance recovered by T_j = T_{j+1} / (1 - alpha_j) from the
//                              forward's final value (aux_t).  Per Gaussian and warp the nine partial sums are
//                              reduced with shuffles and added to grad2d with float atomics (RED.ADD.F32).
//   project_backward_kernel    per Gaussian: recomputes the projection and applies the chain rule down to
//                              points / scales / quaternions / colours / opacity logits.
// This TU is compiled with FMA contraction on: gradients carry a tolerance, not a bit pattern.
#include <math.h>

#include <type_traits>

#include "gsb_internal.cuh"

namespace gsb {
namespace {

constexpr int kBwdThreads = 64;
constexpr int kBwdBatch = 128;
constexpr int kBwdPerThread = kBwdBatch / kBwdThreads;
constexpr int kG2 = 12;  // floats per row of grad2d

struct BwdArgs {
  int width, height, tiles_x, tiles_y;
  float cull_log2;  // log2 of GsbParams.cull_alpha as the forward kernel used it (-inf: no culling)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP; the operand here is 1 - alpha >= 0.01
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One butterfly step over a PAIR of values: afterwards the lanes whose `bit` is clear hold x summed over the lane
// pair (L, L ^ bit), the lanes whose bit is set hold y -- one shuffle for two values instead of two.
__device__ __forceinline__ float pair_step(float x, float y, int lane, int bit) {
  const bool up = (lane & bit) != 0;
  const float keep = up ? y : x, send = up ? x : y;
  return keep + __shfl_xor_sync(0xffffffffu, send, bit);
}
__device__ __forceinline__ float single_step(float x, int bit) { return x + __shfl_xor_sync(0xffffffffu, x, bit); }

// Per pixel, with C = sum_j c_j alpha_j T_j and T_{j+1} = T_j (1 - alpha_j) over the blended prefix [0, n):
//   dL/dc_j     = alpha_j T_j g                                  (g = dL/dC of the pixel)
//   dL/dalpha_j = T_j (c_j . g - B_j),   B_j = (sum_{k>j} c_k alpha_k T_k / T_{j+1}) . g   (a scalar per pixel)
//   B_{j-1}     = B_j + alpha_j (c_j . g - B_j)                  (B_{n-1} = 0: nothing behind, no background)
//   alpha = op2 exp(power), power = a dx^2 + (b + c) dx dy + d dy^2, (a b; c d) = -0.5 inverse covariance,
//   (dx, dy) = mean - pixel  =>  dL/dpower = dL/dalpha * alpha = (alpha T)(c . g - B)  (and dL/dop2 = sum / op2).
// A thread's four pixels share dx, so the five geometric sums come from three moments of dL/dpower over dy:
//   P0 = sum dpw, P1 = sum dpw dy, P2 = sum dpw dy^2:
//   d a = dx^2 P0, d(b+c) = dx P1, d d = P2, d mx = 2 a dx P0 + (b+c) P1, d my = (b+c) dx P0 + 2 d P1.
// The nine per-Gaussian sums of a warp are reduced with a 12-shuffle butterfly (pair_step), stored to the warp's
// row of a shared accumulator, and after each batch of 128 Gaussians the two warps' rows are added into grad2d
// with three vector reductions per Gaussian (REDG.ADD.F32x4 x2 + one scalar) instead of 18 scalar atomics.
// A warp walks only the records the forward kernel's warp walked: the same conservative bound of alpha over the
// warp's 16x8 pixels (may_contribute, gsb_internal.cuh; same inputs, same threshold) is evaluated once per record and
// warp rectangle when a batch has landed, the survivors are compacted into a per-warp slot list, and the skipped
// records get exactly the zero gradient their skipped forward step implies.
template <bool kCull>
__global__ void __launch_bounds__(kBwdThreads)
composite_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ payload,
                          const float4* __restrict__ rec, const float* __restrict__ grad_image,
                          const float* __restrict__ aux_t, const uint32_t* __restrict__ aux_n,
                          float* __restrict__ grad2d, const __grid_constant__ BwdArgs a) {
  __shared__ __align__(16) float4 sm[2][kBwdBatch * 3];
  __shared__ uint32_t sm_idx[2][kBwdBatch];
  __shared__ float sm_acc[2][kBwdBatch * 9];  // [warp][slot][value]; zero except between a batch and its flush
  __shared__ uint32_t sm_nmax[2];
  __shared__ __align__(16) uint32_t sm_mask[2][kBwdBatch / 32];  // [warp rectangle][survivor bits of the batch]
  __shared__ uint8_t sm_list[2][kBwdBatch];                      // per warp: surviving slots, ascending

  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int px = tx * kTile + (lane & 15);
  const int py0 = ty * kTile + warp * 8 + (lane >> 4) * 4;
  const float fx = (float)px;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float rx0 = (float)(tx * kTile), rx1 = rx0 + 15.f;
  const float ry0 = (float)(ty * kTile), ry1 = ry0 + 7.f, ry2 = ry0 + 8.f, ry3 = ry0 + 15.f;
  constexpr bool cull_on = kCull;

  const uint2 rg = ranges[tile];
  const uint32_t len = rg.y - rg.x;
  const uint32_t* pl = payload + rg.x;

  // per-pixel state
  float T[4], gr[4], gg[4], gb[4], B[4], fy[4];
  uint32_t npx[4];
  uint32_t nw = 0, nmin_w = 0xFFFFFFFFu;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int py = py0 + k;
    fy[k] = (float)py;
    const bool in = px < a.width && py < a.height;
    const size_t p = (size_t)py * a.width + px;
    npx[k] = in ? min(aux_n[p], len) : 0u;
    T[k] = in ? aux_t[p] : 1.f;
    gr[k] = in ? grad_image[3 * p] : 0.f;
    gg[k] = in ? grad_image[3 * p + 1] : 0.f;
    gb[k] = in ? grad_image[3 * p + 2] : 0.f;
    B[k] = 0.f;
    nw = max(nw, npx[k]);
    nmin_w = min(nmin_w, npx[k]);
  }
  nw = __reduce_max_sync(0xffffffffu, nw);  // the warp's deepest blended Gaussian + 1
  nmin_w = __reduce_min_sync(0xffffffffu, nmin_w);  // below this index every pixel of the warp blends
  if (lane == 0) sm_nmax[warp] = nw;
  for (int s = tid; s < 2 * kBwdBatch * 9; s += kBwdThreads) (&sm_acc[0][0])[s] = 0.f;
  __syncthreads();
  const uint32_t nmax = max(sm_nmax[0], sm_nmax[1]);
  if (nmax == 0) return;
  const int nb = (int)((nmax + kBwdBatch - 1) / kBwdBatch);

  // where the butterfly leaves the totals: value 8 in the lanes with bit 1 set, value 4*b2 + 2*b3 + b4 elsewhere
  const int v_idx = (lane & 2) ? 8 : ((lane >> 2) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1);
  const bool v_writer = (lane & 1) == 0 && ((lane & 2) == 0 || lane == 2);
  float* my_acc = sm_acc[warp] + v_idx;

  uint32_t idx[kBwdPerThread];
  auto load_idx = [&](int b) {
#pragma unroll
    for (int j = 0; j < kBwdPerThread; ++j) {
      const uint32_t slot = (uint32_t)b * kBwdBatch + j * kBwdThreads + tid;
      idx[j] = slot < nmax ? pl[slot] : 0xFFFFFFFFu;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int j = 0; j < kBwdPerThread; ++j) {
      const int s = j * kBwdThreads + tid;
      sm_idx[buf][s] = idx[j];
      if (idx[j] != 0xFFFFFFFFu) {
        float4* dst = &sm[buf][s * 3];
        const float4* src = rec + 3 * (size_t)idx[j];
        cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
      }
    }
    cp_async_commit();
  };

  load_idx(nb - 1);
  stage((nb - 1) & 1);
  if (nb > 1) load_idx(nb - 2);
  for (int b = nb - 1; b >= 0; --b) {
    const int buf = b & 1;
    cp_async_wait_all();
    __syncthreads();  // batch b visible; everyone is done with the other buffer and with the last flush
    if (b > 0) {
      stage(buf ^ 1);
      if (b > 1) load_idx(b - 2);
    }
    // this warp's slots of the batch: [0, hi)
    const int base = b * kBwdBatch;
    const int hi = min((int)nw - base, kBwdBatch);
    // survivors of the batch per warp rectangle: slot j*64 + tid is bit `lane` of word 2j + warp
    {
      const int hi0 = (int)sm_nmax[0] - base, hi1 = (int)sm_nmax[1] - base;
#pragma unroll
      for (int j = 0; j < kBwdPerThread; ++j) {
        const int s = j * kBwdThreads + tid;
        const bool have = sm_idx[buf][s] != 0xFFFFFFFFu;
        bool k0 = have && s < hi0, k1 = have && s < hi1;
        if (cull_on && (k0 || k1)) {
          const float4 q0 = sm[buf][s * 3], q1 = sm[buf][s * 3 + 1];
          const float dxl = q0.x - rx1, dxh = q0.x - rx0;
          if (k0) k0 = may_contribute(q0.z, q0.w, q1.x, q1.y, q1.z, dxl, dxh, q0.y - ry1, q0.y - ry0, a.cull_log2);
          if (k1) k1 = may_contribute(q0.z, q0.w, q1.x, q1.y, q1.z, dxl, dxh, q0.y - ry3, q0.y - ry2, a.cull_log2);
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
        if (lane == 0) { sm_mask[0][2 * j + warp] = m0; sm_mask[1][2 * j + warp] = m1; }
      }
    }
    __syncthreads();  // masks visible
    const uint4 mk = *reinterpret_cast<const uint4*>(sm_mask[warp]);
    const int p1 = __popc(mk.x), p2 = p1 + __popc(mk.y), p3 = p2 + __popc(mk.z), total = p3 + __popc(mk.w);
    {
      uint8_t* lst = sm_list[warp];
      if ((mk.x >> lane) & 1u) lst[__popc(mk.x & lt_mask)] = (uint8_t)lane;
      if ((mk.y >> lane) & 1u) lst[p1 + __popc(mk.y & lt_mask)] = (uint8_t)(32 + lane);
      if ((mk.z >> lane) & 1u) lst[p2 + __popc(mk.z & lt_mask)] = (uint8_t)(64 + lane);
      if ((mk.w >> lane) & 1u) lst[p3 + __popc(mk.w & lt_mask)] = (uint8_t)(96 + lane);
    }
    __syncwarp();
    // entries of the list below `sel_lo` sit under the warp's minimum blended length: no per-pixel masking there
    int sel_lo;
    {
      const int smin = min(max((int)nmin_w - base, 0), kBwdBatch);  // slots >= smin need the per-pixel test
      auto below = [&](uint32_t m, int w) {  // set bits of word w at slots < smin
        const int r = smin - 32 * w;
        return r <= 0 ? 0 : (r >= 32 ? __popc(m) : __popc(m & ((1u << r) - 1u)));
      };
      sel_lo = below(mk.x, 0) + below(mk.y, 1) + below(mk.z, 2) + below(mk.w, 3);
    }
    (void)hi;
    // one Gaussian (slot i of the batch) against the thread's four pixels; kSel: some pixel of the warp stopped
    // blending before this Gaussian, so alpha is masked per pixel (warp-uniformly false below the warp's minimum)
    auto step = [&](int i, auto sel_tag) {
      constexpr bool kSel = decltype(sel_tag)::value;
      const uint32_t j = (uint32_t)(base + i);
      const float4 q0 = sm[buf][i * 3];      // mx, my, a, b
      const float4 q1 = sm[buf][i * 3 + 1];  // c, d, log2(op2), r
      const float2 q2 = *reinterpret_cast<const float2*>(&sm[buf][i * 3 + 2]);  // g, b
      const float dx = q0.x - fx;
      const float bc = q0.w + q1.x;
      const float adx = q0.z * dx, bcdx = bc * dx, adxx = adx * dx;
      float P0, P1, P2, s_r, s_g, s_b, amax = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dy = q0.y - fy[k];
        const float pw = fmaf(fmaf(q1.y, dy, bcdx), dy, adxx);  // a dx^2 + (b+c) dx dy + d dy^2
        float al = ex2_approx(fmaf(pw, 1.4426950408889634f, q1.z));
        if (kSel) al = (j < npx[k]) ? al : 0.f;  // not blended for this pixel: alpha = 0 makes every update a no-op
        if (!kCull) amax = fmaxf(amax, al);
        T[k] *= rcp_approx(1.f - al);  // T_j from T_{j+1}
        const float w = al * T[k];
        const float e = fmaf(q1.w, gr[k], fmaf(q2.x, gg[k], q2.y * gb[k])) - B[k];
        const float dpw = w * e;
        B[k] = fmaf(al, e, B[k]);
        const float t = dpw * dy;
        if (k == 0) {  // (resolved at compile time: the first pixel starts the sums)
          s_r = w * gr[0]; s_g = w * gg[0]; s_b = w * gb[0];
          P0 = dpw; P1 = t; P2 = t * dy;
        } else {
          s_r = fmaf(w, gr[k], s_r); s_g = fmaf(w, gg[k], s_g); s_b = fmaf(w, gb[k], s_b);
          P0 += dpw; P1 += t; P2 = fmaf(t, dy, P2);
        }
      }
      // without culling: exp underflow everywhere makes every sum exactly zero (with it such records never get here)
      if (!kCull && !__any_sync(0xffffffffu, amax != 0.f)) return;
      const float s_a = dx * dx * P0, s_bc = dx * P1;
      const float s_mx = fmaf(bc, P1, 2.f * adx * P0);
      const float s_my = fmaf(bcdx, P0, 2.f * q1.y * P1);
      // butterfly: (mx,my) (a,bc) (d,pw) (r,g) | b
      float a0 = pair_step(s_mx, s_my, lane, 16), a1 = pair_step(s_a, s_bc, lane, 16);
      float a2 = pair_step(P2, P0, lane, 16), a3 = pair_step(s_r, s_g, lane, 16), a4 = single_step(s_b, 16);
      float b0 = pair_step(a0, a1, lane, 8), b1 = pair_step(a2, a3, lane, 8), b2 = single_step(a4, 8);
      float c0 = pair_step(b0, b1, lane, 4), c1 = single_step(b2, 4);
      float d0 = pair_step(c0, c1, lane, 2);
      d0 = single_step(d0, 1);
      if (v_writer) my_acc[i * 9] = d0;
    };
    int n = total - 1;
    for (; n >= sel_lo; --n) step((int)sm_list[warp][n], std::true_type{});
    for (; n >= 0; --n) step((int)sm_list[warp][n], std::false_type{});
    __syncthreads();  // both warps are done with the batch: flush its sums
#pragma unroll
    for (int jj = 0; jj < kBwdPerThread; ++jj) {
      const int s = jj * kBwdThreads + tid;
      const uint32_t g = sm_idx[buf][s];
      if (g == 0xFFFFFFFFu) continue;
      float v[9];
      bool nz = false;
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        v[q] = sm_acc[0][s * 9 + q] + sm_acc[1][s * 9 + q];
        sm_acc[0][s * 9 + q] = 0.f; sm_acc[1][s * 9 + q] = 0.f;
        nz |= v[q] != 0.f;
      }
      if (!nz) continue;
      float* dst = grad2d + (size_t)g * kG2;
      red_add_v4(dst, v[0], v[1], v[2], v[3]);
      red_add_v4(dst + 4, v[4], v[5], v[6], v[7]);
      atomicAdd(dst + 8, v[8]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// projection backward
// ------------------------------------------------------------------------------------------------
struct CamConst {
  float V[16], P[16];
  float f_x, f_y, limx, limy, half_wm1, half_hm1, det_min, minimum_z;
};

__global__ void __launch_bounds__(256)
project_backward_kernel(const float* __restrict__ planes, int64_t n, int64_t n_pad,
                        const __grid_constant__ CamConst c, const uint32_t* __restrict__ depth_key,
                        const uint32_t* __restrict__ count, const float* __restrict__ grad2d,
                        float* __restrict__ g_points, float* __restrict__ g_scales, float* __restrict__ g_quats,
                        float* __restrict__ g_colors, float* __restrict__ g_opacity) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float dp[3] = {0.f, 0.f, 0.f}, ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f}, dc[3] = {0.f, 0.f, 0.f};
  float dl = 0.f;
  if (depth_key[i] != 0xFFFFFFFFu && count[i] != 0) {
    const float* g2 = grad2d + (size_t)i * kG2;
    const float g_mx = g2[0], g_my = g2[1], g_a = g2[2], g_bc = g2[3], g_d = g2[4], g_pw = g2[5];
    dc[0] = g2[6]; dc[1] = g2[7]; dc[2] = g2[8];
    const float x = planes[PX * n_pad + i], y = planes[PY * n_pad + i], z = planes[PZ * n_pad + i];
    const float s[3] = {planes[PSX * n_pad + i], planes[PSY * n_pad + i], planes[PSZ * n_pad + i]};
    const float qin[4] = {planes[PQW * n_pad + i], planes[PQX * n_pad + i], planes[PQY * n_pad + i],
                          planes[PQZ * n_pad + i]};
    const float logit = planes[POP * n_pad + i];
    const float* V = c.V;
    const float* P = c.P;

    // ---- opacity: alpha = op2 * w, op2 = sigmoid(sigmoid(logit)); g_pw = sum dL/dalpha * alpha = dL/dop2 * op2
    const float op1 = 1.f / (1.f + expf(-logit));
    const float op2 = 1.f / (1.f + expf(-op1));
    dl = g_pw * (1.f - op2) * op1 * (1.f - op1);

    // ---- mean: pixel = (clip.xy / clip.w + 1) * (S - 1) / 2
    const float cx = x * P[0] + y * P[4] + z * P[8] + P[12];
    const float cy = x * P[1] + y * P[5] + z * P[9] + P[13];
    const float cw = x * P[3] + y * P[7] + z * P[11] + P[15];
    const float icw = 1.f / cw;
    const float dndx = g_mx * c.half_wm1, dndy = g_my * c.half_hm1;
    const float dcx = dndx * icw, dcy = dndy * icw;
    const float dcw = -(dndx * cx + dndy * cy) * icw * icw;
#pragma unroll
    for (int k = 0; k < 3; ++k) dp[k] = dcx * P[4 * k] + dcy * P[4 * k + 1] + dcw * P[4 * k + 3];

    // ---- forward recompute: view space, rotation, covariances
    const float vx = x * V[0] + y * V[4] + z * V[8] + V[12];
    const float vy = x * V[1] + y * V[5] + z * V[9] + V[13];
    const float vz = x * V[2] + y * V[6] + z * V[10] + V[14];
    float n1 = sqrtf(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
    const float d1 = fmaxf(n1, 1e-12f);
    const float q1[4] = {qin[0] / d1, qin[1] / d1, qin[2] / d1, qin[3] / d1};
    const float n2 = sqrtf(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
    const float qr = q1[0] / n2, qx = q1[1] / n2, qy = q1[2] / n2, qz = q1[3] / n2;
    float R[9];
    R[0] = 1.f - 2.f * (qy * qy + qz * qz); R[1] = 2.f * (qx * qy - qr * qz); R[2] = 2.f * (qx * qz + qr * qy);
    R[3] = 2.f * (qx * qy + qr * qz); R[4] = 1.f - 2.f * (qx * qx + qz * qz); R[5] = 2.f * (qy * qz - qr * qx);
    R[6] = 2.f * (qx * qz - qr * qy); R[7] = 2.f * (qy * qz + qr * qx); R[8] = 1.f - 2.f * (qx * qx + qy * qy);
    float M[9], S3[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) M[r * 3 + k] = R[r * 3 + k] * s[k];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        S3[r * 3 + k] = M[r * 3] * M[k * 3] + M[r * 3 + 1] * M[k * 3 + 1] + M[r * 3 + 2] * M[k * 3 + 2];
    const float iz = 1.f / vz, iz2 = iz * iz;
    const float rx = vx * iz, ry = vy * iz;
    const bool in_x = rx >= -c.limx && rx <= c.limx, in_y = ry >= -c.limy && ry <= c.limy;
    const float cxr = fminf(fmaxf(rx, -c.limx), c.limx), cyr = fminf(fmaxf(ry, -c.limy), c.limy);
    const float tx = cxr * vz, ty = cyr * vz;
    const float J00 = c.f_x * iz, J02 = -c.f_x * tx * iz2, J11 = c.f_y * iz, J12 = -c.f_y * ty * iz2;
    // T = (J W)[:2], W[l][k] = V[k][l]  =>  T[i][k] = sum_l J[i][l] V[k][l]   (V row-major, stride 4)
    float Tm[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      Tm[k] = J00 * V[4 * k] + J02 * V[4 * k + 2];
      Tm[3 + k] = J11 * V[4 * k + 1] + J12 * V[4 * k + 2];
    }
    // TS = T Sigma (2x3); cov2 = TS T^T
    float TS[6];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        TS[r * 3 + k] = Tm[r * 3] * S3[k] + Tm[r * 3 + 1] * S3[3 + k] + Tm[r * 3 + 2] * S3[6 + k];
    const float cA = TS[0] * Tm[0] + TS[1] * Tm[1] + TS[2] * Tm[2];
    const float cB = TS[0] * Tm[3] + TS[1] * Tm[4] + TS[2] * Tm[5];
    const float cC = TS[3] * Tm[0] + TS[4] * Tm[1] + TS[5] * Tm[2];
    const float cD = TS[3] * Tm[3] + TS[4] * Tm[4] + TS[5] * Tm[5];
    const float det = cA * cD - cB * cC;
    const bool det_free = det >= c.det_min;  // torch.clamp(min=): gradient passes where det >= min
    const float idet = 1.f / (det_free ? det : c.det_min);

    // ---- inverse covariance: i00 = D/det, i01 = -B/det, i10 = -C/det, i11 = A/det; record holds -0.5 * inverse
    const float G00 = -0.5f * g_a, G01 = -0.5f * g_bc, G11 = -0.5f * g_d;  // G10 = G01
    float dA = G11 * idet, dD = G00 * idet, dB = -G01 * idet, dC = -G01 * idet;
    if (det_free) {
      const float ddet = -(G00 * cD - G01 * cB - G01 * cC + G11 * cA) * idet * idet;
      dA += ddet * cD; dD += ddet * cA; dB -= ddet * cC; dC -= ddet * cB;
    }
    // ---- cov2 = T Sigma T^T:  dSigma = T^T G T,  dT = G T Sigma + G^T T Sigma   (Sigma symmetric)
    float GT[6], GtT[6];  // G T and G^T T (2x3)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      GT[k] = dA * Tm[k] + dB * Tm[3 + k];
      GT[3 + k] = dC * Tm[k] + dD * Tm[3 + k];
      GtT[k] = dA * Tm[k] + dC * Tm[3 + k];
      GtT[3 + k] = dB * Tm[k] + dD * Tm[3 + k];
    }
    float dS[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) dS[r * 3 + k] = Tm[r] * GT[k] + Tm[3 + r] * GT[3 + k];
    float dT[6];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* g1 = GT + 3 * r;
        const float* g2t = GtT + 3 * r;
        dT[r * 3 + k] = (g1[0] + g2t[0]) * S3[k] + (g1[1] + g2t[1]) * S3[3 + k] + (g1[2] + g2t[2]) * S3[6 + k];
      }
    // ---- T = J W:  dJ[i][l] = sum_k dT[i][k] V[k][l]
    const float dJ00 = dT[0] * V[0] + dT[1] * V[4] + dT[2] * V[8];
    const float dJ02 = dT[0] * V[2] + dT[1] * V[6] + dT[2] * V[10];
    const float dJ11 = dT[3] * V[1] + dT[4] * V[5] + dT[5] * V[9];
    const float dJ12 = dT[3] * V[2] + dT[4] * V[6] + dT[5] * V[10];
    const float iz3 = iz2 * iz;
    float dvz = -dJ00 * c.f_x * iz2 - dJ11 * c.f_y * iz2 + 2.f * dJ02 * c.f_x * tx * iz3 + 2.f * dJ12 * c.f_y * ty * iz3;
    const float dtx = -dJ02 * c.f_x * iz2, dty = -dJ12 * c.f_y * iz2;
    // t = clamp(v/z) * z: inside the clamp t = v (d/dv = 1, d/dz = 0); outside t = +-lim * z
    float dvx = 0.f, dvy = 0.f;
    if (in_x) dvx = dtx; else dvz += cxr * dtx;
    if (in_y) dvy = dty; else dvz += cyr * dty;
#pragma unroll
    for (int k = 0; k < 3; ++k) dp[k] += dvx * V[4 * k] + dvy * V[4 * k + 1] + dvz * V[4 * k + 2];

    // ---- Sigma = M M^T, M = R diag(s)
    float dM[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        dM[r * 3 + k] = (dS[r * 3] + dS[r]) * M[k] + (dS[r * 3 + 1] + dS[3 + r]) * M[3 + k] +
                        (dS[r * 3 + 2] + dS[6 + r]) * M[6 + k];
    float dR[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        ds[k] += dM[r * 3 + k] * R[r * 3 + k];
        dR[r * 3 + k] = dM[r * 3 + k] * s[k];
      }
    // ---- R(q), q = (r, x, y, z) unit
    float dqn[4];
    dqn[0] = 2.f * (-qz * dR[1] + qy * dR[2] + qz * dR[3] - qx * dR[5] - qy * dR[6] + qx * dR[7]);
    dqn[1] = 2.f * (qy * dR[1] + qz * dR[2] + qy * dR[3] - 2.f * qx * dR[4] - qr * dR[5] + qz * dR[6] + qr * dR[7] -
                    2.f * qx * dR[8]);
    dqn[2] = 2.f * (-2.f * qy * dR[0] + qx * dR[1] + qr * dR[2] + qx * dR[3] + qz * dR[5] - qr * dR[6] + qz * dR[7] -
                    2.f * qy * dR[8]);
    dqn[3] = 2.f * (-2.f * qz * dR[0] - qr * dR[1] + qx * dR[2] + qr * dR[3] - 2.f * qz * dR[4] + qy * dR[5] +
                    qx * dR[6] + qy * dR[7]);
    // two normalisations (F.normalize, then build_rotation's own): d u = (d v - v (v . d v)) / |u|
    const float qn[4] = {qr, qx, qy, qz};
    float dot = qn[0] * dqn[0] + qn[1] * dqn[1] + qn[2] * dqn[2] + qn[3] * dqn[3];
    float dq1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) dq1[k] = (dqn[k] - qn[k] * dot) / n2;
    if (n1 > 1e-12f) {
      dot = q1[0] * dq1[0] + q1[1] * dq1[1] + q1[2] * dq1[2] + q1[3] * dq1[3];
#pragma unroll
      for (int k = 0; k < 4; ++k) dq[k] = (dq1[k] - q1[k] * dot) / n1;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) dq[k] = dq1[k] / d1;  // x / eps: the norm is clamped, a constant divisor
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g_points[3 * i + k] = dp[k];
    g_scales[3 * i + k] = ds[k];
    g_colors[3 * i + k] = dc[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) g_quats[4 * i + k] = dq[k];
  g_opacity[i] = dl;
}

}  // namespace

int launch_composite_backward(const uint2* ranges, const uint32_t* payload, const float4* rec,
                              const float* grad_image, const float* aux_t, const uint32_t* aux_n, float* grad2d,
                              FrameGeom geom, const GsbParams& prm, cudaStream_t st) {
  (void)prm;
  const int tiles = geom.tiles_x * geom.tiles_y;
  if (tiles <= 0) return 0;
  BwdArgs a{geom.width, geom.height, geom.tiles_x, geom.tiles_y, cull_threshold_log2(prm)};
  if (a.cull_log2 > -INFINITY)
    composite_backward_kernel<true><<<tiles, kBwdThreads, 0, st>>>(ranges, payload, rec, grad_image, aux_t, aux_n, grad2d, a);
  else
    composite_backward_kernel<false><<<tiles, kBwdThreads, 0, st>>>(ranges, payload, rec, grad_image, aux_t, aux_n, grad2d, a);
  return (int)cudaGetLastError();
}

int launch_project_backward(const float* planes, int64_t n, int64_t n_pad, const GsbCamera& cam, const GsbParams& prm,
                            const uint32_t* depth_key, const uint32_t* count, const float* grad2d, float* g_points,
                            float* g_scales, float* g_quats, float* g_colors, float* g_opacity, cudaStream_t st) {
  if (n <= 0) return 0;
  CamConst c;
  for (int k = 0; k < 16; ++k) { c.V[k] = cam.world2view[k]; c.P[k] = cam.full_proj[k]; }
  c.f_x = cam.f_x; c.f_y = cam.f_y;
  c.limx = prm.fov_clamp * cam.tan_fovx; c.limy = prm.fov_clamp * cam.tan_fovy;
  c.half_wm1 = 0.5f * ((float)cam.width - 1.f); c.half_hm1 = 0.5f * ((float)cam.height - 1.f);
  c.det_min = prm.det_min; c.minimum_z = prm.minimum_z;
  const int blocks = (int)((n + 255) / 256);
  project_backward_kernel<<<blocks, 256, 0, st>>>(planes, n, n_pad, c, depth_key, count, grad2d, g_points, g_scales,
                                                  g_quats, g_colors, g_opacity);
  return (int)cudaGetLastError();
}

}  // namespace gsb
