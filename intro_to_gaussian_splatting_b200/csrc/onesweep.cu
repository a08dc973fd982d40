// onesweep.cu -- hand-written stable LSD radix sort ("onesweep": one global histogram pass, then ONE
// read+write sweep per 8-bit digit with decoupled look-back instead of a separate scan/scatter pair).
//
// Replaces the reference's depth ordering, `torch.argsort(points_view[:, 2])` + 11 gathers
// (splat/gaussian_scene.py:117-129), and supplies the per-tile ordering that its per-tile boolean
// masks (splat/gaussian_scene.py:209-226) produce implicitly: sorting (tile_id<<32 | depth) keys
// groups instances by tile, front to back.  The sort is STABLE, so depth ties keep emission
// (Gaussian-index) order -- the tie contract of SURVEY.md section 7.2.
//
// Structure of one digit pass (kernel `onesweep_kernel`), per CTA = one tile of 256 x kItems keys:
//   1. ticket = atomicAdd(counter)          -> logical tile id; predecessors are resident => look-back
//                                              can never wait on a CTA that has not started
//   2. coalesced load, warp-striped
//   3. warp-private digit counters in smem; stable in-warp ranks from __match_any_sync
//   4. thread d owns digit d: scans the 8 warp counters, publishes the tile's count for d
//      (AGGREGATE), walks back over earlier tiles until it meets an inclusive PREFIX, publishes its own
//   5. keys scattered to smem in locally sorted order, then written with one coalesced run per digit;
//      payload takes the same route through the same smem.
//
// Roofline: HBM.  Histogram: sizeof(Key) read per key.  Each pass: (sizeof(Key)+4) read + written.
#include "gsb_internal.cuh"

namespace gsb {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;
constexpr int kSortTile = kThreads * kItems;  // 4096 keys per CTA

constexpr uint32_t kFlagShift = 30;
constexpr uint32_t kFlagAggregate = 1u << kFlagShift;
constexpr uint32_t kFlagPrefix = 2u << kFlagShift;
constexpr uint32_t kValueMask = (1u << kFlagShift) - 1u;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift, uint32_t mask) {
  return (uint32_t)(k >> shift) & mask;
}

// ---- upfront histogram of every digit position --------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
histogram_kernel(const KeyT* __restrict__ keys, int64_t n, int begin_bit, int end_bit, int passes,
                 uint32_t* __restrict__ hist /* [passes][256] */) {
  __shared__ uint32_t s_hist[kMaxPasses][kRadix];
  for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kThreads) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t per_block = (int64_t)kSortTile;
  for (int64_t base = (int64_t)blockIdx.x * per_block; base < n; base += (int64_t)gridDim.x * per_block) {
#pragma unroll 4
    for (int it = 0; it < kItems; ++it) {
      const int64_t i = base + (int64_t)it * kThreads + threadIdx.x;
      const bool ok = i < n;
      const KeyT k = ok ? keys[i] : (KeyT)0;
      for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * kRadixBits;
        const int bits = min(kRadixBits, end_bit - shift);
        const uint32_t d = ok ? digit_of(k, shift, (1u << bits) - 1u) : 0xFFFFu;
        // warp-aggregated shared atomic: neighbouring keys share their high digits, so without
        // aggregation every lane would hit the same bank word
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (ok && lane == (__ffs(peers) - 1)) atomicAdd(&s_hist[p][d], (uint32_t)__popc(peers));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kRadix; i += kThreads) {
    const uint32_t v = (&s_hist[0][0])[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// ---- one digit pass ------------------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                uint32_t* __restrict__ vals_out, int64_t n, int shift, int bits,
                const uint32_t* __restrict__ hist /* [256] of this pass */, uint32_t* ticket,
                uint32_t* status /* [tiles][256] */) {
  __shared__ union {
    KeyT keys[kSortTile];
    uint32_t vals[kSortTile];
  } exch;
  __shared__ uint32_t s_cnt[kWarps][kRadix];
  __shared__ uint32_t s_tile_start[kRadix];
  __shared__ uint32_t s_gofs[kRadix];
  __shared__ uint32_t s_scan[2][kWarps];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kWarps * kRadix; i += kThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t base = (int64_t)tile * kSortTile;
  const int valid = (int)min((int64_t)kSortTile, n - base);
  const uint32_t mask = (1u << bits) - 1u;
  const KeyT kPad = ~(KeyT)0;

  // 2. warp-striped coalesced loads; payload prefetched alongside
  KeyT key[kItems];
  uint32_t val[kItems];
  const int64_t wbase = base + (int64_t)warp * (32 * kItems) + lane;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t idx = wbase + i * 32;
    key[i] = idx < n ? keys_in[idx] : kPad;
  }
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t idx = wbase + i * 32;
    val[i] = idx < n ? vals_in[idx] : 0u;
  }

  // 3. stable ranks inside the warp
  uint32_t pos[kItems];
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const uint32_t d = digit_of(key[i], shift, mask);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) {
      old = s_cnt[warp][d];
      s_cnt[warp][d] = old + (uint32_t)__popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    pos[i] = old + (uint32_t)__popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // 4. thread d owns digit d
  {
    const int d = tid;
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t t = s_cnt[w][d];
      s_cnt[w][d] = sum;  // exclusive over warps
      sum += t;
    }
    // padding keys (all ones) sit in the highest used bin and, being last in tile order, last in it
    uint32_t real = sum;
    if ((uint32_t)d == mask) real -= (uint32_t)(kSortTile - valid);

    // block-wide exclusive scans: local tile counts (-> smem layout) and the global histogram (-> bin bases)
    uint32_t a = sum, b = hist[d];
    const uint32_t a_in = a, b_in = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o);
      const uint32_t tb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += ta; b += tb; }
    }
    if (lane == 31) { s_scan[0][warp] = a; s_scan[1][warp] = b; }
    __syncthreads();
    uint32_t wa = 0, wb = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
      if (w < warp) { wa += s_scan[0][w]; wb += s_scan[1][w]; }
    const uint32_t tile_start = wa + a - a_in;
    const uint32_t bin_base = wb + b - b_in;

    // decoupled look-back over earlier tiles for this digit
    uint32_t excl = 0;
    uint32_t* st = status + (size_t)tile * kRadix + d;
    if (tile == 0) {
      st_relaxed(st, kFlagPrefix | real);
    } else {
      st_relaxed(st, kFlagAggregate | real);
      const uint32_t* q = st - kRadix;
      for (uint32_t t = tile; t > 0; --t, q -= kRadix) {
        uint32_t s = ld_relaxed(q);
        while ((s >> kFlagShift) == 0) s = ld_relaxed(q);
        excl += s & kValueMask;
        if ((s >> kFlagShift) == 2u) break;
      }
      st_relaxed(st, kFlagPrefix | ((excl + real) & kValueMask));
    }
    s_tile_start[d] = tile_start;
    s_gofs[d] = bin_base + excl - tile_start;  // global index = s_gofs[d] + local position (mod 2^32)
  }
  __syncthreads();

  // 5a. keys -> smem in locally sorted order
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const uint32_t d = digit_of(key[i], shift, mask);
    pos[i] += s_tile_start[d] + s_cnt[warp][d];
    exch.keys[pos[i]] = key[i];
  }
  __syncthreads();
  // 5b. keys -> global: consecutive threads write consecutive addresses inside each digit run
  uint32_t dst[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int j = tid + i * kThreads;
    dst[i] = 0;
    if (j < valid) {
      const KeyT k = exch.keys[j];
      dst[i] = s_gofs[digit_of(k, shift, mask)] + (uint32_t)j;
      keys_out[dst[i]] = k;
    }
  }
  __syncthreads();
  // 5c. payload by the same route
#pragma unroll
  for (int i = 0; i < kItems; ++i) exch.vals[pos[i]] = val[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int j = tid + i * kThreads;
    if (j < valid) vals_out[dst[i]] = exch.vals[j];
  }
}

}  // namespace

template <typename KeyT>
SortPlan make_sort_plan(int64_t n, int begin_bit, int end_bit) {
  SortPlan p;
  p.begin_bit = begin_bit;
  p.end_bit = end_bit;
  p.passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
  if (p.passes < 0) p.passes = 0;
  p.n = n;
  p.tiles = (n + kSortTile - 1) / kSortTile;
  // [hist: passes*256][tickets: 8][status: passes * tiles * 256]
  p.control_words = (size_t)kMaxPasses * kRadix + 8 + (size_t)p.passes * (size_t)p.tiles * kRadix;
  return p;
}
template SortPlan make_sort_plan<uint32_t>(int64_t, int, int);
template SortPlan make_sort_plan<uint64_t>(int64_t, int, int);

template <typename KeyT>
int launch_sort(const SortPlan& plan, KeyT* keys_a, uint32_t* vals_a, KeyT* keys_b, uint32_t* vals_b,
                uint32_t* control, bool* result_in_a, int* launches, cudaStream_t st) {
  *result_in_a = true;
  if (plan.n == 0 || plan.passes == 0) return 0;
  uint32_t* hist = control;
  uint32_t* tickets = control + (size_t)kMaxPasses * kRadix;
  uint32_t* status = tickets + 8;
  int hist_blocks = plan.tiles < 148 * 4 ? (int)plan.tiles : 148 * 4;
  histogram_kernel<KeyT><<<hist_blocks, kThreads, 0, st>>>(keys_a, plan.n, plan.begin_bit, plan.end_bit, plan.passes, hist);
  if (launches) ++*launches;
  KeyT* kin = keys_a; KeyT* kout = keys_b;
  uint32_t* vin = vals_a; uint32_t* vout = vals_b;
  for (int p = 0; p < plan.passes; ++p) {
    const int shift = plan.begin_bit + p * kRadixBits;
    const int bits = (plan.end_bit - shift) < kRadixBits ? (plan.end_bit - shift) : kRadixBits;
    onesweep_kernel<KeyT><<<(unsigned)plan.tiles, kThreads, 0, st>>>(
        kin, vin, kout, vout, plan.n, shift, bits, hist + (size_t)p * kRadix, tickets + p,
        status + (size_t)p * (size_t)plan.tiles * kRadix);
    if (launches) ++*launches;
    KeyT* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  *result_in_a = (plan.passes % 2) == 0;
  return (int)cudaGetLastError();
}
template int launch_sort<uint32_t>(const SortPlan&, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*, bool*, int*, cudaStream_t);
template int launch_sort<uint64_t>(const SortPlan&, uint64_t*, uint32_t*, uint64_t*, uint32_t*, uint32_t*, bool*, int*, cudaStream_t);

}  // namespace gsb
