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
#ifndef GSB_RANK_GROUP
#define GSB_RANK_GROUP 8
#endif
constexpr int kWarps = kThreads / 32;
constexpr int kLookBatch = 8;   // look-back loads in flight per lane (32 in flight for the one-wave depth sort, two
                                // CTAs per SM and no spills: 17.9 us per pass instead of 15.6 -- measured, dropped)
constexpr int kHistItems = 16;
constexpr int kSortTile = kThreads * kHistItems;  // keys per histogram chunk


__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Look-back status word: 2 flag bits + value.  32-bit words carry 30-bit counts (enough below 2^30 keys);
// sorts of 2^30 .. 2^32-1 keys use 64-bit words.
template <typename W> struct StatusWord;
template <> struct StatusWord<uint32_t> {
  static constexpr int kShift = 30;
  static __device__ __forceinline__ uint32_t ld(const uint32_t* p) { return ld_relaxed(p); }
  static __device__ __forceinline__ void st(uint32_t* p, uint32_t v) { st_relaxed(p, v); }
};
template <> struct StatusWord<uint64_t> {
  static constexpr int kShift = 62;
  static __device__ __forceinline__ uint64_t ld(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
  }
  static __device__ __forceinline__ void st(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  }
};

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
    for (int it = 0; it < kHistItems; ++it) {
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
// kMode: what travels with the key
//   kPairs            (key, u32 payload) pairs; vals_in == nullptr means "payload = the key's index"
//                     (first pass of the per-Gaussian depth sort: saves an iota kernel and a key copy)
//   kKeysOnly         the payload is packed into key bits that are not being sorted (tile<<32 | gaussian)
//   kKeysOnlyLowOut   same, and only the packed payload is written (to vals_out): the last tile pass of SPLIT
//                     mode, after which nothing reads the tile field of the key any more.  The payload is the
//                     key's low word masked with low_mask; with a table (vals_in, otherwise unused in this mode)
//                     it is an emission position and table[position] -- the Gaussian index -- is written.
//   kKeysOnlyEntryOut same as kKeysOnlyLowOut with a table, but what is written per key is the super-tile list ENTRY
//                     {Gaussian index, tile mask}: bit (ly << lw | lx) of the mask is set when the Gaussian's tile
//                     rect covers tile (lx, ly) of the super-tile the key names (its sorted bits); the compositing
//                     kernel filters a super-tile's list down to one tile's with that bit (binning.cu, composite.cu)
enum SortMode { kPairs = 0, kKeysOnly = 1, kKeysOnlyLowOut = 2, kKeysOnlyEntryOut = 3 };

// what kKeysOnlyEntryOut needs to turn (Gaussian, super-tile) into a tile mask
struct EntryOut {
  const ushort4* rect;  // tile rects tx0,tx1,ty0,ty1 per Gaussian
  int lw, lh, snx;      // log2 super-tile size in tiles, super-tiles per row
  int key_shift;        // super-tile id = key >> key_shift
};

__device__ __forceinline__ uint32_t tile_mask_of(ushort4 r, uint32_t super, const EntryOut& e) {
  const uint32_t sy = super / (uint32_t)e.snx, sx = super - sy * (uint32_t)e.snx;
  const uint32_t bx = sx << e.lw, by = sy << e.lh;
  const uint32_t x0 = max((uint32_t)r.x, bx) - bx, x1 = min((uint32_t)r.y, bx + (1u << e.lw) - 1u) - bx;
  const uint32_t y0 = max((uint32_t)r.z, by) - by, y1 = min((uint32_t)r.w, by + (1u << e.lh) - 1u) - by;
  const uint32_t row = ((2u << (x1 - x0)) - 1u) << x0;  // x1 - x0 <= 31; 2u << 31 wraps to 0 -> all ones
  uint32_t m = 0u;
  for (uint32_t y = y0; y <= y1; ++y) m |= row << (y << e.lw);
  return m;
}

__device__ __forceinline__ uint32_t digit32(uint32_t k, int shift, uint32_t mask) { return (k >> shift) & mask; }
__device__ __forceinline__ uint32_t digit64(uint64_t k, int shift, uint32_t mask) {
  const uint32_t lo = (uint32_t)k, hi = (uint32_t)(k >> 32);
  const uint32_t v = shift >= 32 ? (hi >> (shift - 32)) : __funnelshift_r(lo, hi, shift);  // warp-uniform branch
  return v & mask;
}
// Lanes holding the same digit, from one ballot per digit bit.  `__match_any_sync` compiles to MATCH.ANY, which
// costs ~30 cycles of a per-SM unit per warp instruction here (clock64 instrumentation: 11 000 of a tile's 23 600
// cycles went into 16 matches per thread); a ballot costs ~4.
__device__ __forceinline__ unsigned match_digit(uint32_t d, int bits) {
  unsigned differ = 0u;  // lanes whose digit differs from mine in some bit
  // differ |= ballot(bit) ^ (bit ? ~0 : 0), spelled in PTX so that it stays ~3 SASS instructions per bit (one R2P
  // for seven predicates, VOTE, predicated complement, 3-input ORs); the C++ form is canonicalised into 6 per bit
#define GSB_MATCH_BIT(B)                                                           \
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t, m, s;\n\t"                               \
      "and.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"                              \
      "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"                                 \
      "selp.b32 s, 0xffffffff, 0, p;\n\t"                                          \
      "xor.b32 m, m, s;\n\tor.b32 %0, %0, m;\n\t}"                                  \
      : "+r"(differ) : "r"(d), "r"(1u << (B)))
  if (bits > 5) {  // warp-uniform.  Two complete sequences: splitting one into 5 + 3 bits costs the 8-bit passes 13 %
    GSB_MATCH_BIT(0); GSB_MATCH_BIT(1); GSB_MATCH_BIT(2); GSB_MATCH_BIT(3); GSB_MATCH_BIT(4);
    GSB_MATCH_BIT(5); GSB_MATCH_BIT(6); GSB_MATCH_BIT(7);
  } else {         // the bits above a narrow digit are zero in every lane and would match trivially
    GSB_MATCH_BIT(0); GSB_MATCH_BIT(1); GSB_MATCH_BIT(2); GSB_MATCH_BIT(3); GSB_MATCH_BIT(4);
  }
#undef GSB_MATCH_BIT
  return ~differ;
}

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_fast(KeyT k, int shift, uint32_t mask) {
  if constexpr (sizeof(KeyT) == 8) return digit64((uint64_t)k, shift, mask);
  else return digit32((uint32_t)k, shift, mask);
}

template <typename KeyT, int kItems>
struct SortSmem {
  union {
    KeyT keys[kThreads * kItems];
    uint32_t vals[kThreads * kItems];
  } exch;
  alignas(16) uint32_t cnt[kWarps][kRadix];
  uint32_t gofs[kRadix];
  uint32_t scan[2][kWarps];
  uint32_t tile;
};

// Opt-in phase clocks (-DGSB_PHASE_CLOCKS, tools/phase_clocks.py): thread 0 of every tile of one pass flavour (GSB_PHASE_MODE)
// adds the cycles between consecutive marks to g_phase[mark]; g_phase[15] counts tiles.  Off in the shipped library.
#ifdef GSB_PHASE_CLOCKS
#ifndef GSB_PHASE_MODE
#define GSB_PHASE_MODE kKeysOnlyEntryOut  // which pass flavour is clocked: 0 = the depth sort's (key, payload) passes
#endif
__device__ unsigned long long g_phase[16];
#define GSB_PHASE(k)                                                        \
  do {                                                                      \
    if (kMode == GSB_PHASE_MODE && threadIdx.x == 0) {                      \
      const long long _t = clock64();                                       \
      atomicAdd(&g_phase[k], (unsigned long long)(_t - _t_prev));           \
      _t_prev = _t;                                                         \
    }                                                                       \
  } while (0)
#else
#define GSB_PHASE(k) do { } while (0)
#endif

// kFull: the tile holds kThreads*kItems keys (every tile but the last) -> straight-line code, no bounds checks
template <typename KeyT, int kItems, int kMode, bool kFull, typename W, int kLook>
__device__ __forceinline__ void onesweep_tile(SortSmem<KeyT, kItems>& sm, const KeyT* __restrict__ keys_in,
                                              const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                                              uint32_t* __restrict__ vals_out, int64_t n, int shift, int bits,
                                              uint32_t mask, const uint32_t* __restrict__ hist, W* status,
                                              uint32_t tile, int valid, uint32_t low_mask, const EntryOut& entry) {
  using SW = StatusWord<W>;
  constexpr W kWAgg = (W)1 << SW::kShift, kWPre = (W)2 << SW::kShift, kWMask = ((W)1 << SW::kShift) - 1;
  constexpr int kTileKeys = kThreads * kItems;
  auto& exch = sm.exch;
  auto& s_cnt = sm.cnt;
  auto& s_gofs = sm.gofs;
  auto& s_scan = sm.scan;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)tile * kTileKeys;
  constexpr bool full = kFull;
  const KeyT kPad = ~(KeyT)0;

#ifdef GSB_PHASE_CLOCKS
  long long _t_prev = clock64();
  if (kMode == GSB_PHASE_MODE && threadIdx.x == 0) atomicAdd(&g_phase[15], 1ull);
#endif
  // 2. warp-striped coalesced loads
  KeyT key[kItems];
  const int64_t wbase = base + (int64_t)warp * (32 * kItems) + lane;
  if (full) {
#pragma unroll
    for (int i = 0; i < kItems; ++i) key[i] = keys_in[wbase + i * 32];
  } else {
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int64_t idx = wbase + i * 32;
      key[i] = idx < n ? keys_in[idx] : kPad;
    }
  }

  // 3. stable ranks inside the warp, in groups of kRankGroup items: the group's ballot matches first
  //    (independent), then its running counts (leaders only, in item order), then its broadcasts -- enough ILP
  //    to cover the ATOMS / SHFL latencies while only kRankGroup peer masks are live at a time.
  // Running per-digit count of this warp, kept by the group leaders (one leader per digit and item) with a
  // shared-memory atomic that returns the count before this item.  (A plain load + store by the leader is what
  // the arithmetic needs, but a different lane may lead the same digit in the next item: racecheck rightly flags
  // that, and ordering it with a __syncwarp() per item measured 6-8 us per frame slower than the atomic.)
  constexpr int kRankGroup = GSB_RANK_GROUP < kItems ? GSB_RANK_GROUP : kItems;
  uint32_t pos[kItems];
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t* wcnt = s_cnt[warp];
#pragma unroll
  for (int g = 0; g < kItems; g += kRankGroup) {
    unsigned peers[kRankGroup];
#pragma unroll
    for (int i = 0; i < kRankGroup; ++i) peers[i] = match_digit(digit_fast(key[g + i], shift, mask), bits);
#pragma unroll
    for (int i = 0; i < kRankGroup; ++i) {
      pos[g + i] = 0;
      if ((peers[i] & lt_mask) == 0u)  // lowest lane of its peer group == the leader
        pos[g + i] = atomicAdd(&wcnt[digit_fast(key[g + i], shift, mask)], (uint32_t)__popc(peers[i]));
    }
#pragma unroll
    for (int i = 0; i < kRankGroup; ++i)
      pos[g + i] = __shfl_sync(0xffffffffu, pos[g + i], __ffs(peers[i]) - 1) + (uint32_t)__popc(peers[i] & lt_mask);
  }
  GSB_PHASE(0);  // load + rank (thread 0's warp)
  __syncthreads();
  GSB_PHASE(1);  // wait for the slowest warp

  // 4a. thread d owns digit d: scan the warp counters, publish the tile's aggregate as early as possible
  uint32_t real, tile_start, bin_base;
  bool bin_live;
  W* st = status + (size_t)tile * kRadix + tid;
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
    real = sum;
    if ((uint32_t)d == mask) real -= (uint32_t)(kTileKeys - valid);
    // bins above a narrow digit's mask, and bins that are empty in the whole array (known from the histogram:
    // e.g. all but a handful of exponent bytes in the last depth pass), take no part in the look-back
    const uint32_t b_hist = hist[d];
    bin_live = (uint32_t)d <= mask && b_hist != 0u;
    if (bin_live) SW::st(st, (tile == 0 ? kWPre : kWAgg) | (W)real);

    // block-wide exclusive scans: local tile counts (-> smem layout) and the global histogram (-> bin bases)
    uint32_t a = sum, b = b_hist;
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
    tile_start = wa + a - a_in;
    bin_base = wb + b - b_in;
    // fold the digit's start inside the tile into the per-warp offsets: one shared load per key in 5a
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s_cnt[w][d] += tile_start;
  }
  __syncthreads();
  GSB_PHASE(2);  // digit scan + publish

  // 5a. keys -> smem in locally sorted order (gives earlier tiles time to publish before the look-back)
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    pos[i] += s_cnt[warp][digit_fast(key[i], shift, mask)];
    exch.keys[pos[i]] = key[i];
  }

  GSB_PHASE(3);  // scatter to shared memory
  // 4b. decoupled look-back for this thread's digit.  Predecessor words are fetched in BATCHES of
  //     independent loads (2 first -- in steady state the nearest tiles already hold an inclusive prefix --
  //     then 8 at a time): at the start of a pass several hundred tiles are in flight with only their
  //     aggregates published, and a one-at-a-time walk pays one L2 round trip for each of them.  (Rounds of 32
  //     were measured slower: most of this phase is spent WAITING for a slow predecessor's aggregate, not walking.)
  {
    uint32_t excl = 0;
    if (tile != 0 && bin_live) {  // e.g. a 5-bit pass has 32 live digits: only warp 0 looks back
      const W* col = status + tid;  // status is [tile][256]
      int64_t t = (int64_t)tile - 1;
      bool found = false;
#ifdef GSB_PHASE_CLOCKS
      unsigned long long n_words = 0, n_spins = 0;
#endif
      {
        W v[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) v[k] = (t - k >= 0) ? SW::ld(col + (size_t)(t - k) * kRadix) : kWPre;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (found) break;
          W x = v[k];
          while ((x >> SW::kShift) == 0) {
#ifdef GSB_PHASE_CLOCKS
            ++n_spins;
#endif
            x = SW::ld(col + (size_t)(t - k) * kRadix);
          }
#ifdef GSB_PHASE_CLOCKS
          ++n_words;
#endif
          excl += (uint32_t)(x & kWMask);
          found = (x >> SW::kShift) == 2u;
        }
        t -= 2;
      }
      while (!found) {
        W v[kLook];
#pragma unroll
        for (int k = 0; k < kLook; ++k)
          v[k] = (t - k >= 0) ? SW::ld(col + (size_t)(t - k) * kRadix) : kWPre;
#pragma unroll
        for (int k = 0; k < kLook; ++k) {
          if (found) break;
          W x = v[k];
          while ((x >> SW::kShift) == 0) {
#ifdef GSB_PHASE_CLOCKS
            ++n_spins;
#endif
            x = SW::ld(col + (size_t)(t - k) * kRadix);
          }
#ifdef GSB_PHASE_CLOCKS
          ++n_words;
#endif
          excl += (uint32_t)(x & kWMask);
          found = (x >> SW::kShift) == 2u;
        }
        t -= kLook;
      }
      SW::st(st, kWPre | (((W)excl + (W)real) & kWMask));
#ifdef GSB_PHASE_CLOCKS
      if (kMode == GSB_PHASE_MODE && threadIdx.x == 0) { atomicAdd(&g_phase[8], n_words); atomicAdd(&g_phase[9], n_spins); }
#endif
    }
    s_gofs[tid] = bin_base + excl - tile_start;  // global index = s_gofs[d] + local position (mod 2^32)
  }
  GSB_PHASE(4);  // look-back of digit 0
  __syncthreads();
  GSB_PHASE(5);  // wait for the slowest digit

  // 5b. keys -> global: consecutive threads write consecutive addresses inside each digit run
  uint32_t dst[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int j = tid + i * kThreads;
    dst[i] = 0;
    if (full || j < valid) {
      const KeyT k = exch.keys[j];
      dst[i] = s_gofs[digit_fast(k, shift, mask)] + (uint32_t)j;
      if constexpr (kMode == kKeysOnlyLowOut) {
        const uint32_t low = (uint32_t)k & low_mask;
        vals_out[dst[i]] = vals_in ? vals_in[low] : low;
      } else if constexpr (kMode == kKeysOnlyEntryOut) {
        const uint32_t low = (uint32_t)k & low_mask;
        const uint32_t g = vals_in ? vals_in[low] : low;
        const uint32_t m = tile_mask_of(entry.rect[g], (uint32_t)(k >> entry.key_shift), entry);
        reinterpret_cast<uint2*>(vals_out)[dst[i]] = make_uint2(g, m);
      } else {
        keys_out[dst[i]] = k;
      }
    }
  }
  if constexpr (kMode == kPairs) {
    // 5c. payload by the same route (loaded late: keeps the register footprint of the ranking phase small)
    uint32_t val[kItems];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int64_t idx = wbase + i * 32;
      val[i] = vals_in ? ((full || idx < n) ? vals_in[idx] : 0u) : (uint32_t)idx;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kItems; ++i) exch.vals[pos[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int j = tid + i * kThreads;
      if (full || j < valid) vals_out[dst[i]] = exch.vals[j];
    }
  }
  GSB_PHASE(6);  // write
}

#ifndef GSB_SORT_MINBLOCKS16
#define GSB_SORT_MINBLOCKS16 3
#endif
template <typename KeyT, int kItems, int kMode, typename W, int kLook>
__global__ void __launch_bounds__(kThreads, kItems == 8 ? 4 : GSB_SORT_MINBLOCKS16)
onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                uint32_t* __restrict__ vals_out, int64_t n, const uint32_t* __restrict__ n_dev,
                const uint32_t* __restrict__ abort, int shift, int bits,
                const uint32_t* __restrict__ hist /* [256] of this pass */, uint32_t* ticket,
                W* status /* [num_tiles][256] */, uint32_t low_mask, const EntryOut entry) {
  constexpr int kTileKeys = kThreads * kItems;
  __shared__ SortSmem<KeyT, kItems> sm;
  const int tid = threadIdx.x;
  // device-side key count: the grid was sized for a capacity; CTAs whose ticket lies past the last tile leave
  // (they come after every working CTA in ticket order, so no look-back ever waits for them)
  if (abort && *abort) return;
  if (n_dev) n = (int64_t)*n_dev;
  if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
  {
    uint4* z = reinterpret_cast<uint4*>(&sm.cnt[0][0]);
    for (int i = tid; i < kWarps * kRadix / 4; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const uint32_t tile = sm.tile;
  if ((int64_t)tile * kTileKeys >= n) return;
  const int valid = (int)min((int64_t)kTileKeys, n - (int64_t)tile * kTileKeys);
  const uint32_t mask = (1u << bits) - 1u;
  if (valid == kTileKeys)
    onesweep_tile<KeyT, kItems, kMode, true, W, kLook>(sm, keys_in, vals_in, keys_out, vals_out, n, shift, bits, mask, hist, status, tile, valid, low_mask, entry);
  else
    onesweep_tile<KeyT, kItems, kMode, false, W, kLook>(sm, keys_in, vals_in, keys_out, vals_out, n, shift, bits, mask, hist, status, tile, valid, low_mask, entry);
}

}  // namespace

#ifdef GSB_PHASE_CLOCKS
// debug build only (not declared in include/gsb.h): copy out and reset the phase clocks
extern "C" int gsb_debug_phase_clocks(unsigned long long* out) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_phase, z, sizeof(z));
  return (int)e;
}
#endif
static int g_sort_items = 0;  // keys per thread of the onesweep tile forced by GSB_SORT_ITEMS (8 or 16); 0: per sort
void set_sort_items(int items) { g_sort_items = (items == 8 || items == 16) ? items : 0; }
static int g_force_wide = 0;   // test knob: use the 64-bit look-back words regardless of the key count
void set_force_wide_status(int on) { g_force_wide = on ? 1 : 0; }

// items: keys per thread (8 or 16) the caller would like; GSB_SORT_ITEMS overrides it for every sort
template <typename KeyT>
SortPlan make_sort_plan(int64_t n, int begin_bit, int end_bit, int items) {
  SortPlan p;
  p.begin_bit = begin_bit;
  p.end_bit = end_bit;
  p.passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
  if (p.passes < 0) p.passes = 0;
  p.n = n;
  p.keys_only = 0;
  p.low_bits = 0;
  p.gather_table = nullptr;
  p.n_dev = nullptr;
  p.abort = nullptr;
  p.entry_rect = nullptr;
  p.entry_lw = p.entry_lh = 0;
  p.entry_snx = 1;
  p.items = g_sort_items ? g_sort_items : (items == 8 ? 8 : 16);
  const int64_t tile_keys = (int64_t)kThreads * p.items;
  p.tiles = (n + tile_keys - 1) / tile_keys;
  // [tickets: 8, padded to one 128-byte line][status: passes * tiles * 256 words; 64-bit words from 2^30 keys on].
  // With a 128-byte-aligned control block every tile's status row starts on a line: a look-back read of 32
  // neighbouring words is one L2 line, not two (depth sort of config 3: 59 -> 53 us).
  p.wide_status = (g_force_wide || n >= ((int64_t)1 << 30)) ? 1 : 0;
  p.control_words = kTicketWords + (size_t)p.passes * (size_t)p.tiles * kRadix * (p.wide_status ? 2 : 1);
  return p;
}
template SortPlan make_sort_plan<uint32_t>(int64_t, int, int, int);
template SortPlan make_sort_plan<uint64_t>(int64_t, int, int, int);

// Histogram of the keys themselves (only the stand-alone sort entry needs it; the frame pipeline gets its
// histograms from the projection kernel and tile_stats_kernel).  hist: kMaxPasses*256 zeroed words.
template <typename KeyT>
int launch_key_histogram(const SortPlan& plan, const KeyT* keys, uint32_t* hist, cudaStream_t st) {
  if (plan.n == 0 || plan.passes == 0) return 0;
  int64_t chunks = (plan.n + kSortTile - 1) / kSortTile;
  const int cap = sm_count() * 4;
  int hist_blocks = chunks < cap ? (int)chunks : cap;
  histogram_kernel<KeyT><<<hist_blocks, kThreads, 0, st>>>(keys, plan.n, plan.begin_bit, plan.end_bit, plan.passes, hist);
  return (int)cudaGetLastError();
}
template int launch_key_histogram<uint32_t>(const SortPlan&, const uint32_t*, uint32_t*, cudaStream_t);
template int launch_key_histogram<uint64_t>(const SortPlan&, const uint64_t*, uint32_t*, cudaStream_t);

template <typename KeyT, int kItems, typename W>
static void launch_pass(int mode, unsigned tiles, cudaStream_t st, const KeyT* kin, const uint32_t* vin, KeyT* kout,
                        uint32_t* vout, int64_t n, const uint32_t* n_dev, const uint32_t* abort, int shift, int bits,
                        const uint32_t* hist, uint32_t* ticket, W* status, uint32_t low_mask, const EntryOut& entry) {
  if (mode == kPairs) {
    onesweep_kernel<KeyT, kItems, kPairs, W, kLookBatch><<<tiles, kThreads, 0, st>>>(kin, vin, kout, vout, n, n_dev, abort, shift, bits, hist, ticket, status, low_mask, entry);
  } else if (mode == kKeysOnly) {
    onesweep_kernel<KeyT, kItems, kKeysOnly, W, kLookBatch><<<tiles, kThreads, 0, st>>>(kin, vin, kout, vout, n, n_dev, abort, shift, bits, hist, ticket, status, low_mask, entry);
  } else if (mode == kKeysOnlyLowOut) {
    onesweep_kernel<KeyT, kItems, kKeysOnlyLowOut, W, kLookBatch><<<tiles, kThreads, 0, st>>>(kin, vin, kout, vout, n, n_dev, abort, shift, bits, hist, ticket, status, low_mask, entry);
  } else {
    onesweep_kernel<KeyT, kItems, kKeysOnlyEntryOut, W, kLookBatch><<<tiles, kThreads, 0, st>>>(kin, vin, kout, vout, n, n_dev, abort, shift, bits, hist, ticket, status, low_mask, entry);
  }
}

template <typename KeyT>
int launch_sort(const SortPlan& plan, const KeyT* keys_src, const uint32_t* vals_src, KeyT* keys_a, uint32_t* vals_a,
                KeyT* keys_b, uint32_t* vals_b, const uint32_t* hist, uint32_t* control, bool* result_in_a,
                int* launches, cudaStream_t st) {
  // pass 0 reads (keys_src, vals_src) and writes the b-buffers; later passes ping-pong b -> a -> b ...
  // keys_src may alias keys_a (in-place use) or be a read-only array that must survive (depth keys).
  // plan.keys_only: no payload arrays; the LAST pass writes only the low 32 bits of each key into the
  // vals buffer of its destination side (vals_a or vals_b).
  *result_in_a = true;
  if (plan.n == 0 || plan.passes == 0) return 0;
  uint32_t* tickets = control;
  uint32_t* status = control + kTicketWords;
  const KeyT* kin = keys_src; const uint32_t* vin = vals_src;
  KeyT* kout = keys_b; uint32_t* vout = vals_b;
  for (int p = 0; p < plan.passes; ++p) {
    const int shift = plan.begin_bit + p * kRadixBits;
    const int bits = (plan.end_bit - shift) < kRadixBits ? (plan.end_bit - shift) : kRadixBits;
    const size_t per_pass = (size_t)plan.tiles * kRadix;
    const int mode = !plan.keys_only ? kPairs
                     : (p == plan.passes - 1 ? (plan.entry_rect ? kKeysOnlyEntryOut : kKeysOnlyLowOut) : kKeysOnly);
    const EntryOut entry{plan.entry_rect, plan.entry_lw, plan.entry_lh, plan.entry_snx,
                         sizeof(KeyT) == 8 ? 32 : plan.low_bits};
    const uint32_t* h = hist + (size_t)p * kRadix;
    const uint32_t low_mask = (plan.low_bits > 0 && plan.low_bits < 32) ? ((1u << plan.low_bits) - 1u) : 0xFFFFFFFFu;
    if (mode == kKeysOnlyLowOut || mode == kKeysOnlyEntryOut) vin = plan.gather_table;  // the one keys-only pass that reads `vals_in`: as a table
    if (plan.wide_status) {
      uint64_t* stp = reinterpret_cast<uint64_t*>(status) + (size_t)p * per_pass;  // status starts 8-byte aligned
      if (plan.items == 16)
        launch_pass<KeyT, 16, uint64_t>(mode, (unsigned)plan.tiles, st, kin, vin, kout, vout, plan.n, plan.n_dev, plan.abort, shift, bits, h, tickets + p, stp, low_mask, entry);
      else
        launch_pass<KeyT, 8, uint64_t>(mode, (unsigned)plan.tiles, st, kin, vin, kout, vout, plan.n, plan.n_dev, plan.abort, shift, bits, h, tickets + p, stp, low_mask, entry);
    } else {
      uint32_t* stp = status + (size_t)p * per_pass;
      if (plan.items == 16)
        launch_pass<KeyT, 16, uint32_t>(mode, (unsigned)plan.tiles, st, kin, vin, kout, vout, plan.n, plan.n_dev, plan.abort, shift, bits, h, tickets + p, stp, low_mask, entry);
      else
        launch_pass<KeyT, 8, uint32_t>(mode, (unsigned)plan.tiles, st, kin, vin, kout, vout, plan.n, plan.n_dev, plan.abort, shift, bits, h, tickets + p, stp, low_mask, entry);
    }
    if (launches) ++*launches;
    kin = kout; vin = vout;
    if (kout == keys_b) { kout = keys_a; vout = vals_a; } else { kout = keys_b; vout = vals_b; }
  }
  *result_in_a = (plan.passes % 2) == 0;
  return (int)cudaGetLastError();
}
template int launch_sort<uint32_t>(const SortPlan&, const uint32_t*, const uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                                   uint32_t*, const uint32_t*, uint32_t*, bool*, int*, cudaStream_t);
template int launch_sort<uint64_t>(const SortPlan&, const uint64_t*, const uint32_t*, uint64_t*, uint32_t*, uint64_t*,
                                   uint32_t*, const uint32_t*, uint32_t*, bool*, int*, cudaStream_t);

}  // namespace gsb
