// binning.cu -- tile binning: tile statistics (ranges, histograms, K), prefix sum of tile counts, key emission.
//
// Replaces the per-tile boolean masks of GaussianScene.render_image
// (splat/gaussian_scene.py:208-226): instead of an O(tiles*M) mask sweep, every in-view Gaussian is
// expanded into one (key, payload) pair per tile of its rect,
//     key = tile_id << 32 | float_as_uint(z_view),  tile_id = ty*tiles_x + tx,  payload = Gaussian index
// (SURVEY.md Appendix A.8).  The per-tile [start,end) ranges of the sorted array are known before any key exists:
// tile_stats_kernel derives them from the 2-D difference grid of tile rects written by the projection kernel.
//
//
// SPLIT mode bins in two levels: the Gaussians, already in depth order, are expanded into
// SUPER-TILE instances (a super-tile is 8 x 4 tiles: 255 of them at 1080p), those few keys take one stable radix
// pass per 8 bits of super-tile id (one at 1080p), and the last pass leaves, per instance, the Gaussian index and a
// 32-bit mask of the tiles of the super-tile that the rect covers.  A tile's list is its super-tile's list filtered
// by one bit; the compositing kernel does that on the fly, so no key is ever written per tile instance, nothing
// K-sized is sorted, and nothing K-sized is even stored.
//
// Roofline: HBM.  scan: 12 B read + 4 B written per Gaussian.  emit: 4-12 B written per key (+ 24 B per Gaussian
// read).  tile stats: 4 B per grid cell read, 8 B per tile written.  expand: 4 B written per tile instance.
#include "gsb_internal.cuh"

namespace gsb {

namespace {

__device__ __forceinline__ uint64_t ld_relaxed64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// emit: output-parallel, load-balanced expansion.  A block owns kEmitChunk consecutive OUTPUT slots, whatever
// Gaussians they belong to: two warps locate the first and last Gaussian of the chunk with a 32-way search of the
// offsets array (4-5 rounds of one coalesced probe per lane), the block stages those Gaussians (offset, index,
// rect) in shared memory in windows of kEmitWindow, and every thread fills aligned groups of four slots: one
// binary search in shared memory and one division locate the first slot of a group, the other three follow by
// stepping through the rect (and on to the next Gaussian); a full group leaves as one 16/32-byte vector store.
// A Gaussian covering 2 000 tiles costs the same per key as one covering 4, and a block of near, huge Gaussians
// costs the same as a block of far, tiny ones (blocks that own 256 GAUSSIANS do not: in depth order the first
// blocks emit 100x more than the last; measured 110 us vs 36 us once the tile-less Gaussians moved to the end).
//   kKind = kEmitFull    (FULL):  keys[o] = tile << 32 | depth bits,  payload[o] = Gaussian index
//   kKind = kEmitSplit64 (SPLIT): keys[o] = tile << 32 | Gaussian index  (depth order is the emission order;
//                                 the radix passes sort the tile field only and carry the index inside the key)
//   kKind = kEmitSplit32 (SPLIT): 32-bit keys[o] = tile << rank_bits | emission position.  The projection keys
//                                 Gaussians without tiles to the end of the depth order, so the position of an
//                                 emitting Gaussian is < V and rank_bits = ceil(log2 V); chosen by the host when
//                                 tile bits + rank_bits <= 32 -- half the bytes through every radix pass.
// ------------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;

// ------------------------------------------------------------------------------------------------
// exclusive scan of the tile counts in emission order: single pass, decoupled look-back, 8 192 items per CTA
// (about 120 CTAs at N = 1 M, so the prefix ripples through the grid in 4 window rounds; fusing the scan into
// emit_kernel's 3 900 CTAs was measured 20 us slower).  status[0] = ticket, ((u64*)status)[1 + b] = flag << 62
// | value of logical CTA b; 64-bit words because K may need more than the 30 bits a u32 leaves next to the flags.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

size_t scan_status_words(int64_t n) { return 2 * ((size_t)((n + kScanTile - 1) / kScanTile) + 2); }

// kFromRect: the scanned value is the number of SUPER-TILES the row's rect touches (rect.y < rect.x marks "no
// tiles"), computed on the fly; rows at emission positions >= *v_limit (the Gaussians without tiles, which the depth
// sort leaves at the end) are not even read.
__device__ __forceinline__ uint32_t coarse_count(ushort4 r, int lw, int lh) {
  if (r.y < r.x) return 0u;
  return (uint32_t)((r.y >> lw) - (r.x >> lw) + 1) * (uint32_t)((r.w >> lh) - (r.z >> lh) + 1);
}

template <bool kFromRect>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const uint32_t* __restrict__ count, const ushort4* __restrict__ rect, int lw, int lh,
            const uint32_t* __restrict__ perm, int64_t n, const uint32_t* __restrict__ v_limit,
            uint32_t* __restrict__ offsets, uint32_t* status) {
  constexpr uint64_t kAgg = 1ull << 62, kPre = 2ull << 62, kMask = (1ull << 62) - 1;
  __shared__ uint32_t s_block, s_excl;
  __shared__ uint32_t s_warp[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_block = atomicAdd(&status[0], 1u);  // ticket: predecessors are already running
  __syncthreads();
  const uint32_t b = s_block;
  uint64_t* st = reinterpret_cast<uint64_t*>(status) + 1;
  if (v_limit) {  // rows past the limit carry nothing: a CTA that starts there has no successor that needs it
    const int64_t lim = (int64_t)*v_limit;
    n = lim < n ? lim : n;
    if ((int64_t)b * kScanTile >= n) return;
  }
  const int64_t base = (int64_t)b * kScanTile + (int64_t)tid * kScanItems;
  uint32_t v[kScanItems];
  uint32_t local = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    uint32_t c = 0;
    if (i < n) {
      const int64_t g = perm ? (int64_t)perm[i] : i;
      c = kFromRect ? coarse_count(rect[g], lw, lh) : count[g];
    }
    v[k] = local;  // exclusive within the thread
    local += c;
  }
  uint32_t inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t warp_off = 0, block_sum = 0;
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) warp_off += t;
    block_sum += t;
  }
  if (warp == 0) {
    uint64_t excl = 0;
    if (b == 0) {
      if (lane == 0) st_relaxed64(&st[0], kPre | (uint64_t)block_sum);
    } else {
      if (lane == 0) st_relaxed64(&st[b], kAgg | (uint64_t)block_sum);
      for (int64_t idx = (int64_t)b - 1;; idx -= 32) {  // a window of 32 predecessors per round
        const int64_t j = idx - lane;
        uint64_t sv = kPre;  // virtual CTA -1: inclusive prefix 0
        if (j >= 0) {
          sv = ld_relaxed64(&st[j]);
          while ((sv >> 62) == 0) sv = ld_relaxed64(&st[j]);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (sv >> 62) == 2u);
        uint64_t contrib = sv & kMask;
        if (pm && lane > __ffs(pm) - 1) contrib = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (pm) break;
      }
      if (lane == 0) st_relaxed64(&st[b], kPre | ((excl + block_sum) & kMask));
    }
    if (lane == 0) s_excl = (uint32_t)excl;
  }
  __syncthreads();
  const uint32_t off = s_excl + warp_off + inc - local;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    if (i < n) offsets[i] = off + v[k];
  }
}

int launch_scan(const uint32_t* count, const uint32_t* perm, int64_t n, uint32_t* offsets, uint32_t* status,
                cudaStream_t st) {
  if (n == 0) return 0;
  scan_kernel<false><<<(unsigned)((n + kScanTile - 1) / kScanTile), kScanThreads, 0, st>>>(
      count, nullptr, 0, 0, perm, n, nullptr, offsets, status);
  return (int)cudaGetLastError();
}

int launch_scan_coarse(const ushort4* rect, SuperGeom sg, const uint32_t* perm, int64_t n, const uint32_t* v_limit,
                       uint32_t* offsets, uint32_t* status, cudaStream_t st) {
  if (n == 0) return 0;
  scan_kernel<true><<<(unsigned)((n + kScanTile - 1) / kScanTile), kScanThreads, 0, st>>>(
      nullptr, rect, sg.lw, sg.lh, perm, n, v_limit, offsets, status);
  return (int)cudaGetLastError();
}

enum EmitKind { kEmitFull = 0, kEmitSplit64 = 1, kEmitSplit32 = 2 };
// output slots per block: 8 192 for the tile-instance keys of FULL mode (~350 Gaussians per block: one staging
// window), 2 048 for the super-tile keys of SPLIT mode (a Gaussian emits 3-4 of those, so a block of 8 192 slots
// would walk 4-16 windows of dependent order -> rect gathers back to back: measured 46 us for 1.35 M keys)
constexpr int kEmitChunkFull = 8192;
constexpr int kEmitChunkSuper = 2048;
constexpr int kEmitWindow = 512;   // Gaussians staged per window
constexpr uint32_t kEmitRun = 16;  // consecutive output slots per thread and search

// The SPLIT kinds emit one key per SUPER-TILE of the rect (rect >> (lw, lh), tiles_x = super-tiles per row); with
// lw = lh = 0 that is one key per tile.  The grid is sized from the CAPACITY of the key buffer, not from the key
// count (which only the device knows when the kernel is queued): blocks past *total leave at once, and nothing
// runs when *abort is set (the count exceeded the capacity; the host re-queues the frame's tail after growing).
template <int kKind, int kEmitChunk>
__global__ void __launch_bounds__(kEmitThreads)
emit_kernel(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ total,
            int64_t n, const uint32_t* __restrict__ v_limit, const uint32_t* __restrict__ abort,
            const uint32_t* __restrict__ depth_key, const ushort4* __restrict__ rect, int tiles_x, int lw, int lh,
            int rank_bits, uint64_t* __restrict__ keys, uint32_t* __restrict__ payload) {
  constexpr bool kCombined = kKind != kEmitFull;
  __shared__ uint32_t s_off[kEmitWindow + 1];
  __shared__ uint32_t s_gid[kEmitWindow];
  __shared__ uint32_t s_low[kEmitWindow];  // low key word: depth bits (FULL), Gaussian index or emission position
  __shared__ ushort4 s_rect[kEmitWindow];
  __shared__ int64_t s_bound[2];
  if (abort && *abort) return;
  const uint32_t k_total = *total;
  const uint32_t c0 = blockIdx.x * (uint32_t)kEmitChunk;
  if (c0 >= k_total) return;
  if (v_limit) {  // rows past the limit touch no tile and their offsets were never written
    const int64_t lim = (int64_t)*v_limit;
    n = lim < n ? lim : n;
  }
  const uint32_t c1 = (k_total - c0 > (uint32_t)kEmitChunk) ? c0 + kEmitChunk : k_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 2) {
    // largest i with offsets[i] <= target (equal offsets = Gaussians without tiles: the last one wins, i.e. the one
    // that emits).  Invariant: offsets[lo] <= target < offsets[hi], with the virtual offsets[n] = K.
    const uint32_t target = warp == 0 ? c0 : c1 - 1u;
    int64_t lo = 0, hi = n;
    while (hi - lo > 1) {
      const int64_t step = (hi - lo + 32) / 33;
      const int64_t p = lo + (int64_t)(lane + 1) * step;
      const bool le = p < hi && offsets[p] <= target;
      const int cnt = __popc(__ballot_sync(0xffffffffu, le));  // monotone: the first cnt lanes
      const int64_t nhi = lo + (int64_t)(cnt + 1) * step;
      lo += (int64_t)cnt * step;
      hi = nhi < hi ? nhi : hi;
    }
    if (lane == 0) s_bound[warp] = lo;
  }
  __syncthreads();
  const int64_t g0 = s_bound[0], g1 = s_bound[1];
  for (int64_t w0 = g0; w0 <= g1; w0 += kEmitWindow) {
    const int cntw = (int)((g1 - w0 + 1 < (int64_t)kEmitWindow) ? g1 - w0 + 1 : (int64_t)kEmitWindow);
    if (w0 != g0) __syncthreads();  // everyone is done with the previous window
    for (int j = threadIdx.x; j < cntw; j += kEmitThreads) {
      const int64_t i = w0 + j;
      const uint32_t g = perm ? perm[i] : (uint32_t)i;
      s_off[j] = offsets[i];
      s_gid[j] = g;
      s_low[j] = kKind == kEmitSplit32 ? (uint32_t)i : (kKind == kEmitSplit64 ? g : depth_key[g]);
      const ushort4 rf = rect[g];
      s_rect[j] = kCombined ? make_ushort4(rf.x >> lw, rf.y >> lw, rf.z >> lh, rf.w >> lh) : rf;
    }
    if (threadIdx.x == 0) s_off[cntw] = (w0 + cntw < n) ? offsets[w0 + cntw] : k_total;
    __syncthreads();
    const uint32_t begin = s_off[0] > c0 ? s_off[0] : c0;
    const uint32_t end = s_off[cntw] < c1 ? s_off[cntw] : c1;
    // a thread owns kEmitRun consecutive slots (four aligned groups of four): ONE search and one division, then it
    // steps through rects; a warp's stores cover 32 * kEmitRun consecutive slots
    for (uint32_t o16 = (begin & ~(kEmitRun - 1u)) + kEmitRun * threadIdx.x; o16 < end; o16 += kEmitRun * kEmitThreads) {
      const uint32_t first = o16 > begin ? o16 : begin;
      const uint32_t last = (o16 + kEmitRun < end) ? o16 + kEmitRun : end;  // exclusive
      if (first >= last) continue;
      // largest j with s_off[j] <= first  (entries with count 0 share their successor's offset and lose)
      int lo = 0, hi = cntw;  // invariant: s_off[lo] <= first < s_off[hi]
#pragma unroll
      for (int step = 0; step < 9; ++step) {  // 2^9 = kEmitWindow
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= first) lo = mid; else hi = mid;
      }
      ushort4 r = s_rect[lo];
      uint32_t nxt_off = s_off[lo + 1];
      const uint32_t t = first - s_off[lo];
      const uint32_t w = (uint32_t)r.y - (uint32_t)r.x + 1u;
      uint32_t ty = (uint32_t)r.z + t / w;
      uint32_t tx = (uint32_t)r.x + t % w;
      uint32_t low = s_low[lo], gid = s_gid[lo];
#pragma unroll
      for (uint32_t o4 = o16; o4 < o16 + kEmitRun; o4 += 4u) {
        if (o4 + 4u <= first || o4 >= last) continue;
        uint32_t tv[4], lv[4], pv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t slot = o4 + q;
          tv[q] = 0; lv[q] = 0; pv[q] = 0;
          if (slot >= first && slot < last) {
            while (slot >= nxt_off) {  // move on to the next Gaussian with a non-empty rect
              ++lo;
              nxt_off = s_off[lo + 1];
              r = s_rect[lo];
              low = s_low[lo]; gid = s_gid[lo];
              tx = r.x; ty = r.z;
            }
            tv[q] = ty * (uint32_t)tiles_x + tx;
            lv[q] = low;
            pv[q] = gid;
            if (++tx > (uint32_t)r.y) { tx = r.x; ++ty; }
          }
        }
        const bool whole = o4 >= first && o4 + 4u <= last;
        if (kKind == kEmitSplit32) {
          // 32-bit keys: tile << rank_bits | emission position
          uint32_t* k32 = reinterpret_cast<uint32_t*>(keys);
          uint32_t v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = (tv[q] << rank_bits) | lv[q];
          if (whole) {
            *reinterpret_cast<uint4*>(k32 + o4) = make_uint4(v[0], v[1], v[2], v[3]);  // o4 % 4 == 0: 16-byte aligned
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t slot = o4 + q;
              if (slot >= first && slot < last) k32[slot] = v[q];
            }
          }
        } else {
          uint64_t kv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) kv[q] = ((uint64_t)tv[q] << 32) | (uint64_t)lv[q];
          if (whole) {
            ulonglong2* kp = reinterpret_cast<ulonglong2*>(keys + o4);  // o4 % 4 == 0: 32-byte aligned
            kp[0] = make_ulonglong2(kv[0], kv[1]);
            kp[1] = make_ulonglong2(kv[2], kv[3]);
            if (!kCombined) *reinterpret_cast<uint4*>(payload + o4) = make_uint4(pv[0], pv[1], pv[2], pv[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t slot = o4 + q;
              if (slot >= first && slot < last) {
                keys[slot] = kv[q];
                if (!kCombined) payload[slot] = pv[q];
              }
            }
          }
        }
      }
    }
  }
}

int launch_emit(const uint32_t* offsets, const uint32_t* perm, const uint32_t* total, int64_t n, int64_t k,
                const uint32_t* depth_key, const ushort4* rect, int tiles_x, uint64_t* keys, uint32_t* payload,
                cudaStream_t st) {
  if (n == 0 || k <= 0) return 0;
  unsigned blocks = (unsigned)((k + kEmitChunkFull - 1) / kEmitChunkFull);
  emit_kernel<kEmitFull, kEmitChunkFull><<<blocks, kEmitThreads, 0, st>>>(offsets, perm, total, n, nullptr, nullptr, depth_key, rect,
                                                          tiles_x, 0, 0, 0, keys, payload);
  return (int)cudaGetLastError();
}

int launch_emit_coarse(const uint32_t* offsets, const uint32_t* perm, const uint32_t* total, int64_t n,
                       const uint32_t* v_limit, const uint32_t* abort, int64_t capacity, const ushort4* rect,
                       SuperGeom sg, int rank_bits, void* keys, cudaStream_t st) {
  if (n == 0 || capacity <= 0) return 0;
  unsigned blocks = (unsigned)((capacity + kEmitChunkSuper - 1) / kEmitChunkSuper);
  if (rank_bits > 0)
    emit_kernel<kEmitSplit32, kEmitChunkSuper><<<blocks, kEmitThreads, 0, st>>>(offsets, perm, total, n, v_limit, abort, nullptr, rect,
                                                               sg.nx, sg.lw, sg.lh, rank_bits,
                                                               reinterpret_cast<uint64_t*>(keys), nullptr);
  else
    emit_kernel<kEmitSplit64, kEmitChunkSuper><<<blocks, kEmitThreads, 0, st>>>(offsets, perm, total, n, v_limit, abort, nullptr, rect,
                                                               sg.nx, sg.lw, sg.lh, 0,
                                                               reinterpret_cast<uint64_t*>(keys), nullptr);
  return (int)cudaGetLastError();
}

// Debug surface of SPLIT mode: materialise the sorted 64-bit keys tile << 32 | depth bits, which the
// production path never needs (compositing reads payload + ranges only).  One warp per tile.
__global__ void __launch_bounds__(256)
rebuild_keys_kernel(const uint2* __restrict__ ranges, int tiles, const uint32_t* __restrict__ payload,
                    const uint32_t* __restrict__ depth_key, uint64_t* __restrict__ keys) {
  const int t = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (t >= tiles) return;
  const uint2 rg = ranges[t];
  for (uint32_t j = rg.x + (threadIdx.x & 31); j < rg.y; j += 32)
    keys[j] = ((uint64_t)(uint32_t)t << 32) | (uint64_t)depth_key[payload[j]];
}

int launch_rebuild_keys(const uint2* ranges, int tiles, const uint32_t* payload, const uint32_t* depth_key,
                        uint64_t* keys, cudaStream_t st) {
  if (tiles <= 0) return 0;
  rebuild_keys_kernel<<<(unsigned)((tiles * 32 + 255) / 256), 256, 0, st>>>(ranges, tiles, payload, depth_key, keys);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// expand: per-tile lists from per-super-tile lists (SPLIT mode; ON DEMAND only).
//
// After the radix pass(es) over the super-tile ids, clist[ranges_s[s].x .. ranges_s[s].y) holds one entry
// {Gaussian index, tile mask} per Gaussian whose rect touches super-tile s, in depth order; bit ly << lw | lx of the
// mask says whether the rect covers tile (lx, ly) of the super-tile (at most 32 tiles: one word).  The list of tile
// t is exactly the entries of its super-tile whose bit t is set, in the same order.  The compositing kernel reads
// the super-tile lists directly and filters on the fly (composite.cu), so the frame path never materialises the
// per-tile lists; this kernel does it for whoever wants them as arrays: the parity surface (gsb_debug_sorted_keys)
// and nothing else.  One warp per tile: 32 entries per step, one ballot, the hits leave as one run of consecutive
// 4-byte stores at payload[ranges[t].x ..] -- the starts are known from tile_stats_kernel.  No atomics, no keys.
// ------------------------------------------------------------------------------------------------
constexpr int kExpWarps = 16;
constexpr int kExpThreads = kExpWarps * 32;

__global__ void __launch_bounds__(kExpThreads)
expand_kernel(const uint2* __restrict__ ranges_s, const uint2* __restrict__ clist, const uint2* __restrict__ ranges,
              uint32_t* __restrict__ payload, int tiles_x, int tiles_y, int snx, int lw, int lh,
              const uint32_t* __restrict__ abort) {
  __shared__ uint2 s_e[2][kExpThreads];
  if (abort && *abort) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_super = ((1 << (lw + lh)) + kExpWarps - 1) / kExpWarps;  // CTAs per super-tile
  const int s = blockIdx.x / per_super, sub = blockIdx.x - s * per_super;
  const int lt = sub * kExpWarps + warp;     // tile inside the super-tile, row-major
  const uint32_t tx = (uint32_t)(((s % snx) << lw) + (lt & ((1 << lw) - 1)));
  const uint32_t ty = (uint32_t)(((s / snx) << lh) + (lt >> lw));
  const bool valid = lt < (1 << (lw + lh)) && tx < (uint32_t)tiles_x && ty < (uint32_t)tiles_y;
  const uint2 rs = ranges_s[s];
  const uint32_t len = rs.y - rs.x;
  const uint2* list = clist + rs.x;
  uint32_t out = valid ? ranges[ty * (uint32_t)tiles_x + tx].x : 0u;
  const unsigned lt_mask = (1u << lane) - 1u;

  uint2 e_next = make_uint2(0u, 0u);
  if ((uint32_t)tid < len) e_next = list[tid];
  for (uint32_t w0 = 0, it = 0; w0 < len; w0 += kExpThreads, ++it) {
    const int buf = (int)(it & 1u);
    s_e[buf][tid] = e_next;
    __syncthreads();  // also orders the reads of this buffer two windows ago before the writes above
    if (w0 + kExpThreads + tid < len) e_next = list[w0 + kExpThreads + tid];
    const uint32_t cnt = len - w0 < (uint32_t)kExpThreads ? len - w0 : (uint32_t)kExpThreads;
    if (valid) {
      for (uint32_t c = 0; c < cnt; c += 32) {
        const uint32_t e = c + lane;
        uint2 v = make_uint2(0u, 0u);
        if (e < cnt) v = s_e[buf][e];
        const bool hit = e < cnt && ((v.y >> lt) & 1u);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) payload[out + (uint32_t)__popc(m & lt_mask)] = v.x;
        out += (uint32_t)__popc(m);
      }
    }
  }
}

int launch_expand(const uint2* ranges_s, const uint2* clist, const uint2* ranges, uint32_t* payload, FrameGeom geom,
                  SuperGeom sg, const uint32_t* abort, cudaStream_t st) {
  const int supers = sg.nx * sg.ny;
  if (supers <= 0) return 0;
  const int per_super = ((1 << (sg.lw + sg.lh)) + kExpWarps - 1) / kExpWarps;
  expand_kernel<<<(unsigned)(supers * per_super), kExpThreads, 0, st>>>(ranges_s, clist, ranges, payload, geom.tiles_x,
                                                                       geom.tiles_y, sg.nx, sg.lw, sg.lh, abort);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// tile statistics from the 2-D difference grid written by the projection kernel.
// One CTA (the grid has (tiles_x+1)*(tiles_y+1) cells: 8 349 at 1080p, 33 k at 4K; it lives in L2).
//   1. prefix sum along x (one warp per row), 2. prefix sum along y (one thread per column)
//      -> cell (ty,tx) = number of instances of tile (ty,tx)  [exact: integer arithmetic]
//   3. exclusive scan over tiles in row-major order -> ranges[t] = (start, start+count), (0,0) if empty
//   4. histograms of the tile-id digits for the radix passes over the tile field; total K.
// ------------------------------------------------------------------------------------------------
constexpr int kStatThreads = 1024;

__global__ void __launch_bounds__(kStatThreads)
tile_stats_kernel(int32_t* __restrict__ grid_global, int use_smem, int tiles_x, int tiles_y,
                  uint32_t* __restrict__ tile_hist, uint2* __restrict__ ranges, uint32_t* ctl, StatsPost post) {
  extern __shared__ int32_t s_grid[];
  __shared__ uint32_t s_hist[4][kRadix];
  __shared__ uint32_t s_warp[kStatThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gw = tiles_x + 1;
  for (int t = tid; t < 4 * kRadix; t += kStatThreads) (&s_hist[0][0])[t] = 0;
  // the column pass is a chain of dependent accesses: keep the grid in shared memory whenever it fits
  int32_t* grid = grid_global;
  if (use_smem) {
    const int cells = gw * (tiles_y + 1);
    for (int t = tid; t < cells; t += kStatThreads) s_grid[t] = grid_global[t];
    grid = s_grid;
    __syncthreads();
  }
  // 1. rows
  for (int r = warp; r < tiles_y; r += kStatThreads / 32) {
    int32_t carry = 0;
    for (int x0 = 0; x0 < tiles_x; x0 += 32) {
      const int x = x0 + lane;
      int32_t v = x < tiles_x ? grid[r * gw + x] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      v += carry;
      if (x < tiles_x) grid[r * gw + x] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  // 2. columns
  for (int x = tid; x < tiles_x; x += kStatThreads) {
    int32_t acc = 0;
    for (int r = 0; r < tiles_y; ++r) {
      acc += grid[r * gw + x];
      grid[r * gw + x] = acc;
    }
  }
  __syncthreads();
  // 3. exclusive scan in row-major tile order; each thread owns a contiguous chunk of tiles
  const int tiles = tiles_x * tiles_y;
  const int per = (tiles + kStatThreads - 1) / kStatThreads;
  const int t0 = tid * per, t1 = min(t0 + per, tiles);
  uint32_t local = 0;
  for (int t = t0; t < t1; ++t) local += (uint32_t)grid[(t / tiles_x) * gw + (t % tiles_x)];
  uint32_t inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t off = 0;
  for (int w = 0; w < kStatThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) off += t;
  }
  // exact 64-bit total (the u32 scan above wraps when K >= 2^32; the host then rejects the frame)
  __shared__ unsigned long long s_total;
  if (tid == 0) s_total = 0;
  __syncthreads();
  {
    unsigned long long l64 = local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l64 += __shfl_xor_sync(0xffffffffu, l64, o);
    if (lane == 0) atomicAdd(&s_total, l64);
  }
  __syncthreads();
  const unsigned long long total64 = s_total;
  uint32_t run = off + inc - local;
  // histogram rows: row p counts digit (tile >> 8p) & 255.  A thread's tiles are consecutive, so the
  // higher digits repeat: accumulate runs locally and issue one shared atomic per run.
  uint32_t acc[3] = {0, 0, 0};
  uint32_t cur[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  for (int t = t0; t < t1; ++t) {
    const uint32_t c = (uint32_t)grid[(t / tiles_x) * gw + (t % tiles_x)];
    ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
    if (c) {
      atomicAdd(&s_hist[0][(uint32_t)t & 255u], c);
#pragma unroll
      for (int p = 1; p < 4; ++p) {
        const uint32_t d = ((uint32_t)t >> (8 * p)) & 255u;
        if (d != cur[p - 1]) {
          if (acc[p - 1]) atomicAdd(&s_hist[p][cur[p - 1]], acc[p - 1]);
          cur[p - 1] = d; acc[p - 1] = 0;
        }
        acc[p - 1] += c;
      }
    }
    run += c;
  }
#pragma unroll
  for (int p = 1; p < 4; ++p)
    if (acc[p - 1]) atomicAdd(&s_hist[p][cur[p - 1]], acc[p - 1]);
  __syncthreads();
  if (tile_hist)
    for (int t = tid; t < 4 * kRadix; t += kStatThreads) tile_hist[t] = (&s_hist[0][0])[t];
  if (tid == 0) {  // 64-bit totals: the host rejects K >= 2^32 (positions are u32)
    uint32_t* mine = ctl + (post.level ? kCtlKs : kCtlK);
    mine[0] = (uint32_t)total64;
    mine[1] = (uint32_t)(total64 >> 32);
    if (post.enabled) {
      // K = tile instances, Ks = super-tile instances (= K when the frame is binned in one level)
      unsigned long long k = total64, ks = total64;
      if (post.level) k = ((unsigned long long)ctl[kCtlK + 1] << 32) | ctl[kCtlK];  // written by the previous launch
      else { ctl[kCtlKs] = (uint32_t)total64; ctl[kCtlKs + 1] = (uint32_t)(total64 >> 32); }
      // The kernels that consume these counts are ALREADY queued, with grids and buffers sized from the context's
      // capacities: if a count does not fit they must not run (the host grows the buffers and re-queues them).
      const uint32_t ab = (k > post.cap_k || ks > post.cap_ks) ? 1u : 0u;
      ctl[kCtlAbort] = ab;
      if (post.mailbox) {
        // Mailbox in mapped pinned host memory: the host learns M, K and the verdict without draining the stream.
        volatile uint32_t* box = post.mailbox;
        box[0] = ctl[kCtlM];
        box[1] = ctl[kCtlVisible];  // V: Gaussians with tiles
        box[2] = (uint32_t)k;
        box[3] = (uint32_t)(k >> 32);
        box[5] = ab;
        box[6] = (uint32_t)ks;
        box[7] = (uint32_t)(ks >> 32);
        __threadfence_system();
        box[4] = post.seq;
      }
    }
  }
}

int launch_tile_stats(int32_t* diff_grid, int tiles_x, int tiles_y, uint32_t* tile_hist, uint2* ranges, uint32_t* ctl,
                      const StatsPost& post, cudaStream_t st) {
  if (tiles_x <= 0 || tiles_y <= 0) return 0;
  const size_t bytes = (size_t)(tiles_x + 1) * (size_t)(tiles_y + 1) * sizeof(int32_t);
  const int use_smem = bytes <= 200 * 1024;
  if (use_smem && bytes > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(tile_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  tile_stats_kernel<<<1, kStatThreads, use_smem ? bytes : 0, st>>>(diff_grid, use_smem, tiles_x, tiles_y, tile_hist,
                                                                  ranges, ctl, post);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------
__global__ void iota_kernel(uint32_t* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}
int launch_iota(uint32_t* p, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n);
  return (int)cudaGetLastError();
}

// (H,W,3) -> (W,H,3): the layout GaussianScene.render_image returns (splat/gaussian_scene.py:206,:227)
__global__ void hwc_to_whc_kernel(const float* __restrict__ src, float* __restrict__ dst, int W, int H) {
  __shared__ float tile[32][33 * 3];
  int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int y = y0 + r;
    for (int c = threadIdx.x; c < 96; c += blockDim.x) {
      int x = x0 + c / 3;
      if (x < W && y < H) tile[r][c] = src[((size_t)y * W + x0) * 3 + c];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r = x within tile
    int x = x0 + r;
    for (int c = threadIdx.x; c < 96; c += blockDim.x) {
      int yy = c / 3, ch = c % 3;
      int y = y0 + yy;
      if (x < W && y < H) dst[((size_t)x * H + y0) * 3 + c] = tile[yy][r * 3 + ch];
    }
  }
}
int launch_hwc_to_whc(const float* src, float* dst, int W, int H, cudaStream_t st) {
  dim3 grid((W + 31) / 32, (H + 31) / 32), block(32, 8);
  hwc_to_whc_kernel<<<grid, block, 0, st>>>(src, dst, W, H);
  return (int)cudaGetLastError();
}

__global__ void to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = fminf(fmaxf(src[i], 0.f), 1.f);
    dst[i] = (uint8_t)__float2int_rn(v * 255.f);
  }
}
int launch_to_u8(const float* src, uint8_t* dst, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  to_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, n);
  return (int)cudaGetLastError();
}

// PreprocessedScene rows in depth order (gsb_preprocess): gather through `order`.
__global__ void gather_preprocess_kernel(const uint32_t* __restrict__ order, int64_t m, const float4* __restrict__ rec,
                                         const float* __restrict__ planes, int64_t n_pad,
                                         const uint32_t* __restrict__ depth_key, DebugOut dbg, float* points_xy,
                                         float* colors, float* cov2d, float* depths, float* conic, float* radius,
                                         float* min_x, float* min_y, float* max_x, float* max_y, float* sig_op,
                                         int32_t* src_index) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint32_t g = order[j];
  const float4 r0 = rec[3 * (size_t)g], r2 = rec[3 * (size_t)g + 2];
  if (points_xy) { points_xy[2 * j] = r0.x; points_xy[2 * j + 1] = r0.y; }
  if (colors) {
    colors[3 * j] = planes[PR * n_pad + g];
    colors[3 * j + 1] = planes[PG * n_pad + g];
    colors[3 * j + 2] = planes[PB * n_pad + g];
  }
  if (depths) depths[j] = __uint_as_float(depth_key[g]);
  if (radius) radius[j] = r2.z;
  if (sig_op) sig_op[j] = r2.w;
  if (src_index) src_index[j] = (int32_t)g;
  // destinations are packed back to back in one staging block: only 4-byte alignment is guaranteed
  if (cov2d) {
    const float4 v = reinterpret_cast<const float4*>(dbg.cov2d)[g];
    cov2d[4 * j] = v.x; cov2d[4 * j + 1] = v.y; cov2d[4 * j + 2] = v.z; cov2d[4 * j + 3] = v.w;
  }
  if (conic) {
    const float4 v = reinterpret_cast<const float4*>(dbg.conic)[g];
    conic[4 * j] = v.x; conic[4 * j + 1] = v.y; conic[4 * j + 2] = v.z; conic[4 * j + 3] = v.w;
  }
  const float4 bb = reinterpret_cast<const float4*>(dbg.bbox)[g];
  if (min_x) min_x[j] = bb.x;
  if (min_y) min_y[j] = bb.y;
  if (max_x) max_x[j] = bb.z;
  if (max_y) max_y[j] = bb.w;
}

int launch_gather_preprocess(const uint32_t* order, int64_t m, const float4* rec, const float* planes, int64_t n_pad,
                             const uint32_t* depth_key, const DebugOut& dbg, float* points_xy, float* colors,
                             float* cov2d, float* depths, float* conic, float* radius, float* min_x, float* min_y,
                             float* max_x, float* max_y, float* sig_op, int32_t* src_index, cudaStream_t st) {
  if (m == 0) return 0;
  gather_preprocess_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(order, m, rec, planes, n_pad, depth_key, dbg,
                                                                      points_xy, colors, cov2d, depths, conic, radius,
                                                                      min_x, min_y, max_x, max_y, sig_op, src_index);
  return (int)cudaGetLastError();
}

}  // namespace gsb
