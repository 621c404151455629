// Hand-written radix sort of the MSM's (bucket key, point index) pairs: LSD passes (below), and for the
// partitioned lists of the large windows an MSD pass followed by a group-local sort in shared memory
// (radix_sort_pairs_msd at the end of the file).
//
// 8 bits per pass, ceil(c/8) passes for a c-bit key.  Each pass is three kernels:
//   sort_tile_hist    per-tile digit histogram (tile = 256 threads x 16 entries),
//                     written digit-major so one exclusive scan gives every
//                     (digit, tile) its global base
//   scan_*            exclusive scan of the 256 x tiles counters
//   sort_tile_scatter stable multi-split of the tile in shared memory (per-warp
//                     digit counters + __match_any_sync ranking), then a coalesced
//                     copy-out of each digit run to its global base
// Per pass and entry: 4 B (histogram read) + 8 B read + 8 B write = 20 B of HBM traffic.
// The scatter kernel is persistent (2 blocks per SM) and prefetches the next tile into
// registers; peer masks come from ballots (MATCH.ANY was the bottleneck of an earlier version).
// Stability is not needed by the MSM (bucket sums commute) but keeps the sort a plain
// LSD radix sort whose result is independent of scheduling.
#include <stdlib.h>
#include "ctx.cuh"

namespace mz {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 entries
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRadix = 256;

// ---------------------------------------------------------------------------
// exclusive scan of a uint32 array (reduce - scan - apply)
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += o;
  }
  return v;
}
// block-wide exclusive scan of one value per thread; returns the exclusive prefix and the block total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
    uint32_t si = warp_incl_scan(s);
    warp_sums[lane] = si - s;  // exclusive warp bases
    if (lane == 31) warp_sums[32] = si;
  }
  __syncthreads();
  total = warp_sums[32];
  uint32_t r = warp_sums[w] + incl - v;
  __syncthreads();
  return r;
}

// d_tiles (optional): the array really holds *d_tiles x kRadix counters (known on the device only); chunks past
// that are skipped
__global__ void __launch_bounds__(kScanThreads) scan_chunk_sums(const uint32_t* __restrict__ in, size_t n,
                                                                uint32_t* __restrict__ sums,
                                                                const uint32_t* __restrict__ d_tiles) {
  __shared__ uint32_t ws[33];
  size_t base = (size_t)blockIdx.x * kScanChunk;
  if (d_tiles) {
    const size_t lim = (size_t)*d_tiles * 256;
    if (lim < n) n = lim;
    if (base >= n) {
      if (threadIdx.x == 0) sums[blockIdx.x] = 0;
      return;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    size_t idx = base + (size_t)i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint32_t total;
  block_excl_scan(s, ws, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// single block: in-place exclusive scan of the chunk sums
__global__ void __launch_bounds__(kScanThreads) scan_sums_inplace(uint32_t* sums, size_t m) {
  __shared__ uint32_t ws[33];
  uint32_t carry = 0;
  for (size_t base = 0; base < m; base += kScanThreads) {
    size_t idx = base + threadIdx.x;
    uint32_t v = idx < m ? sums[idx] : 0;
    uint32_t total;
    uint32_t ex = block_excl_scan(v, ws, total);
    if (idx < m) sums[idx] = carry + ex;
    carry += total;
  }
}
__global__ void __launch_bounds__(kScanThreads) scan_apply(uint32_t* __restrict__ data, size_t n,
                                                           const uint32_t* __restrict__ sums,
                                                           const uint32_t* __restrict__ d_tiles) {
  __shared__ uint32_t ws[33];
  if (d_tiles) {
    const size_t lim = (size_t)*d_tiles * 256;
    if (lim < n) n = lim;
    if ((size_t)blockIdx.x * kScanChunk >= n) return;
  }
  size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * kScanItems;  // blocked: 8 consecutive per thread
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = (base + i < n) ? data[base + i] : 0;
    s += v[i];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan(s, ws, total) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) data[base + i] = ex;
    ex += v[i];
  }
}

// ---------------------------------------------------------------------------
// pass kernels
// ---------------------------------------------------------------------------
// Lanes of the warp holding the same 8-bit digit (and the same validity): nine ballots and a
// few logic ops.  MATCH.ANY computes the same mask in one instruction, but its throughput
// (~1 per 64 cycles per SM, measured through the scatter kernel) made it the bottleneck.
__device__ __forceinline__ uint32_t warp_match_digit(uint32_t d, bool valid) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
  if (!valid) peers = ~peers;
#pragma unroll
  for (int b = 0; b < 8; b++) {
    const bool bit = (d >> b) & 1u;
    const uint32_t ball = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? ball : ~ball;
  }
  return peers;
}

// Partitioned sort: the entry list is a sequence of P partitions (MSD split done by the recode kernel,
// msm.cu) that are sorted independently by their low key bits.  parts = [part_base[P + 1] | tile_start[P + 1]]
// on the device: partition p owns entries [part_base[p], part_base[p + 1]) and tiles
// [tile_start[p], tile_start[p + 1]).  Its histogram block is laid out [digit][tile of p] right behind the
// blocks of the partitions before it, so ONE exclusive scan over everything yields, for every (partition,
// digit, tile), the global position of that run: everything in front of a partition sums to part_base[p].
// parts == nullptr: a single partition [0, n) with ntiles tiles (plain LSD radix sort).
struct TileRef {
  uint64_t base;       // first entry of the tile
  uint32_t count;      // entries in the tile (0: the tile does not exist)
  size_t hist_base;    // index of (digit 0, this tile) in the histogram array
  uint32_t hist_stride;  // distance between consecutive digits
  // The scanned histogram counts only the entries that HAVE tiles.  When every partition has them, the scanned value
  // of a partition's first slot equals its base; when some have none (the oversize-group pass of the MSD sort) it does
  // not, so a run's global position is  scanned[slot] - scanned[first_slot] + part_base.
  size_t first_slot;
  uint32_t part_base;
};
__device__ __forceinline__ TileRef locate_tile(uint32_t tile, const uint32_t* __restrict__ parts, int P, uint64_t n,
                                               uint32_t ntiles) {
  TileRef r;
  if (!parts) {
    r.base = (uint64_t)tile * kSortTile;
    r.count = tile < ntiles ? (uint32_t)((n - r.base) < (uint64_t)kSortTile ? (n - r.base) : kSortTile) : 0u;
    r.hist_base = tile;
    r.hist_stride = ntiles;
    r.first_slot = 0;
    r.part_base = 0;
    return r;
  }
  const uint32_t* part_base = parts;
  const uint32_t* tile_start = parts + (P + 1);
  if (tile >= tile_start[P]) {
    r.base = 0; r.count = 0; r.hist_base = 0; r.hist_stride = 0; r.first_slot = 0; r.part_base = 0;
    return r;
  }
  int lo = 0, hi = P;  // tile_start[lo] <= tile < tile_start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tile_start[mid] <= tile) lo = mid; else hi = mid;
  }
  const uint32_t t0 = tile_start[lo], tl = tile - t0;
  r.base = (uint64_t)part_base[lo] + (uint64_t)tl * kSortTile;
  const uint64_t left = (uint64_t)part_base[lo + 1] - r.base;
  r.count = (uint32_t)(left < (uint64_t)kSortTile ? left : kSortTile);
  r.hist_stride = tile_start[lo + 1] - t0;
  r.hist_base = (size_t)kRadix * t0 + tl;
  r.first_slot = (size_t)kRadix * t0;
  r.part_base = part_base[lo];
  return r;
}

__global__ void __launch_bounds__(kSortThreads) sort_tile_hist(const uint32_t* __restrict__ keys, uint64_t n, int shift,
                                                               uint32_t* __restrict__ hist, uint32_t ntiles_arg,
                                                               const uint32_t* __restrict__ parts, int P) {
  __shared__ uint32_t h[kRadix];
  __shared__ TileRef s_ref;
  // grid-stride over the tiles; in partition mode the real tile count is on the device (tile_start[P]), so a
  // launch sized for the upper bound costs nothing when few (or no) tiles exist
  const uint32_t ntiles = parts ? parts[2 * P + 1] : ntiles_arg;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    h[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_ref = locate_tile(tile, parts, P, n, ntiles);
    __syncthreads();
    const TileRef ref = s_ref;
    if (ref.count != 0) {
      const uint64_t base = ref.base;
      // all loads first (16-byte vectors when the tile is whole and aligned; order inside the tile is
      // irrelevant for counting), then the shared-memory atomics
      uint32_t k[kSortItems];
      if (ref.count == kSortTile && (base & 3) == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(keys + base);
#pragma unroll
        for (int i = 0; i < kSortItems / 4; i++) {
          uint4 q = __ldg(src + i * kSortThreads + threadIdx.x);
          k[4 * i] = q.x; k[4 * i + 1] = q.y; k[4 * i + 2] = q.z; k[4 * i + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < kSortItems; i++) atomicAdd(&h[(k[i] >> shift) & (kRadix - 1)], 1u);
      } else {
        // partial or unaligned tile (partitions start anywhere): scalar loads, still all issued before the atomics
#pragma unroll
        for (int i = 0; i < kSortItems; i++) {
          uint32_t p = (uint32_t)i * kSortThreads + threadIdx.x;
          k[i] = p < ref.count ? __ldg(keys + base + p) : 0u;
        }
#pragma unroll
        for (int i = 0; i < kSortItems; i++) {
          uint32_t p = (uint32_t)i * kSortThreads + threadIdx.x;
          if (p < ref.count) atomicAdd(&h[(k[i] >> shift) & (kRadix - 1)], 1u);
        }
      }
      __syncthreads();
      hist[ref.hist_base + (size_t)threadIdx.x * ref.hist_stride] = h[threadIdx.x];
    }
    __syncthreads();  // h and s_ref are reused by the next tile
  }
}

// Persistent: each block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  In the stable form the next tile's
// keys, values and digit bases are fetched into registers before the current tile is ranked, so the DRAM latency is
// covered by the ranking work rather than by occupancy; the unordered form needs fewer registers and does better
// with a third resident block instead (kOcc).
// kStable = false: entries of equal digit may leave in any order (allowed for a pass whose input order carries
// no information - the first pass after the recode): ranks come from one shared-memory atomic per entry
// instead of the ballot match and the per-warp counters, about a third of the instructions.
// kOcc = resident blocks per SM the kernel is compiled for: 2 = with the register prefetch of the next tile; 3 / 4 = no
// prefetch, the latency is covered by more resident warps instead (A/B on B200 in profiles/experiments_r2.md).
template <bool kStable, int kOcc>
__global__ void __launch_bounds__(kSortThreads, kOcc)
    sort_tile_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t n, int shift,
                      const uint32_t* __restrict__ hist_scanned, uint32_t ntiles_arg, uint32_t* __restrict__ keys_out,
                      uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ parts, int P) {
  __shared__ uint32_t s_keys[kSortTile];
  __shared__ uint32_t s_vals[kSortTile];
  __shared__ uint32_t cnt[kSortWarps][kRadix];  // per-warp digit counters, then per-warp bases
  __shared__ uint32_t digit_start[kRadix];      // start of each digit run inside the tile
  __shared__ uint32_t gbase[kRadix];            // global base of each digit run of this tile
  __shared__ uint32_t ws[33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // warp-blocked arrangement: warp w owns tile entries [w*512, (w+1)*512) in index order
  const uint32_t wbase = (uint32_t)w * (32 * kSortItems);

  const uint32_t ntiles = parts ? parts[2 * P + 1] : ntiles_arg;  // tile_start[P]
  uint32_t k[kSortItems], v[kSortItems], kn[kSortItems], vn[kSortItems], rank[kSortItems];
  uint32_t gb = 0, gbn = 0, tn_next = 0;
  auto fetch = [&](uint32_t tile, uint32_t* kk, uint32_t* vv, uint32_t& g) {
    const TileRef ref = locate_tile(tile, parts, P, n, ntiles);
    const uint64_t tb = ref.base;
    const uint32_t tn = ref.count;
    tn_next = tn;
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      uint32_t p = wbase + i * 32 + lane;
      kk[i] = p < tn ? keys_in[tb + p] : 0xffffffffu;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      uint32_t p = wbase + i * 32 + lane;
      vv[i] = p < tn ? vals_in[tb + p] : 0u;
    }
    g = hist_scanned[ref.hist_base + (size_t)threadIdx.x * ref.hist_stride];
    if (parts && tn) g = g - hist_scanned[ref.first_slot] + ref.part_base;
  };
  constexpr bool kPrefetch = kOcc <= 2;
  uint32_t tile = blockIdx.x;
  if (kPrefetch && tile < ntiles) fetch(tile, kn, vn, gbn);
  for (; tile < ntiles; tile += gridDim.x) {
    if (kPrefetch) {
#pragma unroll
      for (int i = 0; i < kSortItems; i++) { k[i] = kn[i]; v[i] = vn[i]; }
      gb = gbn;
    } else {
      fetch(tile, k, v, gb);
    }
    const uint32_t tile_n = tn_next;  // set by the fetch of THIS tile (before the next one is prefetched)
    if (kPrefetch && tile + gridDim.x < ntiles) fetch(tile + gridDim.x, kn, vn, gbn);  // in flight during the ranking below
#pragma unroll
    for (int i = 0; i < kSortWarps; i++) cnt[i][threadIdx.x] = 0;
    gbase[threadIdx.x] = gb;
    __syncthreads();
    if (kStable) {
      // all 16 matches first (they only depend on the keys, so they pipeline), then the
      // sequential per-warp counter updates
#pragma unroll
      for (int i = 0; i < kSortItems; i++) {
        uint32_t p = wbase + i * 32 + lane;
        uint32_t d = (k[i] >> shift) & (kRadix - 1);
        rank[i] = warp_match_digit(d, p < tile_n);  // peers: lanes of this warp with the same digit in this round
      }
#pragma unroll
      for (int i = 0; i < kSortItems; i++) {
        uint32_t p = wbase + i * 32 + lane;
        bool ok = p < tile_n;
        uint32_t d = (k[i] >> shift) & (kRadix - 1);
        uint32_t peers = rank[i];
        uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (ok && before == 0) {  // lowest lane of the peer group bumps the warp counter
          prev = cnt[w][d];
          cnt[w][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
        rank[i] = prev + before;
        __syncwarp();
      }
    } else {
#pragma unroll
      for (int i = 0; i < kSortItems; i++) {
        uint32_t p = wbase + i * 32 + lane;
        uint32_t d = (k[i] >> shift) & (kRadix - 1);
        rank[i] = p < tile_n ? atomicAdd(&cnt[0][d], 1u) : 0u;
      }
    }
    __syncthreads();
    // digit totals -> exclusive scan over digits -> per-warp bases
    {
      uint32_t tot = 0;
#pragma unroll
      for (int i = 0; i < kSortWarps; i++) tot += cnt[i][threadIdx.x];
      uint32_t total;
      uint32_t ex = block_excl_scan(tot, ws, total);
      digit_start[threadIdx.x] = ex;
      uint32_t run = ex;
#pragma unroll
      for (int i = 0; i < kSortWarps; i++) {
        uint32_t c = cnt[i][threadIdx.x];
        cnt[i][threadIdx.x] = run;
        run += c;
      }
    }
    __syncthreads();
    // place into shared memory in digit order
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      uint32_t p = wbase + i * 32 + lane;
      if (p < tile_n) {
        uint32_t d = (k[i] >> shift) & (kRadix - 1);
        uint32_t pos = cnt[kStable ? w : 0][d] + rank[i];
        s_keys[pos] = k[i];
        s_vals[pos] = v[i];
      }
    }
    __syncthreads();
    // coalesced copy-out of the digit runs
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
      uint32_t p = (uint32_t)i * kSortThreads + threadIdx.x;
      if (p < tile_n) {
        uint32_t key = s_keys[p];
        uint32_t d = (key >> shift) & (kRadix - 1);
        uint64_t g = (uint64_t)gbase[d] + (p - digit_start[d]);
        keys_out[g] = key;
        vals_out[g] = s_vals[p];
      }
    }
    __syncthreads();  // shared arrays are reused by the next tile
  }
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// MSD variant for the partitioned MSM sort: one 256-way pass inside each partition, then every group
// (partition, digit) - a few thousand entries that share all key bits above the last r - is sorted by ONE block
// in shared memory and written back in place, fully coalesced: no histogram, no scan and no scattered run
// copies for the last digit.
// ---------------------------------------------------------------------------
constexpr int kGroupThreads = 512;
constexpr int kGroupItems = 26;
constexpr int kGroupCap = kGroupThreads * kGroupItems;  // 13312 entries = 104 KB of shared memory, two blocks per SM
static int g_group_cap = kGroupCap;                     // lowered by a test hook to reach the oversize path
void sort_set_group_cap(int cap) { g_group_cap = cap > 0 && cap < kGroupCap ? cap : kGroupCap; }
int sort_group_cap() { return g_group_cap; }

// One block: groups[G + 1] = start of every group after the 256-way pass (read off that pass's scanned
// histogram: the base of (partition, digit, first tile)), and the tile table of the OVERSIZE groups (more than
// `cap` entries - skewed scalars), laid out like a partition table [group_base[G + 1] | tile_start[G + 1]] so that
// the generic histogram / scan / scatter kernels can sort exactly those; groups of the last partition (the sentinel
// keys of zero digits, all equal) need no sorting and get no tiles.
__global__ void __launch_bounds__(1024) sort_group_table(const uint32_t* __restrict__ parts, int P,
                                                         const uint32_t* __restrict__ hist_scanned, uint32_t cap,
                                                         uint32_t* __restrict__ parts2) {
  __shared__ uint32_t ws[33];
  const int G = P * kRadix;
  uint32_t* gstart = parts2;
  uint32_t* tile_start2 = parts2 + (G + 1);
  const uint32_t* part_base = parts;
  const uint32_t* tile_start = parts + (P + 1);
  for (int g = threadIdx.x; g <= G; g += blockDim.x) {
    uint32_t v;
    if (g == G) {
      v = part_base[P];
    } else {
      const int p = g >> 8, d = g & 255;
      const uint32_t t0 = tile_start[p], stride = tile_start[p + 1] - t0;
      v = stride ? hist_scanned[(size_t)kRadix * t0 + (size_t)d * stride] : part_base[p];
    }
    gstart[g] = v;
  }
  __syncthreads();
  // exclusive scan of the oversize groups' tile counts: thread t owns groups [t * per, (t + 1) * per)
  const int per = (G + (int)blockDim.x - 1) / (int)blockDim.x;
  const int g0 = (int)threadIdx.x * per, g1 = g0 + per < G ? g0 + per : G;
  const int g_sent = (P - 1) * kRadix;
  uint32_t local = 0;
  for (int g = g0; g < g1; g++) {
    const uint32_t m = gstart[g + 1] - gstart[g];
    if (m > cap && g < g_sent) local += (m + kSortTile - 1) / kSortTile;
  }
  uint32_t total;
  uint32_t run = block_excl_scan(local, ws, total);
  for (int g = g0; g < g1; g++) {
    tile_start2[g] = run;
    const uint32_t m = gstart[g + 1] - gstart[g];
    if (m > cap && g < g_sent) run += (m + kSortTile - 1) / kSortTile;
  }
  if (threadIdx.x == 0) tile_start2[G] = total;
}

// Block per group (grid-stride).  All keys of a group agree above their last r bits (partition + 256-way pass), so
// after the counting step a thread keeps only the r-bit digits of its keys, four to a register; that frees the
// registers to have ALL its value loads in flight during the scan.  Digit counts and placement use shared-memory
// atomics; the sorted group leaves through shared memory in one coalesced sweep.
__global__ void __launch_bounds__(kGroupThreads, 2)
    sort_group_local(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     const uint32_t* __restrict__ gstart, int G_sort, uint32_t mask, uint32_t cap) {
  extern __shared__ uint32_t s_grp[];  // keys[kGroupCap] | vals[kGroupCap]
  __shared__ uint32_t cnt[kRadix];
  __shared__ uint32_t ws[33];
  __shared__ uint32_t s_common;
  uint32_t* s_keys = s_grp;
  uint32_t* s_vals = s_grp + kGroupCap;
  constexpr int kPacked = (kGroupItems + 3) / 4;
  for (int g = blockIdx.x; g < G_sort; g += gridDim.x) {
    const uint32_t gs = gstart[g], m = gstart[g + 1] - gs;
    if (m == 0 || m > cap) continue;  // uniform per block
    uint32_t k[kGroupItems];
    if (threadIdx.x < kRadix) cnt[threadIdx.x] = 0;
#pragma unroll
    for (int i = 0; i < kGroupItems; i++) {
      const uint32_t p = (uint32_t)i * kGroupThreads + threadIdx.x;
      k[i] = p < m ? keys_in[gs + p] : 0u;
    }
    if (threadIdx.x == 0) s_common = k[0] & ~mask;
    __syncthreads();
    uint32_t dg[kPacked];
#pragma unroll
    for (int j = 0; j < kPacked; j++) dg[j] = 0;
#pragma unroll
    for (int i = 0; i < kGroupItems; i++) {
      const uint32_t p = (uint32_t)i * kGroupThreads + threadIdx.x;
      const uint32_t d = k[i] & mask;
      dg[i >> 2] |= d << (8 * (i & 3));
      if (p < m) atomicAdd(&cnt[d], 1u);
    }
    // the values: every load issued before the scan's barriers
    uint32_t v[kGroupItems];
#pragma unroll
    for (int i = 0; i < kGroupItems; i++) {
      const uint32_t p = (uint32_t)i * kGroupThreads + threadIdx.x;
      v[i] = p < m ? vals_in[gs + p] : 0u;
    }
    __syncthreads();
    {
      const uint32_t c = threadIdx.x < kRadix ? cnt[threadIdx.x] : 0u;
      uint32_t total;
      const uint32_t ex = block_excl_scan(c, ws, total);  // ends with a barrier
      if (threadIdx.x < kRadix) cnt[threadIdx.x] = ex;    // running cursor of the digit
    }
    __syncthreads();
    const uint32_t common = s_common;
#pragma unroll
    for (int i = 0; i < kGroupItems; i++) {
      const uint32_t p = (uint32_t)i * kGroupThreads + threadIdx.x;
      if (p < m) {
        const uint32_t d = (dg[i >> 2] >> (8 * (i & 3))) & 255u;
        const uint32_t pos = atomicAdd(&cnt[d], 1u);
        s_keys[pos] = common | d;
        s_vals[pos] = v[i];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kGroupItems; i++) {
      const uint32_t p = (uint32_t)i * kGroupThreads + threadIdx.x;
      if (p < m) {
        keys_out[gs + p] = s_keys[p];
        vals_out[gs + p] = s_vals[p];
      }
    }
    __syncthreads();  // shared arrays are reused by the next group
  }
}

// the last partition holds only sentinel keys (zero digits): nothing to sort, and the accumulate never reads
// their values - the keys alone must be in the output buffer
__global__ void sort_copy_sentinel_keys(const uint32_t* __restrict__ parts, int P, uint32_t* __restrict__ keys_out) {
  const uint32_t lo = parts[P - 1], hi = parts[P];
  for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x)
    keys_out[i] = 0xffffffffu;
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
// one 8-bit pass: histogram, exclusive scan, multi-split.  parts / P: partition table or nullptr; d_tiles: device
// word holding the real tile count (nullptr: ntiles is exact or an upper bound whose excess is harmless)
static int sort_pass(myzkp_ctx* ctx, const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint64_t n, int shift,
                     uint32_t ntiles, uint32_t* hist, uint32_t* sums, const uint32_t* d_parts, int P, bool stable,
                     const uint32_t* d_tiles) {
  const size_t hist_len = (size_t)ntiles * kRadix;
  const size_t nchunks = (hist_len + kScanChunk - 1) / kScanChunk;
  // tiles past the last real one (partition mode) leave their slots untouched: those lie behind every
  // used slot in scan order, so whatever they hold cannot reach a used prefix
  const uint32_t hblocks = ntiles < (uint32_t)ctx->sm_count * 64 ? ntiles : (uint32_t)ctx->sm_count * 64;
  sort_tile_hist<<<hblocks, kSortThreads, 0, ctx->stream>>>(ki, n, shift, hist, ntiles, d_parts, P);
  MZ_LAUNCH_CHECK(ctx);
  scan_chunk_sums<<<(unsigned)nchunks, kScanThreads, 0, ctx->stream>>>(hist, hist_len, sums, d_tiles);
  MZ_LAUNCH_CHECK(ctx);
  scan_sums_inplace<<<1, kScanThreads, 0, ctx->stream>>>(sums, nchunks);
  MZ_LAUNCH_CHECK(ctx);
  scan_apply<<<(unsigned)nchunks, kScanThreads, 0, ctx->stream>>>(hist, hist_len, sums, d_tiles);
  MZ_LAUNCH_CHECK(ctx);
  static const int env_occ = getenv("MZ_SORT_OCC") ? atoi(getenv("MZ_SORT_OCC")) : 0;  // experiment knob
  // unordered passes: 3 blocks per SM without the prefetch (80 registers) beat 2 with it (116): 1.29 -> 1.0 ms at 2^24
  const int occ = stable ? 2 : (env_occ >= 2 && env_occ <= 4) ? env_occ : 3;
  const uint32_t sblocks = ntiles < (uint32_t)(ctx->sm_count * occ) ? ntiles : (uint32_t)(ctx->sm_count * occ);
  if (stable)
    sort_tile_scatter<true, 2><<<sblocks, kSortThreads, 0, ctx->stream>>>(ki, vi, n, shift, hist, ntiles, ko, vo, d_parts, P);
  else if (occ == 4)
    sort_tile_scatter<false, 4><<<sblocks, kSortThreads, 0, ctx->stream>>>(ki, vi, n, shift, hist, ntiles, ko, vo, d_parts, P);
  else if (occ == 3)
    sort_tile_scatter<false, 3><<<sblocks, kSortThreads, 0, ctx->stream>>>(ki, vi, n, shift, hist, ntiles, ko, vo, d_parts, P);
  else
    sort_tile_scatter<false, 2><<<sblocks, kSortThreads, 0, ctx->stream>>>(ki, vi, n, shift, hist, ntiles, ko, vo, d_parts, P);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

// Sorts n (key, val) pairs by the low `bits` key bits.  The result lands in
// (*out_keys, *out_vals), which alias either the a- or the b-buffers.
// first_pass_unordered: the caller does not care about the order of equal keys (the MSM does not), so the
// first pass may be unstable; the later passes are always stable, as LSD radix sort needs.
// d_parts / P (optional): the list is already split into P partitions by its high key bits (layout in the
// comment above locate_tile); each partition is then sorted on its own by the low `bits` bits.
int radix_sort_pairs(myzkp_ctx* ctx, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint64_t n,
                     int bits, uint32_t** out_keys, uint32_t** out_vals, const uint32_t* d_parts, int P,
                     bool first_pass_unordered) {
  *out_keys = keys_a;
  *out_vals = vals_a;
  if (n == 0 || bits <= 0) return MYZKP_OK;
  if (n >= (1ull << 32)) return fail(ctx, MYZKP_ERR_INVALID_ARG, "too many entries to sort");
  // with partitions every partition may end in a partial tile: an upper bound on the tile count
  const uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile) + (d_parts ? (uint32_t)P : 0u);
  const size_t hist_len = (size_t)ntiles * kRadix;
  const size_t nchunks = (hist_len + kScanChunk - 1) / kScanChunk;
  MZ_CUDA_TRY(ctx, ctx->sort_tmp.ensure((hist_len + nchunks + 64) * sizeof(uint32_t)));
  uint32_t* hist = ctx->sort_tmp.as<uint32_t>();
  uint32_t* sums = hist + hist_len;
  uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  for (int shift = 0; shift < bits; shift += 8) {
    MZ_TRY(sort_pass(ctx, ki, vi, ko, vo, n, shift, ntiles, hist, sums, d_parts, P, !(shift == 0 && first_pass_unordered),
                     nullptr));
    uint32_t* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  *out_keys = ki;
  *out_vals = vi;
  return MYZKP_OK;
}

// MSD form for a partitioned list whose partitions share all key bits above `low_bits` = 8 + r: a 256-way pass on
// bits [r, r + 8) inside each partition (equal keys may land in any order: bucket sums commute), then the group-local
// sort on the last r bits.  Oversize groups go through one more generic pass restricted to them.  The order of
// equal keys is unspecified; the result is in (*out_keys, *out_vals) = the a-buffers.
int radix_sort_pairs_msd(myzkp_ctx* ctx, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint64_t n,
                         int r, uint32_t** out_keys, uint32_t** out_vals, const uint32_t* d_parts, int P) {
  *out_keys = keys_a;
  *out_vals = vals_a;
  if (n == 0) return MYZKP_OK;
  if (n >= (1ull << 32) || !d_parts || P < 2 || r < 1 || r > 8) return fail(ctx, MYZKP_ERR_INVALID_ARG, "msd sort: bad plan");
  const uint32_t cap = (uint32_t)g_group_cap;
  const int G = P * kRadix;
  const uint32_t tiles_a = (uint32_t)((n + kSortTile - 1) / kSortTile) + (uint32_t)P;
  const uint32_t tiles_b = (uint32_t)((n + kSortTile - 1) / kSortTile) + (uint32_t)(n / cap) + 2;  // oversize groups only
  const uint32_t tiles_max = tiles_a > tiles_b ? tiles_a : tiles_b;
  const size_t hist_len = (size_t)tiles_max * kRadix;
  const size_t nchunks = (hist_len + kScanChunk - 1) / kScanChunk;
  MZ_CUDA_TRY(ctx, ctx->sort_tmp.ensure((hist_len + nchunks + 64) * sizeof(uint32_t)));
  MZ_CUDA_TRY(ctx, ctx->sort_groups.ensure((size_t)(2 * (G + 1) + 8) * sizeof(uint32_t)));
  uint32_t* hist = ctx->sort_tmp.as<uint32_t>();
  uint32_t* sums = hist + hist_len;
  uint32_t* parts2 = ctx->sort_groups.as<uint32_t>();
  // 1. 256-way pass inside the partitions: a -> b
  MZ_TRY(sort_pass(ctx, keys_a, vals_a, keys_b, vals_b, n, r, tiles_a, hist, sums, d_parts, P, /*stable=*/false, nullptr));
  // 2. group starts (+ the tile table of oversize groups)
  sort_group_table<<<1, 1024, 0, ctx->stream>>>(d_parts, P, hist, cap, parts2);
  MZ_LAUNCH_CHECK(ctx);
  // 3. group-local sort of everything that fits: b -> a
  static bool attr_set[64] = {};
  constexpr int smem = 2 * kGroupCap * (int)sizeof(uint32_t);
  if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
    MZ_CUDA_TRY(ctx, cudaFuncSetAttribute(sort_group_local, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[ctx->device] = true;
  }
  const int G_sort = (P - 1) * kRadix;  // the last partition is the sentinel's
  const unsigned gblocks = (unsigned)(G_sort < ctx->sm_count * 8 ? G_sort : ctx->sm_count * 8);
  sort_group_local<<<gblocks, kGroupThreads, smem, ctx->stream>>>(keys_b, vals_b, keys_a, vals_a, parts2, G_sort,
                                                                  (1u << r) - 1u, cap);
  MZ_LAUNCH_CHECK(ctx);
  sort_copy_sentinel_keys<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(d_parts, P, keys_a);
  MZ_LAUNCH_CHECK(ctx);
  // 4. oversize groups (none for uniform scalars: these launches then find zero tiles and return at once)
  MZ_TRY(sort_pass(ctx, keys_b, vals_b, keys_a, vals_a, n, 0, tiles_b, hist, sums, parts2, G, /*stable=*/false,
                   parts2 + (2 * G + 1)));
  return MYZKP_OK;
}

}  // namespace mz
