// Pippenger bucket MSM over the resident SRS table - the GPU form of
// Polynomial::eval_with_powers_on_curve (polynomial.rs:156-165), which in the
// reference is N sequential double-and-add scalar multiplications
// (curve.rs:163-191) of sanitized coefficients (field.rs:260-270).
//
// Pipeline (all on the ctx stream, no host round trips):
//   1. msm_recode      scalar -> W signed c-bit digits; one (bucket key, table
//                      index | sign) entry per non-zero digit.  Because the table
//                      holds 2^(c w) P_i for every window w (rows at the bit offsets
//                      of ctx->row_bits), every window shares ONE set of 2^(c-1)
//                      buckets and no doublings are ever needed.
//   2. radix sort      entries by bucket key (sort.cu).  Small windows: LSD passes of 8 bits.  Windows 20 / 22 / 24:
//                      the recode leaves the entries partitioned by their high key bits, one 256-way pass and a
//                      group-local shared-memory sort finish the job (radix_sort_pairs_msd).
//   3. msm_accumulate  fixed-length segments of the sorted entry list, one
//                      thread each, XYZZ mixed adds; load-balanced for any
//                      scalar distribution (a heavy bucket just spans segments).
//   4. msm_merge_heads segments that start inside a bucket hand their first
//                      partial to the bucket's owner.
//   5. msm_bucket_chunks + msm_bucket_reduce + xyzz_tree_reduce   sum_k k * B_k by chunked
//                      running sums in two levels, then a tree sum.
#include <stdlib.h>

#include <vector>

#include <cmath>
#include "ctx.cuh"
#include "inv.cuh"

namespace mz {

// ---------------------------------------------------------------------------
// small load/store helpers (16-byte vector accesses)
// ---------------------------------------------------------------------------
__device__ __forceinline__ Affine load_affine(const Affine* p) {
  Affine r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
  r.x.v[0] = a.x; r.x.v[1] = a.y; r.x.v[2] = a.z; r.x.v[3] = a.w;
  r.x.v[4] = b.x; r.x.v[5] = b.y; r.x.v[6] = b.z; r.x.v[7] = b.w;
  r.y.v[0] = c.x; r.y.v[1] = c.y; r.y.v[2] = c.z; r.y.v[3] = c.w;
  r.y.v[4] = d.x; r.y.v[5] = d.y; r.y.v[6] = d.z; r.y.v[7] = d.w;
  return r;
}
// the same with an L2 fetch-size hint of 64 bytes (experiment: a 64-byte gather pulls a 128-byte line from HBM)
__device__ __forceinline__ uint4 ldg_l2_64(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ Affine load_affine_64(const Affine* p) {
  Affine r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = ldg_l2_64(q), b = ldg_l2_64(q + 1), c = ldg_l2_64(q + 2), d = ldg_l2_64(q + 3);
  r.x.v[0] = a.x; r.x.v[1] = a.y; r.x.v[2] = a.z; r.x.v[3] = a.w;
  r.x.v[4] = b.x; r.x.v[5] = b.y; r.x.v[6] = b.z; r.x.v[7] = b.w;
  r.y.v[0] = c.x; r.y.v[1] = c.y; r.y.v[2] = c.z; r.y.v[3] = c.w;
  r.y.v[4] = d.x; r.y.v[5] = d.y; r.y.v[6] = d.z; r.y.v[7] = d.w;
  return r;
}
__device__ __forceinline__ void store_xyzz(XYZZ* p, const XYZZ& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
}
__device__ __forceinline__ XYZZ load_xyzz(const XYZZ* p) {
  XYZZ v;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint32_t* s = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint4 a = q[i];
    s[4 * i] = a.x; s[4 * i + 1] = a.y; s[4 * i + 2] = a.z; s[4 * i + 3] = a.w;
  }
  return v;
}

// ---------------------------------------------------------------------------
// 1. signed-digit recode
// ---------------------------------------------------------------------------
// digits d_w in (-2^(c-1), 2^(c-1)], sum_w d_w 2^(c w) = scalar.  W*c >= 255
// guarantees the top window absorbs the last carry for any scalar < 2^254.
// key = |d| - 1 in [0, 2^(c-1)), or `sentinel` = 2^(c-1) for d == 0 (sorted to
// the end and ignored).  val = sign << 31 | (row_of_bit[c w] * srs_n + srs_off + i).
// One launch recodes a batch of polynomials (blockIdx.y; a single MSM is a batch of one): polynomial
// y has its own entry region [entry_off, entry_off + W n) and its own bucket range starting at
// key_base; zero digits of every polynomial share the batch-wide sentinel key.
struct RecodeDesc {
  const uint32_t* scalars;
  uint64_t n;
  uint64_t entry_off;
  uint32_t key_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(256) msm_recode(RecodeDesc single, const RecodeDesc* __restrict__ descs, int c, int W,
                                                  const uint8_t* __restrict__ row_of_bit, uint32_t srs_n,
                                                  uint32_t srs_off, uint32_t sentinel_key, uint32_t* __restrict__ keys_all,
                                                  uint32_t* __restrict__ vals_all, int* __restrict__ flag,
                                                  uint32_t win_stride) {
  const RecodeDesc dsc = descs ? descs[blockIdx.y] : single;  // a batch of one needs no descriptor upload
  const uint32_t* __restrict__ scalars = dsc.scalars;
  const size_t n = dsc.n;
  uint32_t* __restrict__ keys = keys_all + dsc.entry_off;
  uint32_t* __restrict__ vals = vals_all + dsc.entry_off;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr sc;
  {
    const uint4* q = reinterpret_cast<const uint4*>(scalars + i * 8);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    sc.v[0] = a.x; sc.v[1] = a.y; sc.v[2] = a.z; sc.v[3] = a.w;
    sc.v[4] = b.x; sc.v[5] = b.y; sc.v[6] = b.z; sc.v[7] = b.w;
  }
  if (!fe_is_canonical(sc)) atomicOr(flag, 1);
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (c == 32) ? 0xffffffffu : ((1u << c) - 1u);
  uint32_t carry = 0;
  for (int w = 0; w < W; w++) {
    int bit = w * c;
    int limb = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (limb < 8) {
      uint64_t two = sc.v[limb];
      if (limb + 1 < 8) two |= (uint64_t)sc.v[limb + 1] << 32;
      raw = (uint32_t)(two >> sh) & mask;
    }
    uint32_t d = raw + carry;
    uint32_t neg = d > half ? 1u : 0u;
    uint32_t mag = neg ? ((1u << c) - d) : d;
    carry = neg;
    // win_stride != 0: plain points without a table of multiples - every window owns its bucket range and
    // reads row 0; the window weights 2^(c w) are applied after the bucket reduce (msm_window_combine)
    uint32_t key = mag ? dsc.key_base + (uint32_t)w * win_stride + mag - 1 : sentinel_key;
    uint32_t idx = (win_stride ? 0u : (uint32_t)row_of_bit[bit] * srs_n) + srs_off + (uint32_t)i;
    keys[(size_t)w * n + i] = key;
    vals[(size_t)w * n + i] = idx | (neg << 31);
  }
}

// ---------------------------------------------------------------------------
// 1b. recode with an MSD split: entries leave the recode already partitioned by their high key bits
// ---------------------------------------------------------------------------
// A 22-bit bucket key costs three 8-bit radix passes over 1.6 GB of entries.  The recode kernel produces
// the entries on chip anyway, so it can do the first (most significant) split for free: partition =
// key >> low_bits (the last partition holds the zero digits), and only low_bits remain for the sort, which then
// runs inside each partition (sort.cu, locate_tile).  One polynomial: low_bits = 8 + r for the MSD sort (r = 7 and
// 65 partitions at 2^24 points, c = 22); batches: low_bits = 16 (33 partitions for c = 22) and two LSD passes.
//   msm_recode_count    per-partition entry counts (shared-memory counters, one global add per block and bin)
//   msm_partition_plan  partition bases, tile bases, zeroed cursors (one block)
//   msm_recode_scatter  recodes (digits stay in registers), groups the block's entries by partition in shared
//                       memory, reserves room in every partition with one atomic add each, copies the runs out
// Order inside a partition is arbitrary (bucket sums commute).
constexpr int kMaxParts = 260;
// device layout of ctx->sort_parts (uint32): part_base[P+1] | tile_start[P+1] | counts[P] | cursor[P]
constexpr int kSortTileEntries = 4096;  // = kSortTile of sort.cu

// row index of the table row holding 2^(c w) P for window w (kernel parameter: lives in the constant bank)
struct RecodeRows {
  uint8_t row[64];
};

// digits of one scalar, least significant window first: f(w, key, sign).  C is a compile-time window so
// the loop unrolls and every shift is an immediate (the run-time version costs ~60 instructions per digit).
template <int C, class F>
__device__ __forceinline__ void recode_digits(const Fr& sc, uint32_t key_base, uint32_t sentinel_key, F f) {
  constexpr int W = (255 + C - 1) / C;
  constexpr uint32_t half = 1u << (C - 1);
  constexpr uint32_t mask = (1u << C) - 1u;
  uint32_t carry = 0;
#pragma unroll
  for (int w = 0; w < W; w++) {
    const int bit = w * C;
    const int limb = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (limb < 8) {
      if (sh + C <= 32 || limb + 1 >= 8) raw = (sc.v[limb] >> sh) & mask;
      else raw = __funnelshift_r(sc.v[limb], sc.v[limb + 1], sh) & mask;
    }
    const uint32_t d = raw + carry;
    const uint32_t neg = d > half ? 1u : 0u;
    const uint32_t mag = neg ? ((1u << C) - d) : d;
    carry = neg;
    f(w, mag ? key_base + mag - 1 : sentinel_key, neg);
  }
}
__device__ __forceinline__ Fr load_scalar(const uint32_t* scalars, size_t i) {
  Fr sc;
  const uint4* q = reinterpret_cast<const uint4*>(scalars + i * 8);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  sc.v[0] = a.x; sc.v[1] = a.y; sc.v[2] = a.z; sc.v[3] = a.w;
  sc.v[4] = b.x; sc.v[5] = b.y; sc.v[6] = b.z; sc.v[7] = b.w;
  return sc;
}

struct RecodeCtl {
  int low_bits, P;
  uint32_t srs_n, srs_off, sentinel_key;
};
constexpr int kPartScalars = 512;  // scalars per block of the partitioned recode (2 per thread)

// grid (blocks of kPartScalars scalars over the longest polynomial, K)
template <int C>
__global__ void __launch_bounds__(256) msm_recode_count(RecodeDesc single, const RecodeDesc* __restrict__ descs,
                                                        RecodeCtl ctl, uint32_t* __restrict__ counts,
                                                        int* __restrict__ flag) {
  __shared__ uint32_t h[kMaxParts];
  const RecodeDesc dsc = descs ? descs[blockIdx.y] : single;
  for (int p = threadIdx.x; p < ctl.P; p += blockDim.x) h[p] = 0;
  __syncthreads();
  const size_t lo = (size_t)blockIdx.x * kPartScalars;
  const size_t hi = lo + kPartScalars < dsc.n ? lo + kPartScalars : dsc.n;
  bool bad = false;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const Fr sc = load_scalar(dsc.scalars, i);
    bad |= !fe_is_canonical(sc);
    recode_digits<C>(sc, dsc.key_base, ctl.sentinel_key,
                     [&](int, uint32_t key, uint32_t) { atomicAdd(&h[key >> ctl.low_bits], 1u); });
  }
  if (bad) atomicOr(flag, 1);
  __syncthreads();
  for (int p = threadIdx.x; p < ctl.P; p += blockDim.x)
    if (h[p]) atomicAdd(&counts[p], h[p]);
}

// one block: part_base / tile_start from the counts; cursors cleared
__global__ void msm_partition_plan(uint32_t* __restrict__ parts, int P) {
  uint32_t* part_base = parts;
  uint32_t* tile_start = parts + (P + 1);
  const uint32_t* counts = parts + 2 * (P + 1);
  uint32_t* cursor = parts + 2 * (P + 1) + P;
  if (threadIdx.x == 0) {
    uint32_t b = 0, t = 0;
    for (int p = 0; p < P; p++) {
      part_base[p] = b;
      tile_start[p] = t;
      b += counts[p];
      t += (counts[p] + kSortTileEntries - 1) / kSortTileEntries;
    }
    part_base[P] = b;
    tile_start[P] = t;
  }
  for (int p = threadIdx.x; p < P; p += blockDim.x) cursor[p] = 0;
}

// Every thread recodes its two scalars ONCE into registers; the block then groups the entries by partition
// in shared memory and copies each partition's run to the room it reserved with one atomic add.
template <int C>
__global__ void __launch_bounds__(256) msm_recode_scatter(RecodeDesc single, const RecodeDesc* __restrict__ descs,
                                                          RecodeCtl ctl, RecodeRows rows, uint32_t* __restrict__ parts,
                                                          uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  constexpr int W = (255 + C - 1) / C;
  constexpr int kPer = kPartScalars / 256;
  extern __shared__ uint32_t stage[];  // keys[kPartScalars * W] | vals[kPartScalars * W]
  __shared__ uint32_t h[kMaxParts];      // counts, then running local cursors
  __shared__ uint32_t loc_off[kMaxParts + 1];
  __shared__ uint32_t gbase[kMaxParts];
  __shared__ uint32_t ws[33];
  uint32_t* s_keys = stage;
  uint32_t* s_vals = stage + kPartScalars * W;
  const RecodeDesc dsc = descs ? descs[blockIdx.y] : single;
  const int P = ctl.P;
  const uint32_t* part_base = parts;
  uint32_t* cursor = parts + 2 * (P + 1) + P;
  for (int p = threadIdx.x; p < kMaxParts; p += blockDim.x) h[p] = 0;
  __syncthreads();
  const size_t lo = (size_t)blockIdx.x * kPartScalars;
  if (lo >= dsc.n) return;
  uint32_t keys_r[kPer][W], vals_r[kPer][W];
#pragma unroll
  for (int s = 0; s < kPer; s++) {
    const size_t i = lo + (size_t)s * 256 + threadIdx.x;
    if (i < dsc.n) {
      const Fr sc = load_scalar(dsc.scalars, i);
      recode_digits<C>(sc, dsc.key_base, ctl.sentinel_key, [&](int w, uint32_t key, uint32_t neg) {
        keys_r[s][w] = key;
        vals_r[s][w] = ((uint32_t)rows.row[w] * ctl.srs_n + ctl.srs_off + (uint32_t)i) | (neg << 31);
        atomicAdd(&h[key >> ctl.low_bits], 1u);
      });
    } else {
#pragma unroll
      for (int w = 0; w < W; w++) keys_r[s][w] = 0xffffffffu;  // no entry
    }
  }
  __syncthreads();
  // exclusive scan of the partition counts (P <= kMaxParts <= 2 x 256: two bins per thread)
  {
    const int p0 = 2 * threadIdx.x, p1 = p0 + 1;
    const uint32_t c0 = p0 < P ? h[p0] : 0u, c1 = p1 < P ? h[p1] : 0u;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    uint32_t incl = c0 + c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) ws[wrp] = incl;
    __syncthreads();
    if (wrp == 0) {
      uint32_t v = lane < 8 ? ws[lane] : 0u, sc = v;
#pragma unroll
      for (int d = 1; d < 8; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, sc, d);
        if (lane >= d) sc += o;
      }
      if (lane < 8) ws[lane] = sc - v;
      if (lane == 7) ws[32] = sc;
    }
    __syncthreads();
    const uint32_t ex = ws[wrp] + incl - (c0 + c1);
    if (p0 < P) loc_off[p0] = ex;
    if (p1 < P) loc_off[p1] = ex + c0;
    if (threadIdx.x == 0) loc_off[P] = ws[32];
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const uint32_t cnt = h[p];
    gbase[p] = cnt ? part_base[p] + atomicAdd(&cursor[p], cnt) : 0u;
    h[p] = loc_off[p];  // becomes the running local cursor
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kPer; s++) {
#pragma unroll
    for (int w = 0; w < W; w++) {
      const uint32_t key = keys_r[s][w];
      if (key != 0xffffffffu) {
        const uint32_t pos = atomicAdd(&h[key >> ctl.low_bits], 1u);
        s_keys[pos] = key;
        s_vals[pos] = vals_r[s][w];
      }
    }
  }
  __syncthreads();
  const uint32_t total = loc_off[P];
  for (uint32_t pos = threadIdx.x; pos < total; pos += blockDim.x) {
    const uint32_t key = s_keys[pos];
    const uint32_t p = key >> ctl.low_bits;
    const uint32_t g = gbase[p] + (pos - loc_off[p]);
    keys_out[g] = key;
    vals_out[g] = s_vals[pos];
  }
}

template <int C>
static int launch_partitioned_recode(myzkp_ctx* ctx, const RecodeDesc& single, const RecodeDesc* d_descs, size_t K, size_t n,
                                     const RecodeCtl& ctl, const RecodeRows& rows, uint32_t* parts, uint32_t* counts,
                                     uint32_t* keys_a, uint32_t* vals_a, int* flag) {
  constexpr int W = (255 + C - 1) / C;
  constexpr int smem = 2 * kPartScalars * W * (int)sizeof(uint32_t);
  const unsigned gx = (unsigned)((n + kPartScalars - 1) / kPartScalars);
  msm_recode_count<C><<<dim3(gx, (unsigned)K), 256, 0, ctx->stream>>>(single, d_descs, ctl, counts, flag);
  MZ_LAUNCH_CHECK(ctx);
  msm_partition_plan<<<1, 256, 0, ctx->stream>>>(parts, ctl.P);
  MZ_LAUNCH_CHECK(ctx);
  static bool attr_set[64] = {};
  if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
    MZ_CUDA_TRY(ctx, cudaFuncSetAttribute(msm_recode_scatter<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[ctx->device] = true;
  }
  msm_recode_scatter<C><<<dim3(gx, (unsigned)K), 256, smem, ctx->stream>>>(single, d_descs, ctl, rows, parts, keys_a, vals_a);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

// ---------------------------------------------------------------------------
// 3. segment accumulate
// ---------------------------------------------------------------------------
// Thread t owns sorted entries [t*L, (t+1)*L).  Its first run (bucket of its
// first entry) may continue a run of the previous segment, so it always goes to
// heads[t]; every later run starts inside the segment, making this thread the
// unique first writer of that bucket.  Buckets are pre-zeroed (= infinity).
constexpr int kAccThreads = 128;

// kOnto: the buckets already hold the sums of earlier scalar chunks (upload pipeline): a run's unique writer
// starts from the bucket's value instead of infinity, so no second bucket set and no dense bucket addition
// are needed; runs that go to `heads` are folded in by the merge kernels, which read-modify-write anyway.
template <bool kHint64, bool kOnto>
__global__ void __launch_bounds__(kAccThreads, 4)
    msm_accumulate(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t M, uint32_t L,
                   uint32_t sentinel, const Affine* __restrict__ tbl, XYZZ* __restrict__ buckets,
                   XYZZ* __restrict__ heads, uint32_t* __restrict__ head_keys, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint64_t s = t * L;
  uint64_t e = s + L < M ? s + L : M;
  uint32_t cur = keys[s];
  if (cur >= sentinel) {
    head_keys[t] = sentinel;
    return;
  }
  head_keys[t] = cur;
  XYZZ acc = xyzz_inf();
  bool first_run = true;
  // software pipeline: the next entry's point is in flight while this one is added
  uint32_t v = vals[s];
  Affine pt = kHint64 ? load_affine_64(tbl + (v & 0x7fffffffu)) : load_affine(tbl + (v & 0x7fffffffu));
  for (uint64_t i = s; i < e; i++) {
    uint32_t k_next = sentinel, v_next = 0;
    Affine pt_next;
    pt_next.x = Fq::zero(); pt_next.y = Fq::zero();
    if (i + 1 < e) {
      k_next = keys[i + 1];
      if (k_next < sentinel) {
        v_next = vals[i + 1];
        pt_next = kHint64 ? load_affine_64(tbl + (v_next & 0x7fffffffu)) : load_affine(tbl + (v_next & 0x7fffffffu));
      }
    }
    if (v >> 31) pt.y = fe_neg(pt.y);
    xyzz_madd(acc, pt);
    if (k_next != cur) {
      if (first_run) store_xyzz(heads + t, acc);
      else store_xyzz(buckets + cur, acc);
      first_run = false;
      cur = k_next;
      if (k_next >= sentinel) break;
      acc = kOnto ? load_xyzz(buckets + k_next) : xyzz_inf();
    }
    v = v_next;
    pt = pt_next;
  }
}

// ---------------------------------------------------------------------------
// 3b. segment accumulate with batched-affine pair sums (fused, one kernel)
// ---------------------------------------------------------------------------
// Same contract as msm_accumulate (segments of the sorted entry list, heads / buckets), but two
// entries of the same bucket are first added in AFFINE coordinates and only the pair sum goes through
// the XYZZ mixed addition:  per two entries  (5M + 1S) + (8M + 2S)  instead of  2 x (8M + 2S).
// The affine addition needs 1 / (x_b - x_a); the thread shares ONE inversion among the kBaaBatch pairs
// of a batch (Montgomery's trick, lane-private, so no synchronisation):
//   phase 1  walk the batch backwards: suf[j] = prod_{i > j} dx_i (local memory), S = prod of all dx
//   invert   I = 1 / S with the branch-free safegcd inverse (inv.cuh; integer-add pipe, no divergence)
//   phase 2  walk forwards: 1 / dx_j = I * suf[j], I *= dx_j; lambda, x3, y3; xyzz_madd(acc, pair sum)
// A pair that cannot be added in affine form - different buckets, an operand at infinity, equal x
// (P + P or P - P), the sentinel - contributes dx = 1 to the product and its entries go through the
// general mixed addition one by one, so every case of curve.rs:131-145 keeps its meaning.  In the
// common split case (a ends a run, b starts the next) b simply becomes the new accumulator.
// Points are gathered twice (x in phase 1, x and y in phase 2); prefetch.global.L2 a few pairs ahead
// keeps both walks off the DRAM latency.
// The loop is far larger than the 32 KB instruction cache when every multiply is expanded in place (ncu of
// the first version: "no instruction" was the top stall, the multiply pipe 58 % busy), so this kernel's
// multiplies are out-of-line calls - one copy each, arguments and result in registers.
struct CallOps {
  static __device__ __noinline__ Fq mul(Fq a, Fq b) { return fe_mul(a, b); }
  static __device__ __noinline__ Fq sqr(Fq a) { return fe_sqr(a); }
  static __device__ __noinline__ Fq mul2(Fq a, Fq b, Fq c, Fq d) { return fe_mul2(a, b, c, d); }
};
constexpr int kBaaBatch = 64;
constexpr int kBaaPrefetch1 = 4;
constexpr int kBaaAutoMinL = 1 << 30;  // automatic selection threshold on the segment length (off until measured)  // pairs ahead, phase 1 (1 multiply per pair)

__device__ __forceinline__ Fq load_fq_ldg(const Fq* p) {
  Fq r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int kMinBlocks>
__global__ void __launch_bounds__(kAccThreads, kMinBlocks)
    msm_accumulate_baa(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t M, uint32_t L,
                       uint32_t sentinel, const Affine* __restrict__ tbl, XYZZ* __restrict__ buckets,
                       XYZZ* __restrict__ heads, uint32_t* __restrict__ head_keys, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint64_t s = t * L;  // L is even: pairs are 8-byte aligned in keys / vals
  const uint64_t e = s + L < M ? s + L : M;
  uint32_t cur = keys[s];
  if (cur >= sentinel) {
    head_keys[t] = sentinel;
    return;
  }
  head_keys[t] = cur;
  XYZZ acc = xyzz_inf();
  bool first_run = true;
  auto flush = [&]() {
    if (first_run) store_xyzz(heads + t, acc);
    else store_xyzz(buckets + cur, acc);
    first_run = false;
  };
  auto load_pair = [&](uint64_t i0, uint32_t& ka, uint32_t& kb, uint32_t& va, uint32_t& vb) {
    if (i0 + 1 < e) {
      const uint2 k2 = *reinterpret_cast<const uint2*>(keys + i0);
      const uint2 v2 = *reinterpret_cast<const uint2*>(vals + i0);
      ka = k2.x; kb = k2.y; va = v2.x; vb = v2.y;
    } else {
      ka = keys[i0]; va = vals[i0];
      kb = sentinel; vb = 0;
    }
  };
  Fq suf[kBaaBatch];
  const uint32_t npairs = (uint32_t)((e - s + 1) / 2);
  for (uint32_t base = 0; base < npairs; base += kBaaBatch) {
    const uint32_t cnt = npairs - base < (uint32_t)kBaaBatch ? npairs - base : (uint32_t)kBaaBatch;
    const uint64_t b0 = s + 2 * (uint64_t)base;
    // ---- phase 1: suffix products of the x differences ----
    Fq S = Fq::one();
#pragma unroll 1
    for (int j = (int)cnt - 1; j >= 0; j--) {
      if (j >= kBaaPrefetch1) {
        uint32_t ka, kb, va, vb;
        load_pair(b0 + 2 * (uint64_t)(j - kBaaPrefetch1), ka, kb, va, vb);
        if (ka == kb && kb < sentinel) {
          prefetch_l2(tbl + (va & 0x7fffffffu));
          prefetch_l2(tbl + (vb & 0x7fffffffu));
        }
      }
      uint32_t ka, kb, va, vb;
      load_pair(b0 + 2 * (uint64_t)j, ka, kb, va, vb);
      suf[j] = S;
      Fq dx = Fq::one();
      if (ka == kb && kb < sentinel) {
        const Fq xa = load_fq_ldg(&tbl[va & 0x7fffffffu].x);
        const Fq xb = load_fq_ldg(&tbl[vb & 0x7fffffffu].x);
        const Fq d = fe_sub(xb, xa);
        if (!xa.is_zero() && !xb.is_zero() && !d.is_zero()) dx = d;
      }
      S = CallOps::mul(S, dx);
    }
    // ---- one inversion for the batch ----
    Fq I = fe_inv_safegcd(S);
    // ---- phase 2: pair sums, then the mixed addition ----
    {
      uint32_t ka, kb, va, vb;
      load_pair(b0, ka, kb, va, vb);
      if (ka < sentinel) prefetch_l2(tbl + (va & 0x7fffffffu));
      if (kb < sentinel) prefetch_l2(tbl + (vb & 0x7fffffffu));
    }
#pragma unroll 1
    for (uint32_t j = 0; j < cnt; j++) {
      uint32_t ka, kb, va, vb;
      load_pair(b0 + 2 * (uint64_t)j, ka, kb, va, vb);
      if (ka >= sentinel) break;  // nothing but sentinels from here on
      if (j + 1 < cnt) {
        uint32_t ka2, kb2, va2, vb2;
        load_pair(b0 + 2 * (uint64_t)(j + 1), ka2, kb2, va2, vb2);
        if (ka2 < sentinel) prefetch_l2(tbl + (va2 & 0x7fffffffu));
        if (kb2 < sentinel) prefetch_l2(tbl + (vb2 & 0x7fffffffu));
      }
      const bool has_b = kb < sentinel;
      Affine a = load_affine(tbl + (va & 0x7fffffffu));
      Affine b;
      b.x = Fq::zero(); b.y = Fq::zero();
      if (has_b) b = load_affine(tbl + (vb & 0x7fffffffu));
      bool valid = false;
      Fq dx = Fq::one();
      if (ka == kb && has_b) {
        const Fq d = fe_sub(b.x, a.x);
        if (!a.x.is_zero() && !b.x.is_zero() && !d.is_zero()) {
          dx = d;
          valid = true;
        }
      }
      if (va >> 31) a.y = fe_neg(a.y);
      if (vb >> 31) b.y = fe_neg(b.y);
      const Fq inv = CallOps::mul(I, suf[j]);
      I = CallOps::mul(I, dx);
      Affine op = a;
      if (valid) {
        const Fq lam = CallOps::mul(fe_sub(b.y, a.y), inv);
        const Fq x3 = fe_sub(fe_sub(CallOps::sqr(lam), a.x), b.x);
        op.y = fe_sub(CallOps::mul(lam, fe_sub(a.x, x3)), a.y);
        op.x = x3;
      }
      if (ka != cur) {
        flush();
        cur = ka;
        acc = xyzz_inf();
      }
      xyzz_madd_t<CallOps>(acc, op);
      if (!valid) {
        if (!has_b) break;  // b is the sentinel or past the end of the segment
        if (kb != ka) {     // a closed its run, b opens the next one
          flush();
          cur = kb;
          acc = xyzz_from_affine(b);
        } else {
          xyzz_madd_t<CallOps>(acc, b);  // same bucket, not addable in affine form (equal x, infinity): general path
        }
      }
    }
  }
  flush();
}

// 4. head merge.  Heads with the same key form a chain (consecutive, since segments are
// in key order).  Typical chains have one or two heads, but skewed scalars - and the top
// window, which holds only the few leading scalar bits and so concentrates n entries on a
// handful of buckets - make chains of thousands.  Chains are cut at multiples of
// kMergeFan: the chain's first head folds the heads up to the next cut into the bucket
// (its only writer on this level); a head sitting on a cut sums its kMergeFan-piece into a
// slot of the next level, where the same rule applies.  Depth log_fan(longest chain); fan 16, or 8 for large inputs.
constexpr int kMergeFanDefault = 16;

__global__ void __launch_bounds__(128) msm_merge_level(XYZZ* __restrict__ buckets, const XYZZ* __restrict__ heads,
                                                       const uint32_t* __restrict__ head_keys, uint64_t T,
                                                       uint32_t sentinel, XYZZ* __restrict__ next_heads,
                                                       uint32_t* __restrict__ next_keys, uint32_t kMergeFan) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint32_t k = head_keys[t];
  if (k >= sentinel) return;
  const bool chain_start = (t == 0) || head_keys[t - 1] != k;
  const bool on_cut = (t % kMergeFan) == 0;
  if (!chain_start && !on_cut) return;
  uint64_t end = (t / kMergeFan + 1) * kMergeFan;
  if (end > T) end = T;
  XYZZ acc = xyzz_inf();
  for (uint64_t j = t; j < end && head_keys[j] == k; j++) {
    XYZZ h = load_xyzz(heads + j);
    xyzz_add(acc, h);
  }
  if (chain_start) {
    XYZZ b = load_xyzz(buckets + k);
    xyzz_add(b, acc);
    store_xyzz(buckets + k, b);
  } else {
    store_xyzz(next_heads + t / kMergeFan, acc);
    next_keys[t / kMergeFan] = k;
  }
}

// ---------------------------------------------------------------------------
// 5. bucket reduce: sum_{idx} (idx + 1) * B[idx]
// ---------------------------------------------------------------------------
// Two levels.  Level 1 (msm_bucket_chunks): thread j owns buckets [j*Lb, (j+1)*Lb); running sums give
// A_j = sum B and S_j = sum (idx - j*Lb + 1) B.  The whole sum is sum_j S_j + Lb * sum_j j A_j, and the second
// term is the same problem over the n1 = nb / Lb values A_j: level 2 (msm_bucket_reduce on A_1 ..) solves it with
// short chunks and pays the offset of a chunk with a double-and-add by the (small, public) chunk start; the factor Lb
// is one more small double-and-add on each level-2 partial.  Level 1 therefore runs 2 additions
// per bucket and nothing else (round 1 ran the double-and-add in every level-1 thread: +22 % at c = 22, +100 % at
// c = 20, and could not use more than two warps per sub-partition because of it).
// run = sum B, sum = sum (idx - lo + 1) B over [lo, hi), walking down from hi.  ONE instance of the general addition
// in the loop: a step either adds the next bucket to `run` or `run` to `sum`, operands picked by selects - two inlined
// additions (~33 KB of SASS each) do not fit the 32 KB instruction cache, and the kernel then waits on instruction
// fetches instead of the multiply pipe.
__device__ __forceinline__ XYZZ xyzz_select(bool c, const XYZZ& a, const XYZZ& b) {
  XYZZ r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.x.v[i] = c ? a.x.v[i] : b.x.v[i];
    r.y.v[i] = c ? a.y.v[i] : b.y.v[i];
    r.zz.v[i] = c ? a.zz.v[i] : b.zz.v[i];
    r.zzz.v[i] = c ? a.zzz.v[i] : b.zzz.v[i];
  }
  return r;
}
__device__ __forceinline__ void running_sums(const XYZZ* __restrict__ buckets, uint64_t lo, uint64_t hi, XYZZ& run, XYZZ& sum) {
  if (hi <= lo) return;
  uint64_t idx = hi;
  XYZZ q = load_xyzz(buckets + (idx - 1));
  bool to_run = true;
#pragma unroll 1
  while (true) {
    XYZZ acc = xyzz_select(to_run, run, sum);
    xyzz_add(acc, q);
    if (to_run) {
      run = acc;
      q = acc;  // next step: sum += run
    } else {
      sum = acc;
      if (--idx == lo) break;
      q = load_xyzz(buckets + (idx - 1));
    }
    to_run = !to_run;
  }
}

// blockIdx.y = bucket set of a batch (sets of nb buckets back to back)
template <int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks)
    msm_bucket_chunks(const XYZZ* __restrict__ buckets_all, uint32_t nb, uint32_t Lb, XYZZ* __restrict__ out_s_all,
                      uint64_t s_stride, XYZZ* __restrict__ out_a_all, uint32_t n1) {
  const XYZZ* __restrict__ buckets = buckets_all + (size_t)blockIdx.y * nb;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n1) return;
  uint64_t lo = (uint64_t)j * Lb;
  uint64_t hi = lo + Lb < nb ? lo + Lb : nb;
  XYZZ run = xyzz_inf(), sum = xyzz_inf();
  running_sums(buckets, lo, hi, run, sum);
  store_xyzz(out_s_all + (size_t)blockIdx.y * s_stride + j, sum);
  store_xyzz(out_a_all + (size_t)blockIdx.y * n1 + j, run);
}

// thread j owns values [j*Lb, (j+1)*Lb) of its set: S = sum (idx - j*Lb + 1) V plus (j*Lb) * sum V by
// double-and-add, then scale * that.  Set y starts at in_all + y * in_stride and holds nb values.
__global__ void __launch_bounds__(128) msm_bucket_reduce(const XYZZ* __restrict__ in_all, uint64_t in_stride, uint32_t nb,
                                                         uint32_t Lb, int scale, XYZZ* __restrict__ out_all,
                                                         uint64_t out_stride, uint32_t nchunks) {
  const XYZZ* __restrict__ buckets = in_all + (size_t)blockIdx.y * in_stride;
  XYZZ* __restrict__ out = out_all + (size_t)blockIdx.y * out_stride;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nchunks) return;
  uint64_t lo = (uint64_t)j * Lb;
  uint64_t hi = lo + Lb < nb ? lo + Lb : nb;
  XYZZ run = xyzz_inf(), sum = xyzz_inf();
  running_sums(buckets, lo, hi, run, sum);
  // sum += lo * run  (MSB-first double-and-add on the small public scalar lo)
  if (lo != 0 && !xyzz_is_inf(run)) {
    XYZZ m = xyzz_inf();
    int top = 63 - __clzll((unsigned long long)lo);
#pragma unroll 1
    for (int bit = top; bit >= 0; bit--) {
      xyzz_dbl(m);
      if ((lo >> bit) & 1) xyzz_add(m, run);
    }
    xyzz_add(sum, m);
  }
  // sum *= scale (the level-1 chunk length, a small public constant)
  if (scale > 1 && !xyzz_is_inf(sum)) {
    XYZZ m = sum;
#pragma unroll 1
    for (int bit = 30 - __clz(scale); bit >= 0; bit--) {
      xyzz_dbl(m);
      if ((scale >> bit) & 1) xyzz_add(m, sum);
    }
    sum = m;
  }
  store_xyzz(out + j, sum);
}

// block-wide tree sum of XYZZ values; out[blockIdx.x] = sum of in[block range]
constexpr int kTreeThreads = 128;
// blockIdx.y = independent segment (in: in_stride values apart, out: out_stride)
__global__ void __launch_bounds__(kTreeThreads) xyzz_tree_reduce(const XYZZ* __restrict__ in_all, uint64_t n,
                                                                 XYZZ* __restrict__ out_all, uint64_t in_stride,
                                                                 uint64_t out_stride) {
  __shared__ XYZZ sm[kTreeThreads];
  const XYZZ* __restrict__ in = in_all + (size_t)blockIdx.y * in_stride;
  XYZZ* __restrict__ out = out_all + (size_t)blockIdx.y * out_stride;
  uint64_t i = (uint64_t)blockIdx.x * (2 * kTreeThreads) + threadIdx.x;
  XYZZ acc = xyzz_inf();
  if (i < n) acc = load_xyzz(in + i);
  if (i + kTreeThreads < n) {
    XYZZ b = load_xyzz(in + i + kTreeThreads);
    xyzz_add(acc, b);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
#pragma unroll 1
  for (int d = kTreeThreads / 2; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) {
      XYZZ b = sm[threadIdx.x + d];
      xyzz_add(acc, b);
      sm[threadIdx.x] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) store_xyzz(out + blockIdx.x, acc);
}

// one thread per point (each does its own Fermat inversion)
__global__ void xyzz_to_affine_bytes(const XYZZ* in, size_t count, uint32_t* out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  XYZZ p = load_xyzz(in + t);
  Affine a = xyzz_to_affine(p);
  Fq x = fe_from_mont(a.x), y = fe_from_mont(a.y);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[t * 16 + i] = x.v[i];
    out[t * 16 + 8 + i] = y.v[i];
  }
}

__global__ void xyzz_set_inf(XYZZ* out) { store_xyzz(out, xyzz_inf()); }

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
// Window choice from measurements on B200 (profiles/phase_sweep_r1*.json): per entry the
// accumulate costs 0.161 ns and the sort 0.02 ns; the bucket reduce costs ~0.74 ns per
// bucket with a ~0.35 ms latency floor; long head chains make small windows a bad deal
// for small inputs.
static int pick_window(const myzkp_ctx* ctx, size_t n) {
  const int forced = ctx->window_bits;
  if (forced >= 1 && forced <= 24 && ((ctx->windows >> forced) & 1)) return forced;
  auto has = [&](int c) { return ((ctx->windows >> c) & 1) != 0; };
  const int want = n < ((size_t)1 << 11) ? 8 : n < ((size_t)1 << 19) ? 16 : n < ((size_t)1 << 23) ? 20 : 22;
  if (has(want)) return want;
  // a restricted table (myzkp_ctx_set_table_windows, or the lean rows of a very large SRS): the available
  // window with the least modelled time - entries at 0.16 ns, buckets at 0.74 ns (numbers above)
  int best = 0;
  double best_t = 0;
  for (int c = 1; c <= 24; c++) {
    if (!has(c)) continue;
    const double t = 0.16 * (double)((255 + c - 1) / c) * (double)n + 0.74 * (double)((size_t)1 << (c - 1));
    if (!best || t < best_t) { best = c; best_t = t; }
  }
  return best ? best : want;
}

// tree-sum `count` XYZZ values living in buffer `a` (ping-pong with `b`); the
// single result ends in *d_out
// K independent segments of `count` values each (segment y starts at a + y * count); results in d_out[y]
static int tree_sum(myzkp_ctx* ctx, XYZZ* a, XYZZ* b, uint64_t count, XYZZ* d_out, uint32_t K = 1) {
  while (true) {
    uint64_t blocks = (count + 2 * kTreeThreads - 1) / (2 * kTreeThreads);
    XYZZ* dst = (blocks == 1) ? d_out : b;
    xyzz_tree_reduce<<<dim3((unsigned)blocks, K), kTreeThreads, 0, ctx->stream>>>(a, count, dst, count, blocks);
    MZ_LAUNCH_CHECK(ctx);
    if (blocks == 1) return MYZKP_OK;
    count = blocks;
    XYZZ* tmp = a; a = b; b = tmp;
  }
}

int msm_pick_window(const myzkp_ctx* ctx, size_t n) { return pick_window(ctx, n); }

// a[k] += b[k] over a whole bucket set (dense: every lane has an addition to do)
__global__ void __launch_bounds__(128) msm_bucket_add(XYZZ* __restrict__ a, const XYZZ* __restrict__ b, uint32_t nb) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nb) return;
  XYZZ x = load_xyzz(a + k);
  XYZZ y = load_xyzz(b + k);
  xyzz_add(x, y);
  store_xyzz(a + k, x);
}

int msm_add_buckets(myzkp_ctx* ctx, XYZZ* a, const XYZZ* b, int c) {
  const uint32_t nb = 1u << (c - 1);
  msm_bucket_add<<<(nb + 127) / 128, 128, 0, ctx->stream>>>(a, b, nb);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

void msm_chunk_range(size_t n, int K, int pos, double ratio, size_t* lo, size_t* hi) {
  auto cum = [&](int p) -> size_t {
    if (p <= 0) return 0;
    if (p >= K) return n;
    if (ratio == 1.0) return (size_t)(((unsigned __int128)n * (unsigned)p) / (unsigned)K);
    if (ratio == 4.0)
      return (size_t)(((unsigned __int128)n * ((((unsigned __int128)1) << (2 * p)) - 1)) / ((((unsigned __int128)1) << (2 * K)) - 1));
    const double f = (pow(ratio, p) - 1.0) / (pow(ratio, K) - 1.0);
    size_t v = (size_t)((double)n * f);
    return v > n ? n : v;
  };
  size_t a = cum(pos), b = cum(pos + 1);
  if (b < a) b = a;
  *lo = a;
  *hi = b;
}

// Sort / accumulate pipeline of one large MSM (EXPERIMENT, off by default: measured slower on B200, see pipe_chunks).
// The recode and the radix sort are bandwidth- and latency-bound and
// leave the multiply pipe idle; the accumulate is bound by that pipe alone (DRAM 12 % busy).  So the scalars are cut
// into chunks growing by kPipeRatio: chunk k+1 is recoded and sorted on a high-priority child stream (own scratch)
// while chunk k is accumulated on this stream onto the same bucket set (msm_accumulate<kOnto>); only the small first
// chunk's sort is exposed.  Returns the number of chunks to use for n scalars (1 = plain pipeline).
static int pipe_chunks(const myzkp_ctx* ctx, size_t n) {
  if (ctx->is_child) return 1;       // children are the pipeline's (and Gemini's) workers
  if (ctx->peer_same_device) return 1;  // emulated ranks on one device (tests): scratch is frozen, keep to one stream
  static const int env = getenv("MZ_PIPE_CHUNKS") ? atoi(getenv("MZ_PIPE_CHUNKS")) : 0;  // experiment knob
  int K = ctx->pipe_chunks > 0 ? ctx->pipe_chunks : env;
  if (K <= 0) K = 1;  // off by default: measured slower on B200 (profiles/experiments_r2.md)
  if (K > myzkp_ctx::kMaxPipe) K = myzkp_ctx::kMaxPipe;
  if ((size_t)K > n) K = 1;
  return K;
}
static double pipe_ratio() {
  static const double env = getenv("MZ_PIPE_RATIO") ? atof(getenv("MZ_PIPE_RATIO")) : 0.0;  // experiment knob
  return env >= 1.0 ? env : 4.0;
}

static int msm_xyzz_pipelined(myzkp_ctx* ctx, const uint32_t* d_scalars, size_t n, size_t srs_off, int c, int K,
                              XYZZ* buckets) {
  for (int k = 0; k < K; k++) {
    if (!ctx->pipe_sorted_ev[k]) MZ_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_sorted_ev[k], cudaEventDisableTiming));
    if (!ctx->pipe_acc_ev[k]) MZ_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_acc_ev[k], cudaEventDisableTiming));
  }
  if (!ctx->fork_ev) MZ_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  MZ_CUDA_TRY(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));  // the scalars are ready from here on
  const double ratio = pipe_ratio();
  SortedEntries se[myzkp_ctx::kMaxPipe];
  myzkp_ctx* used[2] = {nullptr, nullptr};
  int rc = MYZKP_OK;
  auto enqueue_sort = [&](int k) -> int {
    myzkp_ctx* ch = nullptr;
    MZ_TRY(get_child(ctx, kPipeChild0 + (k & 1), &ch));
    used[k & 1] = ch;
    // a child's scratch is free again once the accumulate of its previous chunk has run
    MZ_CUDA_TRY(ctx, cudaStreamWaitEvent(ch->stream, k < 2 ? ctx->fork_ev : ctx->pipe_acc_ev[k - 2], 0));
    size_t lo, hi;
    msm_chunk_range(n, K, k, ratio, &lo, &hi);
    MsmItem it{d_scalars + lo * 8, hi - lo};
    int r = msm_sort_entries(ch, ctx, k, &it, 1, srs_off + lo, c, false, &se[k]);
    if (r != MYZKP_OK && ctx->err.empty()) ctx->err = ch->err;
    ctx->launches += ch->launches;
    ch->launches = 0;
    if (r != MYZKP_OK) return r;
    MZ_CUDA_TRY(ctx, cudaEventRecord(ctx->pipe_sorted_ev[k], ch->stream));
    return MYZKP_OK;
  };
  rc = enqueue_sort(0);
  for (int k = 0; k < K && rc == MYZKP_OK; k++) {
    if (k + 1 < K) rc = enqueue_sort(k + 1);
    if (rc != MYZKP_OK) break;
    rc = cudaStreamWaitEvent(ctx->stream, ctx->pipe_sorted_ev[k], 0) == cudaSuccess ? MYZKP_OK : MYZKP_ERR_CUDA;
    if (rc != MYZKP_OK) break;
    rc = msm_accumulate_sorted(ctx, k, se[k], buckets, /*onto=*/k > 0);
    if (rc != MYZKP_OK) break;
    rc = cudaEventRecord(ctx->pipe_acc_ev[k], ctx->stream) == cudaSuccess ? MYZKP_OK : MYZKP_ERR_CUDA;
  }
  if (rc != MYZKP_OK) {
    // error path: nothing of the children may still be reading the scalars or writing scratch when the caller goes on
    for (myzkp_ctx* ch : used)
      if (ch) cudaStreamSynchronize(ch->stream);
    cudaGetLastError();
    ctx->chunk_idx = 0;
    if (ctx->err.empty()) ctx->err = "sort / accumulate pipeline failed";
  }
  return rc;
}

int msm_xyzz(myzkp_ctx* ctx, const uint32_t* d_scalars, size_t n, size_t srs_off, XYZZ* d_out) {
  if (n == 0) {
    xyzz_set_inf<<<1, 1, 0, ctx->stream>>>(d_out);
    MZ_LAUNCH_CHECK(ctx);
    return MYZKP_OK;
  }
  const int c = pick_window(ctx, n);
  MZ_CUDA_TRY(ctx, ctx->buckets.ensure(((size_t)1 << (c - 1)) * sizeof(XYZZ)));
  const int Kp = pipe_chunks(ctx, n);
  if (Kp > 1) {
    if (srs_off + n > ctx->srs_n)
      return fail(ctx, MYZKP_ERR_INVALID_ARG, "polynomial longer than the SRS (reference panics at polynomial.rs:162)");
    MZ_TRY(msm_xyzz_pipelined(ctx, d_scalars, n, srs_off, c, Kp, ctx->buckets.as<XYZZ>()));
  } else {
    MZ_TRY(msm_fill_buckets(ctx, d_scalars, n, srs_off, c, ctx->buckets.as<XYZZ>()));
  }
  return msm_reduce_buckets(ctx, c, ctx->buckets.as<XYZZ>(), d_out);
}

// Window for a batch of K polynomials run as one pipeline: latency does not matter there, so take
// the supported window with the least work - 10 multiplies per entry (W(c) entries per coefficient)
// against 2 x 14 multiplies per bucket (2^(c-1) buckets per polynomial) - within a 2 GiB bucket budget.
int msm_pick_window_batch(const myzkp_ctx* ctx, const MsmItem* items, size_t K) {
  {  // a forced window (myzkp_ctx_set_msm_params) applies to batches too, when the bucket sets fit
    const int forced = ctx->window_bits;
    if (forced >= 1 && forced <= 24 && ((ctx->windows >> forced) & 1) && ((uint64_t)K << (forced - 1)) < (1ull << 32) - 1 &&
        ((uint64_t)K << (forced - 1)) * sizeof(XYZZ) <= (8ull << 30))
      return forced;
  }
  uint64_t total = 0;
  for (size_t y = 0; y < K; y++) total += items[y].n;
  int best = 0;
  double best_cost = 0;
  for (int c = 4; c <= 22; c++) {
    if (!((ctx->windows >> c) & 1)) continue;
    const uint64_t nbk = (uint64_t)K << (c - 1);
    if (nbk * sizeof(XYZZ) > (2ull << 30) || nbk >= (1ull << 32) - 1) continue;
    const int W = (255 + c - 1) / c;
    const double cost = 10.0 * W * (double)total + 28.0 * (double)nbk;
    if (!best || cost < best_cost) { best = c; best_cost = cost; }
  }
  return best;
}

// K commitments in one pipeline: d_out[y] = sum_i items[y].scalars[i] * SRS[srs_off + i]
int msm_batch_xyzz(myzkp_ctx* ctx, const MsmItem* items, size_t K, size_t srs_off, XYZZ* d_out) {
  if (K == 0) return MYZKP_OK;
  if (K == 1) return msm_xyzz(ctx, items[0].d_scalars, items[0].n, srs_off, d_out);
  const int c = msm_pick_window_batch(ctx, items, K);
  if (!c) return fail(ctx, MYZKP_ERR_INVALID_ARG, "batch too large for one pipeline");
  MZ_CUDA_TRY(ctx, ctx->buckets.ensure(((size_t)K << (c - 1)) * sizeof(XYZZ)));
  MZ_TRY(msm_fill_buckets_batch(ctx, items, K, srs_off, c, ctx->buckets.as<XYZZ>()));
  return msm_reduce_buckets(ctx, c, ctx->buckets.as<XYZZ>(), d_out, K);
}

// steps 1-4: recode, sort, accumulate, merge -> `buckets` (2^(c-1) XYZZ, overwritten) holds the
// bucket sums of sum_i scalars[i] * SRS[srs_off + i] for window c
int msm_fill_buckets(myzkp_ctx* ctx, const uint32_t* d_scalars, size_t n, size_t srs_off, int c, XYZZ* buckets, bool onto) {
  MsmItem one{d_scalars, n};
  return msm_fill_buckets_batch(ctx, &one, 1, srs_off, c, buckets, false, onto);
}

// The same for a batch of K polynomials sharing the window c (each against SRS[srs_off ...]) in ONE
// pipeline: polynomial y owns the bucket range [y nb, (y+1) nb) of `buckets` (K 2^(c-1) XYZZ), so one
// sort, one accumulate and one merge serve the whole batch - many small commitments cost their
// entries, not K latency-bound pipelines.
int msm_fill_buckets_batch(myzkp_ctx* ctx, const MsmItem* items, size_t K, size_t srs_off, int c, XYZZ* buckets,
                           bool per_window, bool onto) {
  const int chunk = ctx->chunk_idx;
  SortedEntries se;
  MZ_TRY(msm_sort_entries(ctx, ctx, chunk, items, K, srs_off, c, per_window, &se));
  return msm_accumulate_sorted(ctx, chunk, se, buckets, onto);
}

#define MZ_PHASE_ON(tctx, chunk) ((tctx)->phase_timing && (tctx)->phase_ev[0][0][0] && (chunk) < myzkp_ctx::kMaxPipe)

// steps 1-2: recode and sort on actx (stream, scratch); events and the non-canonical flag belong to tctx
int msm_sort_entries(myzkp_ctx* actx, myzkp_ctx* tctx, int chunk, const MsmItem* items, size_t K, size_t srs_off, int c,
                     bool per_window, SortedEntries* out) {
  myzkp_ctx* ctx = actx;
  if (!ctx->table) return fail(ctx, MYZKP_ERR_NO_SRS, "no SRS loaded");
  if (c < 1 || c > 24 || (!per_window && !((ctx->windows >> c) & 1)))
    return fail(ctx, MYZKP_ERR_INVALID_ARG, "window not supported by the table");
  if (K == 0 || K > 65535) return fail(ctx, MYZKP_ERR_INVALID_ARG, "batch of 1..65535 polynomials");
  const int W = (255 + c - 1) / c;
  const uint32_t nb1 = 1u << (c - 1);
  const uint64_t sets = (uint64_t)K * (per_window ? (uint64_t)W : 1u);  // bucket sets of nb1 buckets each
  if (sets * nb1 >= (1ull << 32) - 1) return fail(ctx, MYZKP_ERR_INVALID_ARG, "batch needs more than 2^32 buckets");
  const uint32_t nb = (uint32_t)(sets * nb1);  // buckets of the whole batch; also the sentinel key
  uint64_t M = 0;
  size_t n = 0;  // longest polynomial
  std::vector<RecodeDesc> descs(K);
  for (size_t y = 0; y < K; y++) {
    if (srs_off + items[y].n > ctx->srs_n)
      return fail(ctx, MYZKP_ERR_INVALID_ARG, "polynomial longer than the SRS (reference panics at polynomial.rs:162)");
    descs[y] = RecodeDesc{items[y].d_scalars, items[y].n, M, (uint32_t)(y * (per_window ? (uint64_t)W : 1u)) * nb1, 0};
    M += (uint64_t)W * items[y].n;
    if (items[y].n > n) n = items[y].n;
  }
  if (M >= (1ull << 32)) return fail(ctx, MYZKP_ERR_INVALID_ARG, "MSM too large for 32-bit entry indices");
  int sort_bits = c;  // c-1 bucket bits + the sentinel bit
  while (sort_bits < 32 && (1ull << sort_bits) <= nb) sort_bits++;

  MZ_CUDA_TRY(ctx, ctx->keys_a.ensure(M * 4));
  MZ_CUDA_TRY(ctx, ctx->keys_b.ensure(M * 4));
  MZ_CUDA_TRY(ctx, ctx->vals_a.ensure(M * 4));
  MZ_CUDA_TRY(ctx, ctx->vals_b.ensure(M * 4));
  MZ_CUDA_TRY(tctx, tctx->small.ensure(4096));
  int* flag = reinterpret_cast<int*>(tctx->small.as<uint8_t>() + 512);

  uint32_t* keys_a = ctx->keys_a.as<uint32_t>();
  uint32_t* keys_b = ctx->keys_b.as<uint32_t>();
  uint32_t* vals_a = ctx->vals_a.as<uint32_t>();
  uint32_t* vals_b = ctx->vals_b.as<uint32_t>();

  const int slot = (int)(tctx->msm_count % myzkp_ctx::kPhaseSlots);
  const bool timing = MZ_PHASE_ON(tctx, chunk);
#define MZ_PHASE(i) do { if (timing) MZ_CUDA_TRY(ctx, cudaEventRecord(tctx->phase_ev[slot][chunk][i], ctx->stream)); } while (0)
  tctx->phase_valid[slot] = false;
  MZ_PHASE(0);
  // 1. recode
  const RecodeDesc* d_descs = nullptr;
  if (K > 1) {
    MZ_CUDA_TRY(ctx, ctx->descs.ensure(K * sizeof(RecodeDesc)));
    // pageable source: staged before the call returns
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->descs.p, descs.data(), K * sizeof(RecodeDesc), cudaMemcpyHostToDevice, ctx->stream));
    d_descs = ctx->descs.as<RecodeDesc>();
  }
  // MSD split inside the recode when the key has more than 16 bits (windows 20, 22, 24: the recode kernels are
  // compiled per window): the sort then only handles the low 16 (or 24) bits, inside each partition
  int low_bits = 0, P = 0, msd_r = 0;
  if (sort_bits > 16 && !per_window && (c == 20 || c == 22 || c == 24) && !getenv("MZ_NO_PARTITION")) {
    low_bits = 16;
    if (((uint64_t)nb >> low_bits) + 1 > (uint64_t)(kMaxParts - 3)) low_bits = 24;
    // One polynomial: MSD sort (sort.cu: radix_sort_pairs_msd).  The partition keeps the key bits above 8 + r, a
    // 256-way pass follows, and the groups that remain - they share all but the last r key bits - are sorted in
    // shared memory: r is the largest value whose expected group size M / 2^(c-1-r) fits the group sort.
    static const bool no_msd = getenv("MZ_SORT_LSD") != nullptr;  // experiment knob: the two-pass LSD form
    if (K == 1 && !no_msd) {
      static const int fill_pct = getenv("MZ_SORT_GROUP_FILL") ? atoi(getenv("MZ_SORT_GROUP_FILL")) : 95;  // experiment knob
      int r = 8;
      while (r > 4 && (M >> (c - 1 - r)) > (uint64_t)sort_group_cap() * (uint64_t)fill_pct / 100) r--;
      const int a = c - 1 - 8 - r;  // key bits decided by the partition
      if (a >= 1 && a <= 8) {
        msd_r = r;
        low_bits = 8 + r;
      }
    }
    if (sort_bits > low_bits) P = (int)(((uint64_t)nb >> low_bits) + 1);
    if (P < 2 || P > kMaxParts - 3) { P = 0; low_bits = 0; msd_r = 0; }
  }
  const uint32_t* d_parts = nullptr;
  if (n && P) {
    RecodeCtl ctl{low_bits, P, (uint32_t)ctx->srs_n, (uint32_t)srs_off, nb};
    RecodeRows rows;
    for (int w = 0; w < 64; w++) rows.row[w] = w < W ? ctx->row_of_bit[w * c] : 0;
    MZ_CUDA_TRY(ctx, ctx->sort_parts.ensure((size_t)(4 * P + 8) * sizeof(uint32_t)));
    uint32_t* parts = ctx->sort_parts.as<uint32_t>();
    uint32_t* counts = parts + 2 * (P + 1);
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, (size_t)P * sizeof(uint32_t), ctx->stream));
    if (c == 20) MZ_TRY(launch_partitioned_recode<20>(ctx, descs[0], d_descs, K, n, ctl, rows, parts, counts, keys_a, vals_a, flag));
    else if (c == 22) MZ_TRY(launch_partitioned_recode<22>(ctx, descs[0], d_descs, K, n, ctl, rows, parts, counts, keys_a, vals_a, flag));
    else MZ_TRY(launch_partitioned_recode<24>(ctx, descs[0], d_descs, K, n, ctl, rows, parts, counts, keys_a, vals_a, flag));
    d_parts = parts;
  } else if (n) {
    msm_recode<<<dim3((unsigned)((n + 255) / 256), (unsigned)K), 256, 0, ctx->stream>>>(
        descs[0], d_descs, c, W, ctx->d_row_of_bit, (uint32_t)ctx->srs_n, (uint32_t)srs_off, nb, keys_a, vals_a, flag,
        per_window ? nb1 : 0u);
    MZ_LAUNCH_CHECK(ctx);
  }

  MZ_PHASE(1);
  // 2. sort by bucket key (c bits: c-1 bucket bits + the sentinel bit; only the low bits when partitioned)
  uint32_t *keys_s = nullptr, *vals_s = nullptr;
  if (d_parts && msd_r)
    MZ_TRY(radix_sort_pairs_msd(ctx, keys_a, vals_a, keys_b, vals_b, M, msd_r, &keys_s, &vals_s, d_parts, P));
  else
    MZ_TRY(radix_sort_pairs(ctx, keys_a, vals_a, keys_b, vals_b, M, d_parts ? low_bits : sort_bits, &keys_s, &vals_s,
                            d_parts, P, !getenv("MZ_SORT_STABLE")));
  MZ_PHASE(2);
#undef MZ_PHASE
  out->keys = keys_s;
  out->vals = vals_s;
  out->M = M;
  out->nb = nb;
  out->c = c;
  out->W = W;
  return MYZKP_OK;
}

// Segment length of the accumulate: enough segments to fill the GPU several times over, but not much
// shorter than the average bucket run - every extra segment inside a run costs a head that the merge must fold in.
static uint32_t pick_segment_len(const myzkp_ctx* ctx, uint64_t M) {
  uint32_t L = (uint32_t)ctx->segment_len;
  if (L == 0) {
    // Measured on B200 (profiles/phase_sweep_r1d.json): the head merge costs ~0.64 ns per
    // segment while doubling L slows the accumulate by ~2 %, so aim at four waves of
    // resident threads (~300 k segments) with L between 32 (16 for tiny inputs) and 256.
    uint64_t target_threads = (uint64_t)ctx->sm_count * 512 * 4;
    uint64_t l = (M + target_threads - 1) / target_threads;
    const uint64_t l_min = M >= ((uint64_t)1 << 19) ? 32 : 16;
    L = (uint32_t)(l < l_min ? l_min : (l > 256 ? 256 : l));
    // quantise to whole waves of resident blocks (4 blocks of 128 threads per SM) so
    // the last wave is not half empty
    const uint64_t wave = (uint64_t)ctx->sm_count * 4 * kAccThreads;
    uint64_t waves = (M / L + wave - 1) / wave;
    if (waves >= 2 && waves <= 64) {
      uint64_t l2 = (M + waves * wave - 1) / (waves * wave);
      if (l2 >= 8 && l2 <= 256) L = (uint32_t)l2;
    }
  }
  L += L & 1;  // even: the fused batched-affine accumulate reads (key, val) pairs as aligned 8-byte words
  return L;
}
uint64_t msm_segments_for(const myzkp_ctx* ctx, int c, size_t n) {
  const uint64_t M = (uint64_t)((255 + c - 1) / c) * n;
  if (M == 0) return 0;
  const uint32_t L = pick_segment_len(ctx, M);
  return (M + L - 1) / L;
}

// step 4: merge T segment heads (keys in head_keys, sentinel = no head) into the buckets, level by level (levels
// beyond the first are nearly empty unless some bucket spans more than 16 segments)
int msm_merge_heads(myzkp_ctx* ctx, XYZZ* buckets, uint32_t nb, uint64_t T) {
  if (T == 0) return MYZKP_OK;
  static const int env_fan = getenv("MZ_MERGE_FAN") ? atoi(getenv("MZ_MERGE_FAN")) : 0;  // experiment knob
  // measured (scripts/mergefan_ab.sh): 2^24 points, 832 k heads: 0.51 / 0.38 / 0.36 ms for a fan of 16 / 8 / 4 (the top
  // window's 34-head chains are folded serially, one lane per cut); 2^21 points, 303 k heads: 0.124 / 0.134 / 0.149 ms
  // (the extra levels are launches of pure latency)
  const uint32_t kMergeFan = env_fan >= 2 && env_fan <= 64 ? (uint32_t)env_fan
                             : T >= ((uint64_t)1 << 19) ? 8u : (uint32_t)kMergeFanDefault;
  const XYZZ* cur_heads = ctx->heads.as<XYZZ>();
  const uint32_t* cur_keys = ctx->head_keys.as<uint32_t>();
  uint64_t Tc = T;
  const uint64_t T2max = (T + kMergeFan - 1) / kMergeFan;
  const size_t lvl_stride = (((size_t)T2max * (sizeof(XYZZ) + sizeof(uint32_t)) + 255) / 256 + 1) * 256;
  MZ_CUDA_TRY(ctx, ctx->heads2.ensure(2 * lvl_stride));
  int pp = 0;
  while (true) {
    const uint64_t T2 = (Tc + kMergeFan - 1) / kMergeFan;
    uint8_t* base = ctx->heads2.as<uint8_t>() + (size_t)pp * lvl_stride;
    XYZZ* nh = reinterpret_cast<XYZZ*>(base);
    uint32_t* nk = reinterpret_cast<uint32_t*>(base + (size_t)T2max * sizeof(XYZZ));
    if (Tc > kMergeFan) MZ_CUDA_TRY(ctx, cudaMemsetAsync(nk, 0xff, T2 * sizeof(uint32_t), ctx->stream));
    msm_merge_level<<<(unsigned)((Tc + 127) / 128), 128, 0, ctx->stream>>>(buckets, cur_heads, cur_keys, Tc, nb, nh, nk, kMergeFan);
    MZ_LAUNCH_CHECK(ctx);
    if (Tc <= kMergeFan) break;  // every head was inside the first cut: nothing was forwarded
    cur_heads = nh;
    cur_keys = nk;
    Tc = T2;
    pp ^= 1;
  }
  return MYZKP_OK;
}

// steps 3-4: accumulate the sorted entries into `buckets` and fold the segment heads, on ctx's stream.
int msm_accumulate_sorted(myzkp_ctx* ctx, int chunk, const SortedEntries& se, XYZZ* buckets, bool onto) {
  const uint32_t* keys_s = se.keys;
  const uint32_t* vals_s = se.vals;
  const uint64_t M = se.M;
  const uint32_t nb = se.nb;
  const int slot = (int)(ctx->msm_count % myzkp_ctx::kPhaseSlots);
  const bool timing = MZ_PHASE_ON(ctx, chunk);
#define MZ_PHASE(i) do { if (timing) MZ_CUDA_TRY(ctx, cudaEventRecord(ctx->phase_ev[slot][chunk][i], ctx->stream)); } while (0)
  // 3. accumulate
  const uint32_t L = pick_segment_len(ctx, M ? M : 1);
  auto finish = [&](uint64_t T) {
    ctx->msm_info[slot][0] = (uint64_t)se.c;
    ctx->msm_info[slot][1] = (uint64_t)se.W;
    ctx->msm_info[slot][2] = (chunk ? ctx->msm_info[slot][2] : 0) + M;
    ctx->msm_info[slot][3] = L;
    ctx->msm_info[slot][4] = (chunk ? ctx->msm_info[slot][4] : 0) + T;
    ctx->msm_info[slot][5] = nb;
    ctx->phase_valid[slot] = false;  // set by msm_reduce_buckets once event 5 is recorded
    ctx->phase_pending = timing;
    if (chunk < myzkp_ctx::kMaxPipe) ctx->phase_chunks[slot] = chunk + 1;
    ctx->chunk_idx = chunk + 1;
  };
  if (M == 0) {  // nothing but empty polynomials: every bucket is the point at infinity
    if (!onto) MZ_CUDA_TRY(ctx, cudaMemsetAsync(buckets, 0, (size_t)nb * sizeof(XYZZ), ctx->stream));
    MZ_PHASE(6);
    MZ_PHASE(3);
    MZ_PHASE(4);
    finish(0);
    return MYZKP_OK;
  }
  const uint64_t T = (M + L - 1) / L;
  MZ_CUDA_TRY(ctx, ctx->heads.ensure(T * sizeof(XYZZ)));
  MZ_CUDA_TRY(ctx, ctx->head_keys.ensure(T * 4));
  XYZZ* heads_p = ctx->heads.as<XYZZ>();
  uint32_t* hkeys_p = ctx->head_keys.as<uint32_t>();
  if (!onto) MZ_CUDA_TRY(ctx, cudaMemsetAsync(buckets, 0, (size_t)nb * sizeof(XYZZ), ctx->stream));
  MZ_PHASE(6);
  // Accumulate variants (myzkp_ctx_set_baa_rounds): 0 = XYZZ mixed additions only; -2 = fused batched-affine
  // pair sums (msm_accumulate_baa); -1 = automatic: fused when segments are long enough for the lane-private
  // batch inversion to amortise; 1..3 = the older multi-pass rounds (baa.cu, kept for comparison).
  int baa = ctx->baa_rounds;
  if (onto) baa = 0;  // only the XYZZ kernel knows how to start from existing bucket values
  const bool fused = (baa == -2 && L >= 4) || (baa == -1 && L >= (uint32_t)kBaaAutoMinL);
  if (fused) {
    static const int minb = getenv("MZ_BAA_MINB") ? atoi(getenv("MZ_BAA_MINB")) : 4;  // experiment knob
    if (minb == 3)
      msm_accumulate_baa<3><<<(unsigned)((T + kAccThreads - 1) / kAccThreads), kAccThreads, 0, ctx->stream>>>(
          keys_s, vals_s, M, L, nb, ctx->table, buckets, heads_p, hkeys_p, T);
    else
      msm_accumulate_baa<4><<<(unsigned)((T + kAccThreads - 1) / kAccThreads), kAccThreads, 0, ctx->stream>>>(
          keys_s, vals_s, M, L, nb, ctx->table, buckets, heads_p, hkeys_p, T);
    MZ_LAUNCH_CHECK(ctx);
  } else if (baa > 0 && L >= 4) {
    MZ_TRY(baa_accumulate(ctx, keys_s, vals_s, M, L, nb, baa, buckets, heads_p,
                          hkeys_p, T));
  } else {
    static const bool hint64 = getenv("MZ_GATHER_L2_64B") != nullptr;  // experiment knob
    const unsigned blocks = (unsigned)((T + kAccThreads - 1) / kAccThreads);
    if (onto)
      msm_accumulate<false, true><<<blocks, kAccThreads, 0, ctx->stream>>>(
          keys_s, vals_s, M, L, nb, ctx->table, buckets, heads_p, hkeys_p, T);
    else if (hint64)
      msm_accumulate<true, false><<<blocks, kAccThreads, 0, ctx->stream>>>(
          keys_s, vals_s, M, L, nb, ctx->table, buckets, heads_p, hkeys_p, T);
    else
      msm_accumulate<false, false><<<blocks, kAccThreads, 0, ctx->stream>>>(
          keys_s, vals_s, M, L, nb, ctx->table, buckets, heads_p, hkeys_p, T);
    MZ_LAUNCH_CHECK(ctx);
  }

  MZ_PHASE(3);
  // 4. merge the segment heads.  (Every chunk of the upload pipeline merges its own: the heads of several chunks
  // cannot share one merge launch, because the same bucket then has a chain in every chunk's range and their
  // first heads would read-modify-write it concurrently - tried, profiles/experiments_r2.md.)
  MZ_TRY(msm_merge_heads(ctx, buckets, nb, T));
  MZ_PHASE(4);
#undef MZ_PHASE
  finish(T);
  return MYZKP_OK;
}

// step 5: sum_k (k+1) * buckets[k] -> *d_out (XYZZ); K bucket sets back to back -> d_out[0..K)
int msm_reduce_buckets(myzkp_ctx* ctx, int c, const XYZZ* buckets, XYZZ* d_out, size_t K) {
  const uint32_t nb = 1u << (c - 1);
  const int slot = (int)(ctx->msm_count % myzkp_ctx::kPhaseSlots);
  const bool timing = ctx->phase_pending && ctx->phase_timing && ctx->phase_ev[0][0][0];
  // Level 1: chunks of Lb buckets - the kernel is two additions per bucket and nothing else.
  static const int env_lb = getenv("MZ_REDUCE_LB") ? atoi(getenv("MZ_REDUCE_LB")) : 0;        // experiment knobs
  static const int env_lb2 = getenv("MZ_REDUCE_LB2") ? atoi(getenv("MZ_REDUCE_LB2")) : 0;
  static const int env_minb = getenv("MZ_REDUCE_MINB") ? atoi(getenv("MZ_REDUCE_MINB")) : 2;
  // exactly one wave of level-1 threads (kReduceMinBlocks blocks of 128 per SM): a power-of-two chunk length left a
  // second wave 15 % full at c = 22 (ncu, profiles/experiments_r2.md)
  const int minb = env_minb >= 4 ? 4 : env_minb == 3 ? 3 : 2;
  const uint64_t wave = (uint64_t)ctx->sm_count * minb * 128;
  uint32_t Lb = (uint32_t)(((uint64_t)nb * K + wave - 1) / wave);
  if (env_lb > 0) Lb = (uint32_t)env_lb;
  if (Lb > nb) Lb = nb;
  if (Lb < 8) Lb = 1;  // small bucket sets: one level (the second level's latency would exceed what the first saves)
  const uint32_t n1 = Lb > 1 ? (nb + Lb - 1) / Lb : 0;
  // Level 2 over A_1 .. A_{n1-1} (A_0 has weight 0) - or the only level, over the buckets themselves: two blocks of
  // 128 threads per SM, at least 4 values per thread (each thread also pays a ~30-operation double-and-add by its
  // chunk offset; sweep in profiles/experiments_r1.md)
  const uint32_t nb2 = n1 ? n1 - 1 : nb;
  const uint64_t target = (uint64_t)ctx->sm_count * 2 * 128;
  uint32_t Lb2 = (uint32_t)(((uint64_t)nb2 * K + target - 1) / target);
  if (Lb2 < 4) Lb2 = 4;
  if (Lb2 > 256) Lb2 = 256;
  if (env_lb2 > 0) Lb2 = (uint32_t)env_lb2;
  if (Lb2 > nb2) Lb2 = nb2;
  const uint32_t n2 = (nb2 + Lb2 - 1) / Lb2;
  const uint64_t per_set = (uint64_t)n1 + n2;  // values the tree sums per set: S_0 .. S_{n1-1}, then the level-2 partials
  MZ_CUDA_TRY(ctx, ctx->red_a.ensure((size_t)K * per_set * sizeof(XYZZ)));
  MZ_CUDA_TRY(ctx, ctx->red_b.ensure((size_t)K * ((size_t)per_set / (2 * kTreeThreads) + 2) * sizeof(XYZZ)));
  XYZZ* d_s = ctx->red_a.as<XYZZ>();
  if (n1) {
    MZ_CUDA_TRY(ctx, ctx->red_c.ensure((size_t)K * n1 * sizeof(XYZZ)));
    XYZZ* d_a = ctx->red_c.as<XYZZ>();
    const dim3 g1((n1 + 127) / 128, (unsigned)K);
    if (minb == 4)
      msm_bucket_chunks<4><<<g1, 128, 0, ctx->stream>>>(buckets, nb, Lb, d_s, per_set, d_a, n1);
    else if (minb == 3)
      msm_bucket_chunks<3><<<g1, 128, 0, ctx->stream>>>(buckets, nb, Lb, d_s, per_set, d_a, n1);
    else
      msm_bucket_chunks<2><<<g1, 128, 0, ctx->stream>>>(buckets, nb, Lb, d_s, per_set, d_a, n1);
    MZ_LAUNCH_CHECK(ctx);
    msm_bucket_reduce<<<dim3((n2 + 127) / 128, (unsigned)K), 128, 0, ctx->stream>>>(d_a + 1, n1, nb2, Lb2, (int)Lb, d_s + n1,
                                                                                  per_set, n2);
  } else {
    msm_bucket_reduce<<<dim3((n2 + 127) / 128, (unsigned)K), 128, 0, ctx->stream>>>(buckets, nb, nb, Lb2, 1, d_s, per_set, n2);
  }
  MZ_LAUNCH_CHECK(ctx);
  MZ_TRY(tree_sum(ctx, d_s, ctx->red_b.as<XYZZ>(), per_set, d_out, (uint32_t)K));
  if (timing) MZ_CUDA_TRY(ctx, cudaEventRecord(ctx->phase_ev[slot][0][5], ctx->stream));
  ctx->phase_valid[slot] = timing;
  ctx->phase_pending = false;
  ctx->chunk_idx = 0;
  ctx->msm_count++;
  return MYZKP_OK;
}

// ---------------------------------------------------------------------------
// MSM over caller-supplied points, without a table of multiples (accumulate_curve_points,
// zksnark/utils.rs:83-93): classic windowed Pippenger.  Window w has its own bucket range, the pipeline
// above produces S_w = sum_k k B_{w,k} for every window, and out = sum_w 2^(c w) S_w by Horner from the
// top window (c doublings + one addition per window; ~254 dependent doublings, about a millisecond of
// latency, independent of n).
// ---------------------------------------------------------------------------
__global__ void msm_window_combine(const XYZZ* __restrict__ sums, int W, int c, XYZZ* __restrict__ out) {
  XYZZ acc = load_xyzz(sums + (W - 1));
  for (int w = W - 2; w >= 0; w--) {
    for (int k = 0; k < c; k++) xyzz_dbl(acc);
    XYZZ sw = load_xyzz(sums + w);
    xyzz_add(acc, sw);
  }
  store_xyzz(out, acc);
}

int msm_points_xyzz(myzkp_ctx* ctx, const uint32_t* d_scalars, const Affine* d_points_mont, size_t n, XYZZ* d_out) {
  if (n == 0) {
    xyzz_set_inf<<<1, 1, 0, ctx->stream>>>(d_out);
    MZ_LAUNCH_CHECK(ctx);
    return MYZKP_OK;
  }
  // window by work: W(c) n mixed additions (10 multiplies) + W(c) 2^(c-1) buckets x 2 additions (28)
  int c = 8;
  {
    double best = 0;
    for (int cc = 4; cc <= 20; cc++) {
      const int Wc = (255 + cc - 1) / cc;
      const double cost = 10.0 * Wc * (double)n + 28.0 * Wc * (double)((size_t)1 << (cc - 1));
      if (best == 0 || cost < best) { best = cost; c = cc; }
    }
  }
  const int W = (255 + c - 1) / c;
  const size_t nbw = (size_t)W << (c - 1);
  // the pipeline reads points through ctx->table / srs_n: point it at the caller's points for this call
  Affine* saved_table = ctx->table;
  const size_t saved_n = ctx->srs_n;
  ctx->table = const_cast<Affine*>(d_points_mont);
  ctx->srs_n = n;
  int rc = [&]() -> int {
    MZ_CUDA_TRY(ctx, ctx->buckets.ensure(nbw * sizeof(XYZZ)));
    MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure((size_t)(W + 2) * sizeof(XYZZ)));
    MsmItem one{d_scalars, n};
    MZ_TRY(msm_fill_buckets_batch(ctx, &one, 1, 0, c, ctx->buckets.as<XYZZ>(), /*per_window=*/true));
    MZ_TRY(msm_reduce_buckets(ctx, c, ctx->buckets.as<XYZZ>(), ctx->xyzz_tmp.as<XYZZ>(), (size_t)W));
    msm_window_combine<<<1, 1, 0, ctx->stream>>>(ctx->xyzz_tmp.as<XYZZ>(), W, c, d_out);
    MZ_LAUNCH_CHECK(ctx);
    return MYZKP_OK;
  }();
  ctx->table = saved_table;
  ctx->srs_n = saved_n;
  return rc;
}

int xyzz_to_bytes(myzkp_ctx* ctx, const XYZZ* d_in, size_t count, uint8_t* d_out64) {
  if (count == 0) return MYZKP_OK;
  xyzz_to_affine_bytes<<<(unsigned)((count + 31) / 32), 32, 0, ctx->stream>>>(d_in, count,
                                                                              reinterpret_cast<uint32_t*>(d_out64));
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

int sum_partials(myzkp_ctx* ctx, const XYZZ* d_partials, size_t k, uint8_t* d_out64) {
  MZ_CUDA_TRY(ctx, ctx->red_a.ensure((k + 2) * sizeof(XYZZ)));
  MZ_CUDA_TRY(ctx, ctx->red_b.ensure((k / (2 * kTreeThreads) + 2) * sizeof(XYZZ)));
  MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
  XYZZ* res = reinterpret_cast<XYZZ*>(ctx->small.as<uint8_t>() + 1024);
  if (k == 0) {
    xyzz_set_inf<<<1, 1, 0, ctx->stream>>>(res);
    MZ_LAUNCH_CHECK(ctx);
  } else {
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->red_a.p, d_partials, k * sizeof(XYZZ), cudaMemcpyDeviceToDevice, ctx->stream));
    MZ_TRY(tree_sum(ctx, ctx->red_a.as<XYZZ>(), ctx->red_b.as<XYZZ>(), k, res));
  }
  return xyzz_to_bytes(ctx, res, 1, d_out64);
}

}  // namespace mz
