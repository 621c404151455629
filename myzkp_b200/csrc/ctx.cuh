// Host-side context shared by the translation units of libmyzkp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/myzkp_b200.h"
#include "g1.cuh"

namespace mz {

// Rows of the resident SRS table: row j holds 2^(b_j) * P_i for every SRS point, for a
// set of bit offsets b_j, so that an MSM with window c needs no doublings at all:
// sum_w 2^(c w) d_w P = sum_w d_w * row[bit = c w].  The offsets are the union of
// {c w} over the supported windows: multiples of 4 (c in {4,8,...,24}) plus the
// multiples of 22 (c = 22, the best window around 2^22..2^24 points); a very large SRS
// falls back to multiples of 8 (c in {8,16,24}).
constexpr int kMaxTableRows = 80;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  // When set and true, the buffer must not grow: cudaMalloc / cudaFree synchronise the whole device, and with
  // several ranks on ONE device (in-process tests) that would wait for a peer's exchange kernel which is itself
  // spinning for this rank - see myzkp_ctx_reserve.
  const bool* frozen = nullptr;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (frozen && *frozen) return cudaErrorNotPermitted;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    // grow geometrically so repeated slightly-larger calls do not thrash
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace mz

struct myzkp_ctx {
  // Child contexts: own stream and scratch, SRS table shared with (and owned by) the parent.
  // Used to run many small, latency-bound MSMs concurrently (Gemini's low levels).
  std::vector<myzkp_ctx*> children;
  bool is_child = false;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  // sort / accumulate pipeline of one large MSM (msm.cu: msm_xyzz): chunk k+1 is recoded and sorted on a child
  // stream while chunk k is accumulated on this one
  cudaEvent_t pipe_sorted_ev[4] = {}, pipe_acc_ev[4] = {};
  int pipe_chunks = 0;  // 0 = automatic, 1 = off

  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;

  // tuning (0 = auto)
  int window_bits = 0;
  int segment_len = 0;

  // resident SRS table: table_rows rows of srs_n affine points (Montgomery)
  mz::Affine* table = nullptr;
  size_t srs_n = 0;
  int table_rows = 0;
  int row_bits[mz::kMaxTableRows] = {};   // sorted bit offsets of the rows
  uint8_t row_of_bit[256] = {};           // bit offset -> row (0xff = absent)
  uint32_t windows = 0;                   // bit c set <=> window c is supported
  uint32_t table_windows = 0;             // caller-chosen window set for the next SRS (0 = automatic)
  uint8_t* d_row_of_bit = nullptr;        // device copies
  int* d_row_bits = nullptr;

  // fixed-base comb table of G for srs_generate: [32][256] affine
  mz::Affine* gcomb = nullptr;

  // scratch (grow-only)
  mz::DevBuf scalars;      // staged scalars / coefficients (n * 32 B)
  mz::DevBuf scalars2;     // quotient / folded coefficients
  mz::DevBuf keys_a, keys_b, vals_a, vals_b, sort_tmp;
  mz::DevBuf sort_parts;   // partition plan of the MSD split (msm.cu, sort.cu)
  mz::DevBuf sort_groups;  // group starts / oversize-group tile table of the MSD sort (sort.cu)
  mz::DevBuf caller_points;  // Montgomery copies of caller-supplied points (myzkp_g1_msm)
  mz::DevBuf buckets;      // XYZZ per bucket
  mz::DevBuf heads, head_keys;
  mz::DevBuf heads2;       // ping-pong levels of the head merge
  mz::DevBuf baa_pts, baa_keys, baa_prefix, baa_meta, baa_trans;  // batched-affine rounds: private lists, prefixes, products
  int baa_rounds = -1;     // -1 = automatic, 0 = off (XYZZ accumulate only)
  mz::DevBuf red_a, red_b, red_c; // reduction partials (XYZZ)
  mz::DevBuf poly_tiles;   // per-tile (mult, add) maps for the quotient scan
  mz::DevBuf small;        // misc small device outputs (flags, y, points)
  bool small_init = false; // the sticky non-canonical flag inside `small` has been zeroed once
  mz::DevBuf descs;        // per-polynomial descriptors of a batched MSM
  mz::DevBuf xyzz_tmp;     // XYZZ temporaries (SRS generation)
  // host-API upload pipeline: chunks of a large polynomial are copied on copy_stream
  // while earlier chunks are already being committed on `stream`
  int upload_chunks = 0;  // 0 = automatic
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[8] = {};
  cudaEvent_t copy_done_ev = nullptr;

  // peer-memory exchange of the range-sharded multi-GPU path (peer.cu): one small buffer per
  // rank, mapped into every peer (CUDA IPC between processes, plain pointers inside one)
  static constexpr int kMaxPeers = 16;
  uint8_t* peer_local = nullptr;
  uint8_t* peer_bufs[kMaxPeers] = {};
  bool peer_ipc[kMaxPeers] = {};
  int peer_rank = -1, peer_world = 0;
  bool peer_same_device = false;  // some attached peer shares this device: scratch is frozen (DevBuf::frozen)
  uint32_t peer_epoch = 0;
  unsigned long long peer_timeout_ns = 10ull * 1000 * 1000 * 1000;

  // optional per-phase CUDA-event timing of the last MSM (bench.py roofline)
  // phases: 0 recode, 1 sort, 2 accumulate, 3 merge heads, 4 bucket reduce + tree sum
  // a ring of kPhaseSlots MSM calls so a timed loop can be read back after its final sync
  // An MSM may run as several chunks (upload pipeline, sort / accumulate pipeline): events per chunk
  // 0 start, 1 after recode, 2 after sort (these three on the stream that sorts - a child's in the pipelined form),
  // 6 before accumulate, 3 after accumulate, 4 after merge (this stream); 5 = after the bucket reduce (chunk 0 only).
  static constexpr int kPhaseSlots = 32;
  static constexpr int kMaxPipe = 4;   // chunks with their own events (and the longest sort / accumulate pipeline)
  bool phase_timing = false;
  cudaEvent_t phase_ev[kPhaseSlots][kMaxPipe][7] = {};
  bool phase_valid[kPhaseSlots] = {};
  int phase_chunks[kPhaseSlots] = {};  // chunks recorded in the slot
  int chunk_idx = 0;                   // chunks filled since the last bucket reduce
  bool phase_pending = false;  // fill_buckets recorded the events of the current slot
  uint64_t msm_count = 0;  // MSMs run so far (slot = count % kPhaseSlots)
  // facts about each MSM: window bits, windows, entries, segment length, segments, buckets
  uint64_t msm_info[kPhaseSlots][6] = {};
};

namespace mz {
template <class F>
inline void for_each_scratch(myzkp_ctx* ctx, F f) {
  DevBuf* bufs[] = {&ctx->scalars, &ctx->scalars2, &ctx->keys_a, &ctx->keys_b, &ctx->vals_a, &ctx->vals_b,
                    &ctx->sort_tmp, &ctx->sort_parts, &ctx->sort_groups, &ctx->caller_points, &ctx->buckets, &ctx->heads, &ctx->head_keys, &ctx->heads2,
                    &ctx->baa_pts, &ctx->baa_keys, &ctx->baa_prefix, &ctx->baa_meta, &ctx->baa_trans,
                    &ctx->red_a, &ctx->red_b, &ctx->red_c, &ctx->poly_tiles, &ctx->small, &ctx->xyzz_tmp, &ctx->descs};
  for (DevBuf* b : bufs) f(b);
}
}  // namespace mz

#define MZ_CUDA_TRY(ctx, expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
      if (_e == cudaErrorNotPermitted)                                                \
        (ctx)->err += " - scratch would have to grow while peers on the same device are attached "  \
                      "(a device-wide synchronisation could deadlock their exchange): call myzkp_ctx_reserve first"; \
      return (_e == cudaErrorMemoryAllocation) ? MYZKP_ERR_OOM : MYZKP_ERR_CUDA;      \
    }                                                                                 \
  } while (0)

#define MZ_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != MYZKP_OK) return _r; \
  } while (0)

#define MZ_LAUNCH_CHECK(ctx)                    \
  do {                                          \
    (ctx)->launches++;                          \
    MZ_CUDA_TRY(ctx, cudaGetLastError());       \
  } while (0)

namespace mz {

inline int fail(myzkp_ctx* ctx, int code, const char* msg) {
  ctx->err = msg;
  return code;
}

// ---- msm.cu ----
// MSM of n canonical scalars (device, 32 B LE each) against SRS points
// [srs_off, srs_off + n); result XYZZ (Montgomery) written to d_out (device).
int msm_xyzz(myzkp_ctx* ctx, const uint32_t* d_scalars, size_t n, size_t srs_off, XYZZ* d_out);
// the two halves of msm_xyzz, for callers that accumulate several scalar chunks into bucket sets of
// the same window c before one final reduce (upload pipeline in capi.cu)
int msm_pick_window(const myzkp_ctx* ctx, size_t n);
// onto: add to what `buckets` already holds (the sums of earlier chunks) instead of overwriting it
int msm_fill_buckets(myzkp_ctx* ctx, const uint32_t* d_scalars, size_t n, size_t srs_off, int c, XYZZ* buckets,
                     bool onto = false);
int msm_add_buckets(myzkp_ctx* ctx, XYZZ* a, const XYZZ* b, int c);
int msm_reduce_buckets(myzkp_ctx* ctx, int c, const XYZZ* buckets, XYZZ* d_out, size_t K = 1);
// K polynomials as ONE pipeline (shared sort / accumulate / merge, bucket range y per polynomial):
// d_out[y] = sum_i items[y].d_scalars[i] * SRS[srs_off + i]
struct MsmItem {
  const uint32_t* d_scalars;
  size_t n;
};
int msm_pick_window_batch(const myzkp_ctx* ctx, const MsmItem* items, size_t K);
// The two halves of msm_fill_buckets_batch.  Steps 1-2 (recode, sort) run on `actx` - its stream and scratch; it may
// be a child of `tctx`, which owns the timing slot and the non-canonical flag - and leave the sorted entries in
// actx's scratch; steps 3-4 (accumulate, head merge) run on ctx's stream into `buckets`.
struct SortedEntries {
  const uint32_t* keys = nullptr;
  const uint32_t* vals = nullptr;
  uint64_t M = 0;   // entries
  uint32_t nb = 0;  // buckets of the whole batch (= the sentinel key)
  int c = 0, W = 0;
};
int msm_sort_entries(myzkp_ctx* actx, myzkp_ctx* tctx, int chunk, const MsmItem* items, size_t K, size_t srs_off, int c,
                     bool per_window, SortedEntries* out);
int msm_accumulate_sorted(myzkp_ctx* ctx, int chunk, const SortedEntries& se, XYZZ* buckets, bool onto);
// segments (= heads) the accumulate of n scalars at window c will use; the merge of T heads into the buckets
uint64_t msm_segments_for(const myzkp_ctx* ctx, int c, size_t n);
int msm_merge_heads(myzkp_ctx* ctx, XYZZ* buckets, uint32_t nb, uint64_t T);
// chunk `pos` of n scalars cut into K chunks whose sizes grow by `ratio` (1 = equal chunks)
void msm_chunk_range(size_t n, int K, int pos, double ratio, size_t* lo, size_t* hi);
// child i of ctx (own stream and scratch), (re)pointed at the parent's current SRS table; children 0-1 serve the
// batched Gemini levels, kPipeChild0 and kPipeChild0 + 1 the sort / accumulate pipeline (high-priority streams)
constexpr int kPipeChild0 = 2;
int get_child(myzkp_ctx* ctx, int i, myzkp_ctx** out);
int msm_fill_buckets_batch(myzkp_ctx* ctx, const MsmItem* items, size_t K, size_t srs_off, int c, XYZZ* buckets,
                           bool per_window = false, bool onto = false);
// sum_i scalars[i] * points[i] for caller-supplied points (Montgomery affine on the device), no table of
// multiples: per-window buckets + Horner over the windows
int msm_points_xyzz(myzkp_ctx* ctx, const uint32_t* d_scalars, const Affine* d_points_mont, size_t n, XYZZ* d_out);
// canonical 64-byte affine points (device) -> Montgomery affine; ORs 1 into *d_flag for a coordinate >= p
int points_import(myzkp_ctx* ctx, const uint32_t* d_in, size_t n, Affine* d_out, int* d_flag);
int msm_batch_xyzz(myzkp_ctx* ctx, const MsmItem* items, size_t K, size_t srs_off, XYZZ* d_out);
// XYZZ (device) -> canonical affine 64 B (device)
int xyzz_to_bytes(myzkp_ctx* ctx, const XYZZ* d_in, size_t count, uint8_t* d_out64);
int sum_partials(myzkp_ctx* ctx, const XYZZ* d_partials, size_t k, uint8_t* d_out64);
// peer.cu: fused exchange (+ sum / carry composition) over peer memory.  mode 0: XYZZ partials ->
// affine bytes in d_out0 (and rank 0's 32 B tail in d_out1); mode 1: (h, u^n) pairs -> carry in d_out0
int peer_exchange(myzkp_ctx* ctx, int mode, const void* d_payload, int bytes, void* d_out0, void* d_out1);
int peer_check(myzkp_ctx* ctx);
void peer_release(myzkp_ctx* ctx);

// ---- baa.cu ----
int baa_accumulate(myzkp_ctx* ctx, const uint32_t* keys_s, const uint32_t* vals_s, uint64_t M, uint32_t L,
                   uint32_t sentinel, int rounds, XYZZ* buckets, XYZZ* heads, uint32_t* head_keys, uint64_t T);

// ---- sort.cu ----
// LSD radix sort of (key, val) pairs by the low `bits` key bits; result in (*out_keys, *out_vals)
int radix_sort_pairs(myzkp_ctx* ctx, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint64_t n,
                     int bits, uint32_t** out_keys, uint32_t** out_vals, const uint32_t* d_parts = nullptr, int P = 0,
                     bool first_pass_unordered = false);

// MSD form for a list partitioned by the key bits above 8 + r (msm.cu's recode): 256-way pass + group-local sort
int radix_sort_pairs_msd(myzkp_ctx* ctx, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint64_t n,
                         int r, uint32_t** out_keys, uint32_t** out_vals, const uint32_t* d_parts, int P);
void sort_set_group_cap(int cap);  // test hook: entries a group may have for the shared-memory sort (0 = default)
int sort_group_cap();

// ---- srs.cu ----
int srs_alloc(myzkp_ctx* ctx, size_t n);
int srs_build_from_row0(myzkp_ctx* ctx);  // fills rows 1.. from row 0 (Montgomery affine)

// ---- poly.cu ----
// (h, u^n) of a coefficient range; both canonical 32 B written to device
int fr_range_eval(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32],
                  uint32_t* d_h, uint32_t* d_upow, int* d_flag = nullptr);
// with c_i = f_{lo+i} + u c_{i+1} and c_n = carry: d_q[i] = c_{i+1} (= q_{lo+i}), *d_c0 = c_0
// (carry_le == NULL means 0); d_flag (optional, both): set to 1 when a coefficient is not canonical
int fr_range_quotient(myzkp_ctx* ctx, const uint32_t* d_coefs, size_t n, const uint8_t u_le[32],
                      const uint8_t carry_le[32], uint32_t* d_q, uint32_t* d_c0, const uint32_t* d_carry = nullptr,
                      int* d_flag = nullptr);
int fr_fold(myzkp_ctx* ctx, const uint32_t* d_in, size_t n_out, const uint32_t* d_rho, uint32_t* d_out);
int fr_check_canonical(myzkp_ctx* ctx, const uint32_t* d_in, size_t n, int* d_flag);

}  // namespace mz
