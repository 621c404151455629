// extern "C" entry points of libmyzkp_b200.so (see include/myzkp_b200.h).
#include <string.h>

#include <cmath>
#include <vector>

#include "ctx.cuh"

using namespace mz;

namespace {

// canonical little-endian moduli for host-side checks of the 32-byte scalars
const uint8_t kRModLE[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                             0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

bool fr_bytes_canonical(const uint8_t* b) {
  for (int i = 31; i >= 0; i--) {
    if (b[i] < kRModLE[i]) return true;
    if (b[i] > kRModLE[i]) return false;
  }
  return false;
}

// layout of ctx->small (4 KiB)
constexpr size_t kSmallFlag = 512;    // int: non-canonical input seen
constexpr size_t kSmallY = 640;       // 32 B: y / c0
constexpr size_t kSmallXyzz = 1024;   // XYZZ result(s): up to 16 slots (2 KiB)
constexpr size_t kSmallPoint = 3072;  // 64 B affine bytes result

// The non-canonical flag is STICKY: kernels only ever OR into it, begin_call does not clear it, and it is read
// (and then cleared) by whichever comes first - the end of a host-buffer call or myzkp_ctx_sync.  So an
// asynchronous device-pointer call with a scalar >= r is reported by the next synchronising call.
int begin_call(myzkp_ctx* ctx) {
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!ctx->small_init) {
    MZ_CUDA_TRY(ctx, ctx->small.ensure(4096));
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(ctx->small.as<uint8_t>() + kSmallFlag, 0, sizeof(int), ctx->stream));
    ctx->small_init = true;
  }
  return MYZKP_OK;
}

// Upload pipeline for the host-buffer entry points: how many chunks to split n into.
// Alone on its host a GPU uploads at ~54 GB/s, a chunk's MSM then takes about four times as long as its upload, and
// chunk sizes growing 4x expose only the small first upload.  With 4 or more ranks attached (one process per GPU of
// the same host) the ranks' uploads share the host's memory and PCIe root - measured at 8 ranks: ~23 GB/s each - so
// an upload takes about half as long as its chunk's MSM and EQUAL chunks expose least.
bool uploads_contended(const myzkp_ctx* ctx) { return ctx->peer_world >= 4; }
int upload_chunks(const myzkp_ctx* ctx, size_t n) {
  if (ctx->upload_chunks > 0) return (size_t)ctx->upload_chunks <= (n ? n : 1) ? ctx->upload_chunks : 1;
  const char* e = getenv("MZ_UPLOAD_CHUNKS");  // experiment knob
  if (e && atoi(e) > 0) return (size_t)atoi(e) <= (n ? n : 1) ? atoi(e) : 1;
  if (n >= ((size_t)1 << 24)) return 4;  // 2^24: 45.2 / 38.5 / 37.2 / 37.0 ms for 1 / 2 / 3 / 4 chunks (resident 35.5)
  if (n >= ((size_t)1 << 21)) return uploads_contended(ctx) ? 4 : 2;  // 2^21 alone: 6.75 / 6.04 / 6.25 / 6.52 ms (resident 5.51)
  return 1;
}
// Chunk `pos` (in processing order) of n coefficients cut into K chunks whose sizes grow by `ratio` (4, or 1 = equal
// chunks).  Descending: the first chunk processed is the top of the polynomial (the quotient scan runs downwards).
// growth ratio of the upload chunks: 4 alone on the host, kContendedRatio when the ranks' uploads share it
constexpr double kContendedRatio = 1.0;
double upload_ratio(const myzkp_ctx* ctx) {
  static const double env = getenv("MZ_UPLOAD_RATIO") ? atof(getenv("MZ_UPLOAD_RATIO")) : 0.0;  // experiment knob
  if (env >= 1.0) return env;
  if (getenv("MZ_UPLOAD_EQUAL")) return 1.0;
  return uploads_contended(ctx) ? kContendedRatio : 4.0;
}
void chunk_range(const myzkp_ctx* ctx, size_t n, int K, int pos, bool descending, size_t* lo, size_t* hi) {
  size_t a, b;
  msm_chunk_range(n, K, pos, upload_ratio(ctx), &a, &b);
  if (descending) {
    *lo = n - b;
    *hi = n - a;
  } else {
    *lo = a;
    *hi = b;
  }
}
int ensure_copy_stream(myzkp_ctx* ctx) {
  if (ctx->copy_stream) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 8; i++) MZ_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming));
  MZ_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->copy_done_ev, cudaEventDisableTiming));
  return MYZKP_OK;
}
// enqueue the H2D copies of K contiguous chunks on the copy stream (chunk order given by `descending`)
int enqueue_chunk_uploads(myzkp_ctx* ctx, const uint8_t* host, uint8_t* dev, size_t n, int K, bool descending) {
  MZ_TRY(ensure_copy_stream(ctx));
  // the copy stream must not overwrite the staging buffer while earlier work on `stream` still reads it
  MZ_CUDA_TRY(ctx, cudaEventRecord(ctx->copy_done_ev, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done_ev, 0));
  for (int pos = 0; pos < K; pos++) {
    size_t lo, hi;
    chunk_range(ctx, n, K, pos, descending, &lo, &hi);
    if (lo < hi)
      MZ_CUDA_TRY(ctx, cudaMemcpyAsync(dev + lo * 32, host + lo * 32, (hi - lo) * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
    MZ_CUDA_TRY(ctx, cudaEventRecord(ctx->copy_ev[pos], ctx->copy_stream));
  }
  return MYZKP_OK;
}

int end_call_check_flag(myzkp_ctx* ctx) {
  int h_flag = 0;
  if (ctx->small_init)
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(&h_flag, ctx->small.as<uint8_t>() + kSmallFlag, sizeof(int), cudaMemcpyDeviceToHost,
                                     ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flag) {
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(ctx->small.as<uint8_t>() + kSmallFlag, 0, sizeof(int), ctx->stream));
    return fail(ctx, MYZKP_ERR_NONCANONICAL, "input scalar >= r (callers must sanitize, field.rs:260-270)");
  }
  return MYZKP_OK;
}

constexpr int kMaxChildren = 8;

void free_ctx_scratch(myzkp_ctx* ctx) {
  for_each_scratch(ctx, [](DevBuf* b) { b->release(); });
}

}  // namespace

// child i of ctx, (re)pointed at the parent's current SRS table
int mz::get_child(myzkp_ctx* ctx, int i, myzkp_ctx** out) {
  if (i < 0 || i >= kMaxChildren) return fail(ctx, MYZKP_ERR_INVALID_ARG, "no such child context");
  while ((int)ctx->children.size() <= i) {
    myzkp_ctx* c = new myzkp_ctx();
    c->is_child = true;
    c->device = ctx->device;
    c->sm_count = ctx->sm_count;
    // the pipeline's sort streams get the highest priority: their short kernels must slip in between the blocks of
    // the accumulate kernel that runs beside them
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio = (int)ctx->children.size() >= kPipeChild0 ? prio_hi : prio_lo;
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming) != cudaSuccess) {
      delete c;
      return fail(ctx, MYZKP_ERR_CUDA, "cannot create a child stream");
    }
    c->own_stream = true;
    ctx->children.push_back(c);
  }
  myzkp_ctx* c = ctx->children[i];
  c->table = ctx->table;
  c->srs_n = ctx->srs_n;
  c->table_rows = ctx->table_rows;
  c->windows = ctx->windows;
  memcpy(c->row_bits, ctx->row_bits, sizeof c->row_bits);
  memcpy(c->row_of_bit, ctx->row_of_bit, sizeof c->row_of_bit);
  c->d_row_of_bit = ctx->d_row_of_bit;
  c->d_row_bits = ctx->d_row_bits;
  c->window_bits = ctx->window_bits;
  c->segment_len = ctx->segment_len;
  c->baa_rounds = ctx->baa_rounds;
  *out = c;
  return MYZKP_OK;
}

namespace {

// MSM of n scalars that arrive in K upload chunks (events ctx->copy_ev[k]).  With u_le != NULL the
// scalars are the quotient of the uploaded polynomial by (x - u): chunks are then consumed top first,
// the scan carry stays on the device, and y is left at kSmallY.
int chunked_msm(myzkp_ctx* ctx, uint32_t* d_coefs, size_t n, int K, bool descending, const uint8_t* u_le,
                uint32_t* d_quot, XYZZ* d_res, size_t srs_off = 0) {
  uint8_t* s = ctx->small.as<uint8_t>();
  uint32_t* c0 = reinterpret_cast<uint32_t*>(s + kSmallY);  // carry chain, ends as y
  uint32_t* c0_prev = reinterpret_cast<uint32_t*>(s + kSmallY + 32);
  int* flag = reinterpret_cast<int*>(s + kSmallFlag);
  const size_t n_msm = u_le ? n - 1 : n;  // the quotient has one coefficient less
  const int c = msm_pick_window(ctx, n_msm);
  const size_t nb = (size_t)1 << (c - 1);
  MZ_CUDA_TRY(ctx, ctx->buckets.ensure(nb * sizeof(XYZZ)));
  XYZZ* b0 = ctx->buckets.as<XYZZ>();
  bool first = true;
  for (int pos = 0; pos < K; pos++) {
    size_t lo, hi;
    chunk_range(ctx, n, K, pos, descending, &lo, &hi);
    MZ_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[pos], 0));
    if (lo >= hi) continue;
    const uint32_t* sc = d_coefs + lo * 8;
    size_t len = hi - lo;
    if (u_le) {
      const bool top = (hi == n);
      if (!top) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(c0_prev, c0, 32, cudaMemcpyDeviceToDevice, ctx->stream));
      // the scan checks canonicity of what it reads (no separate pass over the coefficients)
      MZ_TRY(fr_range_quotient(ctx, sc, len, u_le, nullptr, d_quot + lo * 8, c0, top ? nullptr : c0_prev, flag));
      // q[i] = q_{lo+i} pairs with SRS point lo+i; the global top coefficient q_{n-1} is the zero carry
      sc = d_quot + lo * 8;
      if (top) len -= 1;
    }
    if (len == 0) continue;
    // later chunks accumulate onto the same bucket set (msm_accumulate<kOnto>): no second set, no dense addition
    MZ_TRY(msm_fill_buckets(ctx, sc, len, srs_off + lo, c, b0, /*onto=*/!first));
    first = false;
  }
  if (first) MZ_CUDA_TRY(ctx, cudaMemsetAsync(b0, 0, nb * sizeof(XYZZ), ctx->stream));
  return msm_reduce_buckets(ctx, c, b0, d_res);
}

}  // namespace

extern "C" {

int myzkp_ctx_create(myzkp_ctx** out, int device_id) {
  if (!out) return MYZKP_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device_id < 0 || device_id >= count) return MYZKP_ERR_CUDA;
  if (cudaSetDevice(device_id) != cudaSuccess) return MYZKP_ERR_CUDA;
  myzkp_ctx* ctx = new myzkp_ctx();
  ctx->device = device_id;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  // the MSM gathers 64-byte points at random: do not let L2 over-fetch whole 128-byte lines
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 64);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return MYZKP_ERR_CUDA;
  }
  ctx->own_stream = true;
  for_each_scratch(ctx, [ctx](DevBuf* b) { b->frozen = &ctx->peer_same_device; });
  *out = ctx;
  return MYZKP_OK;
}

int myzkp_ctx_reserve(myzkp_ctx* ctx, size_t n_max) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  MZ_TRY(begin_call(ctx));
  // Run the pipelines once on all-zero coefficients: buffer sizes depend on the length only (window, entries,
  // segments, tiles), so this sizes every scratch buffer exactly as a real commit / open of n_max would.
  const bool was_frozen = ctx->peer_same_device;
  ctx->peer_same_device = false;
  int rc = MYZKP_OK;
  auto run = [&]() -> int {
    const size_t n = n_max ? n_max : 1;
    MZ_CUDA_TRY(ctx, ctx->scalars.ensure(2 * n * 32 + 32 * 66));  // Gemini keeps all levels and the challenges here
    MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 32));
    MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure(66 * (sizeof(XYZZ) + 64)));
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(ctx->scalars.p, 0, n * 32, ctx->stream));
    uint8_t* s = ctx->small.as<uint8_t>();
    const uint8_t zero[32] = {0};
    const size_t n_msm = n_max < ctx->srs_n ? n_max : ctx->srs_n;
    if (ctx->table && n_msm) {
      MZ_CUDA_TRY(ctx, ctx->buckets.ensure(((size_t)1 << (msm_pick_window(ctx, n_msm) - 1)) * sizeof(XYZZ)));
      MZ_TRY(msm_xyzz(ctx, ctx->scalars.as<uint32_t>(), n_msm, 0, reinterpret_cast<XYZZ*>(s + kSmallXyzz)));
    }
    if (n_max) {
      MZ_TRY(fr_range_eval(ctx, ctx->scalars.as<uint32_t>(), n_max, zero, reinterpret_cast<uint32_t*>(s + kSmallY),
                           reinterpret_cast<uint32_t*>(s + kSmallY + 32)));
      MZ_TRY(fr_range_quotient(ctx, ctx->scalars.as<uint32_t>(), n_max, zero, nullptr, ctx->scalars2.as<uint32_t>(),
                               reinterpret_cast<uint32_t*>(s + kSmallY)));
    }
    MZ_TRY(ensure_copy_stream(ctx));
    MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return MYZKP_OK;
  };
  rc = run();
  ctx->peer_same_device = was_frozen;
  return rc;
}

int myzkp_ctx_destroy(myzkp_ctx* ctx) {
  if (!ctx) return MYZKP_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->table) cudaFree(ctx->table);
  if (ctx->gcomb) cudaFree(ctx->gcomb);
  if (ctx->d_row_of_bit) cudaFree(ctx->d_row_of_bit);
  if (ctx->d_row_bits) cudaFree(ctx->d_row_bits);
  for (myzkp_ctx* c : ctx->children) {
    cudaStreamSynchronize(c->stream);
    free_ctx_scratch(c);
    if (c->join_ev) cudaEventDestroy(c->join_ev);
    cudaStreamDestroy(c->stream);
    delete c;
  }
  ctx->children.clear();
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  peer_release(ctx);
  free_ctx_scratch(ctx);
  for (int s = 0; s < myzkp_ctx::kPhaseSlots; s++)
    for (int k = 0; k < myzkp_ctx::kMaxPipe; k++)
      for (int i = 0; i < 7; i++)
        if (ctx->phase_ev[s][k][i]) cudaEventDestroy(ctx->phase_ev[s][k][i]);
  for (int k = 0; k < 4; k++) {
    if (ctx->pipe_sorted_ev[k]) cudaEventDestroy(ctx->pipe_sorted_ev[k]);
    if (ctx->pipe_acc_ev[k]) cudaEventDestroy(ctx->pipe_acc_ev[k]);
  }
  for (int i = 0; i < 8; i++)
    if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
  if (ctx->copy_done_ev) cudaEventDestroy(ctx->copy_done_ev);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return MYZKP_OK;
}

int myzkp_ctx_set_stream(myzkp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return MYZKP_OK;
}

int myzkp_ctx_sync(myzkp_ctx* ctx) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // reports (and clears) the sticky non-canonical flag of the asynchronous device-pointer calls
  MZ_TRY(end_call_check_flag(ctx));
  return peer_check(ctx);
}

const char* myzkp_last_error(const myzkp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
uint64_t myzkp_kernel_launches(const myzkp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int myzkp_ctx_set_msm_params(myzkp_ctx* ctx, int window_bits, int segment_len) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  if (window_bits < 0 || window_bits > 24)
    return fail(ctx, MYZKP_ERR_INVALID_ARG, "window_bits must be 0 (auto) or one of the supported windows (<= 24)");
  if (segment_len < 0 || segment_len > 65536) return fail(ctx, MYZKP_ERR_INVALID_ARG, "bad segment_len");
  ctx->window_bits = window_bits;
  ctx->segment_len = segment_len;
  return MYZKP_OK;
}

int myzkp_ctx_set_baa_rounds(myzkp_ctx* ctx, int rounds) {
  if (!ctx || rounds < -2 || rounds > 16) return MYZKP_ERR_INVALID_ARG;
  ctx->baa_rounds = rounds;
  return MYZKP_OK;
}

int myzkp_ctx_set_upload_chunks(myzkp_ctx* ctx, int chunks) {
  if (!ctx || chunks < 0 || chunks > 8) return MYZKP_ERR_INVALID_ARG;
  ctx->upload_chunks = chunks;
  return MYZKP_OK;
}

int myzkp_ctx_enable_phase_timing(myzkp_ctx* ctx, int on) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ctx->phase_timing = on != 0;
  if (on && !ctx->phase_ev[0][0][0])
    for (int s = 0; s < myzkp_ctx::kPhaseSlots; s++)
      for (int k = 0; k < myzkp_ctx::kMaxPipe; k++)
        for (int i = 0; i < 7; i++) MZ_CUDA_TRY(ctx, cudaEventCreate(&ctx->phase_ev[s][k][i]));
  return MYZKP_OK;
}

int myzkp_ctx_msm_phases(myzkp_ctx* ctx, int back, float out_ms[5], uint64_t out_info[6]) {
  if (!ctx || !out_ms || !out_info || back < 0 || back >= myzkp_ctx::kPhaseSlots) return MYZKP_ERR_INVALID_ARG;
  if ((uint64_t)back >= ctx->msm_count) return fail(ctx, MYZKP_ERR_INVALID_ARG, "no such MSM recorded");
  const int slot = (int)((ctx->msm_count - 1 - back) % myzkp_ctx::kPhaseSlots);
  for (int i = 0; i < 6; i++) out_info[i] = ctx->msm_info[slot][i];
  for (int i = 0; i < 5; i++) out_ms[i] = -1.f;
  if (!ctx->phase_valid[slot] || ctx->phase_chunks[slot] < 1) return MYZKP_OK;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  auto& ev = ctx->phase_ev[slot];
  MZ_CUDA_TRY(ctx, cudaEventSynchronize(ev[0][5]));
  // sums over the chunks of the MSM: recode, sort (on the stream that sorted), accumulate, head merge; then the reduce
  const int pairs[4][2] = {{0, 1}, {1, 2}, {6, 3}, {3, 4}};
  const int chunks = ctx->phase_chunks[slot];
  for (int i = 0; i < 4; i++) out_ms[i] = 0.f;
  for (int k = 0; k < chunks; k++)
    for (int i = 0; i < 4; i++) {
      float ms = 0.f;
      MZ_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ev[k][pairs[i][0]], ev[k][pairs[i][1]]));
      out_ms[i] += ms;
    }
  MZ_CUDA_TRY(ctx, cudaEventElapsedTime(&out_ms[4], ev[chunks - 1][4], ev[0][5]));
  return MYZKP_OK;
}

int myzkp_host_alloc(void** out, size_t bytes) {
  if (!out) return MYZKP_ERR_INVALID_ARG;
  return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? MYZKP_OK : MYZKP_ERR_OOM;
}
int myzkp_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? MYZKP_OK : MYZKP_ERR_CUDA; }

// ---- device-pointer variants ------------------------------------------------
int myzkp_g1_msm_partial_dev(myzkp_ctx* ctx, const void* d_scalars, size_t n, size_t srs_off, void* d_out_xyzz128) {
  if (!ctx || (!d_scalars && n) || !d_out_xyzz128) return MYZKP_ERR_INVALID_ARG;
  MZ_TRY(begin_call(ctx));
  return msm_xyzz(ctx, static_cast<const uint32_t*>(d_scalars), n, srs_off, static_cast<XYZZ*>(d_out_xyzz128));
}

int myzkp_g1_msm_partial(myzkp_ctx* ctx, const uint8_t* scalars_le, size_t n, size_t srs_off, void* d_out_xyzz128) {
  if (!ctx || (!scalars_le && n) || !d_out_xyzz128) return MYZKP_ERR_INVALID_ARG;
  if (srs_off + n > ctx->srs_n) return fail(ctx, n && !ctx->table ? MYZKP_ERR_NO_SRS : MYZKP_ERR_INVALID_ARG,
                                            "scalar range longer than the SRS");
  MZ_TRY(begin_call(ctx));
  const int K = upload_chunks(ctx, n);
  if (n) MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
  if (K == 1) {
    if (n) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, scalars_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return msm_xyzz(ctx, ctx->scalars.as<uint32_t>(), n, srs_off, static_cast<XYZZ*>(d_out_xyzz128));
  }
  MZ_TRY(enqueue_chunk_uploads(ctx, scalars_le, ctx->scalars.as<uint8_t>(), n, K, false));
  return chunked_msm(ctx, ctx->scalars.as<uint32_t>(), n, K, false, nullptr, nullptr, static_cast<XYZZ*>(d_out_xyzz128),
                     srs_off);
}

// ---- range-sharded commit / open with the exchange fused over peer memory (peer.cu) -------------
int myzkp_kzg_commit_sharded_dev(myzkp_ctx* ctx, const void* d_scalars, size_t n_local, void* d_out_c64) {
  if (!ctx || (!d_scalars && n_local) || !d_out_c64) return MYZKP_ERR_INVALID_ARG;
  MZ_TRY(begin_call(ctx));
  XYZZ* res = reinterpret_cast<XYZZ*>(ctx->small.as<uint8_t>() + kSmallXyzz);
  MZ_TRY(msm_xyzz(ctx, static_cast<const uint32_t*>(d_scalars), n_local, 0, res));
  return peer_exchange(ctx, 0, res, 128, d_out_c64, nullptr);
}

int myzkp_g1_exchange_sum_dev(myzkp_ctx* ctx, const void* d_partial_xyzz128, void* d_out_c64) {
  if (!ctx || !d_partial_xyzz128 || !d_out_c64) return MYZKP_ERR_INVALID_ARG;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return peer_exchange(ctx, 0, d_partial_xyzz128, 128, d_out_c64, nullptr);
}

int myzkp_kzg_commit_sharded(myzkp_ctx* ctx, const uint8_t* scalars_le, size_t n_local, uint8_t out_c[64]) {
  if (!ctx || (!scalars_le && n_local) || !out_c) return MYZKP_ERR_INVALID_ARG;
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  MZ_TRY(myzkp_g1_msm_partial(ctx, scalars_le, n_local, 0, s + kSmallXyzz));
  MZ_TRY(peer_exchange(ctx, 0, s + kSmallXyzz, 128, s + kSmallPoint, nullptr));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_c, s + kSmallPoint, 64, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_TRY(end_call_check_flag(ctx));
  return peer_check(ctx);
}

int myzkp_kzg_open_sharded_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n_local, const uint8_t u_le[32],
                               void* d_out_y32, void* d_out_w64) {
  if (!ctx || (!d_coefs && n_local) || !u_le || !d_out_y32 || !d_out_w64) return MYZKP_ERR_INVALID_ARG;
  if (!fr_bytes_canonical(u_le)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u >= r");
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  const uint32_t* coefs = static_cast<const uint32_t*>(d_coefs);
  uint32_t* pair = reinterpret_cast<uint32_t*>(s + kSmallY);         // (h, u^n) of this range
  uint32_t* carry = reinterpret_cast<uint32_t*>(s + kSmallY + 64);   // carry entering from above
  XYZZ* res = reinterpret_cast<XYZZ*>(s + kSmallXyzz);               // partial, then c_0 right behind it
  uint32_t* c0 = reinterpret_cast<uint32_t*>(s + kSmallXyzz + sizeof(XYZZ));
  MZ_TRY(fr_range_eval(ctx, coefs, n_local, u_le, pair, pair + 8, reinterpret_cast<int*>(s + kSmallFlag)));
  MZ_TRY(peer_exchange(ctx, 1, pair, 64, carry, nullptr));
  if (n_local) {
    MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n_local * 32));
    MZ_TRY(fr_range_quotient(ctx, coefs, n_local, u_le, nullptr, ctx->scalars2.as<uint32_t>(), c0, carry));
  } else {
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(c0, carry, 32, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  // q_{lo+i} pairs with local SRS point i (the global top quotient coefficient is the zero carry)
  MZ_TRY(msm_xyzz(ctx, ctx->scalars2.as<uint32_t>(), n_local, 0, res));
  return peer_exchange(ctx, 0, res, 160, d_out_w64, d_out_y32);
}

int myzkp_kzg_open_sharded(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n_local, const uint8_t u_le[32],
                           uint8_t out_y[32], uint8_t out_w[64]) {
  if (!ctx || (!coefs_le && n_local) || !u_le || !out_y || !out_w) return MYZKP_ERR_INVALID_ARG;
  if (n_local > ctx->srs_n) return fail(ctx, MYZKP_ERR_INVALID_ARG, "coefficient slice longer than this rank's SRS range");
  MZ_TRY(begin_call(ctx));
  if (n_local) {
    MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n_local * 32));
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, coefs_le, n_local * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  uint8_t* s = ctx->small.as<uint8_t>();
  // results land behind the scan's own use of kSmallY (pair, carry): y at +128, W at kSmallPoint
  MZ_TRY(myzkp_kzg_open_sharded_dev(ctx, ctx->scalars.p, n_local, u_le, s + kSmallY + 128, s + kSmallPoint));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_y, s + kSmallY + 128, 32, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_w, s + kSmallPoint, 64, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_TRY(end_call_check_flag(ctx));
  return peer_check(ctx);
}

int myzkp_g1_sum_partials_dev(myzkp_ctx* ctx, const void* d_partials, size_t k, void* d_out_c64) {
  if (!ctx || (!d_partials && k) || !d_out_c64) return MYZKP_ERR_INVALID_ARG;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return sum_partials(ctx, static_cast<const XYZZ*>(d_partials), k, static_cast<uint8_t*>(d_out_c64));
}

int myzkp_kzg_commit_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, void* d_out_c64) {
  if (!ctx || (!d_coefs && n) || !d_out_c64) return MYZKP_ERR_INVALID_ARG;
  MZ_TRY(begin_call(ctx));
  XYZZ* res = reinterpret_cast<XYZZ*>(ctx->small.as<uint8_t>() + kSmallXyzz);
  MZ_TRY(msm_xyzz(ctx, static_cast<const uint32_t*>(d_coefs), n, 0, res));
  return xyzz_to_bytes(ctx, res, 1, static_cast<uint8_t*>(d_out_c64));
}

int myzkp_kzg_open_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, const uint8_t u_le[32], void* d_out_y32,
                       void* d_out_w64) {
  if (!ctx || (!d_coefs && n) || !u_le || !d_out_y32 || !d_out_w64) return MYZKP_ERR_INVALID_ARG;
  if (!fr_bytes_canonical(u_le)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u >= r");
  MZ_TRY(begin_call(ctx));
  XYZZ* res = reinterpret_cast<XYZZ*>(ctx->small.as<uint8_t>() + kSmallXyzz);
  if (n == 0) {
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(d_out_y32, 0, 32, ctx->stream));
    MZ_CUDA_TRY(ctx, cudaMemsetAsync(d_out_w64, 0, 64, ctx->stream));
    return MYZKP_OK;
  }
  MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 32));
  int* flag = reinterpret_cast<int*>(ctx->small.as<uint8_t>() + kSmallFlag);
  MZ_TRY(fr_range_quotient(ctx, static_cast<const uint32_t*>(d_coefs), n, u_le, nullptr, ctx->scalars2.as<uint32_t>(),
                           static_cast<uint32_t*>(d_out_y32), nullptr, flag));
  // q has n-1 coefficients (q[n-1] is the zero carry entering from above)
  MZ_TRY(msm_xyzz(ctx, ctx->scalars2.as<uint32_t>(), n - 1, 0, res));
  return xyzz_to_bytes(ctx, res, 1, static_cast<uint8_t*>(d_out_w64));
}

int myzkp_fr_range_eval_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, const uint8_t u_le[32], void* d_out_h32,
                            void* d_out_upow32) {
  if (!ctx || (!d_coefs && n) || !u_le || !d_out_h32 || !d_out_upow32) return MYZKP_ERR_INVALID_ARG;
  if (!fr_bytes_canonical(u_le)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u >= r");
  MZ_TRY(begin_call(ctx));
  return fr_range_eval(ctx, static_cast<const uint32_t*>(d_coefs), n, u_le, static_cast<uint32_t*>(d_out_h32),
                       static_cast<uint32_t*>(d_out_upow32), reinterpret_cast<int*>(ctx->small.as<uint8_t>() + kSmallFlag));
}

int myzkp_fr_range_quotient_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, const uint8_t u_le[32],
                                const uint8_t carry_in_le[32], void* d_q, void* d_c0) {
  if (!ctx || (!d_coefs && n) || !u_le || !carry_in_le || (!d_q && n) || !d_c0) return MYZKP_ERR_INVALID_ARG;
  if (!fr_bytes_canonical(u_le) || !fr_bytes_canonical(carry_in_le)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u or carry >= r");
  MZ_TRY(begin_call(ctx));
  if (n == 0) {
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(d_c0, carry_in_le, 32, cudaMemcpyHostToDevice, ctx->stream));
    return MYZKP_OK;
  }
  return fr_range_quotient(ctx, static_cast<const uint32_t*>(d_coefs), n, u_le, carry_in_le,
                           static_cast<uint32_t*>(d_q), static_cast<uint32_t*>(d_c0), nullptr,
                           reinterpret_cast<int*>(ctx->small.as<uint8_t>() + kSmallFlag));
}

// ---- host-buffer variants ------------------------------------------------
int myzkp_kzg_commit(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, uint8_t out_c[64]) {
  if (!ctx || (!coefs_le && n) || !out_c) return MYZKP_ERR_INVALID_ARG;
  if (n > ctx->srs_n) return fail(ctx, n && !ctx->table ? MYZKP_ERR_NO_SRS : MYZKP_ERR_INVALID_ARG,
                                  "polynomial longer than the SRS (reference panics at polynomial.rs:162)");
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  const int K = upload_chunks(ctx, n);
  if (n) MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
  if (K == 1) {
    if (n) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, coefs_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    MZ_TRY(myzkp_kzg_commit_dev(ctx, ctx->scalars.p, n, s + kSmallPoint));
  } else {
    // chunk k is accumulated (against SRS points [lo_k, hi_k)) while chunk k+1 is still on the bus;
    // all chunks use the window of the whole polynomial and share one final bucket reduce
    MZ_TRY(enqueue_chunk_uploads(ctx, coefs_le, ctx->scalars.as<uint8_t>(), n, K, false));
    XYZZ* res = reinterpret_cast<XYZZ*>(s + kSmallXyzz);
    MZ_TRY(chunked_msm(ctx, ctx->scalars.as<uint32_t>(), n, K, /*descending=*/false, nullptr, nullptr, res));
    MZ_TRY(xyzz_to_bytes(ctx, res, 1, s + kSmallPoint));
  }
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_c, s + kSmallPoint, 64, cudaMemcpyDeviceToHost, ctx->stream));
  return end_call_check_flag(ctx);
}

int myzkp_kzg_open(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32], uint8_t out_y[32],
                   uint8_t out_w[64]) {
  if (!ctx || (!coefs_le && n) || !u_le || !out_y || !out_w) return MYZKP_ERR_INVALID_ARG;
  if (n > ctx->srs_n + 1 || (n > 1 && !ctx->table))
    return fail(ctx, !ctx->table ? MYZKP_ERR_NO_SRS : MYZKP_ERR_INVALID_ARG, "quotient longer than the SRS");
  if (!fr_bytes_canonical(u_le)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u >= r");
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  const int K = upload_chunks(ctx, n);
  if (n) MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
  if (K == 1) {
    if (n) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, coefs_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    MZ_TRY(myzkp_kzg_open_dev(ctx, ctx->scalars.p, n, u_le, s + kSmallY, s + kSmallPoint));
  } else {
    // the quotient scan runs from the top coefficient down, so chunks are uploaded and
    // consumed top first; the carry leaving a chunk (its c_0) stays on the device
    MZ_TRY(enqueue_chunk_uploads(ctx, coefs_le, ctx->scalars.as<uint8_t>(), n, K, true));
    MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 32));
    XYZZ* res = reinterpret_cast<XYZZ*>(s + kSmallXyzz);
    MZ_TRY(chunked_msm(ctx, ctx->scalars.as<uint32_t>(), n, K, /*descending=*/true, u_le, ctx->scalars2.as<uint32_t>(), res));
    MZ_TRY(xyzz_to_bytes(ctx, res, 1, s + kSmallPoint));
  }
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_y, s + kSmallY, 32, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_w, s + kSmallPoint, 64, cudaMemcpyDeviceToHost, ctx->stream));
  return end_call_check_flag(ctx);
}

// Polynomials of at least this many coefficients get an MSM of their own (window tuned to their size);
// everything smaller in a batch shares ONE pipeline (msm_batch_xyzz), so a batch of many small
// polynomials costs its entries rather than one latency-bound pipeline each.
constexpr size_t kBatchBelow = (size_t)1 << 19;
constexpr size_t kBatchMaxPoints = (size_t)1 << 25;

int myzkp_kzg_commit_batch(myzkp_ctx* ctx, const uint8_t* const* coefs, const size_t* ns, size_t k, uint8_t* out) {
  if (!ctx || (k && (!coefs || !ns || !out))) return MYZKP_ERR_INVALID_ARG;
  if (k == 0) return MYZKP_OK;
  size_t total = 0;
  for (size_t i = 0; i < k; i++) {
    if (ns[i] && !coefs[i]) return MYZKP_ERR_INVALID_ARG;
    if (ns[i] > ctx->srs_n) return fail(ctx, ns[i] && !ctx->table ? MYZKP_ERR_NO_SRS : MYZKP_ERR_INVALID_ARG,
                                        "polynomial longer than the SRS (reference panics at polynomial.rs:162)");
    total += ns[i];
  }
  MZ_TRY(begin_call(ctx));
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure((total ? total : 1) * 32));
  MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure(k * (sizeof(XYZZ) + 64)));
  XYZZ* res = ctx->xyzz_tmp.as<XYZZ>();  // results in processing order: small polynomials first, then the big ones
  uint8_t* d_pts = reinterpret_cast<uint8_t*>(res + k);
  std::vector<const uint32_t*> dptr(k);
  {
    // polynomials land back to back on the device; runs that are already back to back on the host
    // (rows of one matrix) travel as one copy
    uint32_t* p = ctx->scalars.as<uint32_t>();
    const uint8_t* run_src = nullptr;
    uint32_t* run_dst = p;
    size_t run_bytes = 0;
    for (size_t i = 0; i < k; i++) {
      dptr[i] = p;
      if (ns[i]) {
        if (run_bytes && coefs[i] == run_src + run_bytes) {
          run_bytes += ns[i] * 32;
        } else {
          if (run_bytes) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, ctx->stream));
          run_src = coefs[i];
          run_dst = p;
          run_bytes = ns[i] * 32;
        }
      }
      p += ns[i] * 8;
    }
    if (run_bytes) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  std::vector<size_t> order;  // order[j] = index of the polynomial whose result is res[j]
  order.reserve(k);
  std::vector<MsmItem> group;
  size_t group_points = 0;
  auto flush = [&]() -> int {
    if (group.empty()) return MYZKP_OK;
    int rc = msm_batch_xyzz(ctx, group.data(), group.size(), 0, res + (order.size() - group.size()));
    group.clear();
    group_points = 0;
    return rc;
  };
  for (size_t i = 0; i < k; i++) {
    if (ns[i] >= kBatchBelow) continue;
    if (group.size() == 65535 || group_points + ns[i] > kBatchMaxPoints) MZ_TRY(flush());
    group.push_back(MsmItem{dptr[i], ns[i]});
    group_points += ns[i];
    order.push_back(i);
  }
  MZ_TRY(flush());
  for (size_t i = 0; i < k; i++) {
    if (ns[i] < kBatchBelow) continue;
    MZ_TRY(msm_xyzz(ctx, dptr[i], ns[i], 0, res + order.size()));
    order.push_back(i);
  }
  MZ_TRY(xyzz_to_bytes(ctx, res, k, d_pts));
  std::vector<uint8_t> host(k * 64);
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(host.data(), d_pts, k * 64, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_TRY(end_call_check_flag(ctx));
  for (size_t j = 0; j < k; j++) memcpy(out + 64 * order[j], host.data() + 64 * j, 64);
  return MYZKP_OK;
}

int myzkp_gemini_fold_commit(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n_pow2, const uint8_t* rhos_le,
                             size_t n_rhos, uint8_t* out, uint8_t* out_folds) {
  if (!ctx || !coefs_le || !out) return MYZKP_ERR_INVALID_ARG;
  if (n_pow2 == 0 || (n_pow2 & (n_pow2 - 1)))
    return fail(ctx, MYZKP_ERR_INVALID_ARG, "coefs.len() must be a power of two (gemini.rs:55-57)");
  int m = 0;
  while (((size_t)1 << m) < n_pow2) m++;
  if (n_rhos != (size_t)m)
    return fail(ctx, MYZKP_ERR_INVALID_ARG, "points.len() must be log2(coefs.len()) (SplitFoldError::PointsLenMismatch, gemini.rs:60-66)");
  if (m && !rhos_le) return MYZKP_ERR_INVALID_ARG;
  if (n_pow2 > ctx->srs_n) return fail(ctx, !ctx->table ? MYZKP_ERR_NO_SRS : MYZKP_ERR_INVALID_ARG, "polynomial longer than the SRS");
  for (int i = 0; i < m; i++)
    if (!fr_bytes_canonical(rhos_le + 32 * i)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "rho >= r");
  MZ_TRY(begin_call(ctx));
  // level buffers: all m+1 polynomials back to back (2n - 1 coefficients)
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure((2 * n_pow2) * 32 + (size_t)(m + 1) * 32));
  uint32_t* base = ctx->scalars.as<uint32_t>();
  uint32_t* d_rhos = base + (2 * n_pow2) * 8;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(base, coefs_le, n_pow2 * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (m) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(d_rhos, rhos_le, (size_t)m * 32, cudaMemcpyHostToDevice, ctx->stream));
  MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure((size_t)(m + 1) * sizeof(XYZZ) + (size_t)(m + 1) * 64));
  XYZZ* res = ctx->xyzz_tmp.as<XYZZ>();
  uint8_t* d_pts = reinterpret_cast<uint8_t*>(res + (m + 1));
  int* flag = reinterpret_cast<int*>(ctx->small.as<uint8_t>() + kSmallFlag);
  MZ_TRY(fr_check_canonical(ctx, base, n_pow2, flag));
  // all folds first (each level depends on the previous one; they are cheap element-wise kernels) ...
  {
    uint32_t* cur = base;
    size_t len = n_pow2;
    for (int lvl = 0; lvl < m; lvl++) {
      uint32_t* nxt = cur + len * 8;
      MZ_TRY(fr_fold(ctx, cur, len / 2, d_rhos + 8 * lvl, nxt));
      cur = nxt;
      len /= 2;
    }
  }
  // ... then the commitments.  Levels of >= 2^20 coefficients get their own MSM on this stream; the levels below run
  // as TWO batched pipelines on child contexts (own stream and scratch, same SRS table), concurrently with the large
  // ones: 2^16 .. 2^19 coefficients together (window 16 by the batch cost model) and everything below 2^16 (window 12).
  // One batch for all small levels would force one window on them: the bucket reduce of 2^15 buckets per polynomial
  // costs more than the commitments of the tiny levels themselves (round 1: 8.5 ms for 2^20, of which ~1 ms was that).
  if (!ctx->fork_ev) MZ_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  MZ_CUDA_TRY(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
  constexpr size_t kGemOwn = (size_t)1 << 20, kGemMid = (size_t)1 << 16;
  std::vector<MsmItem> cls[2];  // 0: mid batch, 1: small batch
  int first_of[2] = {m + 1, m + 1};
  int first_small = m + 1;      // first level that is not an own-stream MSM
  {
    uint32_t* cur = base;
    size_t len = n_pow2;
    for (int lvl = 0; lvl <= m; lvl++) {
      if (len < kGemOwn) {
        const int k = len >= kGemMid ? 0 : 1;
        if (cls[k].empty()) first_of[k] = lvl;
        if (first_small > lvl) first_small = lvl;
        cls[k].push_back(MsmItem{cur, len});
      }
      cur += len * 8;
      len /= 2;
    }
  }
  myzkp_ctx* chs[2] = {nullptr, nullptr};
  int rc_child = MYZKP_OK, rc_parent = MYZKP_OK;
  const char* child_err = "";
  for (int k = 0; k < 2 && rc_child == MYZKP_OK; k++) {
    if (cls[k].empty()) continue;
    myzkp_ctx* ch = nullptr;
    MZ_TRY(get_child(ctx, k, &ch));
    MZ_CUDA_TRY(ctx, cudaStreamWaitEvent(ch->stream, ctx->fork_ev, 0));  // the folds come first
    chs[k] = ch;
    rc_child = msm_batch_xyzz(ch, cls[k].data(), cls[k].size(), 0, res + first_of[k]);
    if (rc_child != MYZKP_OK) child_err = ch->err.c_str();
    ctx->launches += ch->launches;
    ch->launches = 0;
  }
  if (rc_child == MYZKP_OK) {
    uint32_t* cur = base;
    size_t len = n_pow2;
    for (int lvl = 0; lvl < first_small && rc_parent == MYZKP_OK; lvl++) {
      rc_parent = msm_xyzz(ctx, cur, len, 0, res + lvl);
      cur += len * 8;
      len /= 2;
    }
  }
  // join: whatever was enqueued on a child reads `base` and writes `res`, so the parent stream waits for it on
  // every path out of here, error paths included
  for (int k = 0; k < 2; k++) {
    myzkp_ctx* ch = chs[k];
    if (!ch) continue;
    cudaError_t je = cudaEventRecord(ch->join_ev, ch->stream);
    if (je == cudaSuccess) je = cudaStreamWaitEvent(ctx->stream, ch->join_ev, 0);
    if (je != cudaSuccess) {
      cudaStreamSynchronize(ch->stream);
      cudaGetLastError();
    }
  }
  if (rc_child != MYZKP_OK) return fail(ctx, rc_child, child_err);
  if (rc_parent != MYZKP_OK) return rc_parent;
  MZ_TRY(xyzz_to_bytes(ctx, res, (size_t)(m + 1), d_pts));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_pts, (size_t)(m + 1) * 64, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_folds && n_pow2 > 1)
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_folds, base + n_pow2 * 8, (n_pow2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return end_call_check_flag(ctx);
}

int myzkp_kzg_batch_open(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t* us_le, size_t k,
                         uint8_t* out_ys, uint8_t out_w[64]) {
  if (!ctx || (!coefs_le && n) || (k && (!us_le || !out_ys)) || !out_w || k > 64) return MYZKP_ERR_INVALID_ARG;
  for (size_t i = 0; i < k; i++)
    if (!fr_bytes_canonical(us_le + 32 * i)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u >= r");
  const size_t nq = n > k ? n - k : 0;  // quotient length
  if (nq > ctx->srs_n || (nq && !ctx->table))
    return fail(ctx, !ctx->table ? MYZKP_ERR_NO_SRS : MYZKP_ERR_INVALID_ARG, "quotient longer than the SRS");
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  int* flag = reinterpret_cast<int*>(s + kSmallFlag);
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure((n ? n : 1) * 32));
  MZ_CUDA_TRY(ctx, ctx->scalars2.ensure((n ? n : 1) * 32));
  MZ_CUDA_TRY(ctx, ctx->xyzz_tmp.ensure(k * 32 + 64));
  uint32_t* d_ys = ctx->xyzz_tmp.as<uint32_t>();
  uint32_t* d_scratch = d_ys + 8 * k;  // (u^n, c0) sinks
  if (n) {
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, coefs_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    MZ_TRY(fr_check_canonical(ctx, ctx->scalars.as<uint32_t>(), n, flag));
  }
  // ys[i] = f(u_i) (kzg.rs:79)
  for (size_t i = 0; i < k; i++)
    MZ_TRY(fr_range_eval(ctx, ctx->scalars.as<uint32_t>(), n, us_le + 32 * i, d_ys + 8 * i, d_scratch));
  // (f - I)/Z with deg I < k: floor division by prod (x - u_i) = k successive synthetic divisions
  uint32_t* cur = ctx->scalars.as<uint32_t>();
  uint32_t* nxt = ctx->scalars2.as<uint32_t>();
  size_t len = n;
  for (size_t i = 0; i < k && len > 0; i++) {
    MZ_TRY(fr_range_quotient(ctx, cur, len, us_le + 32 * i, nullptr, nxt, d_scratch + 8));
    uint32_t* t = cur; cur = nxt; nxt = t;
    len -= 1;  // q has len-1 coefficients (the top one written is the zero carry)
  }
  XYZZ* res = reinterpret_cast<XYZZ*>(s + kSmallXyzz);
  MZ_TRY(msm_xyzz(ctx, cur, n > k ? len : 0, 0, res));
  MZ_TRY(xyzz_to_bytes(ctx, res, 1, s + kSmallPoint));
  if (k) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_ys, d_ys, k * 32, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_w, s + kSmallPoint, 64, cudaMemcpyDeviceToHost, ctx->stream));
  return end_call_check_flag(ctx);
}

int myzkp_kzg_prove_degree_bound(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, size_t d, uint8_t out_p[64]) {
  if (!ctx || (!coefs_le && n) || !out_p) return MYZKP_ERR_INVALID_ARG;
  if (!ctx->table || ctx->srs_n == 0) return fail(ctx, MYZKP_ERR_NO_SRS, "no SRS loaded");
  const size_t max_d = ctx->srs_n - 1;
  if (d > max_d) return fail(ctx, MYZKP_ERR_INVALID_ARG, "degree bound above max_d (reference underflows at kzg.rs:127)");
  const size_t shift = max_d - d;  // commit to f * x^(max_d - d)  (kzg.rs:126-133)
  if (n + shift > ctx->srs_n) return fail(ctx, MYZKP_ERR_INVALID_ARG, "deg f > d: shifted polynomial longer than the SRS (reference panics at polynomial.rs:162)");
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  if (n) {
    MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, coefs_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  XYZZ* res = reinterpret_cast<XYZZ*>(s + kSmallXyzz);
  MZ_TRY(msm_xyzz(ctx, ctx->scalars.as<uint32_t>(), n, shift, res));
  MZ_TRY(xyzz_to_bytes(ctx, res, 1, s + kSmallPoint));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_p, s + kSmallPoint, 64, cudaMemcpyDeviceToHost, ctx->stream));
  return end_call_check_flag(ctx);
}

int myzkp_g1_msm(myzkp_ctx* ctx, const uint8_t* scalars_le, const uint8_t* points_or_null, size_t n, uint8_t out[64]) {
  if (!ctx || (!scalars_le && n) || !out) return MYZKP_ERR_INVALID_ARG;
  if (!points_or_null) return myzkp_kzg_commit(ctx, scalars_le, n, out);  // against the resident SRS
  if (n == 0) {
    memset(out, 0, 64);
    return MYZKP_OK;
  }
  // caller-supplied points (accumulate_curve_points-style call sites, zksnark/utils.rs:83-93): classic windowed
  // Pippenger over the points as given (msm_points_xyzz) - no table of multiples is built, the resident SRS
  // is not touched, and the call costs about one MSM plus a fixed ~1 ms window combine
  MZ_TRY(begin_call(ctx));  // (scalars >= r are caught by the recode kernel's canonicity flag)
  uint8_t* s = ctx->small.as<uint8_t>();
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
  MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 64));
  MZ_CUDA_TRY(ctx, ctx->caller_points.ensure(n * sizeof(Affine)));
  int* pflag = reinterpret_cast<int*>(s + 528);  // own word: 512 is the sticky scalar flag
  MZ_CUDA_TRY(ctx, cudaMemsetAsync(pflag, 0, sizeof(int), ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, scalars_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars2.p, points_or_null, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  MZ_TRY(points_import(ctx, ctx->scalars2.as<uint32_t>(), n, ctx->caller_points.as<Affine>(), pflag));
  XYZZ* res = reinterpret_cast<XYZZ*>(s + kSmallXyzz);
  uint8_t* d_pt = s + kSmallPoint;
  MZ_TRY(msm_points_xyzz(ctx, ctx->scalars.as<uint32_t>(), ctx->caller_points.as<Affine>(), n, res));
  MZ_TRY(xyzz_to_bytes(ctx, res, 1, d_pt));
  int h_pflag = 0;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_pt, 64, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(&h_pflag, pflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MZ_TRY(end_call_check_flag(ctx));
  if (h_pflag) return fail(ctx, MYZKP_ERR_NONCANONICAL, "point coordinate >= p");
  return MYZKP_OK;
}

int myzkp_fr_eval(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32], uint8_t out_y[32]) {
  if (!ctx || (!coefs_le && n) || !u_le || !out_y) return MYZKP_ERR_INVALID_ARG;
  if (!fr_bytes_canonical(u_le)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u >= r");
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  if (n) {
    MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
    MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, coefs_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  MZ_TRY(fr_range_eval(ctx, ctx->scalars.as<uint32_t>(), n, u_le, reinterpret_cast<uint32_t*>(s + kSmallY),
                       reinterpret_cast<uint32_t*>(s + kSmallY + 32), reinterpret_cast<int*>(s + kSmallFlag)));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_y, s + kSmallY, 32, cudaMemcpyDeviceToHost, ctx->stream));
  return end_call_check_flag(ctx);
}

int myzkp_fr_quotient(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32], uint8_t out_y[32],
                      uint8_t* out_q) {
  if (!ctx || (!coefs_le && n) || !u_le || !out_y || (n > 1 && !out_q)) return MYZKP_ERR_INVALID_ARG;
  if (!fr_bytes_canonical(u_le)) return fail(ctx, MYZKP_ERR_NONCANONICAL, "u >= r");
  MZ_TRY(begin_call(ctx));
  uint8_t* s = ctx->small.as<uint8_t>();
  if (n == 0) {
    memset(out_y, 0, 32);
    return MYZKP_OK;
  }
  MZ_CUDA_TRY(ctx, ctx->scalars.ensure(n * 32));
  MZ_CUDA_TRY(ctx, ctx->scalars2.ensure(n * 32));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scalars.p, coefs_le, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  MZ_TRY(fr_range_quotient(ctx, ctx->scalars.as<uint32_t>(), n, u_le, nullptr, ctx->scalars2.as<uint32_t>(),
                           reinterpret_cast<uint32_t*>(s + kSmallY), nullptr, reinterpret_cast<int*>(s + kSmallFlag)));
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_y, s + kSmallY, 32, cudaMemcpyDeviceToHost, ctx->stream));
  if (n > 1) MZ_CUDA_TRY(ctx, cudaMemcpyAsync(out_q, ctx->scalars2.p, (n - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return end_call_check_flag(ctx);
}

}  // extern "C"
