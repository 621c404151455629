// Peer-memory exchange for the range-sharded multi-GPU path: ONE kernel that pushes this rank's
// small partial result straight into every peer's HBM over NVLink (plain stores to peer-mapped
// memory), waits for the peers' partials to land in its own buffer and finishes the job in the
// same launch:
//   mode 0  sum of the world's XYZZ partials -> canonical affine bytes (the commitment / proof W),
//           optionally also rank 0's 32-byte c_0 (= y of open_kzg)
//   mode 1  composition of the scan carries of the ranks above this one (open_kzg, polynomial.rs:371-405
//           sharded by index range) -> the carry entering this rank's range
// It replaces all_gather(NCCL) + a separate sum kernel: the payload is 64..160 bytes, so the exchange
// is pure latency and a collective library's launch + protocol overhead dominates it.
//
// Buffer of a rank (own cudaMalloc, exported through CUDA IPC, 8 KiB + flag word):
//   slot(parity, src) = (parity * 16 + src) * 256 B : payload (<= 160 B) at +0, epoch flag at +240
// Every exchange has an epoch number (same sequence on every rank).  Writers store the payload,
// fence at system scope and release-store the epoch into the flag; readers acquire-load the flag
// of each source slot in their OWN buffer.  Two parities suffice: a rank can only reach epoch
// e + 2 after every peer has published epoch e + 1, which a peer does after its epoch-e kernel
// (the reader of parity e) has completed in stream order.
// A reader that waits longer than the timeout sets the error word and gives up (a peer died or
// the ranks disagree on the call sequence); the host reports MYZKP_ERR_CUDA on the next check.
#include <string.h>

#include "ctx.cuh"

namespace mz {

constexpr int kPeerSlotBytes = 256;
constexpr int kPeerFlagOff = 240;
constexpr int kPeerMaxVec = 10;  // 160-byte payload
constexpr size_t kPeerSlots = 2 * myzkp_ctx::kMaxPeers * kPeerSlotBytes;
constexpr size_t kPeerErrOff = kPeerSlots;
constexpr size_t kPeerBufBytes = kPeerSlots + 256;

struct PeerArgs {
  uint8_t* bufs[myzkp_ctx::kMaxPeers];
  int rank, world;
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ XYZZ part_xyzz(const uint4* p) {
  XYZZ v;
  uint32_t* s = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 8; i++) { s[4 * i] = p[i].x; s[4 * i + 1] = p[i].y; s[4 * i + 2] = p[i].z; s[4 * i + 3] = p[i].w; }
  return v;
}
__device__ __forceinline__ Fr part_fr(const uint4* p) {
  Fr v;
  v.v[0] = p[0].x; v.v[1] = p[0].y; v.v[2] = p[0].z; v.v[3] = p[0].w;
  v.v[4] = p[1].x; v.v[5] = p[1].y; v.v[6] = p[1].z; v.v[7] = p[1].w;
  return v;
}

// one block, one warp per peer
__global__ void __launch_bounds__(32 * myzkp_ctx::kMaxPeers)
peer_exchange_kernel(PeerArgs a, int mode, const uint4* __restrict__ payload, int nvec, uint32_t* out0, uint32_t* out1,
                     int* err, unsigned long long timeout_ns) {
  __shared__ uint4 parts[myzkp_ctx::kMaxPeers][kPeerMaxVec];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t parity_base = (size_t)(a.epoch & 1) * myzkp_ctx::kMaxPeers;
  if (w < a.world) {
    // push my payload into rank w's buffer (NVLink store when w is another GPU)
    uint8_t* dst = a.bufs[w] + (parity_base + a.rank) * kPeerSlotBytes;
    if (lane < nvec) reinterpret_cast<uint4*>(dst)[lane] = payload[lane];
    __threadfence_system();
    __syncwarp();
    if (lane == 0) st_release_sys(reinterpret_cast<uint32_t*>(dst + kPeerFlagOff), a.epoch);
    // wait for rank w's payload to land in my buffer
    const uint8_t* src = a.bufs[a.rank] + (parity_base + w) * kPeerSlotBytes;
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(src + kPeerFlagOff);
    const unsigned long long t0 = global_ns();
    bool ok = true;
    while (ld_acquire_sys(flag) != a.epoch) {
      if (global_ns() - t0 > timeout_ns) { ok = false; break; }
    }
    if (!ok && lane == 0) atomicExch(err, 1);
    if (lane < nvec) parts[w][lane] = ld_volatile_v4(reinterpret_cast<const uint4*>(src) + lane);
  }
  __syncthreads();
  if (mode == 0) {
    // tree sum of the partials (group law: any order gives the same element)
#pragma unroll 1
    for (int d = myzkp_ctx::kMaxPeers / 2; d > 0; d >>= 1) {
      if (lane == 0 && w < d && w + d < a.world) {
        XYZZ x = part_xyzz(parts[w]), y = part_xyzz(parts[w + d]);
        xyzz_add(x, y);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
        for (int i = 0; i < 8; i++) parts[w][i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      Affine p = xyzz_to_affine(part_xyzz(parts[0]));
      Fq x = fe_from_mont(p.x), y = fe_from_mont(p.y);
#pragma unroll
      for (int i = 0; i < 8; i++) { out0[i] = x.v[i]; out0[8 + i] = y.v[i]; }
    }
  } else {
    if (threadIdx.x == 0) {
      // carry entering this rank = maps of the ranks above composed downwards, starting from 0:
      // range g sends c to h_g + u^{n_g} * c
      Fr c = Fr::zero();
      for (int g = a.world - 1; g > a.rank; g--) {
        Fr h = fe_to_mont(part_fr(parts[g])), m = fe_to_mont(part_fr(parts[g] + 2));
        c = fe_add(h, fe_mul(m, c));
      }
      c = fe_from_mont(c);
#pragma unroll
      for (int i = 0; i < 8; i++) out0[i] = c.v[i];
    }
  }
  // y of the sharded open: rank 0's c_0 rides behind its partial (vectors 8, 9).  parts[0][8..9]
  // is not touched by the tree sum.
  if (out1 && threadIdx.x < 8) out1[threadIdx.x] = reinterpret_cast<const uint32_t*>(&parts[0][8])[threadIdx.x];
}

int peer_exchange(myzkp_ctx* ctx, int mode, const void* d_payload, int bytes, void* d_out0, void* d_out1) {
  if (ctx->peer_world <= 0 || !ctx->peer_local) return fail(ctx, MYZKP_ERR_INVALID_ARG, "no peers attached (myzkp_peer_attach)");
  if (bytes <= 0 || bytes % 16 || bytes > 16 * kPeerMaxVec) return fail(ctx, MYZKP_ERR_INVALID_ARG, "bad exchange payload");
  PeerArgs a;
  memcpy(a.bufs, ctx->peer_bufs, sizeof a.bufs);
  a.rank = ctx->peer_rank;
  a.world = ctx->peer_world;
  a.epoch = ++ctx->peer_epoch;
  peer_exchange_kernel<<<1, 32 * ctx->peer_world, 0, ctx->stream>>>(
      a, mode, static_cast<const uint4*>(d_payload), bytes / 16, static_cast<uint32_t*>(d_out0),
      static_cast<uint32_t*>(d_out1), reinterpret_cast<int*>(ctx->peer_local + kPeerErrOff), ctx->peer_timeout_ns);
  MZ_LAUNCH_CHECK(ctx);
  return MYZKP_OK;
}

// after a stream synchronisation: did an exchange give up waiting?
int peer_check(myzkp_ctx* ctx) {
  if (!ctx->peer_local || ctx->peer_world <= 0) return MYZKP_OK;
  int h = 0;
  MZ_CUDA_TRY(ctx, cudaMemcpyAsync(&h, ctx->peer_local + kPeerErrOff, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h) return fail(ctx, MYZKP_ERR_CUDA, "peer exchange timed out waiting for another rank");
  return MYZKP_OK;
}

static void peer_close(myzkp_ctx* ctx) {
  for (int r = 0; r < myzkp_ctx::kMaxPeers; r++) {
    if (ctx->peer_ipc[r] && ctx->peer_bufs[r]) cudaIpcCloseMemHandle(ctx->peer_bufs[r]);
    ctx->peer_ipc[r] = false;
    ctx->peer_bufs[r] = nullptr;
  }
  ctx->peer_world = 0;
  ctx->peer_rank = -1;
  ctx->peer_same_device = false;
}

void peer_release(myzkp_ctx* ctx) {
  peer_close(ctx);
  if (ctx->peer_local) cudaFree(ctx->peer_local);
  ctx->peer_local = nullptr;
}

}  // namespace mz

using namespace mz;

extern "C" {

int myzkp_peer_export(myzkp_ctx* ctx, uint8_t out_handle[MYZKP_PEER_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == MYZKP_PEER_HANDLE_BYTES, "IPC handle size");
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  MZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  peer_close(ctx);
  if (!ctx->peer_local) MZ_CUDA_TRY(ctx, cudaMalloc(&ctx->peer_local, kPeerBufBytes));
  MZ_CUDA_TRY(ctx, cudaMemset(ctx->peer_local, 0, kPeerBufBytes));
  MZ_CUDA_TRY(ctx, cudaDeviceSynchronize());
  ctx->peer_epoch = 0;
  if (out_handle) {
    cudaIpcMemHandle_t h;
    MZ_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->peer_local));
    memcpy(out_handle, &h, sizeof h);
  }
  return MYZKP_OK;
}

int myzkp_peer_attach(myzkp_ctx* ctx, int rank, int world, const uint8_t* handles) {
  if (!ctx || !handles) return MYZKP_ERR_INVALID_ARG;
  if (world < 1 || world > myzkp_ctx::kMaxPeers || rank < 0 || rank >= world)
    return fail(ctx, MYZKP_ERR_INVALID_ARG, "peer exchange supports 1..16 ranks");
  if (!ctx->peer_local) return fail(ctx, MYZKP_ERR_INVALID_ARG, "myzkp_peer_export first");
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  peer_close(ctx);
  for (int r = 0; r < world; r++) {
    if (r == rank) {
      ctx->peer_bufs[r] = ctx->peer_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * MYZKP_PEER_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      peer_close(ctx);
      ctx->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
      return MYZKP_ERR_CUDA;
    }
    ctx->peer_bufs[r] = static_cast<uint8_t*>(p);
    ctx->peer_ipc[r] = true;
  }
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  return MYZKP_OK;
}

int myzkp_peer_attach_local(myzkp_ctx* ctx, int rank, int world, myzkp_ctx* const* ctxs) {
  if (!ctx || !ctxs) return MYZKP_ERR_INVALID_ARG;
  if (world < 1 || world > myzkp_ctx::kMaxPeers || rank < 0 || rank >= world || ctxs[rank] != ctx)
    return fail(ctx, MYZKP_ERR_INVALID_ARG, "peer exchange supports 1..16 ranks; ctxs[rank] must be ctx");
  MZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  peer_close(ctx);
  for (int r = 0; r < world; r++) {
    if (!ctxs[r] || !ctxs[r]->peer_local) return fail(ctx, MYZKP_ERR_INVALID_ARG, "every context must myzkp_peer_export first");
    if (ctxs[r]->device != ctx->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return MYZKP_ERR_CUDA;
      }
      cudaGetLastError();
    }
    ctx->peer_bufs[r] = ctxs[r]->peer_local;
    // Ranks sharing one device (in-process tests): from now on this ctx's scratch must not grow - cudaMalloc /
    // cudaFree wait for the whole device, i.e. for a peer's exchange kernel that is spinning for this very rank.
    if (r != rank && ctxs[r]->device == ctx->device) ctx->peer_same_device = true;
  }
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  return MYZKP_OK;
}

int myzkp_peer_detach(myzkp_ctx* ctx) {
  if (!ctx) return MYZKP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  peer_close(ctx);
  return MYZKP_OK;
}

int myzkp_peer_set_timeout_ms(myzkp_ctx* ctx, uint32_t ms) {
  if (!ctx || ms == 0) return MYZKP_ERR_INVALID_ARG;
  ctx->peer_timeout_ns = (unsigned long long)ms * 1000000ull;
  return MYZKP_OK;
}

}  // extern "C"
