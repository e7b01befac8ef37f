// Small all-gather over NVLink peer memory for the per-step messages of the multi-GPU greedy selection (SURVEY.md 8e,
// collective 3): 100 dependent steps, one 34 KB message per rank and step -- latency, not bandwidth.  An NCCL all-gather costs
// ~10 us more per step at 8 GPUs (greedy k = 100: 10.5 vs 9.5 ms, profiles/r2_bench_n8_*.json); this kernel is one launch: block d of rank r stores r's message into slot r of rank d's receive
// buffer (peer stores through NVSwitch), publishes flag (d <- r) with a system-scope release, and waits for flag (r <- d).
//
// Every rank allocates one buffer  [flags: 16 x u32, padded to 1 KB][2 parities][world slots][slot bytes]  with cudaMalloc,
// exports it as a CUDA IPC handle, and opens the handles of its peers (the host layer moves the 64-byte handles with
// torch.distributed).  Messages of sequence number q use parity q & 1 and flag value q + 1 (monotonic, compared modulo 2^32):
// a rank can run at most one exchange ahead of its slowest peer -- it needs that peer's message of exchange q to finish q --
// so the slot a late reader still holds (q) is never the one an early writer fills (q + 1).
#include "nnal_common.cuh"
#include "../../include/nnal_b200.h"
#include <cstring>

namespace p2p {

constexpr int MAX_WORLD = 16;
constexpr size_t FLAG_BYTES = 1024;

struct Ptrs { unsigned char* base[MAX_WORLD]; };

struct State {
  unsigned char* base = nullptr;          // own buffer
  size_t slot = 0;
  int world = 0, rank = -1;
  Ptrs peers;
  bool opened[MAX_WORLD] = {};
};

__global__ void __launch_bounds__(256) allgather_kernel(const uint4* __restrict__ send, int n16, Ptrs p, int world, int rank,
                                                         size_t slot, int parity, unsigned flagval) {
  const int dst = blockIdx.x;
  uint4* out = reinterpret_cast<uint4*>(p.base[dst] + FLAG_BYTES + ((size_t)parity * world + rank) * slot);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) out[i] = send[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* theirs = reinterpret_cast<unsigned*>(p.base[dst]) + rank;        // flag (dst <- rank)
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(flagval) : "memory");
    const unsigned* mine = reinterpret_cast<const unsigned*>(p.base[rank]) + dst;   // flag (rank <- dst)
    long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - flagval) >= 0) break;
      if ((spin & 0x3ff) == 0x3ff) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ll) __trap();       // a peer never arrived (~10 s): CUDA error, not a hung GPU
      }
    }
  }
}

static State* get(nnal_ctx* ctx) { return (State*)ctx->p2p_state; }

static void release(nnal_ctx* ctx) {
  State* s = get(ctx);
  if (!s) return;
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < s->world; ++r)
    if (r != s->rank && s->opened[r]) cudaIpcCloseMemHandle(s->peers.base[r]);
  if (s->base) cudaFree(s->base);
  delete s;
  ctx->p2p_state = nullptr;
}

}  // namespace p2p

int nnal_p2p_release(nnal_ctx* ctx) {
  if (ctx) p2p::release(ctx);
  return NNAL_OK;
}

extern "C" int nnal_p2p_alloc(nnal_ctx* ctx, int world, int rank, int64_t slot_bytes, unsigned char* handle_out) {
  if (!ctx || !handle_out || world < 2 || world > p2p::MAX_WORLD || rank < 0 || rank >= world || slot_bytes <= 0 || slot_bytes % 16)
    return NNAL_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  p2p::release(ctx);
  p2p::State* s = new p2p::State();
  ctx->p2p_state = s;
  s->world = world; s->rank = rank; s->slot = (size_t)slot_bytes;
  const size_t bytes = p2p::FLAG_BYTES + 2 * (size_t)world * s->slot;
  CUDA_TRY(ctx, cudaMalloc((void**)&s->base, bytes));
  CUDA_TRY(ctx, cudaMemset(s->base, 0, bytes));
  CUDA_TRY(ctx, cudaDeviceSynchronize());
  memset(&s->peers, 0, sizeof(s->peers));
  s->peers.base[rank] = s->base;
  cudaIpcMemHandle_t h;
  CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, s->base));
  memcpy(handle_out, &h, 64);
  return NNAL_OK;
}

extern "C" int nnal_p2p_open(nnal_ctx* ctx, const unsigned char* handles) {
  if (!ctx || !handles) return NNAL_ERR_INVALID;
  p2p::State* s = p2p::get(ctx);
  if (!s) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_p2p_alloc not called");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void* p = nullptr;
    CUDA_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peers.base[r] = (unsigned char*)p;
    s->opened[r] = true;
  }
  return NNAL_OK;
}

// Same-process form (several contexts in one process -- the GPU tests, or a host that drives all GPUs from one process): the
// peers' buffers are addressable as they are.  own_base: this rank's buffer (for the other contexts' nnal_p2p_open_local).
extern "C" int nnal_p2p_base(nnal_ctx* ctx, void** own_base) {
  if (!ctx || !own_base) return NNAL_ERR_INVALID;
  p2p::State* s = p2p::get(ctx);
  if (!s) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_p2p_alloc not called");
  *own_base = s->base;
  return NNAL_OK;
}
extern "C" int nnal_p2p_open_local(nnal_ctx* ctx, void* const* bases) {
  if (!ctx || !bases) return NNAL_ERR_INVALID;
  p2p::State* s = p2p::get(ctx);
  if (!s) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_p2p_alloc not called");
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) continue;
    if (!bases[r]) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "null peer buffer");
    s->peers.base[r] = (unsigned char*)bases[r];
    s->opened[r] = false;                                  // not an IPC mapping: nothing to close
  }
  return NNAL_OK;
}

extern "C" int nnal_p2p_allgather(nnal_ctx* ctx, const void* d_send, int64_t nbytes, uint64_t seq, void** d_recv) {
  if (!ctx || !d_send || !d_recv || nbytes <= 0 || nbytes % 16) return NNAL_ERR_INVALID;
  p2p::State* s = p2p::get(ctx);
  if (!s) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_p2p_alloc not called");
  if ((size_t)nbytes > s->slot) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "message larger than the slot the exchange was set up with");
  for (int r = 0; r < s->world; ++r)
    if (!s->peers.base[r]) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_p2p_open not called");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int parity = (int)(seq & 1);
  p2p::allgather_kernel<<<s->world, 256, 0, ctx->stream>>>((const uint4*)d_send, (int)(nbytes / 16), s->peers, s->world, s->rank, s->slot,
                                                           parity, (unsigned)(seq + 1));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  *d_recv = s->base + p2p::FLAG_BYTES + (size_t)parity * s->world * s->slot;
  return NNAL_OK;
}
