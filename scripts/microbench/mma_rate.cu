// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, SS mode, no-swizzle K-major operands) as a
// function of N and of the number of TMEM accumulators the instruction stream rotates over.  Answers the question
// "what does a narrow MMA really cost" that decides the tiling of the conv kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include "../../nn-active-learning_b200/csrc/tc_ptx.cuh"
#include <cstdio>
using namespace tcx;

__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int iters, int a_stride16, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_bf16(128, N);
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      const uint32_t a16 = base >> 4, b16 = (base + 64 * 1024) >> 4;
      const uint32_t A_LBO = (2048u >> 4) << 16, B_LBO = ((uint32_t)(N * 16) >> 4) << 16;
      const int stride = 512 / nacc;
      t0 = clock64();
      int acc = 0;                              // no division in the issue loop: it must not be the bottleneck
      for (int i = 0; i < iters; ++i) {
        const uint64_t dA = ((uint64_t)DESC_HI << 32) | ((a16 + (uint32_t)(i & 7) * a_stride16) | A_LBO);
        const uint64_t dB = ((uint64_t)DESC_HI << 32) | ((b16 + (uint32_t)(i & 3) * 512u) | B_LBO);
        umma_bf16(tmem + acc, dA, dB, idesc, 1);
        acc += stride;
        if (acc >= 512) acc = 0;
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}


// conv_tc.cu issue pattern: for every K-step, TG MMAs (N1) into TG accumulators, then TG MMAs (N2) into the same ones
__global__ void __launch_bounds__(128, 1) kcat(int N1, int N2, int TG, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one_sync()) {
      const uint32_t id1 = make_idesc_bf16(128, N1), id2 = make_idesc_bf16(128, N2);
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      const uint32_t a16 = base >> 4, b16 = (base + 64 * 1024) >> 4;
      const uint32_t A_LBO = (2048u >> 4) << 16, B_LBO = ((uint32_t)(N1 * 16) >> 4) << 16;
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint64_t dB = ((uint64_t)DESC_HI << 32) | ((b16 + (uint32_t)(i & 3) * 512u) | B_LBO);
        for (int t = 0; t < TG; ++t)
          umma_bf16(tmem + t * N1, ((uint64_t)DESC_HI << 32) | ((a16 + (uint32_t)t * 128u) | A_LBO), dB, id1, 1);
        if (N2 > 0)
          for (int t = 0; t < TG; ++t)
            umma_bf16(tmem + t * N1, ((uint64_t)DESC_HI << 32) | ((a16 + 2048u + (uint32_t)t * 128u) | A_LBO), dB, id2, 1);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}

// operand alignment: the shift trick of the conv kernels starts the positions operand at base + 16 * shift bytes, so
// its 128-byte core matrices (8 rows x 16 B) straddle two 128-byte lines of shared memory
__global__ void __launch_bounds__(128, 1) kalign(int N, int a_off16, int b_off16, int a_lbo, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_bf16(128, N);
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      const uint32_t a16 = (base >> 4) + a_off16, b16 = ((base + 96 * 1024) >> 4) + b_off16;
      const uint32_t A_LBO = ((uint32_t)a_lbo >> 4) << 16, B_LBO = ((uint32_t)(N * 16) >> 4) << 16;
      int acc = 0;
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint64_t dA = ((uint64_t)DESC_HI << 32) | ((a16 + (uint32_t)(i & 7) * 29u) | A_LBO);   // tap-like shifts of 29 positions
        const uint64_t dB = ((uint64_t)DESC_HI << 32) | ((b16 + (uint32_t)(i & 3) * 512u) | B_LBO);
        umma_bf16(tmem + acc, dA, dB, idesc, 1);
        acc = acc == 256 ? 0 : 256;
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}

// run-ahead of the issuing thread: bursts of `burst` MMAs separated by `gap` cycles in which the thread does something
// else (the conv kernels compute the next stage's descriptors there).  If tcgen05.mma only enqueues, the gap hides
// behind the queued MMAs; if the queue is shallow, it shows up in full.
__global__ void __launch_bounds__(128, 1) kburst(int N, int burst, int gap, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0, t_issue = 0;
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_bf16(128, N);
      constexpr uint32_t DESC_HI = 8u | (1u << 14);
      const uint32_t a16 = base >> 4, b16 = (base + 96 * 1024) >> 4;
      const uint32_t A_LBO = (2048u >> 4) << 16, B_LBO = ((uint32_t)(N * 16) >> 4) << 16;
      int acc = 0;
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const long long i0 = clock64();
        for (int i = 0; i < burst; ++i) {
          const uint64_t dA = ((uint64_t)DESC_HI << 32) | ((a16 + (uint32_t)(i & 7) * 29u) | A_LBO);
          const uint64_t dB = ((uint64_t)DESC_HI << 32) | ((b16 + (uint32_t)(i & 3) * 512u) | B_LBO);
          umma_bf16(tmem + acc, dA, dB, idesc, 1);
          acc = acc == 256 ? 0 : 256;
        }
        const long long g0 = clock64();
        t_issue += g0 - i0;
        while (clock64() - g0 < gap) {}
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
    if (blockIdx.x == 0 && t_issue) out[1] = t_issue;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  printf("cycles per MMA (M=128, K=16, f16, SS, no-swizzle), one CTA per SM on all 148 SMs; A operand: 8 different 4 KB blocks\n");
  printf("%6s | %8s %8s %8s %8s %8s | floor N/2, fetch (4KB + N*32B)/128\n", "N", "acc=1", "acc=2", "acc=4", "acc=8", "same A");
  int Ns[] = {16, 24, 32, 48, 64, 96, 128, 192, 256};
  for (int N : Ns) {
    printf("%6d |", N);
    for (int v = 0; v < 5; ++v) {
      int nacc = v < 4 ? (1 << v) : 4;
      if (512 / nacc < N) { printf(" %8s", "-"); continue; }
      int a_stride16 = v < 4 ? 256 : 0;
      k<<<148, 128, 200 * 1024>>>(N, nacc, iters, a_stride16, d);
      long long h = 0;
      cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf(" err:%s", cudaGetErrorString(e)); return 1; }
      printf(" %8.1f", (double)h / iters);
    }
    printf(" | %5.0f %5.1f\n", N / 2.0, (4096 + N * 32) / 128.0);
  }
  cudaFuncSetAttribute(kalign, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("\noperand alignment (A = 128 rows whose start moves in steps of 29 rows = 464 B like the conv tap shifts; LBO = K-half distance)\n");
  {
    int cfg[][4] = {{48, 0, 0, 2048}, {48, 1, 0, 2048}, {48, 0, 0, 13568}, {48, 1, 0, 13568}, {96, 0, 0, 3712}, {96, 3, 0, 3712},
                    {256, 0, 0, 2048}, {256, 0, 1, 2048}, {256, 1, 1, 2048}, {128, 0, 1, 2048}};
    for (auto& c : cfg) {
      kalign<<<148, 128, 200 * 1024>>>(c[0], c[1], c[2], c[3], 4096, d);
      long long h = 0;
      cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf(" err:%s\n", cudaGetErrorString(e)); return 1; }
      printf("N %3d  A start +%d x16B (+29 rows per MMA)  B start +%d x16B  A LBO %5d : %6.1f cycles per MMA\n", c[0], c[1], c[2], c[3], (double)h / 4096);
    }
  }
  cudaFuncSetAttribute(kburst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("\nrun-ahead of the issuing thread: bursts of MMAs (N = 48, 44 cycles each) separated by a gap on the issuing thread\n");
  {
    int cfg[][2] = {{24, 0}, {24, 300}, {24, 600}, {24, 1200}, {4, 100}, {8, 200}, {64, 600}, {64, 2000}};
    for (auto& c : cfg) {
      kburst<<<148, 128, 200 * 1024>>>(48, c[0], c[1], 256, d);
      long long h[2] = {0, 0};
      cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf(" err:%s\n", cudaGetErrorString(e)); return 1; }
      printf("burst %3d MMAs, gap %4d cycles: %7.1f cycles per burst (MMA work %d, issuing the burst took %.1f)\n", c[0], c[1],
             (double)h[0] / 256, c[0] * 44, (double)h[1] / 256);
    }
  }
  cudaFuncSetAttribute(kcat, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("\nconv_tc.cu issue pattern: per K-step TG x MMA(N1) then TG x MMA(N2) into the same TG accumulators; cycles per K-step per tile\n");
  int pats[][3] = {{64, 32, 3}, {48, 24, 3}, {96, 48, 2}, {96, 48, 4}, {96, 0, 2}, {48, 0, 4}, {192, 96, 1}, {192, 96, 2}, {256, 128, 2}, {128, 64, 2}};
  for (auto& pt : pats) {
    kcat<<<148, 128, 200 * 1024>>>(pt[0], pt[1], pt[2], 1024, d);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf(" err:%s\n", cudaGetErrorString(e)); return 1; }
    printf("N1 %3d N2 %3d TG %d : %7.1f cycles per K-step per tile (%d MMAs)\n", pt[0], pt[1], pt[2], (double)h / 1024 / pt[2], pt[1] > 0 ? 2 : 1);
  }
  return 0;
}
