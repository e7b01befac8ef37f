// Prints the register <-> (TMEM lane, column) mapping of tcgen05.ld shapes .16x256b / .16x128b / .16x64b on sm_100a:
// TMEM is filled through tcgen05.st.32x32b with value = 1000 * lane + column, then read back with the shape under test.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_layout tmem_layout.cu && ./tmem_layout
#include "../../nn-active-learning_b200/csrc/tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
using namespace tcx;

__global__ void __launch_bounds__(128, 1) k(int* out, int c1, int c2) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // fill: each warp writes its 32 lanes, 64 columns: value = 1000 * lane + column
  for (int c = 0; c < 64; ++c) {
    uint32_t v = 1000u * (warp * 32 + lane) + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c), "r"(v));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {     // warp 1 reads from its own quadrant (lanes 32..63)
    uint32_t r[8];
    // .16x256b.x2: 16 lanes x 16 columns -> 8 registers per thread; first half of the quadrant (lane offset 0), column 3
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(32) << 16) + c1));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[lane * 8 + j] = r[j];
    // second half of the quadrant: lane offset 16
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(32 + 16) << 16) + c2));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[256 + lane * 8 + j] = r[j];
    uint32_t q[4];
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(tmem + ((uint32_t)(32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 4; ++j) out[512 + lane * 4 + j] = q[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64)); }
}

int main(int argc, char** argv) {
  const int c1 = argc > 1 ? atoi(argv[1]) : 0, c2 = argc > 2 ? atoi(argv[2]) : 0;
  int* d; cudaMalloc(&d, 4096); cudaMemset(d, 0xff, 4096);
  k<<<1, 128>>>(d, c1, c2);
  printf("column offsets %d %d\n", c1, c2);
  int h[1024];
  cudaError_t e = cudaMemcpy(h, d, 4096, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  printf(".16x256b.x2 from (lane 32, col c1): thread: regs as lane:col\n");
  for (int t = 0; t < 32; ++t) { printf("T%02d:", t); for (int j = 0; j < 8; ++j) printf(" %2d:%-2d", h[t * 8 + j] / 1000, h[t * 8 + j] % 1000); printf("\n"); }
  printf(".16x256b.x2 from (lane 48, col c2):\n");
  for (int t = 0; t < 32; t += 5) { printf("T%02d:", t); for (int j = 0; j < 8; ++j) printf(" %2d:%-2d", h[256 + t * 8 + j] / 1000, h[256 + t * 8 + j] % 1000); printf("\n"); }
  printf(".16x128b.x2 from (lane 32, col 0):\n");
  for (int t = 0; t < 32; ++t) { printf("T%02d:", t); for (int j = 0; j < 4; ++j) printf(" %2d:%-2d", h[512 + t * 4 + j] / 1000, h[512 + t * 4 + j] % 1000); printf("\n"); }
  return 0;
}
