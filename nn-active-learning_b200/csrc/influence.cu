// Last-layer influence recursion (sm_100a).
//
// Reference behaviour replaced: PW_NNAL.stoch_approx_IF (PW_NNAL.py:851-881) on top of NN.LLFC_grads (NN.py:905-955) and
// NN.LLFC_hess (NN.py:874-903).  Upstream materialises, per iteration, the ((d+1)c)^2 Hessian of one random training sample
// (537 MB at PW1) and multiplies it with the ((d+1)c) x n matrix V.  With H = -LLFC_hess = (diag pi - pi pi^T) (x) [u;1][u;1]^T
// (parameters ordered class-major: W rows a*d+k, then the c biases) the product factors:
//     (H v)[a][k] = r_a ut_k,   r = (diag pi - pi pi^T) s,   s_a = sum_k v[a][k] ut_k,   ut = [u;1],
// so one iteration costs 2 c (d+1) multiply-adds per pool sample.  One CTA owns one pool sample for ALL iterations: its
// c x (d+1) slice of V lives in shared memory (float64) when it fits, the training factors stream through L2.
#include "nnal_common.cuh"
#include "dots.cuh"
#include "../../include/nnal_b200.h"
#include <algorithm>

namespace infl {

constexpr int NT = 256;

// V[i] <- grads_i;  T times:  V[i] <- (grads_i + V[i]) - (H_t V[i]) / scale
__global__ void __launch_bounds__(NT) lissa_kernel(const float* __restrict__ pool_post, const float* __restrict__ pool_U,
                                                   const long long* __restrict__ labels, int64_t n, int c, int d, int64_t T,
                                                   const float* __restrict__ tr_post, const float* __restrict__ tr_U, double scale,
                                                   double* __restrict__ Vg, int use_smem) {
  extern __shared__ double sm[];
  const int64_t i = blockIdx.x;
  const int D1 = d + 1;
  double* v = use_smem ? sm : Vg + (size_t)i * c * D1;            // [c][d+1]
  double* sred = use_smem ? sm + (size_t)c * D1 : nullptr;         // [c] dot products, then r
  __shared__ double sfix[64];                                      // used when V lives in global memory (c <= 64)
  if (!sred) sred = sfix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* u = pool_U + (size_t)i * d;
  const long long y = labels[i];
  // grads (NN.py:921-955): dJ/dW = e_y (x) u - float32(pi (x) u)  (the product is formed in float32 upstream), dJ/db = e_y - pi
  for (int e = tid; e < c * D1; e += NT) {
    const int a = e / D1, k = e - a * D1;
    const float pa = pool_post[(size_t)a * n + i];
    double g;
    if (k < d) g = (a == y ? (double)u[k] : 0.0) - (double)__fmul_rn(pa, u[k]);
    else g = (a == y ? 1.0 : 0.0) - (double)pa;
    v[e] = g;
  }
  __syncthreads();
  for (int64_t t = 0; t < T; ++t) {
    const float* ut = tr_U + (size_t)t * d;
    const float* pt = tr_post + (size_t)t * c;
    // s_a = sum_k v[a][k] ut_k
    for (int a = warp; a < c; a += NT / 32) {
      double acc = 0.0;
      const double* va = v + (size_t)a * D1;
      for (int k = lane; k < d; k += 32) acc = fma(va[k], (double)__ldg(ut + k), acc);
      acc = warp_sum(acc);
      if (lane == 0) sred[a] = acc + va[d];
    }
    __syncthreads();
    if (tid == 0) {
      double ps = 0.0;
      for (int b = 0; b < c; ++b) ps = fma((double)pt[b], sred[b], ps);
      for (int a = 0; a < c; ++a) sred[a] = (double)pt[a] * (sred[a] - ps);           // r = (diag pi - pi pi^T) s
    }
    __syncthreads();
    for (int e = tid; e < c * D1; e += NT) {
      const int a = e / D1, k = e - a * D1;
      const float pa = pool_post[(size_t)a * n + i];
      double g;
      if (k < d) g = (a == y ? (double)u[k] : 0.0) - (double)__fmul_rn(pa, u[k]);
      else g = (a == y ? 1.0 : 0.0) - (double)pa;
      const double utk = k < d ? (double)__ldg(ut + k) : 1.0;
      v[e] = (g + v[e]) - (sred[a] * utk) / scale;
    }
    __syncthreads();
  }
  if (use_smem)
    for (int e = tid; e < c * D1; e += NT) Vg[(size_t)i * c * D1 + e] = v[e];
}

// Vg [n][c][d+1] -> out [(d+1)c][n] in the reference's parameter order (W rows a*d+k, then biases c*d+a)
__global__ void __launch_bounds__(256) reorder_kernel(const double* __restrict__ Vg, int64_t n, int c, int d, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int D1 = d + 1;
  const int64_t P = (int64_t)c * D1;
  const int64_t p0 = (int64_t)blockIdx.y * 32, i0 = (int64_t)blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int64_t i = i0 + r, e = p0 + threadIdx.x;                  // e = a*(d+1)+k in the kernel's layout
    tile[r][threadIdx.x] = (i < n && e < P) ? Vg[(size_t)i * P + e] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int64_t e = p0 + r, i = i0 + threadIdx.x;
    if (e < P && i < n) {
      const int a = (int)(e / D1), k = (int)(e - (int64_t)a * D1);
      const int64_t row = k < d ? (int64_t)a * d + k : (int64_t)c * d + a;
      out[(size_t)row * n + i] = tile[threadIdx.x][r];
    }
  }
}

}  // namespace infl

extern "C" int nnal_if_lissa(nnal_ctx* ctx, int64_t n, int c, int d, const float* pool_post, const float* pool_U,
                             const int64_t* labels, int64_t T, const float* tr_post, const float* tr_U, double scale,
                             double* V_out) {
  if (!ctx || n < 0 || c < 2 || c > 64 || d <= 0 || T < 0 || !(scale != 0.0)) return NNAL_ERR_INVALID;
  if (n == 0) return NNAL_OK;
  if (!pool_post || !pool_U || !labels || !V_out || (T > 0 && (!tr_post || !tr_U))) return NNAL_ERR_INVALID;
  for (int64_t i = 0; i < n; ++i)
    if (labels[i] < 0 || labels[i] >= c) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "label outside [0, c)");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t P = (size_t)c * (d + 1);
  const size_t b_post = (size_t)c * n * 4, b_U = (size_t)n * d * 4, b_lab = (size_t)n * 8, b_tp = (size_t)std::max<int64_t>(T, 1) * c * 4,
               b_tu = (size_t)std::max<int64_t>(T, 1) * d * 4, b_V = (size_t)n * P * 8;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  NNAL_TRY(devbuf_reserve(ctx, ctx->fi_ws, al(b_post) + al(b_U) + al(b_lab) + al(b_tp) + al(b_tu)));
  NNAL_TRY(devbuf_reserve(ctx, ctx->act[0], b_V));
  NNAL_TRY(devbuf_reserve(ctx, ctx->act[1], b_V));
  char* w = (char*)ctx->fi_ws.p;
  float* d_post = (float*)w; w += al(b_post);
  float* d_U = (float*)w; w += al(b_U);
  long long* d_lab = (long long*)w; w += al(b_lab);
  float* d_tp = (float*)w; w += al(b_tp);
  float* d_tu = (float*)w;
  CUDA_TRY(ctx, cudaMemcpyAsync(d_post, pool_post, b_post, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_U, pool_U, b_U, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_lab, labels, b_lab, cudaMemcpyHostToDevice, ctx->stream));
  if (T > 0) {
    CUDA_TRY(ctx, cudaMemcpyAsync(d_tp, tr_post, (size_t)T * c * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_tu, tr_U, (size_t)T * d * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  double* Vg = (double*)ctx->act[0].p;
  double* Vo = (double*)ctx->act[1].p;
  const size_t smem = (P + (size_t)c) * sizeof(double);
  const int use_smem = smem <= 200 * 1024 ? 1 : 0;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(infl::lissa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  infl::lissa_kernel<<<(unsigned)n, infl::NT, use_smem ? smem : 0, ctx->stream>>>(d_post, d_U, d_lab, n, c, d, T, d_tp, d_tu, scale, Vg,
                                                                                  use_smem);
  dim3 grid(cdiv(n, 32), cdiv((long long)P, 32)), block(32, 8);
  infl::reorder_kernel<<<grid, block, 0, ctx->stream>>>(Vg, n, c, d, Vo);
  ctx->launches += 2;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(V_out, Vo, b_V, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return NNAL_OK;
}
