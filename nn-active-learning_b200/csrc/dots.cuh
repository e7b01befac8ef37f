// Shared device helpers: warp reductions and the float32-chain / float64-flush inner product.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The three inner products of a (candidate, winner) pair, float64 accumulation of exact float32 products in
// a FIXED order (four independent chains per lane, lane-strided, then butterfly):
//   uu = u.x,  aa = a.y,  mm = sum_k beta2_k 1[u_k > 0] 1[x_k > 0].
// Inner products of float32 factor rows.  Every lane multiplies-and-adds SHORT float32 chains (four
// independent chains of four FMAs: the product inside an FMA is exact, one rounding per add) and flushes
// them into float64 sums, which are then reduced in float64: the result carries ~1e-8 relative error
// (16-term float32 partials, random signs, averaged over d/16 partials) at one FMA per element -- a
// cvt+DFMA loop is bound by the quarter-rate float32->float64 conversions, a compensated float32 dot
// product by its 10 flops per element; this one stays on the HBM roofline.
// Fixed order (lane-strided, then butterfly) => bit-reproducible.
//   uu = u.x,  aa = a.y,  mm = sum_k beta2_k 1[u_k > 0] 1[x_k > 0]
__device__ __forceinline__ void dot_um(const float* __restrict__ u, const float* __restrict__ x,
                                       const float* __restrict__ beta2, int d, bool mask, int lane, double& uu, double& mm) {
  double su = 0.0, sm = 0.0;
  if ((d & 3) == 0) {
    int k = lane * 4;
    for (; k + 3 * 128 < d; k += 4 * 128) {
      float4 p[4], q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p[i] = *reinterpret_cast<const float4*>(u + k + i * 128);
        q[i] = *reinterpret_cast<const float4*>(x + k + i * 128);
      }
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f0 = fmaf(p[i].x, q[i].x, f0);
        f1 = fmaf(p[i].y, q[i].y, f1);
        f2 = fmaf(p[i].z, q[i].z, f2);
        f3 = fmaf(p[i].w, q[i].w, f3);
      }
      su += ((double)f0 + (double)f1) + ((double)f2 + (double)f3);
      if (mask) {
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(beta2 + k + i * 128));
          g0 += (p[i].x > 0.f && q[i].x > 0.f) ? b.x : 0.f;
          g1 += (p[i].y > 0.f && q[i].y > 0.f) ? b.y : 0.f;
          g2 += (p[i].z > 0.f && q[i].z > 0.f) ? b.z : 0.f;
          g3 += (p[i].w > 0.f && q[i].w > 0.f) ? b.w : 0.f;
        }
        sm += ((double)g0 + (double)g1) + ((double)g2 + (double)g3);
      }
    }
    for (; k < d; k += 128) {
      const float4 p = *reinterpret_cast<const float4*>(u + k);
      const float4 q = *reinterpret_cast<const float4*>(x + k);
      su += ((double)(p.x * q.x) + (double)(p.y * q.y)) + ((double)(p.z * q.z) + (double)(p.w * q.w));   // not reached for d % 512 == 0
      if (mask) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta2 + k));
        sm += (double)((p.x > 0.f && q.x > 0.f) ? b.x : 0.f) + (double)((p.y > 0.f && q.y > 0.f) ? b.y : 0.f) +
              (double)((p.z > 0.f && q.z > 0.f) ? b.z : 0.f) + (double)((p.w > 0.f && q.w > 0.f) ? b.w : 0.f);
      }
    }
  } else {
    for (int k = lane; k < d; k += 32) {
      su = fma((double)u[k], (double)x[k], su);
      if (mask && u[k] > 0.f && x[k] > 0.f) sm += (double)__ldg(beta2 + k);
    }
  }
  uu = su;
  mm = sm;
}


// Same inner products with the CANDIDATE row stored as fp16 (u16) and the winner row x in float32: the large-candidate-set
// variant of the kernel column (fi.cu): the pass is bound by the bytes of the candidate rows, fp16 halves them.  The fp16 value
// is exact in float32, products and sums as above.  d % 8 == 0.
#include <cuda_fp16.h>
__device__ __forceinline__ void dot_um_h8(const uint4 raw, const float* __restrict__ x, const float* __restrict__ beta2, int k,
                                          bool mask, double& su, double& sm) {
  const float4 q0 = *reinterpret_cast<const float4*>(x + k), q1 = *reinterpret_cast<const float4*>(x + k + 4);
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  const float2 p01 = __half22float2(h2[0]), p23 = __half22float2(h2[1]), p45 = __half22float2(h2[2]), p67 = __half22float2(h2[3]);
  float f0 = p01.x * q0.x, f1 = p01.y * q0.y, f2 = p23.x * q0.z, f3 = p23.y * q0.w;
  f0 = fmaf(p45.x, q1.x, f0); f1 = fmaf(p45.y, q1.y, f1); f2 = fmaf(p67.x, q1.z, f2); f3 = fmaf(p67.y, q1.w, f3);
  su += ((double)f0 + (double)f1) + ((double)f2 + (double)f3);
  if (mask) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta2 + k)), b1 = __ldg(reinterpret_cast<const float4*>(beta2 + k + 4));
    float g0 = (p01.x > 0.f && q0.x > 0.f) ? b0.x : 0.f, g1 = (p01.y > 0.f && q0.y > 0.f) ? b0.y : 0.f;
    float g2 = (p23.x > 0.f && q0.z > 0.f) ? b0.z : 0.f, g3 = (p23.y > 0.f && q0.w > 0.f) ? b0.w : 0.f;
    g0 += (p45.x > 0.f && q1.x > 0.f) ? b1.x : 0.f; g1 += (p45.y > 0.f && q1.y > 0.f) ? b1.y : 0.f;
    g2 += (p67.x > 0.f && q1.z > 0.f) ? b1.z : 0.f; g3 += (p67.y > 0.f && q1.w > 0.f) ? b1.w : 0.f;
    sm += ((double)g0 + (double)g1) + ((double)g2 + (double)g3);
  }
}
__device__ __forceinline__ void dot_um_h(const __half* __restrict__ u16, const float* __restrict__ x,
                                         const float* __restrict__ beta2, int d, bool mask, int lane, double& uu, double& mm) {
  double su = 0.0, sm = 0.0;
  int k = lane * 8;
  for (; k + 3 * 256 < d; k += 4 * 256) {                 // four 16-byte candidate loads in flight per lane
    uint4 raw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) raw[j] = *reinterpret_cast<const uint4*>(u16 + k + j * 256);
#pragma unroll
    for (int j = 0; j < 4; ++j) dot_um_h8(raw[j], x, beta2, k + j * 256, mask, su, sm);
  }
  for (; k < d; k += 256) dot_um_h8(*reinterpret_cast<const uint4*>(u16 + k), x, beta2, k, mask, su, sm);
  uu = su;
  mm = sm;
}
