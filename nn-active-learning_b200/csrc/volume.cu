// Volume re-layout and patch gather kernels (sm_100a).
//
// Reference behaviour replaced: patch_utils.get_patches (patch_utils.py:1087-1173) and the
// float64 normalisation of PW_NN.batch_eval (PW_NN.py:503-506) / get_patches_multimg
// (patch_utils.py:1203-1207).
//
// HBM layout.  The reference holds each modality as a C-contiguous (X,Y,Z) array with z fastest,
// and a (d1,d2,1) patch is a window in the two SLOW axes: 625 reads >= Z*itemsize bytes apart
// (SURVEY H1).  On upload every subject is re-laid out ONCE into channel-last, z-slowest
// [Z][X][Y][m]; a patch row (d2 voxels x m modalities) is then one contiguous run that maps 1:1
// onto the NHWC patch row the CNN consumes, so gather reads and writes are both coalesced.
#include "nnal_common.cuh"
#include <cmath>

// ------------------------------------------------------------------------------------------
// relayout: staging [m][X][Y][Z] (z fastest)  ->  out [Z+2pz][X+2px][Y+2py][m]
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) relayout_kernel(const T* __restrict__ in, T* __restrict__ out, int m,
                                                        int64_t X, int64_t Y, int64_t Z, int64_t px, int64_t py,
                                                        int64_t pz) {
  __shared__ T tile[32][33];
  const int64_t x = blockIdx.z;
  const int64_t z0 = (int64_t)blockIdx.x * 32, y0 = (int64_t)blockIdx.y * 32;
  const int64_t Xp = X + 2 * px, Yp = Y + 2 * py;
  for (int j = 0; j < m; ++j) {
    const T* src = in + (int64_t)j * X * Y * Z + x * Y * Z;
#pragma unroll
    for (int r = threadIdx.y; r < 32; r += 8) {
      int64_t y = y0 + r, z = z0 + threadIdx.x;
      tile[r][threadIdx.x] = (y < Y && z < Z) ? src[y * Z + z] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int r = threadIdx.y; r < 32; r += 8) {
      int64_t z = z0 + r, y = y0 + threadIdx.x;
      if (y < Y && z < Z) out[(((z + pz) * Xp + (x + px)) * Yp + (y + py)) * m + j] = tile[threadIdx.x][r];
    }
    __syncthreads();
  }
}

int nnal_k_relayout(nnal_ctx* ctx, const void* stage, int dtype, int m, int64_t X, int64_t Y, int64_t Z,
                    int64_t px, int64_t py, int64_t pz, void* out) {
  dim3 grid(cdiv(Z, 32), cdiv(Y, 32), (unsigned)X), block(32, 8);
  if (grid.y > 65535 || grid.z > 65535) NNAL_FAIL(ctx, NNAL_ERR_UNSUPPORTED, "volume extent too large");
  size_t esz = dtype == NNAL_F64 ? 8 : 4;
  if (px || py || pz)
    CUDA_TRY(ctx, cudaMemsetAsync(out, 0, (size_t)((X + 2 * px) * (Y + 2 * py) * (Z + 2 * pz)) * m * esz, ctx->stream));
  if (dtype == NNAL_F64)
    relayout_kernel<double><<<grid, block, 0, ctx->stream>>>((const double*)stage, (double*)out, m, X, Y, Z, px, py, pz);
  else
    relayout_kernel<float><<<grid, block, 0, ctx->stream>>>((const float*)stage, (float*)out, m, X, Y, Z, px, py, pz);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

// ------------------------------------------------------------------------------------------
// gather: one CTA per patch; threads sweep the d1 x (d2*C) output, C = m*d3.
// Output channel ch = j*d3 + dz holds modality j at depth offset dz (patch_utils.py:1156-1165).
// norm_mode 0: raw; 1: PW_NN.batch_eval style -- channel ch < m uses stats[ch] (PW_NN.py:503-506,
// as written: correct only for d3 == 1); 2: get_patches_multimg style -- block ch/d3 uses
// stats[ch/d3] (patch_utils.py:1203-1207).  Arithmetic in float64 like the reference.
//
// The float64 division (x - mu)/sigma is done as Markstein's correctly-rounded sequence
//   q0 = a * y,  r = fma(-q0, sigma, a),  q = fma(r, y, q0),   y = RN(1/sigma) (computed once per channel),
// which equals IEEE division bit for bit (sigma's significand is never all ones for a standard deviation
// computed in floating point; checked on the host) at 3 flops instead of a ~35-instruction DDIV: the
// gather stays on the HBM roofline.
// ------------------------------------------------------------------------------------------
// generic kernel: any C, element-wise sweep with incremental (row, column, channel) indices -- no divisions
template <typename TV, typename TO>
__global__ void __launch_bounds__(256) gather_kernel(const TV* __restrict__ vol, int m, int64_t Xp, int64_t Yp,
                                                      int64_t Zp, const int64_t* __restrict__ inds, int64_t n, int d1,
                                                      int d2, int d3, const double* __restrict__ stats, int norm_mode,
                                                      TO* __restrict__ out) {
  const int C = m * d3;
  const int rowlen = d2 * C;
  const int per_patch = d1 * rowlen;
  const int64_t Y0 = Yp - (d2 - 1), Z0 = Zp - (d3 - 1);
  // per-thread stride decomposition: e -> e + blockDim.x  ==  (i, r) -> (i + qi, r + qr) with carry
  const int qi = blockDim.x / rowlen, qr = blockDim.x - qi * rowlen;
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    int64_t ind = inds[p];
    // np.unravel_index(ind, orig_shape) (patch_utils.py:1144); host validated the range
    int64_t z = ind % Z0;
    int64_t t = ind / Z0;
    int64_t y = t % Y0;
    int64_t x = t / Y0;
    TO* dst = out + p * (int64_t)per_patch;
    int i = threadIdx.x / rowlen, r = threadIdx.x - i * rowlen;
    for (int e = threadIdx.x; e < per_patch; e += blockDim.x) {
      const int k = r / C;
      const int ch = r - k * C;
      int j, dz;
      if (d3 == 1) { j = ch; dz = 0; } else { j = ch / d3; dz = ch - j * d3; }
      double v = (double)vol[(((z + dz) * Xp + (x + i)) * Yp + (y + k)) * m + j];
      if (norm_mode == 1) {
        if (ch < m) v = norm_apply(v, stats[3 * ch], stats[3 * ch + 1], stats[3 * ch + 2]);
      } else if (norm_mode == 2) {
        v = norm_apply(v, stats[3 * j], stats[3 * j + 1], stats[3 * j + 2]);
      }
      dst[e] = (TO)v;
      i += qi; r += qr;
      if (r >= rowlen) { r -= rowlen; ++i; }
    }
  }
}

// d3 == 1 (the reference's production patch shape (25,25,1), run_on_subjects.py:18): a patch row is ONE
// contiguous run of d2*m values in the [Z][X][Y][m] volume and in the output, so the sweep needs one
// multiply-add per element: src = base + i*rowstride + r, dst = e; (i, r, ch) advance incrementally.
template <typename TV, typename TO, bool NORM>
__global__ void __launch_bounds__(256) gather_rows_kernel(const TV* __restrict__ vol, int m, int64_t Xp, int64_t Yp,
                                                           const int64_t* __restrict__ inds, int64_t n, int d1, int d2,
                                                           int64_t Y0, int64_t Z0, const double* __restrict__ stats,
                                                           TO* __restrict__ out) {
  // Thread t owns column r = t % rowlen of the patch rows i = t / rowlen, + R, + 2R, ... (R = rows covered by the block):
  // source and destination advance by constants, the channel (r % m) never changes, so the normalisation constants
  // live in registers and an element costs a load, a store and two pointer increments.  (The first version carried
  // (row, column, channel) indices through a flat sweep: 42 instructions per element, 71 % issue-slot utilisation
  // for a copy kernel -- ncu, profiles/r1_config4.json history.)
  const int rowlen = d2 * m;
  const int per_patch = d1 * rowlen;
  const int R = blockDim.x / rowlen;                     // launch_gather guarantees rowlen <= blockDim.x
  const int i0 = threadIdx.x / rowlen, r = threadIdx.x - i0 * rowlen;
  const bool active = i0 < R;
  const int64_t rowstride = Yp * m;
  double f_mu = 0., f_sg = 1., f_rs = 1.;
  if (NORM) { const int ch = r % m; f_mu = stats[3 * ch]; f_sg = stats[3 * ch + 1]; f_rs = stats[3 * ch + 2]; }
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const int64_t ind = inds[p];
    const int64_t z = ind % Z0;
    const int64_t t = ind / Z0;
    const int64_t y = t % Y0;
    const int64_t x = t / Y0;
    if (!active) continue;
    const TV* src = vol + ((z * Xp + x + i0) * Yp + y) * m + r;
    TO* dst = out + p * (int64_t)per_patch + i0 * rowlen + r;
    // four independent loads in flight per thread
#pragma unroll 4
    for (int i = i0; i < d1; i += R) {
      const TV raw = *src;
      // (measured and dropped: skipping the two correction FMAs unless q0 is within a few ulps of a float32 rounding
      //  midpoint -- 47.7 vs 46.9 ms -- and assembling the float64 bits with integer instructions -- 51.1 ms: the
      //  kernel is not bound by the float64 pipe)
      if (NORM) *dst = (TO)norm_apply((double)raw, f_mu, f_sg, f_rs);
      else *dst = (TO)raw;
      src += R * rowstride;
      dst += R * rowlen;
    }
  }
}

// Fused gather + normalise + fp16 hi/lo split + channel padding: writes the tensor-core conv's input planes
// [n][d1][d2][8] (hi plane, then lo plane) directly -- one thread per patch position reads its C <= 8 channel
// values (12 contiguous bytes for 3 modalities) and writes one 16-byte chunk per plane.  Replaces the fp32
// patch tensor + a separate split/pad pass (7.5 + 20 KB per patch instead of 7.5 + 7.5 + 7.5 + 20).
__global__ void __launch_bounds__(256) gather_split_kernel(const float* __restrict__ vol, int m, int64_t Xp, int64_t Yp,
                                                            int64_t Zp, const int64_t* __restrict__ inds, int64_t n, int d1,
                                                            int d2, int d3, NormTab tab, nnal_h* __restrict__ out_hi,
                                                            nnal_h* __restrict__ out_lo) {
  const int C = m * d3;
  const int npos = d1 * d2;
  const int64_t Y0 = Yp - (d2 - 1), Z0 = Zp - (d3 - 1);
  const int qi = blockDim.x / d2, qk = blockDim.x - qi * d2;
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const int64_t ind = inds[p];
    const int64_t z = ind % Z0;
    const int64_t t = ind / Z0;
    const int64_t y = t % Y0;
    const int64_t x = t / Y0;
    const float* pbase = vol + ((z * Xp + x) * Yp + y) * m;
    int i = threadIdx.x / d2, k = threadIdx.x - i * d2;
    for (int pos = threadIdx.x; pos < npos; pos += blockDim.x) {
      uint32_t hw[4] = {0, 0, 0, 0}, lw[4] = {0, 0, 0, 0};
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        if (ch < C) {
          int j, dz;
          if (d3 == 1) { j = ch; dz = 0; } else { j = ch / d3; dz = ch - j * d3; }
          double v = (double)(d3 == 1 ? pbase[(i * Yp + k) * m + ch] : vol[(((z + dz) * Xp + (x + i)) * Yp + (y + k)) * m + j]);
          if (tab.on[ch]) v = norm_apply(v, tab.mu[ch], tab.sg[ch], tab.rs[ch]);
          nnal_h h, l;
          nnal_split((float)v, h, l);
          hw[ch >> 1] |= (uint32_t)__half_as_ushort(h) << ((ch & 1) * 16);
          lw[ch >> 1] |= (uint32_t)__half_as_ushort(l) << ((ch & 1) * 16);
        }
      }
      const int64_t o = (p * npos + pos) * 8;
      *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      i += qi; k += qk;
      if (k >= d2) { k -= d2; ++i; }
    }
  }
}

template <typename TO>
static int launch_gather(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                         const double* d_stats, int norm_mode, TO* d_out) {
  if (n == 0) return NNAL_OK;
  int grid = (int)(n < (int64_t)ctx->sm_count * 32 ? n : (int64_t)ctx->sm_count * 32);
  if (d3 == 1 && v.m <= 16 && d2 * v.m <= 256) {
    // (norm modes 1 and 2 coincide for d3 == 1: every output channel is one modality)
    const int64_t Y0 = v.Y - (d2 - 1), Z0 = v.Z;
    const bool norm = norm_mode != 0;
    const int blk = 256;                                   // >= d2 * m (checked above): whole patch rows per sweep
    if (v.dtype == NNAL_F64) {
      if (norm) gather_rows_kernel<double, TO, true><<<grid, blk, 0, ctx->stream>>>((const double*)v.data, v.m, v.X, v.Y, d_inds, n, d1, d2, Y0, Z0, d_stats, d_out);
      else gather_rows_kernel<double, TO, false><<<grid, 256, 0, ctx->stream>>>((const double*)v.data, v.m, v.X, v.Y, d_inds, n, d1, d2, Y0, Z0, d_stats, d_out);
    } else {
      if (norm) gather_rows_kernel<float, TO, true><<<grid, blk, 0, ctx->stream>>>((const float*)v.data, v.m, v.X, v.Y, d_inds, n, d1, d2, Y0, Z0, d_stats, d_out);
      else gather_rows_kernel<float, TO, false><<<grid, 256, 0, ctx->stream>>>((const float*)v.data, v.m, v.X, v.Y, d_inds, n, d1, d2, Y0, Z0, d_stats, d_out);
    }
  } else if (v.dtype == NNAL_F64)
    gather_kernel<double, TO><<<grid, 256, 0, ctx->stream>>>((const double*)v.data, v.m, v.X, v.Y, v.Z, d_inds, n, d1,
                                                              d2, d3, d_stats, norm_mode, d_out);
  else
    gather_kernel<float, TO><<<grid, 256, 0, ctx->stream>>>((const float*)v.data, v.m, v.X, v.Y, v.Z, d_inds, n, d1,
                                                             d2, d3, d_stats, norm_mode, d_out);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

// d_stats: [m][3] = (mu, sigma, RN(1/sigma)) per modality (built by the C-ABI layer)
int nnal_k_gather_f64(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                      const double* d_stats, int norm_mode, double* d_out) {
  return launch_gather<double>(ctx, v, d_inds, n, d1, d2, d3, d_stats, norm_mode, d_out);
}

int nnal_k_gather_norm_f32(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                           const double* d_stats, int norm_mode, float* d_out) {
  return launch_gather<float>(ctx, v, d_inds, n, d1, d2, d3, d_stats, norm_mode, d_out);
}

// Fused gather -> conv1's x-im2col'd tensor-core input (conv_tc.cu CfgConv1X): fp16 hi/lo planes [n][d1][d2][16] where
// element dx * C + ch of position (i, k) is the normalised patch value at (i, k + dx - 2, ch) (zero outside the patch,
// element 15 zero): the 5 filter columns are folded into the channel axis, so conv1 needs 5 K-steps instead of 13.
// One patch per block iteration: every value is read, normalised (float64, as get_patches_multimg / batch_eval) and
// split ONCE into shared memory, then each thread assembles the 16-element rows of its positions (64 B per position).
__global__ void __launch_bounds__(256) gather_x16_kernel(const float* __restrict__ vol, int m, int64_t Xp, int64_t Yp,
                                                          int64_t Zp, const int64_t* __restrict__ inds, int64_t n, int d1,
                                                          int d2, NormTab tab, nnal_h* __restrict__ out_hi,
                                                          nnal_h* __restrict__ out_lo) {
  extern __shared__ uint32_t sv[];                       // [d1][d2][m] packed (hi | lo << 16)
  const int C = m, KW = 5;
  const int npos = d1 * d2, row = d2 * C;
  const int64_t Y0 = Yp - (d2 - 1), Z0 = Zp;
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const int64_t ind = inds[p];
    const int64_t z = ind % Z0;
    const int64_t t = ind / Z0;
    const int64_t y = t % Y0;
    const int64_t x = t / Y0;
    const float* pbase = vol + ((z * Xp + x) * Yp + y) * m;
    for (int e = threadIdx.x; e < npos * C; e += blockDim.x) {
      const int i = e / row, r = e - i * row;            // r = k * C + ch: contiguous in the [Z][X][Y][m] volume
      const int ch = r % C;
      double v = (double)pbase[(int64_t)i * Yp * m + r];
      if (tab.on[ch]) v = norm_apply(v, tab.mu[ch], tab.sg[ch], tab.rs[ch]);
      nnal_h h, l;
      nnal_split((float)v, h, l);
      sv[e] = (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
    }
    __syncthreads();
    for (int pos = threadIdx.x; pos < npos; pos += blockDim.x) {
      const int i = pos / d2, k = pos - i * d2;
      uint32_t hw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 15; ++j) {
        const int dx = j / 3, ch = j - dx * 3;
        const int kk = k + dx - KW / 2;
        if (ch < C && kk >= 0 && kk < d2) {
          const uint32_t w = sv[(i * d2 + kk) * C + ch];
          hw[j >> 1] |= (w & 0xffffu) << ((j & 1) * 16);
          lw[j >> 1] |= (w >> 16) << ((j & 1) * 16);
        }
      }
      const int64_t o = (p * npos + pos) * 16;
      uint4* dh = reinterpret_cast<uint4*>(out_hi + o);
      uint4* dl = reinterpret_cast<uint4*>(out_lo + o);
      dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
      dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
    }
    __syncthreads();
  }
}

bool nnal_k_gather_x16_supported(const Volume& v, int d1, int d2, int d3) {
  return v.dtype == NNAL_F32 && d3 == 1 && v.m == 3 && d1 == 25 && d2 == 25;
}

// host copy of the stats ([m][2]) is needed to build the per-output-channel table passed by value
bool nnal_k_gather_split_supported(const Volume& v, int d3) { return v.dtype == NNAL_F32 && v.m * d3 <= 8; }


int nnal_k_gather_x16(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                      const double* h_stats, int norm_mode, nnal_h* out_hi, nnal_h* out_lo) {
  if (n == 0) return NNAL_OK;
  const NormTab tab = nnal_make_norm_tab(v, d3, h_stats, norm_mode);
  int grid = (int)(n < (int64_t)ctx->sm_count * 16 ? n : (int64_t)ctx->sm_count * 16);
  gather_x16_kernel<<<grid, 256, (size_t)d1 * d2 * v.m * sizeof(uint32_t), ctx->stream>>>((const float*)v.data, v.m, v.X, v.Y, v.Z, d_inds, n,
                                                                                          d1, d2, tab, out_hi, out_lo);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

int nnal_k_gather_split(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                        const double* h_stats, int norm_mode, nnal_h* out_hi, nnal_h* out_lo) {
  if (n == 0) return NNAL_OK;
  const NormTab tab = nnal_make_norm_tab(v, d3, h_stats, norm_mode);
  int grid = (int)(n < (int64_t)ctx->sm_count * 32 ? n : (int64_t)ctx->sm_count * 32);
  gather_split_kernel<<<grid, 256, 0, ctx->stream>>>((const float*)v.data, v.m, v.X, v.Y, v.Z, d_inds, n, d1, d2, d3, tab, out_hi,
                                                     out_lo);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

NormTab nnal_make_norm_tab(const Volume& v, int d3, const double* h_stats, int norm_mode) {
  NormTab tab;
  const int C = v.m * d3;
  for (int ch = 0; ch < 8; ++ch) {
    tab.on[ch] = 0; tab.mu[ch] = 0.0; tab.sg[ch] = 1.0; tab.rs[ch] = 1.0;
    if (ch >= C || norm_mode == 0) continue;
    int src = -1;
    if (norm_mode == 1) { if (ch < v.m) src = ch; }
    else src = ch / d3;
    if (src >= 0) { tab.on[ch] = 1; tab.mu[ch] = h_stats[2 * src]; tab.sg[ch] = h_stats[2 * src + 1]; 
      const double sg = tab.sg[ch];
      uint64_t bits;
      memcpy(&bits, &sg, 8);
      const bool ok = std::isfinite(sg) && sg != 0.0 && (bits & 0xfffffffffffffull) != 0xfffffffffffffull;
      tab.rs[ch] = ok ? 1.0 / sg : std::nan("");
    }
  }
  return tab;
}
