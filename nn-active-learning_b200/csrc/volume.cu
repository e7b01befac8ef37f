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
// ------------------------------------------------------------------------------------------
template <typename TV, typename TO>
__global__ void __launch_bounds__(256) gather_kernel(const TV* __restrict__ vol, int m, int64_t Xp, int64_t Yp,
                                                      int64_t Zp, const int64_t* __restrict__ inds, int64_t n, int d1,
                                                      int d2, int d3, const double* __restrict__ stats, int norm_mode,
                                                      TO* __restrict__ out) {
  const int C = m * d3;
  const int rowlen = d2 * C;
  const int per_patch = d1 * rowlen;
  const int64_t X0 = Xp - (d1 - 1), Y0 = Yp - (d2 - 1), Z0 = Zp - (d3 - 1);
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    int64_t ind = inds[p];
    // np.unravel_index(ind, orig_shape) (patch_utils.py:1144); host validated the range
    int64_t z = ind % Z0;
    int64_t t = ind / Z0;
    int64_t y = t % Y0;
    int64_t x = t / Y0;
    TO* dst = out + p * (int64_t)per_patch;
    for (int e = threadIdx.x; e < per_patch; e += blockDim.x) {
      int i = e / rowlen;
      int r = e - i * rowlen;
      int k = r / C;
      int ch = r - k * C;
      int j, dz;
      if (d3 == 1) { j = ch; dz = 0; } else { j = ch / d3; dz = ch - j * d3; }
      double v = (double)vol[(((z + dz) * Xp + (x + i)) * Yp + (y + k)) * m + j];
      if (norm_mode == 1) {
        if (ch < m) v = (v - stats[2 * ch]) / stats[2 * ch + 1];
      } else if (norm_mode == 2) {
        v = (v - stats[2 * j]) / stats[2 * j + 1];
      }
      dst[e] = (TO)v;
    }
  }
}

template <typename TO>
static int launch_gather(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                         const double* d_stats, int norm_mode, TO* d_out) {
  if (n == 0) return NNAL_OK;
  int grid = (int)(n < (int64_t)ctx->sm_count * 32 ? n : (int64_t)ctx->sm_count * 32);
  if (v.dtype == NNAL_F64)
    gather_kernel<double, TO><<<grid, 256, 0, ctx->stream>>>((const double*)v.data, v.m, v.X, v.Y, v.Z, d_inds, n, d1,
                                                              d2, d3, d_stats, norm_mode, d_out);
  else
    gather_kernel<float, TO><<<grid, 256, 0, ctx->stream>>>((const float*)v.data, v.m, v.X, v.Y, v.Z, d_inds, n, d1,
                                                             d2, d3, d_stats, norm_mode, d_out);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

int nnal_k_gather_f64(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                      const double* d_stats, int norm_mode, double* d_out) {
  return launch_gather<double>(ctx, v, d_inds, n, d1, d2, d3, d_stats, norm_mode, d_out);
}

int nnal_k_gather_norm_f32(nnal_ctx* ctx, const Volume& v, const int64_t* d_inds, int64_t n, int d1, int d2, int d3,
                           const double* d_stats, int norm_mode, float* d_out) {
  return launch_gather<float>(ctx, v, d_inds, n, d1, d2, d3, d_stats, norm_mode, d_out);
}
