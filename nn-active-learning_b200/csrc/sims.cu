// Representativeness queries over the feature layer (sm_100a): cosine-similarity GEMMs on the tcgen05 GEMM of
// gemm_tc.cu + greedy selections.  SURVEY.md §8(f) rank 1.
//
// Reference behaviour replaced:
//   * 'rep-entropy' (NNAL.py:466-523, PW_NNAL.py:284-351): sims = cos(F_rem, F_uncertain) ((n-B) x B), then k greedy
//     steps of facility location: argmax_j sum_rows max(cur_row, sims[row][j]).
//   * 'core-set' (PW_NNAL.py:353-451): sims_i = max_j cos(F_T[:,j], F_u[:,i]) over the labeled features, then k steps
//     of k-center: q = argmin(sims); sims = max(sims, cos(F_u[:,q], F_u)); sims[q] = inf.
//   * PW_NNAL.get_cross_sims / get_self_sims (PW_NNAL.py:1041-1136) are the same row-max of a cosine GEMM.
//
// Rows are the samples of the current pool pass (nnal_pool_begin keep >= 1).  Feature rows are normalised to unit
// length (float64 norm, float32 result), scaled by 2^10 and split into fp16 hi/lo planes [rows][d] -- already
// K-major for both GEMM operands, so no transpose is needed.  Rows with a zero norm (a dead feature vector: the
// reference divides 0/0 there) take no part in either selection.
#include "nnal_common.cuh"
#include "dots.cuh"
#include "../../include/nnal_b200.h"
#include <algorithm>
#include <cmath>

namespace sims {

constexpr float PLANE_SCALE = 1024.f;                 // operands scaled by 2^10: lo terms of unit vectors stay normal
constexpr int64_t ROW_CHUNK = 32768;                  // rows per GEMM call

struct RepMsgHeader { double sim; long long gid; double inorm; double reserved; };

struct State {
  // shared
  int d = 0;
  int64_t n = 0;                                      // rows = pool samples of the current pass
  float* inorm = nullptr;                             // [n] 1/|f_i| (0 for zero rows)
  int64_t inorm_cap = 0;
  nnal_h *Ah = nullptr, *Al = nullptr;                // row planes of one chunk
  size_t a_cap = 0;
  nnal_h *Bh = nullptr, *Bl = nullptr;                // column planes
  size_t b_cap = 0;
  float* cols_raw = nullptr;                          // uploaded column features (fp32)
  size_t cols_cap = 0;
  // rep-entropy
  float* S = nullptr;                                 // [n][ldB] cosine similarities
  size_t s_cap = 0;
  int64_t B = 0, ldB = 0;
  unsigned char* rowmask = nullptr;                   // [n] 1 = contributes
  int64_t mask_cap = 0;
  float* cur = nullptr;                               // [n] running row maximum
  int64_t cur_cap = 0;
  unsigned char* taken = nullptr;                     // [B]
  double* partial = nullptr;                          // [nblk][B]
  size_t part_cap = 0;
  int nblk = 0;
  double* scores = nullptr;                           // [B]
  int64_t score_cap = 0;
  long long* sel = nullptr;
  double* selval = nullptr;
  int64_t sel_cap = 0;
  // core-set
  double* cs = nullptr;                               // [n] running max similarity to labeled/selected
  int64_t cs_cap = 0;
  double* blk_val = nullptr;
  long long* blk_idx = nullptr;
  int blk_cap = 0;
  long long* gids = nullptr;
  int64_t gids_cap = 0;
  bool use_gids = false;
  float* win = nullptr;                               // [d] winner feature row
  size_t win_cap = 0;
  double* win_sc = nullptr;                           // [2] winner inorm, best value
  long long* best_idx = nullptr;                      // [1]
  float* rowmax = nullptr;                            // [ROW_CHUNK] scratch
};

static State* get(nnal_ctx* ctx) {
  if (!ctx->sims_state) ctx->sims_state = new State();
  return (State*)ctx->sims_state;
}

template <typename T>
static int grow(nnal_ctx* ctx, T*& p, size_t& cap, size_t want) {
  if (p && cap >= want) return NNAL_OK;
  if (p) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); CUDA_TRY(ctx, cudaFree(p)); p = nullptr; cap = 0; }
  CUDA_TRY(ctx, cudaMalloc(&p, std::max<size_t>(want, 1) * sizeof(T)));
  cap = want;
  return NNAL_OK;
}
template <typename T>
static int grow(nnal_ctx* ctx, T*& p, int64_t& cap, int64_t want) {
  size_t c = (size_t)cap;
  int rc = grow(ctx, p, c, (size_t)want);
  cap = (int64_t)c;
  return rc;
}
template <typename T>
static int grow(nnal_ctx* ctx, T*& p, int& cap, int want) {
  size_t c = (size_t)cap;
  int rc = grow(ctx, p, c, (size_t)want);
  cap = (int)c;
  return rc;
}

// unit-normalise + scale + fp16 hi/lo split of feature rows; one warp per row
__global__ void __launch_bounds__(256) normalize_split_kernel(const float* __restrict__ F, int64_t n, int d, float scale,
                                                               nnal_h* __restrict__ hi, nnal_h* __restrict__ lo,
                                                               float* __restrict__ inorm_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const float* f = F + i * d;
    double uu, dummy;
    dot_um(f, f, nullptr, d, false, lane, uu, dummy);
    uu = warp_sum(uu);
    const float inv = uu > 0.0 ? (float)(1.0 / sqrt(uu)) : 0.f;
    if (lane == 0 && inorm_out) inorm_out[i] = inv;
    const float s = inv * scale;
    for (int k = lane * 2; k < d; k += 64) {
      nnal_h h0, l0, h1, l1;
      nnal_split(f[k] * s, h0, l0);
      nnal_split(k + 1 < d ? f[k + 1] * s : 0.f, h1, l1);
      *reinterpret_cast<uint32_t*>(hi + i * d + k) = nnal_pack2(h0, h1);
      *reinterpret_cast<uint32_t*>(lo + i * d + k) = nnal_pack2(l0, l1);
    }
  }
}

static int normalize_split(nnal_ctx* ctx, const float* F, int64_t n, int d, nnal_h* hi, nnal_h* lo, float* inorm) {
  if (n == 0) return NNAL_OK;
  int64_t blocks = (n + 7) / 8;
  int grid = (int)std::min<int64_t>(blocks, (int64_t)ctx->sm_count * 16);
  normalize_split_kernel<<<grid, 256, 0, ctx->stream>>>(F, n, d, PLANE_SCALE, hi, lo, inorm);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

// ---- rep-entropy ---------------------------------------------------------------------------------------
__global__ void rep_init_kernel(const float* __restrict__ inorm, const unsigned char* __restrict__ excl, int64_t n,
                                unsigned char* __restrict__ rowmask, float* __restrict__ cur) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    rowmask[i] = (inorm[i] > 0.f && !(excl && excl[i])) ? 1 : 0;
    cur[i] = -INFINITY;
  }
}
__global__ void mark_rows_kernel(const int64_t* __restrict__ pos, int64_t m, unsigned char* __restrict__ flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) flags[pos[i]] = 1;
}

// partial[blk][j] = sum over the block's rows of max(cur[row], S[row][j]): thread = column, rows in float32 runs of
// 32 flushed into a float64 sum
__global__ void __launch_bounds__(256) rep_partial_kernel(const float* __restrict__ S, int64_t ldB, int64_t B, int64_t n,
                                                           const unsigned char* __restrict__ rowmask, const float* __restrict__ cur,
                                                           int64_t rows_per_blk, double* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_blk;
  const int64_t r1 = r0 + rows_per_blk < n ? r0 + rows_per_blk : n;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < B; j += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int64_t r = r0; r < r1; r += 32) {
      float f = 0.f;
      const int64_t re = r + 32 < r1 ? r + 32 : r1;
      for (int64_t q = r; q < re; ++q)
        if (rowmask[q]) f += fmaxf(cur[q], S[q * ldB + j]);
      acc += (double)f;
    }
    partial[(int64_t)blockIdx.y * B + j] = acc;
  }
}
__global__ void __launch_bounds__(256) rep_reduce_kernel(const double* __restrict__ partial, int nblk, int64_t B,
                                                          const unsigned char* __restrict__ taken, double* __restrict__ scores) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < B; j += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * B + j];
    scores[j] = taken[j] ? -INFINITY : s;
  }
}
// arg-max over the (all-reduced) scores (ties: lowest column), one CTA
__global__ void __launch_bounds__(1024) rep_pick_kernel(const double* __restrict__ scores, int64_t B, int t,
                                                         unsigned char* __restrict__ taken, long long* __restrict__ sel,
                                                         double* __restrict__ selval, long long* __restrict__ best) {
  double v = -INFINITY;
  long long idx = 0x7fffffffffffffffll;
  for (int64_t j = threadIdx.x; j < B; j += blockDim.x) {
    const double s = taken[j] ? -INFINITY : scores[j];
    if (s > v || (s == v && j < idx)) { v = s; idx = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __shared__ double wv[32];
  __shared__ long long wi[32];
  if ((threadIdx.x & 31) == 0) { wv[threadIdx.x >> 5] = v; wi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
      if (wv[q] > v || (wv[q] == v && wi[q] < idx)) { v = wv[q]; idx = wi[q]; }
    if (!(v > -INFINITY)) idx = -1;
    sel[t] = idx;
    selval[t] = v;
    *best = idx;
    if (idx >= 0) taken[idx] = 1;
  }
}
__global__ void rep_update_kernel(const float* __restrict__ S, int64_t ldB, int64_t n, const long long* __restrict__ best,
                                  float* __restrict__ cur) {
  const long long j = *best;
  if (j < 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    cur[i] = fmaxf(cur[i], S[i * ldB + j]);
}

// ---- core-set ---------------------------------------------------------------------------------------
__global__ void cs_init_kernel(const float* __restrict__ inorm, int64_t n, double* __restrict__ cs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    cs[i] = inorm[i] > 0.f ? -INFINITY : INFINITY;          // zero rows are never selected
}
__global__ void cs_merge_kernel(const double* __restrict__ src, int64_t n, double* __restrict__ cs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (cs[i] < INFINITY) cs[i] = src[i];
}
// cs[row0+r] = max(cs, max_j G[r][j]) over one GEMM chunk (one warp per row)
__global__ void __launch_bounds__(256) cs_rowmax_kernel(const float* __restrict__ G, int64_t ld, int64_t nT, int64_t rows,
                                                         double* __restrict__ cs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float m = -INFINITY;
    for (int64_t j = lane; j < nT; j += 32) m = fmaxf(m, G[r * ld + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && cs[r] < INFINITY) cs[r] = fmax(cs[r], (double)m);
  }
}
__global__ void __launch_bounds__(256) cs_argmin_kernel(const double* __restrict__ cs, int64_t n, double* __restrict__ blk_val,
                                                         long long* __restrict__ blk_idx) {
  double v = INFINITY;
  long long idx = 0x7fffffffffffffffll;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double s = cs[i];
    if (s < v || (s == v && i < idx)) { v = s; idx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __shared__ double wv[8];
  __shared__ long long wi[8];
  if ((threadIdx.x & 31) == 0) { wv[threadIdx.x >> 5] = v; wi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q)
      if (wv[q] < v || (wv[q] == v && wi[q] < idx)) { v = wv[q]; idx = wi[q]; }
    blk_val[blockIdx.x] = v;
    blk_idx[blockIdx.x] = idx;
  }
}
// final arg-min + message: [sim | gid | inorm | - | feature row (d floats)]
__global__ void __launch_bounds__(256) cs_pack_kernel(const double* __restrict__ blk_val, const long long* __restrict__ blk_idx,
                                                       int nblk, const float* __restrict__ F, const float* __restrict__ inorm,
                                                       const long long* __restrict__ gids, int d, long long* __restrict__ best,
                                                       unsigned char* __restrict__ msg) {
  __shared__ long long s_idx;
  __shared__ double s_val;
  if (threadIdx.x == 0) {
    double v = INFINITY;
    long long idx = 0x7fffffffffffffffll;
    for (int b = 0; b < nblk; ++b)
      if (blk_val[b] < v || (blk_val[b] == v && blk_idx[b] < idx)) { v = blk_val[b]; idx = blk_idx[b]; }
    if (!(v < INFINITY)) idx = -1;
    s_idx = idx; s_val = v;
    *best = idx;
  }
  __syncthreads();
  RepMsgHeader* h = reinterpret_cast<RepMsgHeader*>(msg);
  float* fu = reinterpret_cast<float*>(msg + sizeof(RepMsgHeader));
  const long long i = s_idx;
  if (i < 0) {
    if (threadIdx.x == 0) { h->sim = INFINITY; h->gid = 0x7fffffffffffffffll; h->inorm = 0.0; h->reserved = 0.0; }
    return;
  }
  for (int k = threadIdx.x; k < d; k += blockDim.x) fu[k] = F[i * d + k];
  if (threadIdx.x == 0) { h->sim = s_val; h->gid = gids ? gids[i] : i; h->inorm = (double)inorm[i]; h->reserved = 0.0; }
}
__global__ void __launch_bounds__(256) cs_select_kernel(const unsigned char* __restrict__ msgs, size_t msg_bytes, int world, int rank,
                                                         int t, int d, const long long* __restrict__ best, double* __restrict__ cs,
                                                         float* __restrict__ win, double* __restrict__ win_sc,
                                                         long long* __restrict__ sel, double* __restrict__ selval) {
  int bw = 0;
  double bv = INFINITY;
  long long bg = 0x7fffffffffffffffll;
  for (int r = 0; r < world; ++r) {
    const RepMsgHeader* h = reinterpret_cast<const RepMsgHeader*>(msgs + (size_t)r * msg_bytes);
    if (h->sim < bv || (h->sim == bv && h->gid < bg)) { bv = h->sim; bg = h->gid; bw = r; }
  }
  const unsigned char* m = msgs + (size_t)bw * msg_bytes;
  const RepMsgHeader* h = reinterpret_cast<const RepMsgHeader*>(m);
  const float* fu = reinterpret_cast<const float*>(m + sizeof(RepMsgHeader));
  for (int k = threadIdx.x; k < d; k += blockDim.x) win[k] = fu[k];
  if (threadIdx.x == 0) {
    win_sc[0] = h->inorm;
    sel[t] = bv < INFINITY ? bg : -1;
    selval[t] = bv;
    if (bw == rank && bv < INFINITY) cs[*best] = INFINITY;      // the winner leaves the pool (PW_NNAL.py:448)
  }
}
// cs[i] = max(cs[i], cos(f_i, f_win)) for every row still in play; one warp per row
__global__ void __launch_bounds__(256) cs_update_kernel(const float* __restrict__ F, const float* __restrict__ inorm, int64_t n, int d,
                                                         const float* __restrict__ win, const double* __restrict__ win_sc,
                                                         double* __restrict__ cs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double wn = win_sc[0];
  for (int64_t i = warp; i < n; i += nwarps) {
    if (!(cs[i] < INFINITY)) continue;
    double uu, dummy;
    dot_um(F + i * d, win, nullptr, d, false, lane, uu, dummy);
    uu = warp_sum(uu);
    if (lane == 0) cs[i] = fmax(cs[i], uu * (double)inorm[i] * wn);
  }
}

static int check_pool(nnal_ctx* ctx) {
  if (!ctx->pool_feat || ctx->keep < 1) NNAL_FAIL(ctx, NNAL_ERR_STATE, "pool pass did not keep the feature layer");
  if (ctx->feat_dim % 8 != 0) NNAL_FAIL(ctx, NNAL_ERR_UNSUPPORTED, "similarity GEMM needs a feature width that is a multiple of 8");
  return NNAL_OK;
}

// uploads column features [m][d], builds their planes; returns planes in s->Bh/Bl
static int set_columns(nnal_ctx* ctx, State* s, const float* cols, int64_t m) {
  const int d = s->d;
  NNAL_TRY(grow(ctx, s->cols_raw, s->cols_cap, (size_t)m * d));
  size_t bc = s->b_cap;
  NNAL_TRY(grow(ctx, s->Bh, bc, (size_t)m * d));
  NNAL_TRY(grow(ctx, s->Bl, s->b_cap, (size_t)m * d));
  if (m) CUDA_TRY(ctx, cudaMemcpyAsync(s->cols_raw, cols, (size_t)m * d * 4, cudaMemcpyHostToDevice, ctx->stream));
  return normalize_split(ctx, s->cols_raw, m, d, s->Bh, s->Bl, nullptr);
}

// G[rows r0..r0+nr) x m] = cos-sims of pool rows against the current column planes, into out (row stride ld)
static int gemm_rows(nnal_ctx* ctx, State* s, int64_t r0, int64_t nr, int64_t m, float* out, int64_t ld) {
  const int d = s->d;
  size_t ac = s->a_cap;
  NNAL_TRY(grow(ctx, s->Ah, ac, (size_t)ROW_CHUNK * d));
  NNAL_TRY(grow(ctx, s->Al, s->a_cap, (size_t)ROW_CHUNK * d));
  NNAL_TRY(normalize_split(ctx, ctx->pool_feat + r0 * d, nr, d, s->Ah, s->Al, s->inorm + r0));
  return nnal_tc_gemm_planes(ctx, s->Ah, s->Al, d, nr, s->Bh, s->Bl, d, (int)m, d, nullptr, 1.f / (PLANE_SCALE * PLANE_SCALE), 0, 0, out,
                             (int)ld, nullptr, nullptr, 0);
}

}  // namespace sims

using sims::State;

int nnal_sims_release(nnal_ctx* ctx) {
  if (!ctx->sims_state) return NNAL_OK;
  State* s = (State*)ctx->sims_state;
  void* ptrs[] = {s->inorm, s->Ah, s->Al, s->Bh, s->Bl, s->cols_raw, s->S, s->rowmask, s->cur, s->taken, s->partial, s->scores,
                  s->sel, s->selval, s->cs, s->blk_val, s->blk_idx, s->gids, s->win, s->win_sc, s->best_idx, s->rowmax};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete s;
  ctx->sims_state = nullptr;
  return NNAL_OK;
}

static int sims_common_begin(nnal_ctx* ctx, State* s, int64_t k) {
  s->d = ctx->feat_dim;
  s->n = ctx->pool_n;
  NNAL_TRY(sims::grow(ctx, s->inorm, s->inorm_cap, std::max<int64_t>(s->n, 1)));
  if (s->sel_cap < k) {
    int64_t c = s->sel_cap;
    NNAL_TRY(sims::grow(ctx, s->sel, c, std::max<int64_t>(k, 1)));
    NNAL_TRY(sims::grow(ctx, s->selval, s->sel_cap, std::max<int64_t>(k, 1)));
  }
  if (!s->best_idx) { CUDA_TRY(ctx, cudaMalloc(&s->best_idx, 8)); CUDA_TRY(ctx, cudaMalloc(&s->win_sc, 16)); }
  return NNAL_OK;
}

// ---- rep-entropy --------------------------------------------------------------------------------------
extern "C" int nnal_rep_set(nnal_ctx* ctx, const float* cols, int64_t B, const int64_t* excl_pos, int64_t n_excl, int64_t k) {
  if (!ctx || B < 0 || n_excl < 0 || k < 0 || (B > 0 && !cols) || (n_excl > 0 && !excl_pos)) return NNAL_ERR_INVALID;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  NNAL_TRY(sims::check_pool(ctx));
  for (int64_t i = 0; i < n_excl; ++i)
    if (excl_pos[i] < 0 || excl_pos[i] >= ctx->pool_n) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "excluded position outside the pool");
  State* s = sims::get(ctx);
  NNAL_TRY(sims_common_begin(ctx, s, k));
  const int64_t n = s->n;
  s->B = B;
  s->ldB = (B + 7) / 8 * 8;
  NNAL_TRY(sims::grow(ctx, s->S, s->s_cap, (size_t)std::max<int64_t>(n, 1) * std::max<int64_t>(s->ldB, 8)));
  NNAL_TRY(sims::grow(ctx, s->rowmask, s->mask_cap, std::max<int64_t>(n, 1)));
  NNAL_TRY(sims::grow(ctx, s->cur, s->cur_cap, std::max<int64_t>(n, 1)));
  {
    int64_t c = s->score_cap;
    NNAL_TRY(sims::grow(ctx, s->taken, c, std::max<int64_t>(B, 1)));
    NNAL_TRY(sims::grow(ctx, s->scores, s->score_cap, std::max<int64_t>(B, 1)));
  }
  s->nblk = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 2));
  NNAL_TRY(sims::grow(ctx, s->partial, s->part_cap, (size_t)s->nblk * std::max<int64_t>(B, 1)));
  CUDA_TRY(ctx, cudaMemsetAsync(s->taken, 0, (size_t)std::max<int64_t>(B, 1), ctx->stream));
  NNAL_TRY(sims::set_columns(ctx, s, cols, B));
  // excluded rows: flag them in the row mask (reuse `cur` memory? no: a byte array of its own)
  unsigned char* excl = nullptr;
  if (n_excl) {
    NNAL_TRY(devbuf_reserve(ctx, ctx->fi_ws, (size_t)n + (size_t)n_excl * 8 + 16));
    excl = (unsigned char*)ctx->fi_ws.p;
    int64_t* d_pos = (int64_t*)((char*)ctx->fi_ws.p + ((n + 15) / 16) * 16);
    CUDA_TRY(ctx, cudaMemsetAsync(excl, 0, (size_t)n, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_pos, excl_pos, (size_t)n_excl * 8, cudaMemcpyHostToDevice, ctx->stream));
    sims::mark_rows_kernel<<<cdiv(n_excl, 256), 256, 0, ctx->stream>>>(d_pos, n_excl, excl);
    ctx->launches++;
  }
  for (int64_t r0 = 0; r0 < n && B > 0; r0 += sims::ROW_CHUNK) {
    const int64_t nr = std::min(sims::ROW_CHUNK, n - r0);
    NNAL_TRY(sims::gemm_rows(ctx, s, r0, nr, B, s->S + r0 * s->ldB, s->ldB));
  }
  if (B == 0 && n) {                                   // still need the norms for the mask
    size_t ac = s->a_cap;
    NNAL_TRY(sims::grow(ctx, s->Ah, ac, (size_t)sims::ROW_CHUNK * s->d));
    NNAL_TRY(sims::grow(ctx, s->Al, s->a_cap, (size_t)sims::ROW_CHUNK * s->d));
    for (int64_t r0 = 0; r0 < n; r0 += sims::ROW_CHUNK)
      NNAL_TRY(sims::normalize_split(ctx, ctx->pool_feat + r0 * s->d, std::min(sims::ROW_CHUNK, n - r0), s->d, s->Ah, s->Al, s->inorm + r0));
  }
  if (n) {
    sims::rep_init_kernel<<<std::min(cdiv(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(s->inorm, excl, n, s->rowmask, s->cur);
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // cols / excl_pos are caller-owned
  return NNAL_OK;
}

// local partial scores of every column into d_scores [B] (device, float64): the host layer all-reduces them (sum)
extern "C" int nnal_rep_step_scores(nnal_ctx* ctx, double* d_scores) {
  if (!ctx || !d_scores) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_rep_set not called");
  State* s = (State*)ctx->sims_state;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (s->B == 0) return NNAL_OK;
  const int64_t rows_per_blk = std::max<int64_t>(1, (s->n + s->nblk - 1) / s->nblk);
  dim3 grid((unsigned)std::min<int64_t>((s->B + 255) / 256, 65535), (unsigned)s->nblk);
  sims::rep_partial_kernel<<<grid, 256, 0, ctx->stream>>>(s->S, s->ldB, s->B, s->n, s->rowmask, s->cur, rows_per_blk, s->partial);
  sims::rep_reduce_kernel<<<std::min(cdiv(s->B, 256), 1024), 256, 0, ctx->stream>>>(s->partial, s->nblk, s->B, s->taken, d_scores);
  ctx->launches += 2;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

extern "C" int nnal_rep_step_pick(nnal_ctx* ctx, int64_t step, const double* d_scores) {
  if (!ctx || !d_scores || step < 0) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_rep_set not called");
  State* s = (State*)ctx->sims_state;
  if (step >= s->sel_cap) NNAL_FAIL(ctx, NNAL_ERR_STATE, "more steps than announced to nnal_rep_set");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  sims::rep_pick_kernel<<<1, 1024, 0, ctx->stream>>>(d_scores, s->B, (int)step, s->taken, s->sel, s->selval, s->best_idx);
  ctx->launches++;
  if (s->n) {
    sims::rep_update_kernel<<<std::min(cdiv(s->n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(s->S, s->ldB, s->n, s->best_idx, s->cur);
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

extern "C" int nnal_sel_result(nnal_ctx* ctx, int64_t k, int64_t* sel_out, double* val_out) {
  if (!ctx || k < 0 || !sel_out) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "no selection run");
  State* s = (State*)ctx->sims_state;
  if (k > s->sel_cap) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "more steps requested than were run");
  if (k == 0) return NNAL_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaMemcpyAsync(sel_out, s->sel, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (val_out) CUDA_TRY(ctx, cudaMemcpyAsync(val_out, s->selval, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return NNAL_OK;
}

extern "C" int nnal_rep_greedy(nnal_ctx* ctx, int64_t k, int64_t* sel_out, double* val_out) {
  if (!ctx || k < 0 || !sel_out) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_rep_set not called");
  State* s = (State*)ctx->sims_state;
  if (k > s->B) k = s->B;
  for (int64_t t = 0; t < k; ++t) {
    NNAL_TRY(nnal_rep_step_scores(ctx, s->scores));
    NNAL_TRY(nnal_rep_step_pick(ctx, t, s->scores));
  }
  return nnal_sel_result(ctx, k, sel_out, val_out);
}

// row-wise maximum cosine similarity of the pool rows to a set of feature vectors (get_cross_sims, PW_NNAL.py:1093-1136)
extern "C" int nnal_cross_sims(nnal_ctx* ctx, const float* F2, int64_t n2, double* out) {
  if (!ctx || n2 <= 0 || !F2 || !out) return NNAL_ERR_INVALID;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  NNAL_TRY(sims::check_pool(ctx));
  State* s = sims::get(ctx);
  NNAL_TRY(sims_common_begin(ctx, s, 1));
  const int64_t n = s->n;
  NNAL_TRY(sims::grow(ctx, s->cs, s->cs_cap, std::max<int64_t>(n, 1)));
  if (n == 0) return NNAL_OK;
  // first pass fills inorm (needed by cs_init), so run the GEMMs first into a scratch and take the row max after
  const int64_t CB = 2048;                                // labeled columns per GEMM
  const int64_t ld = CB;
  NNAL_TRY(devbuf_reserve(ctx, ctx->fi_ws, (size_t)sims::ROW_CHUNK * ld * 4));
  bool first = true;
  for (int64_t c0 = 0; c0 < n2; c0 += CB) {
    const int64_t m = std::min(CB, n2 - c0);
    NNAL_TRY(sims::set_columns(ctx, s, F2 + c0 * s->d, m));
    for (int64_t r0 = 0; r0 < n; r0 += sims::ROW_CHUNK) {
      const int64_t nr = std::min(sims::ROW_CHUNK, n - r0);
      NNAL_TRY(sims::gemm_rows(ctx, s, r0, nr, m, (float*)ctx->fi_ws.p, ld));
      if (first) {
        sims::cs_init_kernel<<<std::min(cdiv(nr, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(s->inorm + r0, nr, s->cs + r0);
        ctx->launches++;
      }
      sims::cs_rowmax_kernel<<<std::min(cdiv(nr, 8), ctx->sm_count * 16), 256, 0, ctx->stream>>>((const float*)ctx->fi_ws.p, ld, m, nr, s->cs + r0);
      ctx->launches++;
    }
    first = false;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));     // F2 chunk consumed; column planes are rewritten next
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(out, s->cs, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return NNAL_OK;
}

// ---- core-set ------------------------------------------------------------------------------------------
// Starts a k-center selection over the rows of the current pool pass.  init = 0: similarities start at -inf (no
// labeled set); 1: from the host array sims0 [n]; 2: from the device result of the last nnal_cross_sims (the
// labeled-set pass of PW_NNAL.py:399-425).  gids: global ids of the local rows (NULL: local index).
extern "C" int nnal_cs_begin(nnal_ctx* ctx, int init, const double* sims0, const int64_t* gids, int64_t k) {
  if (!ctx || k < 0 || init < 0 || init > 2 || (init == 1 && !sims0)) return NNAL_ERR_INVALID;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  NNAL_TRY(sims::check_pool(ctx));
  State* s = sims::get(ctx);
  if (init == 2 && (s->n != ctx->pool_n || !s->cs)) NNAL_FAIL(ctx, NNAL_ERR_STATE, "no nnal_cross_sims result for this pool pass");
  NNAL_TRY(sims_common_begin(ctx, s, k));
  const int64_t n = s->n;
  NNAL_TRY(sims::grow(ctx, s->cs, s->cs_cap, std::max<int64_t>(n, 1)));
  NNAL_TRY(sims::grow(ctx, s->win, s->win_cap, (size_t)s->d));
  s->nblk = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 4));
  {
    int c = s->blk_cap;
    NNAL_TRY(sims::grow(ctx, s->blk_val, c, s->nblk));
    NNAL_TRY(sims::grow(ctx, s->blk_idx, s->blk_cap, s->nblk));
  }
  s->use_gids = gids != nullptr;
  if (gids && n) {
    NNAL_TRY(sims::grow(ctx, s->gids, s->gids_cap, n));
    CUDA_TRY(ctx, cudaMemcpyAsync(s->gids, gids, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (init != 2 && n) {
    size_t ac = s->a_cap;
    NNAL_TRY(sims::grow(ctx, s->Ah, ac, (size_t)sims::ROW_CHUNK * s->d));
    NNAL_TRY(sims::grow(ctx, s->Al, s->a_cap, (size_t)sims::ROW_CHUNK * s->d));
    for (int64_t r0 = 0; r0 < n; r0 += sims::ROW_CHUNK)
      NNAL_TRY(sims::normalize_split(ctx, ctx->pool_feat + r0 * s->d, std::min(sims::ROW_CHUNK, n - r0), s->d, s->Ah, s->Al, s->inorm + r0));
    sims::cs_init_kernel<<<std::min(cdiv(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(s->inorm, n, s->cs);
    ctx->launches++;
    if (init == 1) {
      // host similarities override -inf where the row is in play (zero rows stay excluded)
      NNAL_TRY(devbuf_reserve(ctx, ctx->fi_ws, (size_t)n * 8));
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->fi_ws.p, sims0, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
      sims::cs_merge_kernel<<<std::min(cdiv(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>((const double*)ctx->fi_ws.p, n, s->cs);
      ctx->launches++;
    }
  }
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return NNAL_OK;
}

extern "C" int nnal_cs_msg_bytes(nnal_ctx* ctx, int64_t* bytes) {
  if (!ctx || !bytes) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_cs_begin not called");
  State* s = (State*)ctx->sims_state;
  *bytes = (int64_t)((sizeof(sims::RepMsgHeader) + (size_t)s->d * 4 + 15) / 16 * 16);
  return NNAL_OK;
}

extern "C" int nnal_cs_step_pack(nnal_ctx* ctx, int64_t step, void* d_msg) {
  if (!ctx || !d_msg || step < 0) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_cs_begin not called");
  State* s = (State*)ctx->sims_state;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (s->n > 0) {
    sims::cs_argmin_kernel<<<s->nblk, 256, 0, ctx->stream>>>(s->cs, s->n, s->blk_val, s->blk_idx);
    ctx->launches++;
  }
  sims::cs_pack_kernel<<<1, 256, 0, ctx->stream>>>(s->blk_val, s->blk_idx, s->n > 0 ? s->nblk : 0, ctx->pool_feat, s->inorm,
                                                  s->use_gids ? s->gids : nullptr, s->d, s->best_idx, (unsigned char*)d_msg);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

extern "C" int nnal_cs_step_apply_gathered(nnal_ctx* ctx, int64_t step, const void* d_msgs, int world, int rank) {
  if (!ctx || !d_msgs || step < 0 || world <= 0 || rank < 0 || rank >= world) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_cs_begin not called");
  State* s = (State*)ctx->sims_state;
  if (step >= s->sel_cap) NNAL_FAIL(ctx, NNAL_ERR_STATE, "more steps than announced to nnal_cs_begin");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int64_t mb;
  NNAL_TRY(nnal_cs_msg_bytes(ctx, &mb));
  sims::cs_select_kernel<<<1, 256, 0, ctx->stream>>>((const unsigned char*)d_msgs, (size_t)mb, world, rank, (int)step, s->d, s->best_idx,
                                                    s->cs, s->win, s->win_sc, s->sel, s->selval);
  ctx->launches++;
  if (s->n > 0) {
    sims::cs_update_kernel<<<std::min(cdiv(s->n, 8), ctx->sm_count * 16), 256, 0, ctx->stream>>>(ctx->pool_feat, s->inorm, s->n, s->d, s->win,
                                                                                                s->win_sc, s->cs);
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return NNAL_OK;
}

extern "C" int nnal_cs_greedy(nnal_ctx* ctx, int64_t k, int64_t* sel_out, double* val_out) {
  if (!ctx || k < 0 || !sel_out) return NNAL_ERR_INVALID;
  if (!ctx->sims_state) NNAL_FAIL(ctx, NNAL_ERR_STATE, "nnal_cs_begin not called");
  State* s = (State*)ctx->sims_state;
  if (k > s->n) k = s->n;
  int64_t mb;
  NNAL_TRY(nnal_cs_msg_bytes(ctx, &mb));
  NNAL_TRY(devbuf_reserve(ctx, ctx->fi_ws, (size_t)mb));
  for (int64_t t = 0; t < k; ++t) {
    NNAL_TRY(nnal_cs_step_pack(ctx, t, ctx->fi_ws.p));
    NNAL_TRY(nnal_cs_step_apply_gathered(ctx, t, ctx->fi_ws.p, 1, 0));
  }
  return nnal_sel_result(ctx, k, sel_out, val_out);
}
