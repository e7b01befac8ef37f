// Query distribution of the reference's FI query.
//
// NNAL_tools.SDP_query_distribution (NNAL_tools.py:612-659, constraints built by inequality_cvx_matrix :661-720; the cvxpy
// twin solve_FIAL_SDP :576-610) hands cvxopt the SDP
//     minimise sum_j t_j   s.t.  [[sum_i q_i A_i, e_j], [e_j^T, t_j]] >= 0 for all j,   q >= 0,  sum_i q_i = 1
// whose Schur complements make t_j = ((sum_i q_i A_i)^-1)_jj at the optimum: it minimises phi(q) = tr(M(q)^-1),
// M(q) = sum_i q_i A_i, over the simplex (A-optimal design).  tau is 7 for PW1, n = B is thousands: an interior-point
// method factors an n x n positivity block per iteration, while phi's gradient needs only the tau x tau inverse:
//     d_i = -d phi / d q_i = tr(M^-1 A_i M^-1) = <M^-2, A_i>,      sum_i q_i d_i = phi.
// Solved here on the device by the multiplicative algorithm for A-optimality  q_i <- q_i (d_i / phi)^gamma / Z
// (gamma = 1/2 decreases phi monotonically for positive semi-definite A_i; gamma = 1 usually halves the iteration count
// and is used until phi increases for the first time, then 1/2).  Convexity gives the certificate
//     phi(q) - phi* <= max_i d_i - phi(q),
// so the loop stops when max_i d_i / phi - 1 <= tol: the objective is then within tol (relative) of the SDP optimum.
// The matrices are symmetric: only the tau (tau + 1) / 2 entries a <= b are stored, summed and multiplied (off-diagonal
// entries weigh twice in <M^-2, A_i>).
// One kernel per iteration; every CTA rebuilds M from the previous iteration's per-CTA partial sums (fixed order:
// deterministic), inverts it in shared memory, updates its samples and writes the next partial sums.  All float64.
#include "nnal_common.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <vector>

namespace cg = cooperative_groups;

namespace {

constexpr int SDP_MAX_TAU = 16;
constexpr int SDP_T2 = SDP_MAX_TAU * SDP_MAX_TAU;
constexpr int SDP_THREADS = 256;
constexpr int SDP_MAX_GRID = 128;

struct SdpState {
  DevBuf At, stage, qu, part[2], out;
};

SdpState* sdp_state(nnal_ctx* ctx) {
  if (!ctx->sdp_state) ctx->sdp_state = new SdpState();
  return (SdpState*)ctx->sdp_state;
}

// layout of one set of per-CTA partial sums: [G][Tu] packed sums of q_i A_i, then [G] sums of q_i, then [G] max_i d_i / phi
struct Parts {
  double* M;
  double* Z;
  double* R;
};
__host__ __device__ inline Parts parts_of(double* base, int G, int Tu) {
  Parts p;
  p.M = base;
  p.Z = base + (size_t)G * Tu;
  p.R = p.Z + G;
  return p;
}

// packed upper-triangular index u <-> (a, b), a <= b, row-major: u = a tau - a (a - 1) / 2 + (b - a)
__host__ __device__ inline int sdp_packed(int tau) { return tau * (tau + 1) / 2; }
__device__ __forceinline__ void sdp_unpack(int u, int tau, int& a, int& b) {
  a = 0;
  while (u >= tau - a) { u -= tau - a; ++a; }
  b = a + u;
}

// dense host matrices [n][tau][tau] -> packed [Tu][n] (the mean of the two mirrored entries: the matrix itself when it
// is exactly symmetric)
__global__ void sdp_transpose_kernel(const double* __restrict__ A, double* __restrict__ At, int64_t n, int tau) {
  const int Tu = sdp_packed(tau), T2 = tau * tau;
  const int64_t total = n * Tu;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(e / n);
    const int64_t i = e - (int64_t)u * n;
    int a, b;
    sdp_unpack(u, tau, a, b);
    At[e] = 0.5 * (A[i * T2 + a * tau + b] + A[i * T2 + b * tau + a]);
  }
}

// A_i of PW_NNAL.gen_A_matrices (PW_NNAL.py:766-814) written straight into the solver's [tau*tau][n] layout:
// p < 1e-6 -> p = 0 and only g0, p > 1-1e-6 -> p = 1 and only g1, A_i = (1-p) g0 g0^T + p g1 g1^T + diag_load I,
// with the reference's order of operations ((1-p) * (g0_a * g0_b), then + p * (g1_a * g1_b), then + diag_load).
__global__ void sdp_binary_A_kernel(const double* __restrict__ g, const double* __restrict__ p1, int64_t n, int tau,
                                    double diag_load, double* __restrict__ At) {
  const int Tu = sdp_packed(tau);
  const int64_t total = n * Tu;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(e / n);
    const int64_t i = e - (int64_t)u * n;
    int a, b;
    sdp_unpack(u, tau, a, b);
    double p = p1[i];
    const bool lo = p < 1e-6, hi = p > 1.0 - 1e-6;
    if (lo) p = 0.0;
    if (hi) p = 1.0;
    const double* g0 = g + i * tau;
    const double* g1 = g + (n + i) * tau;
    const double o0 = hi ? 0.0 : g0[a] * g0[b];
    const double o1 = lo ? 0.0 : g1[a] * g1[b];
    double v = __dadd_rn(__dmul_rn(o0, 1.0 - p), __dmul_rn(o1, p));        // no FMA contraction: bit-identical to NumPy
    if (a == b) v = __dadd_rn(v, diag_load);
    At[e] = v;
  }
}

__global__ void sdp_fill_kernel(double* __restrict__ q, int64_t n, double v) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) q[e] = v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

// M = (sum of the partial sums) / Z; in-place Gauss-Jordan inverse (M is positive definite: no pivoting); P = Minv^2.
// All threads of the CTA call this; returns phi = tr(Minv), Z and the largest ratio of the previous iterate to every
// thread.  The G x (Tu + 2) partial values live in L2 (written by other CTAs): the reads are spread over all threads
// -- entry e of slice s sums the CTAs g = s, s + S, ... with four loads in flight -- because G dependent L2 round
// trips per entry were the longest stretch of an iteration.  Fixed order: every CTA gets bit-identical sums.
__device__ void sdp_build(const Parts in, int G, int tau, double* s_M, double* s_P, double* s_part, double& phi, double& Z,
                          double& rprev) {
  const int T2 = tau * tau, Tu = sdp_packed(tau), tid = threadIdx.x;
  const int NE = Tu + 2;                                   // entries: packed M (Tu), Z, R
  const int S = NE <= SDP_THREADS ? SDP_THREADS / NE : 1;  // slices
  for (int e = tid % NE, sl = tid / NE; sl < S && e < NE; e += SDP_THREADS) {   // (one trip unless NE > 256)
    const double* base = e < Tu ? in.M + e : (e == Tu ? in.Z : in.R);
    const size_t step = e < Tu ? (size_t)Tu : 1;
    const bool is_max = e == Tu + 1;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int g = sl;
    for (; g + 3 * S < G; g += 4 * S) {
      const double v0 = __ldcg(base + (size_t)g * step), v1 = __ldcg(base + (size_t)(g + S) * step);
      const double v2 = __ldcg(base + (size_t)(g + 2 * S) * step), v3 = __ldcg(base + (size_t)(g + 3 * S) * step);
      if (is_max) { a0 = fmax(a0, v0); a1 = fmax(a1, v1); a2 = fmax(a2, v2); a3 = fmax(a3, v3); }
      else { a0 += v0; a1 += v1; a2 += v2; a3 += v3; }
    }
    for (; g < G; g += S) {
      const double v = __ldcg(base + (size_t)g * step);
      if (is_max) a0 = fmax(a0, v); else a0 += v;
    }
    s_part[sl * NE + e] = is_max ? fmax(fmax(a0, a1), fmax(a2, a3)) : (a0 + a1) + (a2 + a3);
    if (NE <= SDP_THREADS) break;
  }
  __syncthreads();
  double z = 0.0, r = 0.0;
  for (int sl = 0; sl < S; ++sl) { z += s_part[sl * NE + Tu]; r = fmax(r, s_part[sl * NE + Tu + 1]); }
  Z = z;
  rprev = r;
  if (tid < Tu) {
    double m = 0.0;
    for (int sl = 0; sl < S; ++sl) m += s_part[sl * NE + tid];
    int pa, pb;
    sdp_unpack(tid, tau, pa, pb);
    m /= z;
    s_M[pa * tau + pb] = m;
    s_M[pb * tau + pa] = m;
  }
  __syncthreads();
  const int a = tid / tau, b = tid % tau;
  for (int k = 0; k < tau; ++k) {
    double v = 0.0;
    if (tid < T2) {
      const double piv = 1.0 / s_M[k * tau + k];
      if (a == k && b == k) v = piv;
      else if (a == k) v = s_M[k * tau + b] * piv;
      else if (b == k) v = -s_M[a * tau + k] * piv;
      else v = s_M[a * tau + b] - s_M[a * tau + k] * s_M[k * tau + b] * piv;
    }
    __syncthreads();
    if (tid < T2) s_M[tid] = v;
    __syncthreads();
  }
  if (tid < Tu) {                       // packed weights of <M^-2, A_i>: off-diagonal entries count twice
    int pa, pb;
    sdp_unpack(tid, tau, pa, pb);
    double v = 0.0;
    for (int k = 0; k < tau; ++k) v += s_M[pa * tau + k] * s_M[k * tau + pb];
    s_P[tid] = pa == pb ? v : 2.0 * v;
  }
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < tau; ++k) t += s_M[k * tau + k];
  phi = t;
}

// per-CTA partial sums of q_i A_i, q_i and the CTA's largest ratio
__device__ void sdp_accumulate(const double* __restrict__ At, const double* __restrict__ qu, int64_t n, int Tu, double rmax,
                               Parts out, double (*s_w)[SDP_T2], double* s_r) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, j0 = (int64_t)blockIdx.x * blockDim.x + tid;
  // eight entries at a time: their shuffle chains are independent, so the reduction is issue-bound, not latency-bound
  for (int ab0 = 0; ab0 < Tu; ab0 += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.0;
    for (int64_t j = j0; j < n; j += stride) {
      const double q = qu[j];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (ab0 + u < Tu) v[u] += q * At[(int64_t)(ab0 + u) * n + j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] += __shfl_down_sync(0xffffffffu, v[u], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (ab0 + u < Tu) s_w[w][ab0 + u] = v[u];
    }
  }
  double z = 0.0;
  for (int64_t j = j0; j < n; j += stride) z += qu[j];
  z = warp_sum(z);
  rmax = warp_max(rmax);
  if (lane == 0) { s_r[w] = z; s_r[8 + w] = rmax; }
  __syncthreads();
  if (tid < Tu) {
    double v = 0.0;
    for (int k = 0; k < SDP_THREADS / 32; ++k) v += s_w[k][tid];
    out.M[(size_t)blockIdx.x * Tu + tid] = v;
  }
  if (tid == 0) {
    double zz = 0.0, rr = 0.0;
    for (int k = 0; k < SDP_THREADS / 32; ++k) { zz += s_r[k]; rr = fmax(rr, s_r[8 + k]); }
    out.Z[blockIdx.x] = zz;
    out.R[blockIdx.x] = rr;
  }
}

__global__ void __launch_bounds__(SDP_THREADS) sdp_init_kernel(const double* __restrict__ At, const double* __restrict__ qu,
                                                                int64_t n, int tau, double* part_out) {
  __shared__ double s_w[SDP_THREADS / 32][SDP_T2];
  __shared__ double s_r[16];
  sdp_accumulate(At, qu, n, sdp_packed(tau), 0.0, parts_of(part_out, gridDim.x, sdp_packed(tau)), s_w, s_r);
}

// one multiplicative update; returns max_i d_i / phi of the PREVIOUS iterate (identical in every CTA)
__device__ double sdp_iterate(const double* __restrict__ At, double* __restrict__ qu, int64_t n, int tau, double gamma,
                              double* part_in, double* part_out, double* s_M, double* s_P, double (*s_w)[SDP_T2],
                              double* s_r, double& phi_out) {
  const int Tu = sdp_packed(tau), G = gridDim.x;
  const Parts in = parts_of(part_in, G, Tu);
  double phi, Z, rprev;
  sdp_build(in, G, tau, s_M, s_P, &s_w[0][0], phi, Z, rprev);      // s_w doubles as the scratch of the partial read
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double rmax = 0.0;
  const double inv_phi = 1.0 / phi, inv_Z = 1.0 / Z;
  for (int64_t j = j0; j < n; j += stride) {
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;      // four chains: a float64 FMA waits ~8 cycles for its predecessor
    int u = 0;
    for (; u + 3 < Tu; u += 4) {
      d0 += s_P[u] * At[(int64_t)u * n + j];
      d1 += s_P[u + 1] * At[(int64_t)(u + 1) * n + j];
      d2 += s_P[u + 2] * At[(int64_t)(u + 2) * n + j];
      d3 += s_P[u + 3] * At[(int64_t)(u + 3) * n + j];
    }
    for (; u < Tu; ++u) d0 += s_P[u] * At[(int64_t)u * n + j];
    const double d = (d0 + d1) + (d2 + d3);
    const double r = d * inv_phi;
    const double q = qu[j] * inv_Z;
    rmax = fmax(rmax, r);          // over ALL i: the optimality condition also binds where q_i has underflowed to 0
    qu[j] = q * (gamma == 0.5 ? sqrt(r) : pow(r, gamma));
  }
  sdp_accumulate(At, qu, n, Tu, rmax, parts_of(part_out, G, Tu), s_w, s_r);
  phi_out = phi;
  return rprev;
}

__global__ void __launch_bounds__(SDP_THREADS) sdp_iter_kernel(const double* __restrict__ At, double* __restrict__ qu,
                                                                int64_t n, int tau, double gamma, double* part_in,
                                                                double* part_out, double* __restrict__ hist) {
  __shared__ double s_M[SDP_T2], s_P[SDP_T2];
  __shared__ double s_w[SDP_THREADS / 32][SDP_T2];
  __shared__ double s_r[16];
  double phi;
  const double rprev = sdp_iterate(At, qu, n, tau, gamma, part_in, part_out, s_M, s_P, s_w, s_r, phi);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hist[0] = phi;            // phi(q_it)
    hist[1] = rprev;          // max_i d_i / phi of the PREVIOUS iterate (0 before the first update)
  }
}

// The whole loop in ONE cooperative launch: a grid-wide barrier between iterations instead of a kernel boundary
// (the loop is latency-bound: ~20 us per iteration as separate launches).  Every CTA sees the same partial sums, so
// the stopping decision is uniform.  hist = {phi, previous ratio, iterations done}.
__global__ void __launch_bounds__(SDP_THREADS) sdp_loop_kernel(const double* __restrict__ At, double* __restrict__ qu,
                                                                int64_t n, int tau, double gamma, double tol,
                                                                long long max_iter, double* part0, double* part1,
                                                                double* __restrict__ hist) {
  __shared__ double s_M[SDP_T2], s_P[SDP_T2];
  __shared__ double s_w[SDP_THREADS / 32][SDP_T2];
  __shared__ double s_r[16];
  cg::grid_group grid = cg::this_grid();
  long long it = 0;
  double phi = 0.0, rprev = 0.0, phi_prev = 1e300, g = gamma;
  for (; it < max_iter; ++it) {
    double* pin = (it & 1) ? part1 : part0;
    double* pout = (it & 1) ? part0 : part1;
    rprev = sdp_iterate(At, qu, n, tau, g, pin, pout, s_M, s_P, s_w, s_r, phi);
    // exponents above 1/2 are not guaranteed to decrease phi: the first increase switches to the monotone 1/2 for good
    // (phi is bit-identical in every CTA, so the switch is uniform)
    if (phi > phi_prev * (1.0 + 1e-13) && g > 0.5) g = 0.5;
    phi_prev = phi;
    grid.sync();
    if (!(phi == phi) || phi <= 0.0) { ++it; break; }              // not positive definite: reported by the host
    if (it >= 1 && rprev - 1.0 <= tol) { ++it; break; }            // the iterate before this update was already certified
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hist[0] = phi;
    hist[1] = rprev;
    hist[2] = (double)it;
  }
}

// single CTA: normalised q, t_j = (M^-1)_jj, phi and the certificate max_i d_i / phi - 1 of the RETURNED q
__global__ void __launch_bounds__(SDP_THREADS) sdp_final_kernel(const double* __restrict__ At, const double* __restrict__ qu,
                                                                 int64_t n, int tau, int G, double* part_in,
                                                                 double* __restrict__ q_out, double* __restrict__ res) {
  __shared__ double s_M[SDP_T2], s_P[SDP_T2], s_part[SDP_T2 + 2 + SDP_THREADS];
  __shared__ double s_r[8];
  const int Tu = sdp_packed(tau);
  double phi, Z, rprev;
  sdp_build(parts_of(part_in, G, Tu), G, tau, s_M, s_P, s_part, phi, Z, rprev);
  double rmax = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    double d = 0.0;
    for (int u = 0; u < Tu; ++u) d += s_P[u] * At[(int64_t)u * n + j];
    rmax = fmax(rmax, d / phi);           // over ALL i: the optimality condition also binds where q_i = 0
    q_out[j] = qu[j] / Z;
  }
  rmax = warp_max(rmax);
  if ((threadIdx.x & 31) == 0) s_r[threadIdx.x >> 5] = rmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int k = 0; k < SDP_THREADS / 32; ++k) r = fmax(r, s_r[k]);
    res[0] = phi;
    res[1] = r - 1.0;
  }
  if (threadIdx.x < tau) res[2 + threadIdx.x] = s_M[threadIdx.x * tau + threadIdx.x];
}

}  // namespace

int nnal_sdp_release(nnal_ctx* ctx) {
  if (!ctx->sdp_state) return NNAL_OK;
  SdpState* st = (SdpState*)ctx->sdp_state;
  DevBuf* bufs[] = {&st->At, &st->stage, &st->qu, &st->part[0], &st->part[1], &st->out};
  for (DevBuf* b : bufs) { if (b->p) cudaFree(b->p); b->p = nullptr; b->cap = 0; }
  delete st;
  ctx->sdp_state = nullptr;
  return NNAL_OK;
}

// mode 0: A = n dense tau x tau matrices on the host; mode 1: src = shrunk gradients g [2][n][tau], p1 = P(class 1) [n]
static int sdp_solve(nnal_ctx* ctx, int mode, const double* src, const double* p1, double diag_load, int64_t n, int tau,
                     double tol, int64_t max_iter, double gamma, double* q_out, double* t_out, double* obj_out,
                     double* gap_out, int64_t* iters_out) {
  if (!ctx || !src || !q_out || n <= 0 || tau <= 0) return NNAL_ERR_INVALID;
  if (tau > SDP_MAX_TAU) NNAL_FAIL(ctx, NNAL_ERR_UNSUPPORTED, "SDP: more than 16 shrunk coordinates");
  if (!(gamma > 0.0 && gamma <= 1.0)) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: gamma must be in (0, 1]");
  if (!(tol > 0.0)) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: tol must be positive");
  if (max_iter < 1) max_iter = 1;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  SdpState* st = sdp_state(ctx);
  const int T2 = tau * tau, Tu = sdp_packed(tau);
  const int G = (int)std::min<int64_t>((n + SDP_THREADS - 1) / SDP_THREADS, SDP_MAX_GRID);
  const size_t part_doubles = (size_t)G * Tu + 2 * (size_t)G;
  NNAL_TRY(devbuf_reserve(ctx, st->stage, mode == 0 ? (size_t)n * T2 * 8 : (size_t)n * (2 * tau + 1) * 8));
  NNAL_TRY(devbuf_reserve(ctx, st->At, (size_t)n * Tu * 8));
  NNAL_TRY(devbuf_reserve(ctx, st->qu, (size_t)n * 8 * 2));           // unnormalised weights, then the normalised result
  NNAL_TRY(devbuf_reserve(ctx, st->part[0], part_doubles * 8));
  NNAL_TRY(devbuf_reserve(ctx, st->part[1], part_doubles * 8));
  NNAL_TRY(devbuf_reserve(ctx, st->out, (size_t)(8 + SDP_MAX_TAU) * 8));
  double* At = (double*)st->At.p;
  double* qu = (double*)st->qu.p;
  double* qn = qu + n;
  double* hist = (double*)st->out.p;         // [0..2] loop record, [4..] result of the final kernel
  double* res = hist + 4;
  const int tg = (int)std::min<int64_t>((n * Tu + 255) / 256, (int64_t)ctx->sm_count * 8);
  if (mode == 0) {
    CUDA_TRY(ctx, cudaMemcpyAsync(st->stage.p, src, (size_t)n * T2 * 8, cudaMemcpyHostToDevice, ctx->stream));
    sdp_transpose_kernel<<<tg, 256, 0, ctx->stream>>>((const double*)st->stage.p, At, n, tau);
  } else {
    double* d_g = (double*)st->stage.p;
    double* d_p = d_g + 2 * n * tau;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_g, src, (size_t)2 * n * tau * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_p, p1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    sdp_binary_A_kernel<<<tg, 256, 0, ctx->stream>>>(d_g, d_p, n, tau, diag_load, At);
  }
  sdp_fill_kernel<<<G, 256, 0, ctx->stream>>>(qu, n, 1.0);
  int cur = 0;
  sdp_init_kernel<<<G, SDP_THREADS, 0, ctx->stream>>>(At, qu, n, tau, (double*)st->part[cur].p);
  ctx->launches += 3;
  CUDA_TRY(ctx, cudaGetLastError());
  int64_t it = 0;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  const bool no_coop = ctx->dbg.sdp_no_coop != 0;
  bool done = false;
  if (coop && !no_coop) {
    const double* a_At = At;
    double* a_qu = qu;
    int64_t a_n = n;
    int a_tau = tau;
    double a_gamma = gamma, a_tol = tol;
    long long a_max = (long long)max_iter;
    double* a_p0 = (double*)st->part[0].p;
    double* a_p1 = (double*)st->part[1].p;
    double* a_hist = hist;
    void* args[] = {&a_At, &a_qu, &a_n, &a_tau, &a_gamma, &a_tol, &a_max, &a_p0, &a_p1, &a_hist};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)sdp_loop_kernel, dim3(G), dim3(SDP_THREADS), args, 0, ctx->stream);
    if (e == cudaSuccess) {
      ctx->launches++;
      double h[3];
      CUDA_TRY(ctx, cudaMemcpyAsync(h, hist, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      if (!(h[0] == h[0]) || h[0] <= 0.0) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: sum_i q_i A_i is not positive definite");
      it = (int64_t)h[2];
      cur = (int)(it & 1);
      done = true;
    } else {
      cudaGetLastError();            // cooperative launch refused (grid not co-resident): one launch per iteration instead
    }
  }
  const int64_t check_every = 32;
  double g_host = gamma, phi_last = 1e300;
  while (!done && it < max_iter) {
    const int64_t stop = std::min(max_iter, it + check_every);
    for (; it < stop; ++it) {
      sdp_iter_kernel<<<G, SDP_THREADS, 0, ctx->stream>>>(At, qu, n, tau, g_host, (double*)st->part[cur].p,
                                                          (double*)st->part[cur ^ 1].p, hist);
      cur ^= 1;
      ctx->launches++;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    double h[2];
    CUDA_TRY(ctx, cudaMemcpyAsync(h, hist, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!(h[0] == h[0]) || h[0] <= 0.0) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: sum_i q_i A_i is not positive definite");
    if (it > 1 && h[1] - 1.0 <= tol) break;
    if (h[0] > phi_last * (1.0 + 1e-13) && g_host > 0.5) g_host = 0.5;     // as in the cooperative loop, checked per batch
    phi_last = h[0];
  }
  sdp_final_kernel<<<1, SDP_THREADS, 0, ctx->stream>>>(At, qu, n, tau, G, (double*)st->part[cur].p, qn, res);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  double hres[2 + SDP_MAX_TAU];
  CUDA_TRY(ctx, cudaMemcpyAsync(hres, res, sizeof(hres), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(q_out, qn, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (obj_out) *obj_out = hres[0];
  if (gap_out) *gap_out = hres[1];
  if (t_out) for (int j = 0; j < tau; ++j) t_out[j] = hres[2 + j];
  if (iters_out) *iters_out = it;
  return NNAL_OK;
}

// =====================================================================================================================
// lambda_ > 0: the feature-regularised programme of NNAL_tools.SDP_query_distribution (NNAL_tools.py:625-644)
//     minimise  Phi(q) = tr(M(q)^-1) - lambda c^T q,  c_i = |x_i|^2      s.t.  q >= 0,  1^T q = 1,  X q = 0
// (X = the zero-mean refined feature matrix [d][n]; upstream's vector c :630-632 and equalities A x = b :635-644).
// Multiplicative natural-gradient method that keeps every iterate feasible (oracle/fi_oracle.py:sdp_solve_reg):
//     h = d + lambda c,  d_i = <M^-2, A_i>;   (X diag(q) X^T) mu = X (q.h);   g = h - X^T mu;   nu = q.g;
//     q_i <- q_i (1 + theta (g_i / nu - 1)),   theta in (0,1] keeps q > 0 and is halved until Phi decreases (Armijo);
//     certificate  Phi(q) - Phi* <= max_i g_i - nu  (convexity; valid for any mu).
// The O(n d^2) weighted normal equations and the O(n tau^2) sums run on the device (float64), the tau x tau algebra and
// the line search -- M(q + theta D) = M(q) + theta M(D) is linear -- on the host.
namespace reg {

constexpr int RT = 256;

// out[u] = sum_i w_i At[u][i] (u < Tu),  out[Tu] = sum_i w_i c_i,  out[Tu+1] = sum_i w_i      (one CTA per entry)
__global__ void __launch_bounds__(RT) msum_kernel(const double* __restrict__ At, const double* __restrict__ w, const double* __restrict__ c,
                                                 int64_t n, int Tu, double* __restrict__ out) {
  const int u = blockIdx.x;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += RT) {
    const double v = u < Tu ? At[(int64_t)u * n + i] : (u == Tu ? c[i] : 1.0);
    acc = fma(w[i], v, acc);
  }
  acc = warp_sum(acc);
  __shared__ double sh[RT / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < RT / 32; ++k) t += sh[k];
    out[u] = t;
  }
}

__global__ void norms_kernel(const double* __restrict__ X, int64_t n, int d, double* __restrict__ c) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int a = 0; a < d; ++a) { const double v = X[(int64_t)a * n + i]; acc = fma(v, v, acc); }
    c[i] = acc;
  }
}

struct PArg { double P[SDP_MAX_TAU * (SDP_MAX_TAU + 1) / 2]; };

// h_i = <M^-2, A_i> + lambda c_i  and  qh_i = q_i h_i
__global__ void h_kernel(const double* __restrict__ At, const double* __restrict__ q, const double* __restrict__ c, int64_t n, int Tu,
                         PArg pa, double lambda_, double* __restrict__ h, double* __restrict__ qh) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double d = 0.0;
    for (int u = 0; u < Tu; ++u) d = fma(pa.P[u], At[(int64_t)u * n + i], d);
    const double hv = d + lambda_ * c[i];
    h[i] = hv;
    qh[i] = q[i] * hv;
  }
}

// W[a][b] = sum_i q_i X[a][i] X[b][i] (+ ridge on the diagonal, identity on the padding), 64 x 64 tile per CTA, float64
__global__ void __launch_bounds__(256) wgram_kernel(const double* __restrict__ X, const double* __restrict__ q, int64_t n, int d, int np,
                                                    double* __restrict__ W) {
  __shared__ double Xa[32][65], Xb[32][65];
  const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  if (b0 > a0) return;                                     // lower triangle; mirrored below
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double acc[4][4] = {};
  for (int64_t i0 = 0; i0 < n; i0 += 32) {
    for (int e = tid; e < 64 * 32; e += 256) {
      const int r = e >> 5, cidx = e & 31;
      const int64_t i = i0 + cidx;
      const double qi = i < n ? q[i] : 0.0;
      Xa[cidx][r] = (a0 + r < d && i < n) ? X[(int64_t)(a0 + r) * n + i] * qi : 0.0;
      Xb[cidx][r] = (b0 + r < d && i < n) ? X[(int64_t)(b0 + r) * n + i] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      double xa[4], xb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { xa[j] = Xa[k][4 * ty + j]; xb[j] = Xb[k][4 * tx + j]; }
#pragma unroll
      for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) acc[ia][ib] = fma(xa[ia], xb[ib], acc[ia][ib]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int ia = 0; ia < 4; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int a = a0 + 4 * ty + ia, b = b0 + 4 * tx + ib;
      double v = acc[ia][ib];
      if (a >= d || b >= d) v = a == b ? 1.0 : 0.0;
      W[(size_t)a * np + b] = v;
      W[(size_t)b * np + a] = v;
    }
}

__global__ void ridge_kernel(double* __restrict__ W, int d, int np, double rel) {
  __shared__ double tr;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int a = 0; a < d; ++a) t += W[(size_t)a * np + a];
    tr = t;
  }
  __syncthreads();
  for (int a = threadIdx.x; a < d; a += blockDim.x) W[(size_t)a * np + a] += rel * tr / d;
}

// y[a] = sum_i X[a][i] v[i]   (one CTA per feature row)
__global__ void __launch_bounds__(RT) xv_kernel(const double* __restrict__ X, const double* __restrict__ v, int64_t n, double* __restrict__ y) {
  const int a = blockIdx.x;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += RT) acc = fma(X[(int64_t)a * n + i], v[i], acc);
  acc = warp_sum(acc);
  __shared__ double sh[RT / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < RT / 32; ++k) t += sh[k];
    y[a] = t;
  }
}

// mu = Winv rhs (one CTA per row)
__global__ void __launch_bounds__(RT) mv_kernel(const double* __restrict__ Winv, int np, int d, const double* __restrict__ rhs, double* __restrict__ mu) {
  const int a = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < d; b += RT) acc = fma(Winv[(size_t)a * np + b], rhs[b], acc);
  acc = warp_sum(acc);
  __shared__ double sh[RT / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < RT / 32; ++k) t += sh[k];
    mu[a] = t;
  }
}

// g_i = h_i - sum_a X[a][i] mu_a;  per-CTA partials of  nu = q.g,  s2 = q.g^2,  gmax = max g
__global__ void __launch_bounds__(RT) g_kernel(const double* __restrict__ X, const double* __restrict__ mu, const double* __restrict__ h,
                                              const double* __restrict__ q, int64_t n, int d, double* __restrict__ g, double* __restrict__ part) {
  double nu = 0.0, s2 = 0.0, gm = -1e300;
  for (int64_t i = (int64_t)blockIdx.x * RT + threadIdx.x; i < n; i += (int64_t)gridDim.x * RT) {
    double acc = 0.0;
    for (int a = 0; a < d; ++a) acc = fma(X[(int64_t)a * n + i], mu[a], acc);
    const double gv = h[i] - acc;
    g[i] = gv;
    nu = fma(q[i], gv, nu);
    s2 = fma(q[i] * gv, gv, s2);
    gm = fmax(gm, gv);
  }
  nu = warp_sum(nu); s2 = warp_sum(s2); gm = warp_max(gm);
  __shared__ double sh[3][RT / 32];
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = nu; sh[1][threadIdx.x >> 5] = s2; sh[2][threadIdx.x >> 5] = gm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = -1e300;
    for (int k = 0; k < RT / 32; ++k) { a += sh[0][k]; b += sh[1][k]; c = fmax(c, sh[2][k]); }
    part[3 * blockIdx.x] = a; part[3 * blockIdx.x + 1] = b; part[3 * blockIdx.x + 2] = c;
  }
}

// D_i = q_i (g_i / nu - 1);  per-CTA minimum of r_i = g_i / nu - 1 over the live entries (q_i > 0)
__global__ void __launch_bounds__(RT) delta_kernel(const double* __restrict__ q, const double* __restrict__ g, int64_t n, double nu,
                                                  double* __restrict__ D, double* __restrict__ part) {
  double rmin = 1e300;
  for (int64_t i = (int64_t)blockIdx.x * RT + threadIdx.x; i < n; i += (int64_t)gridDim.x * RT) {
    const double r = g[i] / nu - 1.0;
    D[i] = q[i] * r;
    if (q[i] > 0.0) rmin = fmin(rmin, r);
  }
  rmin = -warp_max(-rmin);
  __shared__ double sh[RT / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = rmin;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 1e300;
    for (int k = 0; k < RT / 32; ++k) m = fmin(m, sh[k]);
    part[blockIdx.x] = m;
  }
}

__global__ void axpy_kernel(double* __restrict__ q, const double* __restrict__ D, int64_t n, double theta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    q[i] = fmax(fma(theta, D[i], q[i]), 0.0);
}

// host: inverse of the packed symmetric tau x tau matrix, trace, packed weights of <M^-2, .>; false if not positive definite
static bool host_inverse(const double* Mp, int tau, double* Minv, double* P, double* phi) {
  double m[SDP_MAX_TAU][SDP_MAX_TAU];
  int u = 0;
  for (int a = 0; a < tau; ++a)
    for (int b = a; b < tau; ++b, ++u) m[a][b] = m[b][a] = Mp[u];
  for (int k = 0; k < tau; ++k) {
    if (!(m[k][k] > 0.0)) return false;
    const double piv = 1.0 / m[k][k];
    double row[SDP_MAX_TAU], col[SDP_MAX_TAU];
    for (int j = 0; j < tau; ++j) { row[j] = m[k][j]; col[j] = m[j][k]; }
    for (int a = 0; a < tau; ++a)
      for (int b = 0; b < tau; ++b) {
        if (a == k && b == k) m[a][b] = piv;
        else if (a == k) m[a][b] = row[b] * piv;
        else if (b == k) m[a][b] = -col[a] * piv;
        else m[a][b] -= col[a] * row[b] * piv;
      }
  }
  double t = 0.0;
  for (int a = 0; a < tau; ++a) t += m[a][a];
  *phi = t;
  if (!(t == t) || t <= 0.0) return false;
  u = 0;
  for (int a = 0; a < tau; ++a)
    for (int b = a; b < tau; ++b, ++u) {
      double v = 0.0;
      for (int k = 0; k < tau; ++k) v += m[a][k] * m[k][b];
      if (P) P[u] = a == b ? v : 2.0 * v;
    }
  if (Minv)
    for (int a = 0; a < tau; ++a)
      for (int b = 0; b < tau; ++b) Minv[a * tau + b] = m[a][b];
  return true;
}

}  // namespace reg

extern "C" int nnal_sdp_query_distribution_reg(nnal_ctx* ctx, const double* A, int64_t n, int tau, double lambda_, const double* X,
                                               int d, double tol, int64_t max_iter, double* q_out, double* t_out, double* obj_out,
                                               double* gap_out, int64_t* iters_out) {
  if (!ctx || !A || !X || !q_out || n <= 0 || tau <= 0 || d <= 0) return NNAL_ERR_INVALID;
  if (tau > SDP_MAX_TAU) NNAL_FAIL(ctx, NNAL_ERR_UNSUPPORTED, "SDP: more than 16 shrunk coordinates");
  if (!(lambda_ >= 0.0) || !(tol > 0.0)) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: lambda_ must be >= 0 and tol > 0");
  if (d >= n) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: X q = 0 with d >= n equalities leaves no freedom (refine the feature matrix)");
  if (d > 4096 || n > (1 << 20)) NNAL_FAIL(ctx, NNAL_ERR_UNSUPPORTED, "regularised SDP: at most 4096 feature rows");
  if (max_iter < 1) max_iter = 1;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  SdpState* st = sdp_state(ctx);
  const int T2 = tau * tau, Tu = sdp_packed(tau);
  const int np = (d + 63) / 64 * 64;
  const int G = (int)std::min<int64_t>((n + reg::RT - 1) / reg::RT, 64);
  // workspace: [A stage n T2] | At [Tu n] | X [d n] | q, c, h, qh, g, D [6 n] | W [np np] | rhs, mu [2 np] | small [256]
  NNAL_TRY(devbuf_reserve(ctx, st->stage, (size_t)n * T2 * 8));
  NNAL_TRY(devbuf_reserve(ctx, st->At, (size_t)n * Tu * 8));
  const size_t wdoubles = (size_t)d * n + 6 * (size_t)n + (size_t)np * np + 2 * (size_t)np + 512;
  NNAL_TRY(devbuf_reserve(ctx, st->qu, wdoubles * 8));
  double* At = (double*)st->At.p;
  double* Xd = (double*)st->qu.p;
  double* q = Xd + (size_t)d * n;
  double* c = q + n; double* h = c + n; double* qh = h + n; double* g = qh + n; double* D = g + n;
  double* W = D + n;
  double* rhs = W + (size_t)np * np; double* mu = rhs + np;
  double* small = mu + np;                                  // [0..Tu+1] sums, [200..] partials
  CUDA_TRY(ctx, cudaMemcpyAsync(st->stage.p, A, (size_t)n * T2 * 8, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(Xd, X, (size_t)d * n * 8, cudaMemcpyHostToDevice, ctx->stream));
  const int tg = (int)std::min<int64_t>((n * Tu + 255) / 256, (int64_t)ctx->sm_count * 8);
  sdp_transpose_kernel<<<tg, 256, 0, ctx->stream>>>((const double*)st->stage.p, At, n, tau);
  sdp_fill_kernel<<<G, 256, 0, ctx->stream>>>(q, n, 1.0 / (double)n);
  reg::norms_kernel<<<G, 256, 0, ctx->stream>>>(Xd, n, d, c);
  ctx->launches += 3;
  std::vector<double> hs(Tu + 2), hd(Tu + 2), part(3 * 64 + 64);
  double Mp[SDP_MAX_TAU * (SDP_MAX_TAU + 1) / 2], Mtry[SDP_MAX_TAU * (SDP_MAX_TAU + 1) / 2], Minv[SDP_T2];
  reg::PArg pa;
  double phi = 0.0, cq = 0.0, Phi = 0.0, gap = 1e300;
  auto sums = [&](const double* w, std::vector<double>& out) -> int {
    reg::msum_kernel<<<Tu + 2, reg::RT, 0, ctx->stream>>>(At, w, c, n, Tu, small);
    ctx->launches++;
    CUDA_TRY(ctx, cudaMemcpyAsync(out.data(), small, (size_t)(Tu + 2) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NNAL_OK;
  };
  int64_t it = 0;
  bool fresh = true;
  for (;; ++it) {
    if (fresh || it % 64 == 0) {                            // M(q), c.q from scratch (otherwise carried along the line search)
      NNAL_TRY(sums(q, hs));
      for (int u = 0; u < Tu; ++u) Mp[u] = hs[u];
      cq = hs[Tu];
      fresh = false;
    }
    if (!reg::host_inverse(Mp, tau, Minv, pa.P, &phi)) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: sum_i q_i A_i is not positive definite");
    Phi = phi - lambda_ * cq;
    reg::h_kernel<<<G, 256, 0, ctx->stream>>>(At, q, c, n, Tu, pa, lambda_, h, qh);
    reg::wgram_kernel<<<dim3(np / 64, np / 64), 256, 0, ctx->stream>>>(Xd, q, n, d, np, W);
    reg::ridge_kernel<<<1, 256, 0, ctx->stream>>>(W, d, np, 1e-14);
    reg::xv_kernel<<<d, reg::RT, 0, ctx->stream>>>(Xd, qh, n, rhs);
    ctx->launches += 4;
    NNAL_TRY(nnal_gj64_invert(ctx, W, np));
    reg::mv_kernel<<<d, reg::RT, 0, ctx->stream>>>(W, np, d, rhs, mu);
    reg::g_kernel<<<G, reg::RT, 0, ctx->stream>>>(Xd, mu, h, q, n, d, g, small + 200);
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaMemcpyAsync(part.data(), small + 200, (size_t)3 * G * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double nu = 0.0, s2 = 0.0, gmax = -1e300;
    for (int b = 0; b < G; ++b) { nu += part[3 * b]; s2 += part[3 * b + 1]; gmax = std::max(gmax, part[3 * b + 2]); }
    gap = (gmax - nu) / std::max(std::fabs(Phi), phi);
    if (!(gap == gap)) NNAL_FAIL(ctx, NNAL_ERR_INVALID, "SDP: the regularised programme broke down (singular feature normal equations)");
    if (gap <= tol || it >= max_iter) break;
    reg::delta_kernel<<<G, reg::RT, 0, ctx->stream>>>(q, g, n, nu, D, small + 200);
    ctx->launches++;
    NNAL_TRY(sums(D, hd));                                  // M(D), c.D (the sync also covers the partial minima below)
    CUDA_TRY(ctx, cudaMemcpyAsync(part.data(), small + 200, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double rmin = 1e300;
    for (int b = 0; b < G; ++b) rmin = std::min(rmin, part[b]);
    double theta = rmin < 0.0 ? std::min(1.0, 0.99 * (-1.0 / rmin)) : 1.0;
    const double slope = (s2 - nu * nu) / nu;               // Var_q(g) / nu >= 0
    double phin = 0.0;
    for (;;) {
      for (int u = 0; u < Tu; ++u) Mtry[u] = Mp[u] + theta * hd[u];
      const bool ok = reg::host_inverse(Mtry, tau, nullptr, nullptr, &phin);
      const double Pn = phin - lambda_ * (cq + theta * hd[Tu]);
      if ((ok && Pn <= Phi - 1e-4 * theta * slope) || theta < 1e-12) break;
      theta *= 0.5;
    }
    reg::axpy_kernel<<<G, 256, 0, ctx->stream>>>(q, D, n, theta);
    ctx->launches++;
    for (int u = 0; u < Tu; ++u) Mp[u] = Mtry[u];
    cq += theta * hd[Tu];
  }
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(q_out, q, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (obj_out) *obj_out = Phi;
  if (gap_out) *gap_out = gap;
  if (t_out) for (int j = 0; j < tau; ++j) t_out[j] = Minv[j * tau + j];
  if (iters_out) *iters_out = it;
  return NNAL_OK;
}

extern "C" int nnal_sdp_query_distribution(nnal_ctx* ctx, const double* A, int64_t n, int tau, double tol, int64_t max_iter,
                                           double gamma, double* q_out, double* t_out, double* obj_out, double* gap_out,
                                           int64_t* iters_out) {
  return sdp_solve(ctx, 0, A, nullptr, 0.0, n, tau, tol, max_iter, gamma, q_out, t_out, obj_out, gap_out, iters_out);
}

extern "C" int nnal_sdp_from_shrunk(nnal_ctx* ctx, const double* g, const double* p1, int64_t n, int tau, double diag_load,
                                    double tol, int64_t max_iter, double gamma, double* q_out, double* t_out,
                                    double* obj_out, double* gap_out, int64_t* iters_out) {
  if (!p1 || !(diag_load >= 0.0)) return NNAL_ERR_INVALID;
  return sdp_solve(ctx, 1, g, p1, diag_load, n, tau, tol, max_iter, gamma, q_out, t_out, obj_out, gap_out, iters_out);
}
