// Fused VQ step (K7 in SURVEY.md 2b): distance to every code + argmin + gather + usage histogram + EMA statistics in
// ONE kernel, never materialising the (B, K, d, N) tensors the reference builds twice (vqema_bn.py:135,138).
//
// Bit-exactness contract: all distance arithmetic is IEEE fp32 with explicit round-to-nearest
// intrinsics (no FMA contraction) in a FIXED order -- squares summed in blocks of 16 channels, block sums added in
// order, i.e. the order ATen's cascade_sum uses on its vectorised outer-reduction path (SumKernel.cpp multi_row_sum,
// level_step = 16).  oracle/vq_oracle.c restates the identical order in C; indices and min_dist must match it bit for
// bit.  Ties resolve to the smallest code index (torch.min returns the first minimum).
#include <math.h>

#include "host_util.h"

namespace aewn {

constexpr int VQ_THREADS = 256;
constexpr int VQ_MAX_D = 128;

// sum_j v_j^2 in the canonical order; v supplied by functor f(j)
template <typename F>
__device__ __forceinline__ float sumsq_blk16(int d, F f) {
  float total = 0.0f;
  bool first = true;
  int j = 0;
  for (; j + 16 <= d; j += 16) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = f(j + i);
      acc = __fadd_rn(acc, __fmul_rn(v, v));
    }
    total = first ? acc : __fadd_rn(total, acc);
    first = false;
  }
  if (j < d) {
    float rem = 0.0f;
    for (; j < d; ++j) {
      const float v = f(j);
      rem = __fadd_rn(rem, __fmul_rn(v, v));
    }
    total = first ? rem : __fadd_rn(rem, total);
  }
  return total;
}

// One CTA per (b, n) vector.
__global__ void __launch_bounds__(VQ_THREADS) vq_fwd_kernel(const float* __restrict__ ze, long long ze_bs, long long ze_cs,
                                                            const float* __restrict__ emb, int metric,
                                                            long long* __restrict__ min_ind, float* __restrict__ min_dist,
                                                            float* __restrict__ zq, long long zq_bs, long long zq_cs,
                                                            float* __restrict__ hist, float* __restrict__ z_sum,
                                                            float* __restrict__ n_sum, float* __restrict__ ze_norm,
                                                            int d, int N, int K) {
  __shared__ float s_ze[VQ_MAX_D];
  __shared__ float s_best[VQ_THREADS / 32];
  __shared__ int s_idx[VQ_THREADS / 32];
  __shared__ int s_win;
  const int vec = blockIdx.x;
  const int b = vec / N, n = vec - b * N;
  const float* zp = ze + static_cast<long long>(b) * ze_bs + n;
  for (int j = threadIdx.x; j < d; j += blockDim.x) s_ze[j] = zp[static_cast<long long>(j) * ze_cs];
  __syncthreads();

  float a = 0.0f;  // ||ze||
  if (metric == 1) a = __fsqrt_rn(sumsq_blk16(d, [&](int j) { return s_ze[j]; }));

  float best = INFINITY;
  int best_k = 0x7fffffff;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float* e = emb + static_cast<long long>(k) * d;
    float dist = sumsq_blk16(d, [&](int j) { return __fsub_rn(s_ze[j], __ldg(e + j)); });
    if (metric == 1) {
      const float bn = __fsqrt_rn(sumsq_blk16(d, [&](int j) { return __ldg(e + j); }));
      dist = __fdiv_rn(__fsqrt_rn(dist), __fadd_rn(a, bn));
    }
    if (dist < best) {  // strict: the first (smallest) index wins ties within a thread (k ascending)
      best = dist;
      best_k = k;
    }
  }
  // block argmin, ties -> smaller index
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    if (ob < best || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_best[threadIdx.x >> 5] = best;
    s_idx[threadIdx.x >> 5] = best_k;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bb = s_best[0];
    int bk = s_idx[0];
    for (int w = 1; w < VQ_THREADS / 32; ++w) {
      if (s_best[w] < bb || (s_best[w] == bb && s_idx[w] < bk)) {
        bb = s_best[w];
        bk = s_idx[w];
      }
    }
    if (bk == 0x7fffffff) bk = 0;  // all distances NaN: torch.min would return index of the first NaN; K > 0 so 0
    s_win = bk;
    min_ind[vec] = bk;
    min_dist[vec] = bb;
    if (ze_norm) ze_norm[vec] = (metric == 1) ? a : __fsqrt_rn(sumsq_blk16(d, [&](int j) { return s_ze[j]; }));
    if (hist) atomicAdd(hist + bk, 1.0f);
    if (n_sum) atomicAdd(n_sum + bk, 1.0f);
  }
  __syncthreads();
  const int win = s_win;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    zq[static_cast<long long>(b) * zq_bs + static_cast<long long>(j) * zq_cs + n] = emb[static_cast<long long>(win) * d + j];
    if (z_sum) atomicAdd(z_sum + static_cast<long long>(win) * d + j, s_ze[j]);
  }
}

// d(sum_{b,n} g[b,n] * min_dist[b,n]) / d ze  (commitment term, vqema_bn.py:237,246)
//   scaled L2: (ze - q)/(n (a+b)) - n ze / (a (a+b)^2),  n = |ze - q|, a = |ze|, b = |q|
//   squared L2: 2 (ze - q)
__global__ void vq_commit_bwd_kernel(const float* __restrict__ ze, long long ze_bs, long long ze_cs,
                                     const float* __restrict__ emb, const long long* __restrict__ min_ind,
                                     const float* __restrict__ g, int metric, float* __restrict__ g_ze, long long g_bs,
                                     long long g_cs, int accumulate, int d, int N) {
  const int vec = blockIdx.x;
  const int b = vec / N, n = vec - b * N;
  __shared__ float s_red[3][32];
  const float* zp = ze + static_cast<long long>(b) * ze_bs + n;
  const float* e = emb + min_ind[vec] * d;
  float nn = 0.f, aa = 0.f, bb = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float z = zp[static_cast<long long>(j) * ze_cs], q = e[j];
    nn += (z - q) * (z - q);
    aa += z * z;
    bb += q * q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = nn;
    s_red[1][threadIdx.x >> 5] = aa;
    s_red[2][threadIdx.x >> 5] = bb;
  }
  __syncthreads();
  nn = aa = bb = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) {
    nn += s_red[0][w];
    aa += s_red[1][w];
    bb += s_red[2][w];
  }
  const float nrm = sqrtf(nn), a = sqrtf(aa), bq = sqrtf(bb);
  const float gv = g[vec];
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float z = zp[static_cast<long long>(j) * ze_cs], q = e[j];
    float dv;
    if (metric == 1) {
      const float s = a + bq;
      dv = (nrm > 0.f ? (z - q) / (nrm * s) : 0.f) - (a > 0.f ? nrm * z / (a * s * s) : 0.f);
    } else {
      dv = 2.0f * (z - q);
    }
    float* o = g_ze + static_cast<long long>(b) * g_bs + static_cast<long long>(j) * g_cs + n;
    *o = accumulate ? *o + gv * dv : gv * dv;
  }
}

// vqema_bn.py:190-195 (EMA) and :220-222 (codebook refresh)
__global__ void ema_update_kernel(float* __restrict__ numer, float* __restrict__ denom, const float* __restrict__ z_sum,
                                  const float* __restrict__ n_sum, float gamma, float* __restrict__ emb, int K, int d) {
  const float comp = 1.0f - gamma;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * d; i += gridDim.x * blockDim.x) {
    const int k = i / d;
    float dn = denom[k];
    if (z_sum) {
      numer[i] = gamma * numer[i] + comp * z_sum[i];
      dn = gamma * dn + comp * n_sum[k];
    }
    if (emb) emb[i] = numer[i] / dn;
  }
}
__global__ void ema_denom_kernel(float* __restrict__ denom, const float* __restrict__ n_sum, float gamma, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) denom[k] = gamma * denom[k] + (1.0f - gamma) * n_sum[k];
}

}  // namespace aewn

using namespace aewn;

extern "C" {

int aewn_vq_fwd(const float* ze, long long ze_bs, long long ze_cs, const float* emb, int metric, long long* min_ind,
                float* min_dist, float* zq, long long zq_bs, long long zq_cs, float* hist, float* z_sum, float* n_sum,
                float* ze_norm, int batch, int d, int N, int K, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!ze || !emb || !min_ind || !min_dist || !zq || batch <= 0 || d <= 0 || d > VQ_MAX_D || N <= 0 || K <= 0 ||
      (metric != 0 && metric != 1))
    return set_err(AEWN_ERR_INVALID, "vq_fwd: bad arguments (d must be <= %d)", VQ_MAX_D);
  if (z_sum) {
    cudaError_t e = cudaMemsetAsync(z_sum, 0, sizeof(float) * static_cast<size_t>(K) * d, stream);
    if (e != cudaSuccess) return cuda_err(e, "vq_fwd: memset z_sum");
  }
  if (n_sum) {
    cudaError_t e = cudaMemsetAsync(n_sum, 0, sizeof(float) * static_cast<size_t>(K), stream);
    if (e != cudaSuccess) return cuda_err(e, "vq_fwd: memset n_sum");
  }
  vq_fwd_kernel<<<batch * N, VQ_THREADS, 0, stream>>>(ze, ze_bs, ze_cs, emb, metric, min_ind, min_dist, zq, zq_bs, zq_cs,
                                                      hist, z_sum, n_sum, ze_norm, d, N, K);
  count_launch();
  return cuda_err(cudaGetLastError(), "vq_fwd launch");
}

int aewn_vq_commit_bwd(const float* ze, long long ze_bs, long long ze_cs, const float* emb, const long long* min_ind,
                       const float* g_min_dist, int metric, float* g_ze, long long g_bs, long long g_cs, int accumulate,
                       int batch, int d, int N, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!ze || !emb || !min_ind || !g_min_dist || !g_ze || batch <= 0 || d <= 0 || N <= 0)
    return set_err(AEWN_ERR_INVALID, "vq_commit_bwd: bad arguments");
  vq_commit_bwd_kernel<<<batch * N, 64, 0, stream>>>(ze, ze_bs, ze_cs, emb, min_ind, g_min_dist, metric, g_ze, g_bs,
                                                     g_cs, accumulate, d, N);
  count_launch();
  return cuda_err(cudaGetLastError(), "vq_commit_bwd launch");
}

/* ema_numer/denom <- gamma * old + (1-gamma) * z_sum/n_sum (skipped when z_sum == NULL);
 * emb <- numer / denom[:, None] when emb != NULL (VQEMA.update_codebook). */
int aewn_ema_update(float* ema_numer, float* ema_denom, const float* z_sum, const float* n_sum, float gamma, float* emb,
                    int K, int d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!ema_numer || !ema_denom || K <= 0 || d <= 0 || (z_sum && !n_sum))
    return set_err(AEWN_ERR_INVALID, "ema_update: bad arguments");
  ema_update_kernel<<<(K * d + 255) / 256, 256, 0, stream>>>(ema_numer, ema_denom, z_sum, n_sum, gamma, emb, K, d);
  count_launch();
  if (z_sum) {
    ema_denom_kernel<<<(K + 255) / 256, 256, 0, stream>>>(ema_denom, n_sum, gamma, K);
    count_launch();
  }
  return cuda_err(cudaGetLastError(), "ema_update launch");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// The bottleneck's 1x1 projection (vqema_bn.py:92,131; vq_bn.py:18,35: self.linear, a bias-free Conv1d(n_in -> d, 1))
// in EXACT fp32: sequential fmaf over the input channels, no tensor cores.  The nearest-code search that follows is an
// index computation (north star: bit-identical indices); TF32 operand rounding of this 0.05 GFLOP product moved ~3 % of
// the codes across near-ties, fp32 FMA arithmetic reproduces the reference's codes.
//   out[b, n, t] = sum_k W[n * w_rs + k * w_cs] * x[b, k, t]
// The same kernel computes the data gradient (W read transposed).  Thread = (time step, 8 output rows).
// ---------------------------------------------------------------------------------------------------------------
namespace aewn {

__global__ void __launch_bounds__(256) conv1x1_f32_kernel(const float* __restrict__ x, long long x_bs, long long x_cs,
                                                          const float* __restrict__ w, long long w_rs, long long w_cs,
                                                          float* __restrict__ out, long long o_bs, long long o_cs, int N,
                                                          int K, int T) {
  const int b = blockIdx.z;
  const int t = blockIdx.x * 32 + (threadIdx.x & 31);
  const int n0 = blockIdx.y * 64 + (threadIdx.x >> 5);     // this thread: rows n0, n0 + 8, ..., n0 + 56
  if (t >= T) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  const float* xp = x + static_cast<long long>(b) * x_bs + t;
  for (int k = 0; k < K; ++k) {
    const float xv = xp[static_cast<long long>(k) * x_cs];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + 8 * i;
      if (n < N) acc[i] = fmaf(__ldg(w + n * w_rs + k * w_cs), xv, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + 8 * i;
    if (n < N) out[static_cast<long long>(b) * o_bs + static_cast<long long>(n) * o_cs + t] = acc[i];
  }
}

// dW[n, k] = sum_{b, t} g[b, n, t] * x[b, k, t]: one warp per output, lanes stride over time, fixed-order butterfly.
__global__ void __launch_bounds__(256) conv1x1_wgrad_f32_kernel(const float* __restrict__ g, long long g_bs,
                                                                long long g_cs, const float* __restrict__ x,
                                                                long long x_bs, long long x_cs, float* __restrict__ dw,
                                                                int N, int K, int T, int B) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= N * K) return;
  const int n = o / K, k = o - n * K;
  const int lane = threadIdx.x & 31;
  float acc = 0.0f;
  for (int b = 0; b < B; ++b) {
    const float* gp = g + static_cast<long long>(b) * g_bs + static_cast<long long>(n) * g_cs;
    const float* xp = x + static_cast<long long>(b) * x_bs + static_cast<long long>(k) * x_cs;
    for (int t = lane; t < T; t += 32) acc = fmaf(gp[t], xp[t], acc);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) dw[o] = acc;
}

}  // namespace aewn

extern "C" {

int aewn_conv1x1_f32(const float* x, long long x_bs, long long x_cs, const float* w, long long w_rs, long long w_cs,
                     float* out, long long o_bs, long long o_cs, int batch, int N, int K, int T, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w || !out || batch <= 0 || N <= 0 || K <= 0 || T <= 0)
    return aewn::set_err(AEWN_ERR_INVALID, "conv1x1_f32: bad arguments");
  dim3 grid((T + 31) / 32, (N + 63) / 64, batch);
  aewn::conv1x1_f32_kernel<<<grid, 256, 0, stream>>>(x, x_bs, x_cs, w, w_rs, w_cs, out, o_bs, o_cs, N, K, T);
  aewn::count_launch();
  return aewn::cuda_err(cudaGetLastError(), "conv1x1_f32 launch");
}

int aewn_conv1x1_wgrad_f32(const float* g, long long g_bs, long long g_cs, const float* x, long long x_bs, long long x_cs,
                           float* dw, int batch, int N, int K, int T, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!g || !x || !dw || batch <= 0 || N <= 0 || K <= 0 || T <= 0)
    return aewn::set_err(AEWN_ERR_INVALID, "conv1x1_wgrad_f32: bad arguments");
  aewn::conv1x1_wgrad_f32_kernel<<<(N * K + 7) / 8, 256, 0, stream>>>(g, g_bs, g_cs, x, x_bs, x_cs, dw, N, K, T, batch);
  aewn::count_launch();
  return aewn::cuda_err(cudaGetLastError(), "conv1x1_wgrad_f32 launch");
}

}  // extern "C"
