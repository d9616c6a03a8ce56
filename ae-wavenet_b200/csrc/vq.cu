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

// d == 32 (the reference's bottleneck width): the same order, fully unrolled, so that a code row can live in registers
template <typename F>
__device__ __forceinline__ float sumsq_blk16_d32(F f) {
  float total = 0.0f;
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = f(16 * blk + i);
      acc = __fadd_rn(acc, __fmul_rn(v, v));
    }
    total = blk == 0 ? acc : __fadd_rn(total, acc);
  }
  return total;
}

// One CTA per V consecutive (b, n) vectors; the codebook streams through shared memory in tiles of 256 codes.
//
// Round-2 measurement (profiles/r3_vq_latency.txt): the first version -- one CTA per vector, every thread reading ITS
// codes' rows straight from global memory (lane stride = d floats: 32 different 128-byte lines per load instruction,
// the 512 KB codebook re-read by each of the 1040 CTAs) -- took 1.09 ms at the cfg3 shapes, SLOWER than the reference's
// eager op sequence (0.77 ms).  Here a code tile is loaded once per CTA with coalesced reads into a padded [d][257]
// tile (thread = code: conflict-free reads), V vectors share it, and a thread keeps its code's row in registers when
// d == 32.  The arithmetic per (vector, code) pair -- and with it every index and distance bit -- is unchanged.
constexpr int VQ_TILE = 256;

template <int V, bool REG_E>
__global__ void __launch_bounds__(VQ_THREADS) vq_fwd_kernel(const float* __restrict__ ze, long long ze_bs, long long ze_cs,
                                                            const float* __restrict__ emb, int metric,
                                                            long long* __restrict__ min_ind, float* __restrict__ min_dist,
                                                            float* __restrict__ zq, long long zq_bs, long long zq_cs,
                                                            float* __restrict__ hist, float* __restrict__ z_sum,
                                                            float* __restrict__ n_sum, float* __restrict__ ze_norm,
                                                            int d, int N, int K, int n_vec) {
  extern __shared__ float vq_smem[];
  float* s_e = vq_smem;                               // [d][VQ_TILE + 1]
  float* s_ze = vq_smem + d * (VQ_TILE + 1);          // [V][d]
  __shared__ float s_best[V][VQ_THREADS / 32];
  __shared__ int s_idx[V][VQ_THREADS / 32];
  __shared__ int s_win[V];
  const int vec0 = blockIdx.x * V;
  for (int i = threadIdx.x; i < V * d; i += blockDim.x) {
    const int v = i / d, j = i - v * d;
    const int vec = vec0 + v;
    float x = 0.0f;
    if (vec < n_vec) {
      const int b = vec / N, n = vec - b * N;
      x = ze[static_cast<long long>(b) * ze_bs + static_cast<long long>(j) * ze_cs + n];
    }
    s_ze[i] = x;
  }
  __syncthreads();

  float a[V];  // ||ze_v||
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float* zv = s_ze + v * d;
    a[v] = (metric == 1) ? __fsqrt_rn(sumsq_blk16(d, [&](int j) { return zv[j]; })) : 0.0f;
  }
  float best[V];
  int best_k[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    best[v] = INFINITY;
    best_k[v] = 0x7fffffff;
  }
  for (int k0 = 0; k0 < K; k0 += VQ_TILE) {
    __syncthreads();          // the previous tile has been consumed
    const int nk = min(VQ_TILE, K - k0);
    const float* src = emb + static_cast<long long>(k0) * d;
    for (int i = threadIdx.x; i < nk * d; i += blockDim.x) {      // coalesced: the tile is contiguous in memory
      const int c = i / d, j = i - c * d;
      s_e[j * (VQ_TILE + 1) + c] = __ldg(src + i);
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c < nk) {
      const int k = k0 + c;
      float er[REG_E ? 32 : 1];
      if (REG_E) {       // d == 32
#pragma unroll
        for (int j = 0; j < 32; ++j) er[j] = s_e[j * (VQ_TILE + 1) + c];
      }
      float bn = 0.0f;
      if (metric == 1)
        bn = __fsqrt_rn(REG_E ? sumsq_blk16_d32([&](int j) { return er[j]; })
                              : sumsq_blk16(d, [&](int j) { return s_e[j * (VQ_TILE + 1) + c]; }));
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float* zv = s_ze + v * d;
        float dist = REG_E ? sumsq_blk16_d32([&](int j) { return __fsub_rn(zv[j], er[j]); })
                           : sumsq_blk16(d, [&](int j) { return __fsub_rn(zv[j], s_e[j * (VQ_TILE + 1) + c]); });
        if (metric == 1) dist = __fdiv_rn(__fsqrt_rn(dist), __fadd_rn(a[v], bn));
        if (dist < best[v]) {  // strict: the first (smallest) index wins ties within a thread (k ascending)
          best[v] = dist;
          best_k[v] = k;
        }
      }
    }
  }
  // block argmin per vector, ties -> smaller index
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float bb = best[v];
    int bk = best_k[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, bb, o);
      const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
      if (ob < bb || (ob == bb && ok < bk)) {
        bb = ob;
        bk = ok;
      }
    }
    if ((threadIdx.x & 31) == 0) {
      s_best[v][threadIdx.x >> 5] = bb;
      s_idx[v][threadIdx.x >> 5] = bk;
    }
  }
  __syncthreads();
  if (threadIdx.x < V && vec0 + static_cast<int>(threadIdx.x) < n_vec) {
    const int v = threadIdx.x, vec = vec0 + v;
    float bb = s_best[v][0];
    int bk = s_idx[v][0];
    for (int w = 1; w < VQ_THREADS / 32; ++w) {
      if (s_best[v][w] < bb || (s_best[v][w] == bb && s_idx[v][w] < bk)) {
        bb = s_best[v][w];
        bk = s_idx[v][w];
      }
    }
    if (bk == 0x7fffffff) bk = 0;  // all distances NaN: torch.min would return index of the first NaN; K > 0 so 0
    s_win[v] = bk;
    min_ind[vec] = bk;
    min_dist[vec] = bb;
    const float* zv = s_ze + v * d;
    if (ze_norm) ze_norm[vec] = __fsqrt_rn(sumsq_blk16(d, [&](int j) { return zv[j]; }));
    if (hist) atomicAdd(hist + bk, 1.0f);
    if (n_sum) atomicAdd(n_sum + bk, 1.0f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V * d; i += blockDim.x) {
    const int v = i / d, j = i - v * d;
    const int vec = vec0 + v;
    if (vec >= n_vec) continue;
    const int b = vec / N, n = vec - b * N;
    const int win = s_win[v];
    zq[static_cast<long long>(b) * zq_bs + static_cast<long long>(j) * zq_cs + n] = emb[static_cast<long long>(win) * d + j];
    if (z_sum) atomicAdd(z_sum + static_cast<long long>(win) * d + j, s_ze[i]);
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
  // V vectors per CTA share every code tile: the largest V in {8, 4, 2, 1} that still leaves about one CTA per SM
  const int n_vec = batch * N;
  const int sms = sm_count();
  const int V = n_vec >= 8 * (sms - sms / 8) ? 8 : n_vec >= 4 * (sms - sms / 8) ? 4 : n_vec >= 2 * (sms - sms / 8) ? 2 : 1;
  const size_t smem = sizeof(float) * (static_cast<size_t>(d) * (VQ_TILE + 1) + static_cast<size_t>(V) * d);
  using Fn = void (*)(const float*, long long, long long, const float*, int, long long*, float*, float*, long long,
                      long long, float*, float*, float*, float*, int, int, int, int);
  const bool reg_e = d == 32;
  Fn fn = V == 8 ? (reg_e ? vq_fwd_kernel<8, true> : vq_fwd_kernel<8, false>)
          : V == 4 ? (reg_e ? vq_fwd_kernel<4, true> : vq_fwd_kernel<4, false>)
          : V == 2 ? (reg_e ? vq_fwd_kernel<2, true> : vq_fwd_kernel<2, false>)
                   : (reg_e ? vq_fwd_kernel<1, true> : vq_fwd_kernel<1, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_err(e, "vq_fwd: cudaFuncSetAttribute");
  }
  fn<<<(n_vec + V - 1) / V, VQ_THREADS, smem, stream>>>(ze, ze_bs, ze_cs, emb, metric, min_ind, min_dist, zq, zq_bs, zq_cs,
                                                        hist, z_sum, n_sum, ze_norm, d, N, K, n_vec);
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
