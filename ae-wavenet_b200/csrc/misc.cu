// HBM-bound helper kernels of the decoder path: base-layer embedding gather (K3 in SURVEY.md 2b), its backward,
// fills and the ReLU-mask product.  All are coalesced along the time axis (128-bit where alignment allows).
#include "host_util.h"

namespace aewn {

// ---------------------------------------------------------------------------------------------------------------
// base layer, wavenet.py:348-351:  one_hot(wav.long())[.., off0:off1] -> Conv1d(Q -> R, k=1)  ==  a column gather
//   out[b, r, tau] = W[r, code(b, off0 + tau)] + bias[r]
// CTA = (32 output channels) x (1024 time steps of one batch item); the 32 x Q weight slab sits in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BE_ROWS = 32;
constexpr int BE_TT = 1024;

__global__ void __launch_bounds__(256) base_embed_fwd_kernel(const float* __restrict__ wav, long long wav_pitch,
                                                             int off0, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             long long out_bs, long long out_cs,
                                                             float* __restrict__ dup, int dup_toff, int dup_t_hi,
                                                             int R, int Q, int T, int* err) {
  extern __shared__ float sw[];  // [BE_ROWS][Q + 1]
  const int r0 = blockIdx.y * BE_ROWS;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * BE_TT;
  const int qs = Q + 1;
  for (int i = threadIdx.x; i < BE_ROWS * Q; i += blockDim.x) {
    const int rr = i / Q, q = i - rr * Q;
    sw[rr * qs + q] = (r0 + rr < R) ? w[static_cast<long long>(r0 + rr) * Q + q] + (bias ? bias[r0 + rr] : 0.0f) : 0.0f;
  }
  __syncthreads();
  for (int tt = threadIdx.x; tt < BE_TT; tt += blockDim.x) {
    const int t = t0 + tt;
    if (t >= T) break;
    int code = static_cast<int>(wav[static_cast<long long>(b) * wav_pitch + off0 + t]);  // == .long(): truncation
    if (code < 0 || code >= Q) {
      if (err) atomicExch(err, AEWN_ERR_INVALID);
      code = code < 0 ? 0 : Q - 1;
    }
    float* o = out + static_cast<long long>(b) * out_bs + static_cast<long long>(r0) * out_cs + t;
    const int dt = t + dup_toff;
    const bool dup_ok = dup && dt >= 0 && dt < dup_t_hi;
    float* od = dup ? dup + static_cast<long long>(b) * out_bs + static_cast<long long>(r0) * out_cs + dt : nullptr;
#pragma unroll 8
    for (int rr = 0; rr < BE_ROWS; ++rr) {
      if (r0 + rr < R) {
        const float v = sw[rr * qs + code];
        o[static_cast<long long>(rr) * out_cs] = v;
        if (dup_ok) od[static_cast<long long>(rr) * out_cs] = v;
      }
    }
  }
}

// dW[r, q] += sum_{b, tau : code = q} g[b, r, tau];  dbias[r] += sum g[b, r, tau].
// CTA = 8 channels x 2048 steps of one batch item; per-channel 256-bin shared-memory accumulators, flushed with
// one global atomic per non-empty bin.
constexpr int BB_ROWS = 8;
constexpr int BB_TT = 2048;

__global__ void __launch_bounds__(256) base_embed_bwd_kernel(const float* __restrict__ g, long long g_bs, long long g_cs,
                                                             const float* __restrict__ wav, long long wav_pitch,
                                                             int off0, float* __restrict__ dw,
                                                             float* __restrict__ dbias, int R, int Q, int T) {
  extern __shared__ float sacc[];  // [BB_ROWS][Q]
  const int r0 = blockIdx.y * BB_ROWS;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * BB_TT;
  for (int i = threadIdx.x; i < BB_ROWS * Q; i += blockDim.x) sacc[i] = 0.0f;
  __syncthreads();
  for (int tt = threadIdx.x; tt < BB_TT; tt += blockDim.x) {
    const int t = t0 + tt;
    if (t >= T) break;
    int code = static_cast<int>(wav[static_cast<long long>(b) * wav_pitch + off0 + t]);
    code = code < 0 ? 0 : (code >= Q ? Q - 1 : code);
    const float* gp = g + static_cast<long long>(b) * g_bs + static_cast<long long>(r0) * g_cs + t;
#pragma unroll
    for (int rr = 0; rr < BB_ROWS; ++rr)
      if (r0 + rr < R) atomicAdd(&sacc[rr * Q + code], gp[static_cast<long long>(rr) * g_cs]);
  }
  __syncthreads();
  for (int rr = 0; rr < BB_ROWS; ++rr) {
    if (r0 + rr >= R) break;
    float rowsum = 0.0f;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      const float v = sacc[rr * Q + q];
      if (v != 0.0f) atomicAdd(dw + static_cast<long long>(r0 + rr) * Q + q, v);
      rowsum += v;
    }
    if (dbias) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o);
      if ((threadIdx.x & 31) == 0 && rowsum != 0.0f) atomicAdd(dbias + r0 + rr, rowsum);
    }
  }
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// out[b, c, t] = mask[b, c, t] > 0 ? g[b, c, t] : 0   (ReLU backward; wavenet.py:359-360, wave_encoder.py:39)
__global__ void relu_mask_bwd_kernel(const float* __restrict__ g, long long g_bs, long long g_cs,
                                     const float* __restrict__ mask, long long m_bs, long long m_cs,
                                     float* __restrict__ out, long long o_bs, long long o_cs, int C, int T) {
  const int b = blockIdx.z;
  const int c = blockIdx.y;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const float m = mask[static_cast<long long>(b) * m_bs + static_cast<long long>(c) * m_cs + t];
    const float v = g[static_cast<long long>(b) * g_bs + static_cast<long long>(c) * g_cs + t];
    out[static_cast<long long>(b) * o_bs + static_cast<long long>(c) * o_cs + t] = m > 0.0f ? v : 0.0f;
  }
}

// Weight repacking: a table of strided block copies  dst[i*di + j] = src[i*si + j*sj]  (i < ni, j < nj), one CTA per
// block.  Turns the live PyTorch parameters (out, in, tap) into the K-major zero-padded operand matrices of the GEMM
// engines with ONE launch per step (the table is built once; parameter and pack buffers have stable addresses).
__global__ void pack_blocks_kernel(const aewn_copy_block* __restrict__ blocks, int n_blocks) {
  for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
    const aewn_copy_block b = blocks[bi];
    const long long total = static_cast<long long>(b.ni) * b.nj;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
      const long long i = e / b.nj, j = e - i * b.nj;
      b.dst[i * b.di + j] = b.src[i * b.si + j * b.sj];
    }
  }
}

// Same table, accumulating:  dst[i*di + j] += src[i*si + j*sj].  One launch adds every weight gradient of a backward
// pass into the caller's .grad buffers (instead of one clone + one add launch per parameter).
__global__ void add_blocks_kernel(const aewn_copy_block* __restrict__ blocks, int n_blocks) {
  for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
    const aewn_copy_block b = blocks[bi];
    const long long total = static_cast<long long>(b.ni) * b.nj;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
      const long long i = e / b.nj, j = e - i * b.nj;
      b.dst[i * b.di + j] += b.src[i * b.si + j * b.sj];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// max|x| over n floats (bit pattern of a non-negative float orders like the float) and the power-of-two scale that puts it
// just below `target`: scale2[0] = 2^floor(log2(target / max|x|)), scale2[1] = 1 / scale2[0]  (1, 1 for an all-zero tensor)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ work) {
  float m = 0.0f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = x4[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(x[(n4 << 2) + threadIdx.x]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(work, __float_as_uint(m));      // NaN never compares greater: ignored
}

__global__ void amax_finish_kernel(const unsigned int* __restrict__ work, float target, float* __restrict__ scale2) {
  const float m = __uint_as_float(*work);
  float s = 1.0f;
  if (m > 0.0f && m < INFINITY) s = exp2f(floorf(log2f(target / m)));
  if (!(s > 0.0f) || s == INFINITY) s = 1.0f;
  scale2[0] = s;
  scale2[1] = 1.0f / s;
}

}  // namespace aewn

using namespace aewn;

extern "C" {

int aewn_base_embed_fwd(const float* wav, long long wav_pitch, int off0, const float* w, const float* bias, float* out,
                        long long out_bs, long long out_cs, float* dup, int dup_toff, int dup_t_hi, int batch, int R,
                        int Q, int T, int* err, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!wav || !w || !out || batch <= 0 || R <= 0 || Q <= 0 || T <= 0 || Q > 1024)
    return set_err(AEWN_ERR_INVALID, "base_embed_fwd: bad arguments");
  dim3 grid((T + BE_TT - 1) / BE_TT, (R + BE_ROWS - 1) / BE_ROWS, batch);
  const size_t smem = static_cast<size_t>(BE_ROWS) * (Q + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(base_embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_err(e, "base_embed_fwd: cudaFuncSetAttribute");
  }
  base_embed_fwd_kernel<<<grid, 256, smem, stream>>>(wav, wav_pitch, off0, w, bias, out, out_bs, out_cs, dup, dup_toff,
                                                     dup_t_hi, R, Q, T, err);
  count_launch();
  return cuda_err(cudaGetLastError(), "base_embed_fwd launch");
}

int aewn_base_embed_bwd(const float* g, long long g_bs, long long g_cs, const float* wav, long long wav_pitch, int off0,
                        float* dw, float* dbias, int batch, int R, int Q, int T, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!g || !wav || !dw || batch <= 0 || R <= 0 || Q <= 0 || T <= 0 || Q > 1024)
    return set_err(AEWN_ERR_INVALID, "base_embed_bwd: bad arguments");
  dim3 grid((T + BB_TT - 1) / BB_TT, (R + BB_ROWS - 1) / BB_ROWS, batch);
  const size_t smem = static_cast<size_t>(BB_ROWS) * Q * sizeof(float);
  base_embed_bwd_kernel<<<grid, 256, smem, stream>>>(g, g_bs, g_cs, wav, wav_pitch, off0, dw, dbias, R, Q, T);
  count_launch();
  return cuda_err(cudaGetLastError(), "base_embed_bwd launch");
}

int aewn_fill(float* p, long long n, float value, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || n <= 0) return set_err(AEWN_ERR_INVALID, "fill: bad arguments");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p, n, value);
  count_launch();
  return cuda_err(cudaGetLastError(), "fill launch");
}

int aewn_pack_blocks(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!blocks_dev || n_blocks <= 0) return set_err(AEWN_ERR_INVALID, "pack_blocks: bad arguments");
  int grid = n_blocks < 148 * 8 ? n_blocks : 148 * 8;
  pack_blocks_kernel<<<grid, 256, 0, stream>>>(blocks_dev, n_blocks);
  count_launch();
  return cuda_err(cudaGetLastError(), "pack_blocks launch");
}

int aewn_add_blocks(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!blocks_dev || n_blocks <= 0) return set_err(AEWN_ERR_INVALID, "add_blocks: bad arguments");
  int grid = n_blocks < 148 * 8 ? n_blocks : 148 * 8;
  add_blocks_kernel<<<grid, 256, 0, stream>>>(blocks_dev, n_blocks);
  count_launch();
  return cuda_err(cudaGetLastError(), "add_blocks launch");
}

int aewn_relu_mask_bwd(const float* g, long long g_bs, long long g_cs, const float* mask, long long m_bs, long long m_cs,
                       float* out, long long o_bs, long long o_cs, int batch, int C, int T, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!g || !mask || !out || batch <= 0 || C <= 0 || T <= 0) return set_err(AEWN_ERR_INVALID, "relu_mask_bwd: bad arguments");
  int bx = (T + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(bx, C, batch);
  relu_mask_bwd_kernel<<<grid, 256, 0, stream>>>(g, g_bs, g_cs, mask, m_bs, m_cs, out, o_bs, o_cs, C, T);
  count_launch();
  return cuda_err(cudaGetLastError(), "relu_mask_bwd launch");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Reconstruction loss (RecLoss, wavenet.py:536-552): -mean_{b,t} log_softmax(logits[b, :, t])[target[b, t]].
// Forward: one pass over the logits with an online max / sum-exp per (b, t) (lane = time step, so every channel row
// read is a coalesced 128 B line), writes lse[b, t] and adds the per-CTA partial of -(x_target - lse) to `loss_sum`.
// Backward: g_logits[b, q, t] = (exp(x - lse) - [q == target]) * scale.  Replaces log_softmax + gather + mean and their
// three backward kernels (2 reads + 1 write of the logits instead of ~7 passes).
// ---------------------------------------------------------------------------------------------------------------
namespace aewn {

__global__ void __launch_bounds__(256) nll_fwd_kernel(const float* __restrict__ x, long long x_bs, long long x_cs,
                                                      const float* __restrict__ tgt, long long t_bs,
                                                      float* __restrict__ lse, float* __restrict__ loss_sum, int Q,
                                                      int N, int* err) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  float part = 0.0f;
  if (t < N) {
    const float* p = x + static_cast<long long>(b) * x_bs + t;
    int code = static_cast<int>(tgt[static_cast<long long>(b) * t_bs + t]);
    if (code < 0 || code >= Q) {
      if (err) atomicExch(err, AEWN_ERR_INVALID);
      code = code < 0 ? 0 : Q - 1;
    }
    float m = -INFINITY, s = 0.0f, xt = 0.0f;
    for (int q = 0; q < Q; ++q) {
      const float v = __ldg(p);
      p += x_cs;
      if (q == code) xt = v;
      if (v > m) {
        s = s * __expf(m - v) + 1.0f;
        m = v;
      } else {
        s += __expf(v - m);
      }
    }
    const float l = m + __logf(s);
    lse[static_cast<long long>(b) * N + t] = l;
    part = l - xt;
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
    atomicAdd(loss_sum, tot);
  }
}

__global__ void __launch_bounds__(256) nll_bwd_kernel(const float* __restrict__ x, long long x_bs, long long x_cs,
                                                      const float* __restrict__ tgt, long long t_bs,
                                                      const float* __restrict__ lse, const float* __restrict__ g_loss,
                                                      float scale, float* __restrict__ gx, long long g_bs,
                                                      long long g_cs, int Q, int N) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  const float* p = x + static_cast<long long>(b) * x_bs + t;
  float* g = gx + static_cast<long long>(b) * g_bs + t;
  int code = static_cast<int>(tgt[static_cast<long long>(b) * t_bs + t]);
  code = code < 0 ? 0 : (code >= Q ? Q - 1 : code);
  const float l = lse[static_cast<long long>(b) * N + t];
  const float sc = scale * __ldg(g_loss);
#pragma unroll 8
  for (int q = 0; q < Q; ++q) {
    const float v = __ldg(p);
    *g = (__expf(v - l) - (q == code ? 1.0f : 0.0f)) * sc;
    p += x_cs;
    g += g_cs;
  }
}

}  // namespace aewn

extern "C" {

int aewn_nll_fwd(const float* logits, long long x_bs, long long x_cs, const float* target, long long t_bs, float* lse,
                 float* loss_sum, int batch, int Q, int N, int* err, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !lse || !loss_sum || batch <= 0 || Q <= 0 || N <= 0)
    return aewn::set_err(AEWN_ERR_INVALID, "nll_fwd: bad arguments");
  cudaError_t e = cudaMemsetAsync(loss_sum, 0, sizeof(float), stream);
  if (e != cudaSuccess) return aewn::cuda_err(e, "nll_fwd: memset");
  dim3 grid((N + 255) / 256, batch);
  aewn::nll_fwd_kernel<<<grid, 256, 0, stream>>>(logits, x_bs, x_cs, target, t_bs, lse, loss_sum, Q, N, err);
  aewn::count_launch();
  return aewn::cuda_err(cudaGetLastError(), "nll_fwd launch");
}

int aewn_nll_bwd(const float* logits, long long x_bs, long long x_cs, const float* target, long long t_bs,
                 const float* lse, const float* g_loss, float scale, float* g_logits, long long g_bs, long long g_cs,
                 int batch, int Q, int N, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !lse || !g_loss || !g_logits || batch <= 0 || Q <= 0 || N <= 0)
    return aewn::set_err(AEWN_ERR_INVALID, "nll_bwd: bad arguments");
  dim3 grid((N + 255) / 256, batch);
  aewn::nll_bwd_kernel<<<grid, 256, 0, stream>>>(logits, x_bs, x_cs, target, t_bs, lse, g_loss, scale, g_logits, g_bs,
                                                 g_cs, Q, N);
  aewn::count_launch();
  return aewn::cuda_err(cudaGetLastError(), "nll_bwd launch");
}

}  // extern "C"

extern "C" int aewn_amax_pow2_scale(const float* x, long long n, float target, unsigned int* work, float* scale2,
                                    aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !work || !scale2 || n <= 0 || !(target > 0.0f) || (reinterpret_cast<uintptr_t>(x) & 15u))
    return aewn::set_err(AEWN_ERR_INVALID, "amax_pow2_scale: bad arguments");
  if (int rc = aewn::cuda_err(cudaMemsetAsync(work, 0, sizeof(unsigned int), stream), "amax memset")) return rc;
  const int grid = static_cast<int>(std::min<long long>((n / 4 + 255) / 256 + 1, 148 * 8));
  aewn::amax_kernel<<<grid, 256, 0, stream>>>(x, n, work);
  aewn::amax_finish_kernel<<<1, 1, 0, stream>>>(work, target, scale2);
  aewn::count_launch();
  aewn::count_launch();
  return aewn::cuda_err(cudaGetLastError(), "amax launch");
}
