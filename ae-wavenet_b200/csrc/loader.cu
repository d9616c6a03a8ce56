// Data-loader arithmetic on the GPU: mu-law codec (util.py:62-96), jitter indices (jitter.py:21-33)
// and the MFCC + delta features that mfcc.ProcessWav computes with librosa on the host (mfcc.py:39-76).  All of it is small
// (one 16384-step window is ~100 frames of 400 samples): the point is to keep the loader off the host's critical path
// (librosa MFCC per item in Collate), not kernel speed.  Tables (DFT twiddles, Hann window, mel filterbank, DCT
// matrix, Savitzky-Golay rows) are built once by the host side (aewn/loader.py) and passed in.
#include <cmath>

#include "host_util.h"

namespace aewn {

// ---------------------------------------------------------------------------------------------------------------
// mu-law.  encode (util.py:62-67 numpy / :81-86 torch): amp = sign(x) log1p(mu |x|) / log1p(mu); q = (amp + 1) mu / 2 + 0.5;
// numpy truncates (astype(int32)), torch rounds to nearest even (round_()) -- both are offered.  fp32 arithmetic like
// the reference for float32 audio.  decode (util.py:70-78 / :88-96): a = (2 q - 1) / mu - 1; x = sign(a) ((1 + mu)^|a| - 1) / mu.
// ---------------------------------------------------------------------------------------------------------------
__global__ void mu_encode_kernel(const float* __restrict__ x, long long n, int n_quanta, int torch_round,
                                 int* __restrict__ out) {
  const float mu = static_cast<float>(n_quanta - 1);
  const float inv_l = log1pf(mu);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    const float sgn = (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f);
    const float amp = __fdiv_rn(__fmul_rn(sgn, log1pf(__fmul_rn(mu, fabsf(v)))), inv_l);
    const float q = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(amp, 1.0f), 0.5f), mu), 0.5f);
    out[i] = torch_round ? static_cast<int>(rintf(q)) : static_cast<int>(q);
  }
}

__global__ void mu_decode_kernel(const int* __restrict__ q, long long n, int n_quanta, float* __restrict__ out) {
  const float mu = static_cast<float>(n_quanta - 1);
  const float inv_mu = __frcp_rn(mu);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float qf = static_cast<float>(q[i]);
    const float a = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(2.0f, qf), -1.0f), inv_mu), -1.0f);
    const float sgn = (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f);
    out[i] = __fmul_rn(__fmul_rn(sgn, __fadd_rn(powf(1.0f + mu, fabsf(a)), -1.0f)), inv_mu);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// jitter.py:21-33 with the draws made explicit (oracle/loader_oracle.py jitter_from_uniforms): index[0] = 0, index[1] = 1,
// index[t] = t - 1 + #{cdf_i <= u[t-2]} with cdf = cumsum([p, 1 - 2p, p]) / sum  (the table is indexed [p1][p1], so every
// step draws from the same row).  u: (B, win - 2) doubles in [0, 1).
// ---------------------------------------------------------------------------------------------------------------
__global__ void jitter_kernel(const double* __restrict__ u, int B, int win, double p, long long* __restrict__ out) {
  const double s = 1.0 - 2.0 * p;
  double c0 = p, c1 = p + s, c2 = p + s + p;
  c0 /= c2;
  c1 /= c2;
  c2 /= c2;
  const long long n = static_cast<long long>(B) * win;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / win), t = static_cast<int>(i - static_cast<long long>(b) * win);
    long long v;
    if (t < 2) {
      v = t;
    } else {
      const double x = u[static_cast<long long>(b) * (win - 2) + (t - 2)];
      v = t - 1 + (c0 <= x) + (c1 <= x) + (c2 <= x);
    }
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// MFCC, stage 1: one CTA per (frame, item).  Frame f of librosa.stft(center=True, pad_mode='reflect') over
// wav_pad = [left_pad zeros | wav]: samples wav_pad[reflect(f hop + n - n_fft/2)], Hann window, real DFT in DOUBLE (numpy's
// FFT is double; librosa stores complex64), power = |.|^2 in fp32, mel = W . power, dB = 10 log10(max(1e-10, mel)).
// The per-item maximum (librosa.power_to_db's top_db clamp is relative to the maximum of the WHOLE call, trimmed frames
// included) is reduced with an ordered-integer atomicMax.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

template <typename T>
__global__ void __launch_bounds__(256) mfcc_mel_kernel(const T* __restrict__ wav, long long wav_bs, int L, aewn_mfcc_desc d,
                                                       float* __restrict__ db, int* __restrict__ maxbuf) {
  extern __shared__ double sm[];
  double* xs = sm;                         // [n_fft] windowed samples
  double* tw = xs + d.n_fft;               // [n_fft][2] cos, sin of 2 pi j / n_fft
  float* pw = reinterpret_cast<float*>(tw + 2 * d.n_fft);      // [n_bins] power
  const int f = blockIdx.x, b = blockIdx.y;
  const int n_fft = d.n_fft, n_bins = n_fft / 2 + 1;
  const int Lp = L + d.left_pad;           // len(wav_pad)
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    int k = f * d.hop + n - n_fft / 2;     // index into wav_pad before reflection
    if (k < 0) k = -k;
    if (k >= Lp) k = 2 * (Lp - 1) - k;
    double v = 0.0;
    if (k >= d.left_pad && k < Lp) v = static_cast<double>(wav[static_cast<long long>(b) * wav_bs + (k - d.left_pad)]);
    xs[n] = v * d.window[n];
    tw[2 * n] = d.twiddle[2 * n];
    tw[2 * n + 1] = d.twiddle[2 * n + 1];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_bins; k += blockDim.x) {
    double re = 0.0, im = 0.0;
    int j = 0;
    for (int n = 0; n < n_fft; ++n) {
      re = fma(xs[n], tw[2 * j], re);
      im = fma(-xs[n], tw[2 * j + 1], im);
      j += k;
      if (j >= n_fft) j -= n_fft;
    }
    const float a = hypotf(static_cast<float>(re), static_cast<float>(im));      // np.abs(complex64)
    pw[k] = a * a;
  }
  __syncthreads();
  float local_max = -INFINITY;
  for (int m = threadIdx.x; m < d.n_mels; m += blockDim.x) {
    const float* w = d.melw + static_cast<long long>(m) * n_bins;
    float acc = 0.0f;
    for (int k = 0; k < n_bins; ++k) acc = fmaf(w[k], pw[k], acc);
    const float v = 10.0f * log10f(fmaxf(1e-10f, acc));
    db[(static_cast<long long>(b) * d.n_mels + m) * d.n_frames_all + f] = v;
    local_max = fmaxf(local_max, v);
  }
  for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if ((threadIdx.x & 31) == 0 && local_max > -INFINITY) atomicMax(&maxbuf[b], float_ordered(local_max));
}

// stage 2: top_db clamp, DCT-II (ortho) to n_mfcc coefficients, left/right trim (mfcc.py:71): out[b, c, j], j = f - trim_left
__global__ void __launch_bounds__(128) mfcc_dct_kernel(aewn_mfcc_desc d, const float* __restrict__ db,
                                                       const int* __restrict__ maxbuf, float* __restrict__ out,
                                                       long long out_bs, long long out_cs) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d.n_frames) return;
  const float floor_db = ordered_float(maxbuf[b]) - d.top_db;
  const float* col = db + static_cast<long long>(b) * d.n_mels * d.n_frames_all + (j + d.trim_left);
  for (int c = 0; c < d.n_mfcc; ++c) {
    const float* dm = d.dctm + static_cast<long long>(c) * d.n_mels;
    float acc = 0.0f;
    for (int m = 0; m < d.n_mels; ++m) acc = fmaf(dm[m], fmaxf(col[static_cast<long long>(m) * d.n_frames_all], floor_db), acc);
    out[static_cast<long long>(b) * out_bs + c * out_cs + j] = acc;
  }
}

// stage 3: librosa.feature.delta(width 9, order 1 and 2, mode 'interp') = Savitzky-Golay derivative filters: 9-tap rows in
// the interior, polynomial-fit rows at the four first / last frames (tables from the host: sg[order-1][row][9], row 0..3
// left edge, 4 interior, 5..8 right edge).  Rows n_mfcc .. 3 n_mfcc - 1 of out.
__global__ void __launch_bounds__(128) mfcc_delta_kernel(aewn_mfcc_desc d, float* __restrict__ out, long long out_bs,
                                                         long long out_cs) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int F = d.n_frames;
  if (j >= F) return;
  const float* x = out + static_cast<long long>(b) * out_bs + c * out_cs;
  int row, start;
  if (j < 4) {
    row = j;
    start = 0;
  } else if (j >= F - 4) {
    row = 9 - (F - j);       // F-4 -> 5 ... F-1 -> 8
    start = F - 9;
  } else {
    row = 4;
    start = j - 4;
  }
  for (int o = 0; o < 2; ++o) {
    const float* cf = d.sg + (o * 9 + row) * 9;
    double acc = 0.0;
    for (int i = 0; i < 9; ++i) acc += static_cast<double>(cf[i]) * static_cast<double>(x[start + i]);
    out[static_cast<long long>(b) * out_bs + (d.n_mfcc * (o + 1) + c) * out_cs + j] = static_cast<float>(acc);
  }
}

}  // namespace aewn

using namespace aewn;

extern "C" int aewn_mu_encode(const float* x, long long n, int n_quanta, int torch_round, int* out, aewn_stream_t stream) {
  if (n <= 0) return AEWN_OK;
  if (!x || !out || n_quanta < 2) return AEWN_ERR_INVALID;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  mu_encode_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, n_quanta, torch_round, out);
  return cuda_err(cudaGetLastError(), "mu_encode launch");
}

extern "C" int aewn_mu_decode(const int* q, long long n, int n_quanta, float* out, aewn_stream_t stream) {
  if (n <= 0) return AEWN_OK;
  if (!q || !out || n_quanta < 2) return AEWN_ERR_INVALID;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  mu_decode_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, n, n_quanta, out);
  return cuda_err(cudaGetLastError(), "mu_decode launch");
}

extern "C" int aewn_jitter_indices(const double* u, int B, int win, double p, long long* out, aewn_stream_t stream) {
  if (B <= 0 || win <= 0) return AEWN_OK;
  if (!out || (win > 2 && !u) || p < 0.0 || p > 0.5) return AEWN_ERR_INVALID;
  const long long n = static_cast<long long>(B) * win;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  jitter_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(u, B, win, p, out);
  return cuda_err(cudaGetLastError(), "jitter launch");
}

extern "C" int aewn_mfcc(const void* wav, int wav_dtype, long long wav_bs, int B, int L, const aewn_mfcc_desc* d, float* work_db,
                         int* work_max, float* out, long long out_bs, long long out_cs, aewn_stream_t stream) {
  if (!wav || !d || !work_db || !work_max || !out) return AEWN_ERR_INVALID;
  if (B <= 0 || L <= 0 || d->n_fft < 2 || d->n_fft > 4096 || d->hop < 1 || d->n_mels < 1 || d->n_mfcc < 1 ||
      d->n_mfcc > d->n_mels || d->n_frames_all < 1 || d->n_frames < 9 || d->trim_left + d->n_frames > d->n_frames_all ||
      L + d->left_pad <= d->n_fft / 2 || !d->twiddle || !d->window || !d->melw || !d->dctm || !d->sg)
    return AEWN_ERR_INVALID;      // (fewer than 9 frames: scipy's savgol 'interp' mode is undefined, librosa raises)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // -inf in the ordered-integer encoding
  if (int rc = cuda_err(cudaMemsetAsync(work_max, 0x80, sizeof(int) * B, st), "mfcc memset")) return rc;
  const size_t smem = sizeof(double) * 3 * d->n_fft + sizeof(float) * (d->n_fft / 2 + 1);
  dim3 g1(d->n_frames_all, B);
#define AEWN_MEL(T)                                                                                              \
  do {                                                                                                           \
    if (smem > 48 * 1024)                                                                                        \
      cudaFuncSetAttribute(mfcc_mel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
    mfcc_mel_kernel<T><<<g1, 256, smem, st>>>(static_cast<const T*>(wav), wav_bs, L, *d, work_db, work_max);      \
  } while (0)
  switch (wav_dtype) {
    case 0: AEWN_MEL(unsigned char); break;
    case 1: AEWN_MEL(short); break;
    case 2: AEWN_MEL(int); break;
    case 3: AEWN_MEL(float); break;
    default: return AEWN_ERR_INVALID;
  }
#undef AEWN_MEL
  dim3 g2((d->n_frames + 127) / 128, B);
  mfcc_dct_kernel<<<g2, 128, 0, st>>>(*d, work_db, work_max, out, out_bs, out_cs);
  dim3 g3((d->n_frames + 127) / 128, d->n_mfcc, B);
  mfcc_delta_kernel<<<g3, 128, 0, st>>>(*d, out, out_bs, out_cs);
  return cuda_err(cudaGetLastError(), "mfcc launch");
}
