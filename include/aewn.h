/*
 * aewn.h -- C ABI of libaewn.so: the B200 (sm_100a) hot path of hrbigelow/ae-wavenet.
 *
 * The reference has no FFI layer: its hot path is reached through Python nn.Modules
 * (wavenet.py:15-111 GatedResidualCondConv, wavenet.py:323-364 WaveNet.forward_train, wave_encoder.py:34-50
 * ConvReLURes, vqema_bn.py:125-214 VQEMA.forward, vq_bn.py:28-61 VQ.forward).  The drop-in modules in
 * ae-wavenet_b200/aewn call ONLY the entry points declared here (through ctypes).
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller; nothing is allocated, freed or synchronised here;
 *   - every call enqueues on `stream` and returns 0, or a negative AEWN_ERR_* / -cudaError code; never throws;
 *   - activations are fp32, channel-major with time contiguous: elem(b, c, t) = ptr[b*batch_stride + c*row_pitch + t];
 *     tensors consumed by TMA need ptr 16-byte aligned and row_pitch, batch_stride multiples of 4 elements;
 *   - "absolute time": all tensors of one WaveNet stack share one time axis (tau = index into the first layer's
 *     input); a dilated tap is a negative shift on that axis.  TMA box origins must be 16-byte aligned
 *     (measured on B200: an unaligned inner coordinate raises "illegal instruction"), so every segment shift and
 *     every tile origin is a multiple of 4 elements; taps with dilation 1 or 2 read a pre-shifted duplicate that the
 *     producing kernel writes with its `dup` store.
 *   - kernels report device-side faults (e.g. a lost mbarrier arrival) through a caller-provided int32 error word
 *     (`err`), which the caller may read after synchronising; waits inside kernels are bounded, they never hang.
 */
#ifndef AEWN_H_
#define AEWN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* aewn_stream_t; /* cudaStream_t */

#define AEWN_OK 0
#define AEWN_ERR_INVALID (-1001)  /* bad argument (null pointer, misaligned pitch, size out of range) */
#define AEWN_ERR_DRIVER (-1002)   /* cuTensorMapEncodeTiled / driver entry point unavailable */
#define AEWN_ERR_TIMEOUT (-1003)  /* value written to the device error word when a bounded wait expires */
#define AEWN_ERR_RANGE (-1004)    /* device error word: an activation left the fp16 operand range (|x| > 65504) */

int aewn_version(void);
const char* aewn_last_error_string(void);
/* number of kernels launched by this library in this process since load (bench.py's gpu_launches) */
long long aewn_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Time-major GEMM with shifted segments ("tgemm"): the contraction engine behind every conv on the path.
 *
 *   acc[b, tau, n] = sum_s sum_k  act_s[b, k, tau + shift_s] * W[w_row + n, w_koff_s + k]
 *
 * One CTA tile = 128 time steps (UMMA M) x one "n-tile" (<= 256 output channels, UMMA N).  Activations are the
 * MN-major A operand, loaded by TMA straight from the NCT tensors (out-of-range time/channel coordinates read as
 * zero); W is a K-major matrix [rows][kpad] (kpad % 32 == 0, zero padded).  TF32 tcgen05.mma, FP32 accumulate in
 * TMEM.  A dilated conv layer (wavenet.py:100-101) is 3 segments: x@-d, x@0, cond@0.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* ptr;
  int t_extent;           /* valid time extent (TMA zero-fills beyond) */
  int channels;           /* valid channel extent (TMA zero-fills beyond) */
  int batch;
  long long row_pitch;    /* elements between channels */
  long long batch_stride; /* elements between batch items */
} aewn_act;

typedef struct {
  int act;      /* index into acts[] */
  int shift;    /* time shift: coordinate = tau + shift; must be a multiple of 4 (TMA 16-byte rule) */
  int channels; /* K rows consumed (rounded up to 32 inside; rows >= act.channels read as zero) */
  int w_koff;   /* first K column of W for this segment (multiple of 32) */
} aewn_seg;

/* epilogue modes */
#define AEWN_EPI_LINEAR 0   /* out = acc (+bias[n]) (+add[b,n,t]); flags below; optional dup store to out2 */
#define AEWN_EPI_GATE_FWD 1 /* n == 256: cols [0,128) filt, [128,256) gate -> out=tanh, out2=sigmoid, out3=z */
#define AEWN_EPI_GATE_BWD 2 /* acc = g_z; add=tanh, add2=sigmoid -> out=g_filt, out2=g_gate; optional dup store of
                               g_filt to out3 and of g_gate to out3 + (out2 - out) */
/* flags */
#define AEWN_F_ACCUM 1      /* out += value (read-modify-write) */
#define AEWN_F_RELU 2       /* value = max(value, 0) after bias/add */
#define AEWN_F_MASKPOS 4    /* `add` is a mask source, not an addend: value = add[b,n,t] > 0 ? value : 0 */
#define AEWN_F_RELU_FIRST 8 /* value = max(acc + bias, 0) + add  (wave_encoder.py:39-43 order); out3 (optional)
                               receives max(acc + bias, 0), the activation mask source for the backward pass */
#define AEWN_F_AB16 32      /* GATE_BWD: `add` holds {fp16 a, fp16 b} words (aewn_grcc_fwd, save == 2): g_filt = g_z a, g_gate = g_z b;
                               add2 unused */
#define AEWN_F_NO_OUT32 64   /* GATE_BWD with out16: skip the fp32 stores of g_filt / g_gate (every consumer reads the 16-bit copy) */
#define AEWN_F_MERGE_NEXT 16 /* this tile and the NEXT one in ntiles[] share one accumulator: one MMA of n + n_next (<= 256)
                                columns over their contiguous W rows, one pass over the activations.  Pair mode, LINEAR
                                tiles; the partner must be a plain store / AEWN_F_ACCUM tile on the TMA path and its
                                segments a subset of this tile's (W holds zeros where it has none) */

typedef struct {
  int w_row;      /* first W row of this n-tile */
  int n;          /* tile width: multiple of 16, 16..256 */
  int n_valid;    /* columns actually stored (<= n) */
  int mode;       /* AEWN_EPI_* */
  int flags;      /* AEWN_F_* */
  int seg_mask;   /* bit s set = segment s contributes to this tile */
  int t_lo, t_hi; /* store range on the absolute time axis; tiles outside are skipped */
  int t_zero_lo;  /* stores for tau < t_zero_lo write 0 (keeps the aligned-down margin of a tensor finite) */
  float* out;     /* pre-offset to the tile's first channel */
  float* out2;
  float* out3;
  long long out_bs, out_cs; /* batch / channel strides (elements) of out, out2, out3 */
  int out_toff;             /* out time index = tau + out_toff */
  int dup_toff;             /* dup store time index = tau + dup_toff (skipped when outside [0, dup_t_hi)) */
  int dup_t_hi;
  unsigned long long* zero_count; /* optional: += number of stored values equal to 0 (wave_encoder.py:46) */
  const float* add;         /* optional */
  const float* add2;
  long long add_bs, add_cs;
  int add_toff;
  int add_t_lo;             /* LINEAR: the addend / mask applies only for tau >= add_t_lo */
  const float* bias; /* optional, pre-offset, indexed by column */
  void* out16;       /* GATE_BWD, optional: bf16 channels-last copy of [g_filt; g_gate], element (b, t, c) at
                        out16[b * out16_bs + t * out16_cp + c], g_gate at channel offset (out2 - out) / out_cs */
  long long out16_bs;
  int out16_cp;
  const float* out16_scale; /* NULL: out16 holds bf16.  Else (device pointer to ONE float, a power of two): out16 holds
                               fp16(value * *out16_scale); a scaled value beyond the fp16 range raises AEWN_ERR_RANGE in
                               the descriptor's err word (see aewn_amax_pow2_scale, aewn_grcc_dgrad_desc.g_inv_scale) */
} aewn_ntile;

#define AEWN_CLUSTER_PAIR_MMA 102
#define AEWN_TG_DEFAULT_CLUSTER 2
#define AEWN_MAX_ACTS 6
#define AEWN_MAX_SEGS 6
#define AEWN_MAX_NTILES 4

typedef struct {
  aewn_act acts[AEWN_MAX_ACTS];
  int n_acts;
  aewn_seg segs[AEWN_MAX_SEGS];
  int n_segs;
  const float* w; /* K-major [w_rows][w_kpad] */
  int w_rows;
  int w_kpad;
  aewn_ntile ntiles[AEWN_MAX_NTILES];
  int n_ntiles;
  int batch;
  int t_begin, t_end; /* time range covered by tiles: tile i = [t_begin + 128 i, +128), t_begin % 32 == 0 */
  int* err;           /* device error word (may be NULL) */
  int max_ctas;       /* 0 = one per SM */
  int dbg_lbo, dbg_sbo; /* 0 = defaults; descriptor probing only */
  int cluster;        /* 0 = library default (AEWN_TG_DEFAULT_CLUSTER); 1, 2, 4 = CTAs per cluster sharing W by TMA
                         multicast, each CTA issuing cta_group::1 MMAs (M = 128); AEWN_CLUSTER_PAIR_MMA = 2-CTA clusters
                         issuing ONE cta_group::2 MMA stream (M = 256): each CTA stages its own 128 time steps and half
                         of the W rows, which cuts the shared-memory traffic per MMA by a third */
  int no_tma_store;   /* 1 = force the st.global epilogue (default: tiles whose outputs are 16-byte aligned are written
                         through shared memory + TMA store / reduce-add; rows of a partially active tile below t_lo
                         then receive zeros, i.e. the caller's margins must be don't-care or zero) */
} aewn_tgemm_desc;

int aewn_tgemm(const aewn_tgemm_desc* d, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Weight-gradient GEMM ("wgrad"): contraction over (batch, time).
 *
 *   out[m * out_rs + n * out_cs] += sum_b sum_{u in [t_lo,t_hi)}  G[b, g_row + m, u] * X[b, x_row + n, u + shift]
 *
 * Both operands are K-major (time contiguous), TMA-loaded; TF32 tcgen05.mma; split-K over (b, time) with fp32
 * red.global.add (callers zero `out` first; summation order is therefore not deterministic, like cuDNN wgrad).
 * Implements dW of wavenet.py:25-34.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int g_act, x_act;   /* indices into acts[] */
  int g_row, x_row;   /* first channel row of each operand */
  int m_valid;        /* rows stored (<= 128; 0 = padding partner of a pair, stores nothing) */
  int n;              /* columns of the tile: multiple of 16, 16..384 (two MMAs when > 256) */
  int n_valid;        /* columns stored */
  int shift;          /* X time coordinate = u + shift; multiple of 4 */
  int t_lo, t_hi;     /* u range (G's time axis); t_lo multiple of 4 */
  int n_split;        /* split-K factor for this item (>= 1) */
  float* out;
  long long out_rs, out_cs;
} aewn_wgrad_item;

#define AEWN_WGRAD_MAX_ACTS 6
#define AEWN_WGRAD_MAX_ITEMS 32

typedef struct {
  aewn_act acts[AEWN_WGRAD_MAX_ACTS];
  int n_acts;
  aewn_wgrad_item items[AEWN_WGRAD_MAX_ITEMS];
  int n_items;
  int batch;
  int* err;
  int max_ctas;
  int pair_x;  /* 1: items (2i, 2i+1) share their X tile; they run as a 2-CTA cluster and X is TMA-multicast;
                  2: same pairing, but the cluster issues cta_group::2 MMAs (M = 256 = both items' G rows) and each CTA
                     stages only half of the X rows (items must have n <= 256) */
} aewn_wgrad_desc;

int aewn_wgrad(const aewn_wgrad_desc* d, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Weight-gradient GEMM with wide units on CTA pairs ("wgradw"): same contraction as aewn_wgrad, tiled for operand
 * reuse.  A unit is 256 rows of G (m = 0..255) against up to AEWN_WGW_MAX_CHUNKS column chunks of X, each chunk with
 * its own operand, time shift and output matrix:
 *
 *   chunk.out[m * out_rs + n * out_cs] += sum_b sum_{u in [t_lo,t_hi)} G[b, g_row + m, u] * X[b, x_row + n, u + shift]
 *
 * The chunks of a unit need <= 512 accumulator columns (each chunk rounded up to 32) and <= 256 staged rows per CTA
 * (a chunk of n <= 128 columns stages 64, a wider one 128).  Every unit has its own split-K factor so that units of
 * different cost can be balanced over one wave of CTA pairs.  Callers zero the outputs first.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int x_act, x_row;   /* X operand: acts[] index and first channel row */
  int n;              /* columns of the chunk: multiple of 16, 16..256 */
  int n_valid;        /* columns stored */
  int shift;          /* X time coordinate = u + shift; multiple of 4 */
  int reserved;
  float* out;
  long long out_rs, out_cs;
} aewn_wgw_chunk;

#define AEWN_WGW_MAX_CHUNKS 3
#define AEWN_WGW_MAX_UNITS 8

typedef struct {
  int g_act, g_row;   /* G operand: rows [g_row, g_row + 256); rows beyond the tensor read as zero */
  int m_valid;        /* rows stored (1..256) */
  int t_lo, t_hi;     /* u range; t_lo multiple of 4 */
  int n_chunks;
  int n_split;        /* split-K factor of this unit (>= 1) */
  int reserved;
  aewn_wgw_chunk chunk[AEWN_WGW_MAX_CHUNKS];
} aewn_wgw_unit;

typedef struct {
  aewn_act acts[AEWN_WGRAD_MAX_ACTS];
  int n_acts;
  aewn_wgw_unit units[AEWN_WGW_MAX_UNITS];
  int n_units;
  int batch;
  int* err;
  int max_ctas;
} aewn_wgradw_desc;

int aewn_wgradw(const aewn_wgradw_desc* d, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * The same wide-unit weight gradient on 16-bit CHANNELS-LAST operands (csrc/wgradh.cu): fp16 tensors (batch, t_rows,
 * row_pitch) whose element (b, t, c) sits at ptr[b * batch_stride + t * row_pitch + c] -- the copies the fused layer
 * kernels keep (x16, cond16, the scaled fp16 [g_filt; g_gate] copy).  Units and chunks as for aewn_wgradw, with x_row /
 * g_row = first CHANNEL (multiple of 8) and shift = ANY time shift (a row coordinate: no 16-byte rule); t_lo need not be
 * aligned.  Rows of the last 64-step K block beyond t_hi must read as zero (or lie beyond t_rows).  The partial sums are
 * multiplied by *inv_scale (device pointer, may be NULL) before they are added to the outputs.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* ptr;        /* fp16 */
  int t_rows;             /* time extent of the tensor (TMA zero-fills beyond) */
  int channels;           /* channel extent, multiple of 8 (TMA zero-fills beyond) */
  int batch;
  long long row_pitch;    /* elements between time steps, multiple of 8 */
  long long batch_stride; /* elements between batch items, multiple of 8 */
} aewn_act16;

typedef struct {
  aewn_act16 acts[AEWN_WGRAD_MAX_ACTS];
  int n_acts;
  aewn_wgw_unit units[AEWN_WGW_MAX_UNITS];
  int n_units;
  int batch;
  int* err;
  int max_ctas;
  const float* inv_scale;
} aewn_wgradh_desc;

int aewn_wgradh(const aewn_wgradh_desc* d, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Decoder base layer (wavenet.py:348-351): one_hot(wav.long())[..., off0:off0+T] -> Conv1d(Q->R, k=1) evaluated as
 * a column gather  out[b, r, tau] = w[r, code(b, off0 + tau)] + bias[r]  (no one-hot tensor is materialised).
 * `dup` (optional) receives the same values at time index tau + dup_toff (pre-shifted copy for a dilation-1/2 tap).
 * A code outside [0, Q) sets *err = AEWN_ERR_INVALID (the reference raises in F.one_hot).
 * ------------------------------------------------------------------------------------------------------------ */
int aewn_base_embed_fwd(const float* wav, long long wav_pitch, int off0, const float* w, const float* bias, float* out,
                        long long out_bs, long long out_cs, float* dup, int dup_toff, int dup_t_hi, int batch, int R,
                        int Q, int T, int* err, aewn_stream_t stream);
/* dw[r, q] += sum_{b,tau: code == q} g[b, r, tau];  dbias[r] += sum g[b, r, tau]  (dbias may be NULL) */
int aewn_base_embed_bwd(const float* g, long long g_bs, long long g_cs, const float* wav, long long wav_pitch, int off0,
                        float* dw, float* dbias, int batch, int R, int Q, int T, aewn_stream_t stream);
int aewn_fill(float* p, long long n, float value, aewn_stream_t stream);
/* Weight repacking: n_blocks strided block copies dst[i*di + j] = src[i*si + j*sj] (i < ni, j < nj); `blocks_dev` is a
 * DEVICE array.  One launch turns the (out, in, tap) parameters of wavenet.py:25-34 into the K-major operand matrices. */
typedef struct {
  const float* src;
  float* dst;
  int ni, nj;
  long long si, sj, di;
} aewn_copy_block;
int aewn_pack_blocks(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream);
/* Same table, accumulating: dst[i*di + j] += src[i*si + j*sj].  Adds all weight gradients of one backward pass into the
 * caller's gradient buffers with one launch (what autograd does with one add per parameter, chassis.py:157). */
int aewn_add_blocks(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream);
/* out = mask > 0 ? g : 0   (ReLU backward) */
int aewn_relu_mask_bwd(const float* g, long long g_bs, long long g_cs, const float* mask, long long m_bs, long long m_cs,
                       float* out, long long o_bs, long long o_cs, int batch, int C, int T, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused dilation layer, forward (GatedResidualCondConv.forward, wavenet.py:91-111) -- ONE launch per layer:
 *
 *   filt | gate = [Wf0 Wf1 Pf bf | Wg0 Wg1 Pg bg] . [x(t - dil); x(t); cond(t); 1]      (wavenet.py:100-101)
 *   z = tanh(filt) * sigmoid(gate)                                                       (wavenet.py:102)
 *   x_next(t) = Wr . z + x(t)         skip(t) (+)= Ws . z                                (wavenet.py:103-110)
 *
 * Tensor-core operands are FP16 copies (10-bit mantissa like TF32, round-to-nearest), accumulation is FP32 in TMEM and
 * the residual stream x / x_next stays FP32.  Operand copies are CHANNELS-LAST:
 *   x16  (batch, t_rows, x16_cp)  fp16, x16_cp = R rounded up to 64, pad channels zero
 *   c16  (batch, t_rows, c16_cp)  fp16, channels [cond (n_cond1 - 1) | 1.0 | 0..], c16_cp = n_cond1 rounded up to 64
 *   w1h  [2 D][w1_k]              fp16 K-major, rows per 128-channel block: 128 filt rows, 128 gate rows; columns
 *                                 [tap x(t-dil) (x16_cp) | tap x(t) (x16_cp) | cond projection, bias (c16_cp)]
 *   w2h  [(final ? 0 : R) + S][D] fp16 K-major: dil_res rows then dil_skp rows
 * (aewn_cvt_f16_cl / aewn_pack_blocks_f16 produce them).  The kernel writes the next layer's x16 itself.  A value
 * outside the fp16 range sets *err = AEWN_ERR_RANGE (the stored operand saturates at +-65504).
 * Restrictions: D in {128, 256}; R % 8 == 0; S % 32 == 0; t_lo, skp_t_lo multiples of 4; fp32 strides multiples of 4.
 * Shapes outside them run through aewn_tgemm (two launches per layer, TF32).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x16;        /* layer input, fp16 channels-last */
  long long x16_bs;       /* elements between batch items (x16 and xo16) */
  int x16_cp;
  const void* c16;
  long long c16_bs;
  int c16_cp;
  int t_rows;             /* time rows of x16 / c16 / xo16 (>= t_hi) */
  const void* w1h;
  int w1_k;
  const void* w2h;
  const float* x32;       /* layer input, fp32 (batch, R, T): the residual addend */
  float* xo32;            /* x_next fp32, same strides as x32 (unused for the final layer) */
  long long x_bs, x_cs;
  void* xo16;             /* x_next fp16 channels-last, same geometry as x16 */
  float* dup;             /* optional: x_next again at time index t + dup_toff (same strides as x32), for the backward
                             pass's TF32 weight-gradient tap when the NEXT layer's dilation is not a multiple of 4 */
  int dup_toff, dup_t_hi;
  float* th;              /* save == 1: tanh(filt), sigmoid(gate) (batch, D, T) fp32 for the backward pass;
                             save == 2: th receives ONE 32-bit word per element = {fp16 a, fp16 b}, the derivative factors
                             a = sg (1 - th^2), b = th sg (1 - sg) (what AEWN_EPI_GATE_BWD + AEWN_F_AB16 reads); sg unused */
  float* sg;
  float* z;               /* optional: z (batch, D, T) */
  long long a_bs, a_cs;   /* strides of th / sg / z */
  int save;
  float* skp;             /* skip sum (batch, S, T) */
  long long s_bs, s_cs;
  int skp_mode;           /* 0: skp = Ws.z (first layer); 1: skp += Ws.z; 2: skp = relu(skp + Ws.z) (last layer, wavenet.py:359);
                             3: skp = relu(Ws.z) (a one-layer stack) */
  int batch, R, D, S;
  int n_cond1;            /* conditioning channels + 1 (the bias channel) */
  int dil;
  int final_layer;        /* 1: no dil_res, no x_next (wavenet.py:36-37,105-106) */
  int t_lo, t_zero_lo, t_hi;      /* outputs are stored on [t_lo, t_hi); values for t < t_zero_lo are written as 0 */
  int skp_t_lo, skp_zero_lo;      /* same for the skip sum */
  int* err;
  int max_ctas;           /* 0 = one per SM */
  long long* dbg_clock;   /* optional (profiling): 2 x 4 x 8 x 6 values for cluster 0's first four tiles --
                             [MMA issuer | epilogue warp 0][tile][job][clock64 at: job seen, operands / accumulator ready,
                             done; cycles: MMA = waiting for ring stages, -, waiting for z | epilogue = acquiring staging
                             tiles, TMEM loads, fence + TMA store issue] */
  void* z16;              /* optional: fp16 CHANNELS-LAST copy of z, element (b, t, c) at z16[b * z16_bs + t * z16_cp + c]
                             (the weight-gradient operand of aewn_wgradh; replaces the fp32 `z` output when that is NULL) */
  long long z16_bs;
  int z16_cp;             /* multiple of 16 */
} aewn_grcc_fwd_desc;

int aewn_grcc_fwd(const aewn_grcc_fwd_desc* d, aewn_stream_t stream);
/* ------------------------------------------------------------------------------------------------------------
 * Data gradient of a dilation layer on the fused-layer engine (autograd of wavenet.py:100-101 w.r.t. x and cond):
 *   g_x[t] = tap1^T gfg[t] + tap0^T gfg[t + dil] (+ g_sig[t] for t >= add_t_lo);     g_cond[t] += P^T gfg[t]
 * gfg = [g_filt; g_gate] is read from a bf16 CHANNELS-LAST copy (batch, t_rows, g16_cp) that the gate-derivative launch
 * writes (aewn_ntile.out16), the transposed weights from a bf16 K-major matrix w1t16 [R + n_cond][2 g16_cp] =
 * [tap0^T | tap1^T] whose conditioning rows hold zeros under the shifted block (aewn_pack_blocks_bf16).  bf16 because the
 * operands are gradients (range); accumulation and outputs are fp32.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* g16;
  long long g16_bs;        /* elements between batch items */
  int g16_cp;              /* channels per time row: 2 D rounded up to 64 */
  int t_rows;
  const void* w1t16;
  int w_k;                 /* = 2 * g16_cp */
  const float* g_sig;      /* optional addend (batch, R, T): gradient w.r.t. this layer's output (residual path) */
  float* gx;               /* (batch, R, T), same strides as g_sig */
  long long x_bs, x_cs;
  int add_t_lo;
  float* g_cond;           /* (batch, n_cond, T), accumulated */
  long long c_bs, c_cs;
  int n_cond;
  int batch, R, dil;
  int t_lo, t_zero_lo, t_hi;        /* g_x is stored on [t_lo, t_hi), zero below t_zero_lo */
  int cond_t_lo, cond_zero_lo;      /* g_cond receives contributions for t >= cond_zero_lo (cond_t_lo = its 4-aligned floor) */
  int* err;
  int max_ctas;
  void* gx16;               /* optional (scaled variant only): fp16 channels-last copy of gx * scale, element (b, t, c) at
                               gx16[b * gx16_bs + t * gx16_cp + c] -- the next layer's dil_res weight-gradient operand */
  long long gx16_bs;
  int gx16_cp;
  const float* g_inv_scale; /* NULL: g16 and w1t16 are bf16.  Else (device pointer to one float): g16 holds
                               fp16(gfg * scale) and w1t16 fp16 weights; the accumulators are multiplied by *g_inv_scale */
} aewn_grcc_dgrad_desc;

/* ------------------------------------------------------------------------------------------------------------
 * Gate derivative of a dilation layer on the fused-layer engine (autograd of wavenet.py:100-106 w.r.t. the two
 * pre-activations):  g_z = Wr^T g_x + Ws^T g_skp,  g_filt = g_z a,  g_gate = g_z b  with the {fp16 a,
 * fp16 b} words aewn_grcc_fwd saved (save == 2).  Operands: the SCALED fp16 channels-last copies of g_x (optional: absent
 * for the top layer) and g_skp, fp16 weights w2t16 [D][gx16_cp | gs16_cp] (K-major).  Output: the scaled fp16
 * channels-last copy g16 (b, t, [g_filt at c | g_gate at gg_off + c]); rows in [t_lo, t_zero_lo) are written as zeros.
 * |value| > 65504 raises AEWN_ERR_RANGE in *err.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* gx16;        /* (batch, t_rows, gx16_cp) or NULL */
  long long gx16_bs;
  int gx16_cp;             /* multiple of 64 */
  const void* gs16;        /* (batch, t_rows, gs16_cp) */
  long long gs16_bs;
  int gs16_cp;             /* multiple of 64 */
  int t_rows;
  const void* w2t16;
  int w_k;                 /* row pitch of w2t16 (elements) */
  int w_koff_skp;          /* first column of the Ws^T block (used when gx16 == NULL) */
  const void* ab;          /* (batch, D, T) 32-bit words {fp16 a, fp16 b} */
  long long a_bs, a_cs;
  void* g16;
  long long g16_bs;
  int g16_cp, gg_off;
  int batch, D;
  int t_lo, t_zero_lo, t_hi;
  int* err;
  int max_ctas;
} aewn_grcc_gz_desc;

int aewn_grcc_gz(const aewn_grcc_gz_desc* d, aewn_stream_t stream);

/* Power-of-two scale for a 16-bit copy of a gradient tensor: scale2[0] = 2^floor(log2(target / max|x|)) (1 if the tensor
 * is all zero), scale2[1] = 1 / scale2[0].  x: n contiguous floats; work: one unsigned int.  Two tiny launches. */
int aewn_amax_pow2_scale(const float* x, long long n, float target, unsigned int* work, float* scale2, aewn_stream_t stream);
int aewn_grcc_dgrad(const aewn_grcc_dgrad_desc* d, aewn_stream_t stream);
int aewn_pack_blocks_bf16(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream);

/* (batch, C, T) fp32 -> (batch, T, Cp) fp16 channels-last operand copy; channel ones_ch (>= 0) is written as 1.0, other
 * channels in [C, Cp) as 0.  Cp % 8 == 0, dst 16-byte aligned, d_bs (elements between batch items) % 8 == 0. */
/* the same copy with every value multiplied by *scale (device pointer to a power of two) first; |value * scale| > 65504
 * saturates and raises AEWN_ERR_RANGE in *err */
int aewn_cvt_f16_cl_scaled(const float* src, long long s_bs, long long s_cs, void* dst, long long d_bs, int Cp, int C, int T,
                           int batch, const float* scale, int* err, aewn_stream_t stream);
int aewn_cvt_f16_cl(const float* src, long long s_bs, long long s_cs, void* dst, long long d_bs, int Cp, int C, int T,
                    int batch, int ones_ch, int* err, aewn_stream_t stream);
/* aewn_pack_blocks writing fp16: dst is a __half matrix, di counted in halves */
int aewn_pack_blocks_f16(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused VQ step (vqema_bn.py:133-188, vq_bn.py:38-41): nearest code per (b, n) vector of ze (B, d, N).
 *   metric 0: squared L2 (vq_bn.py:39);  metric 1: scaled L2 |z-e| / (|z| + |e|) (vqema_bn.py:67-76)
 * Outputs: min_ind (B*N) int64, min_dist (B*N), zq (B, d, N) gathered codes; optional hist[K] += counts
 * (util.int_hist accumulate, vqema_bn.py:156), z_sum[K][d] / n_sum[K] (overwritten: per-code sums / counts,
 * vqema_bn.py:172-188), ze_norm (B*N).  Distances are IEEE fp32 in a fixed order (see vq.cu); first index wins ties.
 * ------------------------------------------------------------------------------------------------------------ */
int aewn_vq_fwd(const float* ze, long long ze_bs, long long ze_cs, const float* emb, int metric, long long* min_ind,
                float* min_dist, float* zq, long long zq_bs, long long zq_cs, float* hist, float* z_sum, float* n_sum,
                float* ze_norm, int batch, int d, int N, int K, aewn_stream_t stream);
/* The bottleneck's bias-free 1x1 projection (vqema_bn.py:92,131; vq_bn.py:18,35) in exact fp32 (sequential fmaf over k, no
 * tensor cores): out[b, n, t] = sum_k w[n * w_rs + k * w_cs] * x[b, k, t].  With (w_rs, w_cs) = (1, row pitch) it is the
 * data gradient.  The code indices that follow are an index computation; TF32 rounding here moves codes across near-ties. */
int aewn_conv1x1_f32(const float* x, long long x_bs, long long x_cs, const float* w, long long w_rs, long long w_cs,
                     float* out, long long o_bs, long long o_cs, int batch, int N, int K, int T, aewn_stream_t stream);
/* dw[n * K + k] = sum_{b,t} g[b, n, t] * x[b, k, t]  (overwrites dw; deterministic order) */
int aewn_conv1x1_wgrad_f32(const float* g, long long g_bs, long long g_cs, const float* x, long long x_bs, long long x_cs,
                           float* dw, int batch, int N, int K, int T, aewn_stream_t stream);
/* g_ze (+)= g_min_dist[b,n] * d(min_dist)/d(ze)   (commitment-loss gradient) */
int aewn_vq_commit_bwd(const float* ze, long long ze_bs, long long ze_cs, const float* emb, const long long* min_ind,
                       const float* g_min_dist, int metric, float* g_ze, long long g_bs, long long g_cs, int accumulate,
                       int batch, int d, int N, aewn_stream_t stream);
/* EMA update (vqema_bn.py:190-195) when z_sum != NULL; codebook refresh emb = numer/denom (vqema_bn.py:216-222) when
 * emb != NULL. */
int aewn_ema_update(float* ema_numer, float* ema_denom, const float* z_sum, const float* n_sum, float gamma, float* emb,
                    int K, int d, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Reconstruction loss (RecLoss.forward, wavenet.py:541-552): loss_sum = sum_{b,t} ( lse[b,t] - logits[b, target, t] ),
 * lse = log-sum-exp over the Q channels; the caller divides by batch*N.  aewn_nll_bwd writes
 * g_logits[b,q,t] = (exp(logits - lse) - [q == target]) * scale * (*g_loss).  Targets are float mu-law codes.
 * ------------------------------------------------------------------------------------------------------------ */
int aewn_nll_fwd(const float* logits, long long x_bs, long long x_cs, const float* target, long long t_bs, float* lse,
                 float* loss_sum, int batch, int Q, int N, int* err, aewn_stream_t stream);
int aewn_nll_bwd(const float* logits, long long x_bs, long long x_cs, const float* target, long long t_bs,
                 const float* lse, const float* g_loss, float scale, float* g_logits, long long g_bs, long long g_cs,
                 int batch, int Q, int N, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Incremental sampler (wavenet.py:367-531 WaveNet.forward_test): one PERSISTENT kernel generates many samples.
 *
 * The reference runs, per generated sample, 20 GatedResidualCondConv calls on (n_rep, R, d+1) slices plus the
 * post-net, softmax and torch.multinomial -- hundreds of launches and several host syncs per sample
 * (wavenet.py:455-509).  Here one thread-block cluster (`cluster` CTAs, distributed shared memory) owns `n_rep`
 * replicas: every CTA owns 1/cluster of the output rows of every matrix, streams exactly those rows from L2 through
 * a TMA bulk-copy ring (the row order is static, so the stream is a flat per-CTA array prepared once by the host),
 * exchanges the small activation vectors (z, x, h) through remote shared-memory stores and synchronises with a
 * cluster-scope mbarrier.  fp32 FMA arithmetic (no TF32).  State that outlives a launch -- the per-layer history
 * rings x_l[tau-d .. tau] and the generated codes -- is in global memory, so a long utterance is a sequence of
 * launches over [t_begin, t_end).
 *
 * Step tau: code = wav[rep][tau]; x_0 = base_t[code]; for every layer l (dilation d):
 *   v = [x_l[tau-d] | x_l[tau] | cond[tau] | 1],  z = tanh(A_f v) * sigmoid(A_g v),
 *   x_{l+1}[tau] = W_res z + b_res + x_l[tau],  skip += W_skp z + b_skp;
 * if tau+1 >= t_prime: logits = post2(relu(post1(relu(skip)))); wav[rep][tau+1] = first k with
 * cumsum(softmax(logits))[k] > uniforms[rep][tau+1]   (inverse-CDF draw, distributed like torch.multinomial).
 *
 * Slice-padded vector layout.  A vector of N elements owned 1/cluster per CTA (n = N/cluster each) is stored with
 * every CTA's slice padded to n_p = round4(n) floats -- element i lives at (i / n) * n_p + (i % n), total length
 * Np = cluster * n_p -- because slices travel between CTAs as 16-byte-granular bulk copies.  x (R), z (D), h0 (S),
 * h1 (P) and the logits (Q) use it; matrix columns that multiply such a vector are permuted / zero-padded to match.
 *
 * Per-CTA weight stream (floats), CTA rank c, in consumption order:
 *   kind 0 (gate)  : 2*D/cluster rows [filt_j, gate_j interleaved] x KA, KA = 2*Rp + cond_pitch, columns
 *                    [tap x[t-d] (Rp) | tap x[t] (Rp) | cond (C) | bias | 0..]
 *   kind 1 (mix)   : R/cluster residual rows (absent in the final layer) then S/cluster skip rows, x (Dp + 4)
 *   kind 2 (post1) : P/cluster rows x (Sp + 4);   kind 3 (post2): Q/cluster rows x (Pp + 4)
 * where the column at index Xp holds the bias (the vector carries 1.0 there).
 * ------------------------------------------------------------------------------------------------------------ */
#define AEWN_GEN_MAX_LAYERS 64
#define AEWN_GEN_MAX_BLOCKS (2 * AEWN_GEN_MAX_LAYERS + 2)
#define AEWN_GEN_MAX_REP 4

typedef struct {
  int kind;  /* 0 gate, 1 mix, 2 post1, 3 post2 */
  int rows;  /* rows of this CTA in the block */
  int rowf;  /* floats per row (multiple of 4) */
  int off;   /* float offset of the block inside the CTA's stream */
} aewn_gen_block;

typedef struct {
  int n_layers, R, D, S, P, Q;
  int cluster;      /* CTAs per cluster: 1, 2, 4, 8 or 16; must divide R, D, S, P and Q */
  int n_rep;        /* replicas per cluster: 1, 2 or 4 (pad with dummies) */
  int n_groups;     /* clusters; replica index = group * n_rep + rep */
  int t_begin, t_end, t_prime;
  int dil[AEWN_GEN_MAX_LAYERS];
  int hist_off[AEWN_GEN_MAX_LAYERS + 1]; /* slot offset of layer l's ring (ring length d_l + 1); [n_layers] = total */
  int n_blocks;
  aewn_gen_block blocks[AEWN_GEN_MAX_BLOCKS];
  const float* wstream;        /* [cluster][stream_stride] */
  long long stream_stride;
  const float* cond;           /* [cond_len][cond_pitch] time-major, = [cond(C) | 1 | 0..]; shared by all replicas */
  int cond_pitch, cond_len;
  const float* base_t;         /* [Q][base_pitch]: row q = base weight column q + bias (wavenet.py:253, 462-463),
                                  slice-padded x layout */
  int base_pitch;              /* = Rp */
  float* hist;                 /* [n_groups*n_rep][hist_off[n_layers]][Rp] (x layout), zero before the first launch */
  int* wav;                    /* [n_groups*n_rep][wav_pitch] int32 codes, read for tau < t_prime, written after */
  int wav_pitch;
  const float* uniforms;       /* [n_groups*n_rep][wav_pitch] U[0,1) */
  float* logits_out;           /* optional [n_groups*n_rep][wav_pitch][Q] (row tau+1 = logits that drew wav[tau+1]) */
  int stage_bytes, n_stages;   /* weight ring geometry (stage_bytes % 16 == 0, >= 16*KA) */
  int* err;
  long long* dbg_clock;        /* optional: cluster 0 / CTA 0 writes clock64() stamps of step t_begin+8 (profiling) */
} aewn_gen_desc;

/* dynamic shared memory the launch needs for this descriptor (bytes), or a negative error code */
int aewn_gen_smem_bytes(const aewn_gen_desc* d);
/* how many clusters of d->cluster CTAs can be co-resident on the current device (0 = not launchable) */
int aewn_gen_max_clusters(const aewn_gen_desc* d, int* n_out);
int aewn_gen_run(const aewn_gen_desc* d, aewn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Data-loader arithmetic (csrc/loader.cu).  Replaces, on the device:
 *   util.mu_encode_np / mu_encode_torch (util.py:62-67, 81-86), util.mu_decode_np / mu_decode_torch (util.py:70-78, 88-96),
 *   jitter.Jitter.__call__ (jitter.py:21-33) and mfcc.ProcessWav.__call__ (mfcc.py:39-76: librosa.feature.mfcc + two
 *   librosa.feature.delta calls on the host).
 * ------------------------------------------------------------------------------------------------------------ */
/* q = (sign(x) log1p(mu |x|) / log1p(mu) + 1) mu / 2 + 0.5 in fp32, mu = n_quanta - 1; torch_round = 0: truncate like
 * numpy's astype(int32) (util.py:67), 1: round to nearest even like torch's round_() (util.py:86).  out: int32 codes */
int aewn_mu_encode(const float* x, long long n, int n_quanta, int torch_round, int* out, aewn_stream_t stream);
/* x = sign(a) ((1 + mu)^|a| - 1) / mu,  a = (2 q - 1) / mu - 1 */
int aewn_mu_decode(const int* q, long long n, int n_quanta, float* out, aewn_stream_t stream);
/* out[b, 0] = 0, out[b, 1] = 1, out[b, t] = t - 1 + #{i : cdf_i <= u[b, t - 2]}, cdf = cumsum([p, 1 - 2p, p]) (jitter.py
 * indexes its table [p1][p1]: every step draws from that row).  u: (B, win - 2) doubles in [0, 1); out: (B, win) int64 */
int aewn_jitter_indices(const double* u, int B, int win, double p, long long* out, aewn_stream_t stream);

typedef struct {
  int n_fft, hop, n_mels, n_mfcc;   /* mfcc.py:28-29: win_sz, hop_sz, n_mels, n_mfcc */
  int left_pad, trim_left;          /* mfcc.py:48-50 */
  int n_frames_all;                 /* frames librosa computes: 1 + (L + left_pad) / hop (centered, reflect-padded) */
  int n_frames;                     /* frames kept: n_frames_all - trim_left - trim_right (>= 9) */
  float top_db;                     /* librosa.power_to_db: 80 */
  const double* twiddle;            /* [n_fft][2]: cos, sin of 2 pi j / n_fft */
  const double* window;             /* [n_fft]: periodic Hann */
  const float* melw;                /* [n_mels][n_fft/2 + 1]: librosa.filters.mel (Slaney scale, area-normalised) */
  const float* dctm;                /* [n_mfcc][n_mels]: DCT-II, orthonormal */
  const float* sg;                  /* [2][9][9]: Savitzky-Golay derivative rows, order 1 and 2: rows 0-3 left edge,
                                       4 interior, 5-8 right edge (scipy savgol_filter, window 9, mode 'interp') */
} aewn_mfcc_desc;

/* wav: (B, L) samples, wav_dtype 0 = uint8, 1 = int16, 2 = int32, 3 = float32 (the dat file's snd_dtype, data.py:36-41),
 * row pitch wav_bs elements.  work_db: B * n_mels * n_frames_all floats, work_max: B ints.  out[b, c, j]: (B, 3 n_mfcc,
 * n_frames) with strides out_bs / out_cs = MFCC rows, then first and second derivatives (mfcc.py:72-75) */
int aewn_mfcc(const void* wav, int wav_dtype, long long wav_bs, int B, int L, const aewn_mfcc_desc* d, float* work_db,
              int* work_max, float* out, long long out_bs, long long out_cs, aewn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AEWN_H_ */
