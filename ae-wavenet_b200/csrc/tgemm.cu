// tgemm: time-major GEMM with shifted activation segments on tcgen05 (TF32 in, FP32 accumulate in TMEM).
//
// Implements the contraction of every conv on the reference hot path (wavenet.py:100-109 forward, SURVEY.md 9.1
// backward-data) as   acc[b, tau, n] = sum_s sum_k act_s[b, k, tau + shift_s] * W[w_row + n, w_koff_s + k].
//
// CTA = 12 warps, persistent over (batch, 128-step time tile, n-tile) work items:
//   warp 0      TMA producer  : per 32-row K block, 4 boxes {32 t x 32 k} of the activation (MN-major A operand,
//                               SWIZZLE_128B_ATOM_32B) + up to 2 boxes {32 k x 128 n} of W (K-major B, SWIZZLE_128B)
//   warp 1      MMA issuer    : 4 x tcgen05.mma.kind::tf32 (M=128, N=n, K=8) per K block, accumulator in TMEM
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages x 256)
//   warps 4-11  epilogue      : tcgen05.ld -> registers -> fused elementwise -> coalesced global stores
//                               (lane = time step, so each warp store instruction writes one full 128 B line)
// Pipelines: 4-stage smem ring (full/empty mbarriers) and a 2-stage TMEM ring (tfull/tempty), so the epilogue of
// tile i overlaps the MMAs of tile i+1.
#include "host_util.h"
#include "ptx.cuh"

namespace aewn {

constexpr int TG_BM = 128;
constexpr int TG_BK = 32;
constexpr int TG_STAGES = 4;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 4;   // 16 KB
constexpr int TG_WBOX_BYTES = 128 * TG_BK * 4;  // 16 KB per 128-row W box
constexpr int TG_STAGE_BYTES = TG_A_BYTES + 2 * TG_WBOX_BYTES;  // 48 KB
constexpr int TG_THREADS = 384;
constexpr int TG_EPI_WARPS = 8;
constexpr int TG_SMEM_BYTES = TG_STAGES * TG_STAGE_BYTES + 256 + 1024;  // + barriers + alignment slack

struct TgSeg {
  int map;
  int shift;
  int kblocks;
  int w_koff;
};

struct TgParams {
  CUtensorMap a_map[AEWN_MAX_ACTS];
  CUtensorMap w_map;
  TgSeg seg[AEWN_MAX_SEGS];
  int n_segs;
  aewn_ntile nt[AEWN_MAX_NTILES];
  int n_ntiles;
  int batch;
  int t_begin;
  int n_ttiles;
  int* err;
  uint32_t a_lbo, a_sbo;
  int cluster;   // 1, 2 or 4: CTAs of a cluster work on adjacent time tiles of the same (batch, n-tile) and share W
  int n_tgroups; // ceil(n_ttiles / cluster)
};

struct TgItem {
  int ni, tau0, b;
  bool active;
};

// `item` indexes (batch, time-tile GROUP, n-tile); the CTAs of a cluster take consecutive tiles of the group.  The
// activity test is evaluated on the whole group so that every CTA of a cluster walks the same item sequence (they
// exchange multicast data and barrier arrivals); a CTA whose own tile lies outside the store range just stores nothing.
__device__ __forceinline__ TgItem tg_decode(const TgParams& p, int item, int crank) {
  TgItem it;
  it.ni = item % p.n_ntiles;
  int r = item / p.n_ntiles;
  int tg = r % p.n_tgroups;
  it.b = r / p.n_tgroups;
  const int g0 = p.t_begin + tg * p.cluster * TG_BM;
  it.tau0 = g0 + crank * TG_BM;
  const aewn_ntile& nt = p.nt[it.ni];
  it.active = (g0 + p.cluster * TG_BM > nt.t_lo) && (g0 < nt.t_hi);
  return it;
}

// ------------------------------------------------------------------------------------------------ epilogues
// Common conventions: lane = time step tau; stores happen for tau in [t_lo, t_hi); values for tau < t_zero_lo are
// forced to 0 so that the aligned-down margin of every tensor stays finite (TMA reads it, 0 * garbage must be 0).
// Addend / accumulate source of a LINEAR tile, software-pipelined: the loads of column chunk i+1 are issued before
// chunk i is processed, and those of the first chunk BEFORE the epilogue waits for the accumulator, so the global
// load latency hides behind the MMAs instead of stalling each chunk (GEMM2 of a GRCC layer is epilogue-bound).
struct LinSrc {
  const float* p;   // element (b, first channel of tile, tau [+ toff])
  long long cs;
  bool ok;          // this lane may load
  bool is_add;      // true: `add` operand (or mask);  false: previous value of `out` (accumulate)
  float fill;
};

__device__ __forceinline__ LinSrc lin_src(const aewn_ntile& nt, int b, int tau) {
  LinSrc s;
  const bool in_range = (tau >= nt.t_lo) && (tau < nt.t_hi);
  if (nt.add) {
    s.p = nt.add + static_cast<long long>(b) * nt.add_bs + (tau + nt.add_toff);
    s.cs = nt.add_cs;
    s.ok = in_range && tau >= nt.t_zero_lo && tau >= nt.add_t_lo;
    s.is_add = true;
    s.fill = (nt.flags & AEWN_F_MASKPOS) ? 1.0f : 0.0f;
  } else if (nt.flags & AEWN_F_ACCUM) {
    s.p = nt.out + static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
    s.cs = nt.out_cs;
    s.ok = in_range;
    s.is_add = false;
    s.fill = 0.0f;
  } else {
    s.p = nullptr;
    s.cs = 0;
    s.ok = false;
    s.is_add = false;
    s.fill = 0.0f;
  }
  return s;
}

__device__ __forceinline__ void lin_issue(const LinSrc& s, int n_valid, int c0, float (&a)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float* q = s.p + static_cast<long long>(c0 + j) * s.cs;
    a[j] = (s.ok && c0 + j < n_valid) ? (s.is_add ? __ldg(q) : __ldcg(q)) : s.fill;
  }
}

struct LinCtx {
  bool in_range, live, accum, relu, relu_first, maskpos, both, dup_ok;
  float* outp;
  float* dupp;
};

// One 32-column chunk of a LINEAR tile.  `buf` holds the prefetched addend / previous-output values of this chunk.
__device__ __forceinline__ void lin_chunk(const aewn_ntile& nt, const LinCtx& cx, const LinSrc& src, uint32_t taddr, int c0,
                                          int b, int tau, const float (&buf)[32], unsigned int& zeros) {
  uint32_t v[32];
  tmem_ld32(taddr + c0, v);
  tmem_ld_wait();
  if (c0 >= nt.n_valid) return;
  float r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
  if (nt.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < nt.n_valid) r[j] += __ldg(nt.bias + c0 + j);
  }
  if (cx.relu_first) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = fmaxf(r[j], 0.0f);
    if (nt.out3) {  // keep relu(pre) so the backward pass has the exact activation mask (wave_encoder.py:39)
      float* o3 = nt.out3 + static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cx.in_range && c0 + j < nt.n_valid) o3[static_cast<long long>(c0 + j) * nt.out_cs] = cx.live ? r[j] : 0.0f;
    }
  }
  if (src.p && src.is_add) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = cx.maskpos ? (buf[j] > 0.0f ? r[j] : 0.0f) : r[j] + buf[j];
  }
  if (!cx.live) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = 0.0f;
  }
  if (cx.accum) {
    if (cx.both) {
      float prev[32];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        prev[j] = (cx.in_range && c0 + j < nt.n_valid) ? __ldcg(cx.outp + static_cast<long long>(c0 + j) * nt.out_cs) : 0.0f;
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] += prev[j];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] += buf[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (cx.in_range && c0 + j < nt.n_valid) {
      float x = r[j];
      if (cx.relu) x = fmaxf(x, 0.0f);
      cx.outp[static_cast<long long>(c0 + j) * nt.out_cs] = x;
      if (cx.dup_ok) cx.dupp[static_cast<long long>(c0 + j) * nt.out_cs] = x;
      zeros += (x == 0.0f) ? 1u : 0u;
    }
  }
}

// LINEAR epilogue of one tile.  Column chunks c0 = 32*half + 64*i; chunk i uses buffer A (i even) or B (i odd), and
// as soon as a buffer is consumed the loads of chunk i+2 are issued into it: two chunks (2 x 32 x 128 B per warp) stay
// in flight, the first two are issued by the caller BEFORE it waits for the accumulator.
__device__ __forceinline__ void epi_linear(const aewn_ntile& nt, uint32_t taddr, int half, int b, int tau,
                                           const LinSrc& src, float (&bufA)[32], float (&bufB)[32]) {
  LinCtx cx;
  cx.in_range = (tau >= nt.t_lo) && (tau < nt.t_hi);
  cx.live = tau >= nt.t_zero_lo;
  cx.outp = nt.out + static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
  const int dup_t = tau + nt.dup_toff;
  cx.dup_ok = nt.out2 && cx.in_range && dup_t >= 0 && dup_t < nt.dup_t_hi;
  cx.dupp = nt.out2 ? nt.out2 + static_cast<long long>(b) * nt.out_bs + dup_t : nullptr;
  cx.accum = (nt.flags & AEWN_F_ACCUM) != 0;
  cx.relu = (nt.flags & AEWN_F_RELU) != 0;
  cx.relu_first = (nt.flags & AEWN_F_RELU_FIRST) != 0;
  cx.maskpos = (nt.flags & AEWN_F_MASKPOS) != 0;
  cx.both = nt.add && cx.accum;  // rare: addend prefetched, previous value loaded inline
  unsigned int zeros = 0;
  for (int c0 = half * 32; c0 < nt.n; c0 += 128) {
    lin_chunk(nt, cx, src, taddr, c0, b, tau, bufA, zeros);
    if (src.p && c0 + 128 < nt.n && c0 + 128 < nt.n_valid) lin_issue(src, nt.n_valid, c0 + 128, bufA);
    if (c0 + 64 < nt.n) {
      lin_chunk(nt, cx, src, taddr, c0 + 64, b, tau, bufB, zeros);
      if (src.p && c0 + 192 < nt.n && c0 + 192 < nt.n_valid) lin_issue(src, nt.n_valid, c0 + 192, bufB);
    }
  }
  if (nt.zero_count) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(nt.zero_count, static_cast<unsigned long long>(zeros));
  }
}

// wavenet.py:102  z = tanh(filt) * sigmoid(gate); columns [0,128) = filt, [128,256) = gate of the same channels.
__device__ __forceinline__ void epi_gate_fwd(const aewn_ntile& nt, uint32_t taddr, int half, int b, int tau) {
  const bool in_range = (tau >= nt.t_lo) && (tau < nt.t_hi);
  const bool live = tau >= nt.t_zero_lo;
  const long long off = static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
  for (int c0 = half * 32; c0 < 128; c0 += 64) {
    uint32_t vf[32], vg[32];
    tmem_ld32(taddr + c0, vf);
    tmem_ld32(taddr + 128 + c0, vg);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float f = __uint_as_float(vf[j]);
      float g = __uint_as_float(vg[j]);
      if (nt.bias) {
        f += __ldg(nt.bias + c0 + j);
        g += __ldg(nt.bias + 128 + c0 + j);
      }
      const float th = live ? fast_tanh(f) : 0.0f;
      const float sg = live ? fast_sigmoid(g) : 0.0f;
      if (in_range && c0 + j < nt.n_valid) {
        const long long o = off + static_cast<long long>(c0 + j) * nt.out_cs;
        if (nt.out) nt.out[o] = th;
        if (nt.out2) nt.out2[o] = sg;
        nt.out3[o] = th * sg;
      }
    }
  }
}

// g_f = g_z * sg * (1 - th^2), g_g = g_z * th * sg * (1 - sg); acc columns = g_z of all D channels.
struct GateBwdCtx {
  bool in_range, live, dup_ok;
  long long ooff, aoff, doff, g_delta;
};

__device__ __forceinline__ void gbwd_issue(const aewn_ntile& nt, const GateBwdCtx& cx, int c0, float (&th)[32],
                                           float (&sg)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const bool ok = cx.in_range && cx.live && (c0 + j < nt.n_valid);
    th[j] = ok ? __ldg(nt.add + cx.aoff + static_cast<long long>(c0 + j) * nt.add_cs) : 0.0f;
    sg[j] = ok ? __ldg(nt.add2 + cx.aoff + static_cast<long long>(c0 + j) * nt.add_cs) : 0.0f;
  }
}

__device__ __forceinline__ void gbwd_chunk(const aewn_ntile& nt, const GateBwdCtx& cx, uint32_t taddr, int c0,
                                           const float (&th)[32], const float (&sg)[32]) {
  uint32_t v[32];
  tmem_ld32(taddr + c0, v);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (cx.in_range && c0 + j < nt.n_valid) {
      const float gz = cx.live ? __uint_as_float(v[j]) : 0.0f;
      const float gs = gz * sg[j];
      const float gf = cx.live ? gs * (1.0f - th[j] * th[j]) : 0.0f;
      const float gg = cx.live ? gs * th[j] * (1.0f - sg[j]) : 0.0f;
      const long long o = cx.ooff + static_cast<long long>(c0 + j) * nt.out_cs;
      nt.out[o] = gf;
      nt.out2[o] = gg;
      if (cx.dup_ok) {
        const long long od = cx.doff + static_cast<long long>(c0 + j) * nt.out_cs;
        nt.out3[od] = gf;
        nt.out3[od + cx.g_delta] = gg;
      }
    }
  }
}

__device__ __forceinline__ GateBwdCtx gbwd_ctx(const aewn_ntile& nt, int b, int tau) {
  GateBwdCtx cx;
  cx.in_range = (tau >= nt.t_lo) && (tau < nt.t_hi);
  cx.live = tau >= nt.t_zero_lo;
  cx.ooff = static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
  cx.aoff = static_cast<long long>(b) * nt.add_bs + (tau + nt.add_toff);
  const int dup_t = tau + nt.dup_toff;
  cx.dup_ok = nt.out3 && cx.in_range && dup_t >= 0 && dup_t < nt.dup_t_hi;
  cx.doff = static_cast<long long>(b) * nt.out_bs + dup_t;
  cx.g_delta = nt.out2 - nt.out;  // g_gate rows follow g_filt rows in the same tensor
  return cx;
}

// tanh / sigmoid of the first chunk are loaded by the caller before it waits for the accumulator; those of chunk i+1
// are issued before chunk i is processed.
__device__ __forceinline__ void epi_gate_bwd(const aewn_ntile& nt, uint32_t taddr, int half, const GateBwdCtx& cx,
                                             float (&thA)[32], float (&sgA)[32]) {
  float thB[32], sgB[32];
  for (int c0 = half * 32; c0 < nt.n; c0 += 128) {
    const bool hasB = c0 + 64 < nt.n;
    if (hasB) gbwd_issue(nt, cx, c0 + 64, thB, sgB);
    gbwd_chunk(nt, cx, taddr, c0, thA, sgA);
    if (c0 + 128 < nt.n) gbwd_issue(nt, cx, c0 + 128, thA, sgA);
    if (hasB) gbwd_chunk(nt, cx, taddr, c0 + 64, thB, sgB);
  }
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(TG_THREADS, 1) tgemm_kernel(const __grid_constant__ TgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TG_STAGES * TG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + TG_STAGES;
  uint64_t* tfull_bar = empty_bar + TG_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int i = 0; i < TG_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], p.cluster);   // one tcgen05.commit arrival from every CTA that reads the stage's W
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], TG_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < AEWN_MAX_ACTS; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.w_map);
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();   // peers' barriers must be initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int crank = p.cluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cid = blockIdx.x / p.cluster;            // cluster index; all CTAs of a cluster walk the same items
  const int n_clusters = gridDim.x / p.cluster;
  const uint16_t cmask = static_cast<uint16_t>((1u << p.cluster) - 1u);
  const int total = p.batch * p.n_tgroups * p.n_ntiles;

  // Register reallocation: warps 0-3 (TMA / MMA / TMEM-alloc roles, one warpgroup) shrink to 88 registers and the 8
  // epilogue warps grow to 208, so 32-wide column chunks + prefetch buffers stay in registers
  // (128*88 + 256*208 = 64512 <= 65536).  Each setmaxnreg dominates its role code (no merge of limits).
  if (warp < 4) {
  reg_dealloc<40>();
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      for (int item = cid; item < total && ok; item += n_clusters) {
        const TgItem it = tg_decode(p, item, crank);
        if (!it.active) continue;
        const aewn_ntile& nt = p.nt[it.ni];
        // W rows are split into `cluster` slices of wrows each; CTA r loads slice r and multicasts it to all peers
        const int wrows = 256 / p.cluster;
        const int wslices = (nt.n + wrows - 1) / wrows;
        for (int s = 0; s < p.n_segs && ok; ++s) {
          if (!((nt.seg_mask >> s) & 1)) continue;
          const TgSeg sg = p.seg[s];
          for (int kb = 0; kb < sg.kblocks; ++kb) {
            if (!mbar_wait(&empty_bar[stage], phase ^ 1u, abort_flag)) { ok = false; break; }
            uint8_t* sa = smem + stage * TG_STAGE_BYTES;
            uint8_t* sw = sa + TG_A_BYTES;
            const int wboxes = (nt.n + 127) >> 7;
            mbar_expect_tx(&full_bar[stage], TG_A_BYTES + (p.cluster == 1 ? wboxes * TG_WBOX_BYTES : wslices * wrows * 128));
#pragma unroll
            for (int i = 0; i < 4; ++i)
              tma_load_3d(sa + i * 4096, &p.a_map[sg.map], &full_bar[stage], it.tau0 + sg.shift + 32 * i, kb * TG_BK,
                          it.b);
            if (p.cluster == 1) {
              for (int j = 0; j < wboxes; ++j)
                tma_load_2d(sw + j * TG_WBOX_BYTES, &p.w_map, &full_bar[stage], sg.w_koff + kb * TG_BK,
                            nt.w_row + j * 128);
            } else if (crank < wslices) {
              tma_load_2d_mcast(sw + crank * wrows * 128, &p.w_map, &full_bar[stage], sg.w_koff + kb * TG_BK,
                                nt.w_row + crank * wrows, cmask);
            }
            if (++stage == TG_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ok = true;
      for (int item = cid; item < total && ok; item += n_clusters) {
        const TgItem it = tg_decode(p, item, crank);
        if (!it.active) continue;
        const aewn_ntile& nt = p.nt[it.ni];
        if (!mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, abort_flag)) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        const uint32_t idesc = make_idesc_tf32(TG_BM, nt.n, /*a_mn=*/1, /*b_mn=*/0);
        uint32_t kiter = 0;
        for (int s = 0; s < p.n_segs && ok; ++s) {
          if (!((nt.seg_mask >> s) & 1)) continue;
          const int kblocks = p.seg[s].kblocks;
          for (int kb = 0; kb < kblocks; ++kb) {
            if (!mbar_wait(&full_bar[stage], phase, abort_flag)) { ok = false; break; }
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * TG_STAGE_BYTES);
            const uint32_t w_addr = a_addr + TG_A_BYTES;
#pragma unroll
            for (int ks = 0; ks < TG_BK / 8; ++ks) {
              // A: MN-major, 128B swizzle with 32B atoms: 4-row groups 512 B apart (SBO), 32-step chunks 4 KB apart (LBO)
              const uint64_t adesc = make_smem_desc(a_addr + ks * 1024, p.a_lbo, p.a_sbo, kLayoutSW128Base32);
              // W: K-major, 128B swizzle: 8-row groups 1 KB apart (SBO); K advances 32 B inside the swizzle row
              const uint64_t bdesc = make_smem_desc(w_addr + ks * 32, 16, 1024, kLayoutSW128);
              umma_tf32_ss(d_tmem, adesc, bdesc, idesc, (kiter | ks) != 0u);
            }
            if (p.cluster == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_mcast(&empty_bar[stage], cmask);
            ++kiter;
            if (++stage == TG_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        if (!ok) break;
        umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }
  } else {
    reg_alloc<232>();
    // ===================================================== epilogue
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t acc = 0, acc_phase = 0;
    for (int item = cid; item < total; item += n_clusters) {
      const TgItem it = tg_decode(p, item, crank);
      if (!it.active) continue;
      const aewn_ntile& nt = p.nt[it.ni];
      const int tau = it.tau0 + q * 32 + lane;
      // issue the first chunk's addend / accumulate loads before waiting for the accumulator
      LinSrc src;
      src.p = nullptr;
      float pre[32], pre2[32];
      GateBwdCtx gcx;
      if (nt.mode == AEWN_EPI_LINEAR) {
        src = lin_src(nt, it.b, tau);
        if (src.p && half * 32 < nt.n_valid) lin_issue(src, nt.n_valid, half * 32, pre);
        if (src.p && half * 32 + 64 < nt.n_valid) lin_issue(src, nt.n_valid, half * 32 + 64, pre2);
      } else if (nt.mode == AEWN_EPI_GATE_BWD) {
        gcx = gbwd_ctx(nt, it.b, tau);
        gbwd_issue(nt, gcx, half * 32, pre, pre2);
      }
      if (!mbar_wait(&tfull_bar[acc], acc_phase, abort_flag)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256u + (static_cast<uint32_t>(q * 32) << 16);
      if (nt.mode == AEWN_EPI_LINEAR) epi_linear(nt, taddr, half, it.b, tau, src, pre, pre2);
      else if (nt.mode == AEWN_EPI_GATE_FWD) epi_gate_fwd(nt, taddr, half, it.b, tau);
      else epi_gate_bwd(nt, taddr, half, gcx, pre, pre2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();   // no CTA may exit while a peer can still multicast into its smem
  if (threadIdx.x == 0 && *abort_flag && p.err) atomicExch(p.err, AEWN_ERR_TIMEOUT);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace aewn

using namespace aewn;

extern "C" int aewn_tgemm(const aewn_tgemm_desc* d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_err(AEWN_ERR_INVALID, "tgemm: null descriptor");
  if (d->n_acts < 1 || d->n_acts > AEWN_MAX_ACTS || d->n_segs < 1 || d->n_segs > AEWN_MAX_SEGS ||
      d->n_ntiles < 1 || d->n_ntiles > AEWN_MAX_NTILES)
    return set_err(AEWN_ERR_INVALID, "tgemm: n_acts/n_segs/n_ntiles out of range (%d/%d/%d)", d->n_acts, d->n_segs,
                   d->n_ntiles);
  if (d->batch <= 0 || d->t_end <= d->t_begin || (d->t_begin & 31))
    return set_err(AEWN_ERR_INVALID, "tgemm: need batch>0, t_end>t_begin, t_begin%%32==0 (b=%d t=[%d,%d))", d->batch,
                   d->t_begin, d->t_end);

  {
    cudaError_t e = cudaFuncSetAttribute(tgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES);
    if (e != cudaSuccess) return cuda_err(e, "tgemm: cudaFuncSetAttribute");
  }

  TgParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < d->n_acts; ++i) {
    int rc = encode_act_map(&p.a_map[i], d->acts[i], TG_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (d->acts[i].batch < d->batch) return set_err(AEWN_ERR_INVALID, "tgemm: act %d batch smaller than problem batch", i);
  }
  for (int i = d->n_acts; i < AEWN_MAX_ACTS; ++i) p.a_map[i] = p.a_map[0];
  int cluster = d->cluster > 0 ? d->cluster : 2;
  if (cluster != 1 && cluster != 2 && cluster != 4) return set_err(AEWN_ERR_INVALID, "tgemm: cluster must be 1, 2 or 4");
  int rc = encode_w_map(&p.w_map, d->w, d->w_rows, d->w_kpad, cluster == 1 ? 128 : 256 / cluster);
  if (rc) return rc;

  for (int s = 0; s < d->n_segs; ++s) {
    const aewn_seg& sg = d->segs[s];
    if (sg.act < 0 || sg.act >= d->n_acts || sg.channels <= 0 || (sg.w_koff & 31) || sg.w_koff < 0)
      return set_err(AEWN_ERR_INVALID, "tgemm: bad segment %d", s);
    if (sg.shift & 3)
      return set_err(AEWN_ERR_INVALID, "tgemm: segment %d shift %d is not a multiple of 4 (TMA 16-byte origin rule)", s,
                     sg.shift);
    p.seg[s].map = sg.act;
    p.seg[s].shift = sg.shift;
    p.seg[s].kblocks = (sg.channels + TG_BK - 1) / TG_BK;
    p.seg[s].w_koff = sg.w_koff;
    if (sg.w_koff + p.seg[s].kblocks * TG_BK > d->w_kpad)
      return set_err(AEWN_ERR_INVALID, "tgemm: segment %d overruns W's K extent (%d + %d > %d)", s, sg.w_koff,
                     p.seg[s].kblocks * TG_BK, d->w_kpad);
  }
  p.n_segs = d->n_segs;
  const int all_mask = (1 << d->n_segs) - 1;
  for (int i = 0; i < d->n_ntiles; ++i) {
    aewn_ntile nt = d->ntiles[i];
    if (nt.n < 16 || nt.n > 256 || (nt.n & 15) || nt.n_valid < 1 || nt.n_valid > nt.n)
      return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d width invalid (n=%d n_valid=%d)", i, nt.n, nt.n_valid);
    if ((nt.seg_mask & all_mask) == 0) return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d uses no segment", i);
    if (nt.mode == AEWN_EPI_LINEAR) {
      if (!nt.out) return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d has no output", i);
    } else if (nt.mode == AEWN_EPI_GATE_FWD) {
      if (nt.n != 256 || !nt.out3 || nt.n_valid > 128)
        return set_err(AEWN_ERR_INVALID, "tgemm: GATE_FWD tile %d needs n=256, n_valid<=128 and out3", i);
    } else if (nt.mode == AEWN_EPI_GATE_BWD) {
      if (!nt.out || !nt.out2 || !nt.add || !nt.add2)
        return set_err(AEWN_ERR_INVALID, "tgemm: GATE_BWD tile %d needs out, out2, add, add2", i);
    } else {
      return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d unknown mode %d", i, nt.mode);
    }
    if (nt.w_row < 0 || nt.w_row >= d->w_rows) return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d w_row out of range", i);
    if (nt.t_lo < d->t_begin) nt.t_lo = d->t_begin;
    if (nt.t_hi > d->t_end) nt.t_hi = d->t_end;
    p.nt[i] = nt;
  }
  p.n_ntiles = d->n_ntiles;
  p.batch = d->batch;
  p.t_begin = d->t_begin;
  p.n_ttiles = (d->t_end - d->t_begin + TG_BM - 1) / TG_BM;
  p.err = d->err;
  p.a_lbo = d->dbg_lbo > 0 ? d->dbg_lbo : 4096;
  p.a_sbo = d->dbg_sbo > 0 ? d->dbg_sbo : 512;

  p.cluster = cluster;
  p.n_tgroups = (p.n_ttiles + cluster - 1) / cluster;
  const long long total = static_cast<long long>(p.batch) * p.n_tgroups * p.n_ntiles;   // items per cluster walk
  int clusters = (d->max_ctas > 0 ? d->max_ctas : sm_count()) / cluster;
  if (clusters > total) clusters = static_cast<int>(total);
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * cluster);
  cfg.blockDim = dim3(TG_THREADS);
  cfg.dynamicSmemBytes = TG_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, tgemm_kernel, p);
  count_launch();
  if (le != cudaSuccess) return cuda_err(le, "tgemm launch");
  return cuda_err(cudaGetLastError(), "tgemm launch");
}
