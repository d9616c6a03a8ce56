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
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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
constexpr int TG_STG_BYTES = TG_EPI_WARPS * 4096;  // one 32 ch x 32 t fp32 staging tile per epilogue warp (TMA stores)
constexpr int TG_PAIR_STG_BYTES = 2 * TG_STG_BYTES;  // pair mode: TWO tiles per warp (the next box is written while the
                                                    // TMA engine still reads the previous one), paid for with one ring stage
constexpr int TG_RING_BYTES = TG_STAGES * TG_STAGE_BYTES;      // 192 KB: 4 x 48 KB, or 6 x 32 KB in pair-MMA mode
constexpr int TG_PAIR_STAGES = 6;      // pair mode, one staging tile per warp (long K loops: ring depth hides HBM latency)
constexpr int TG_PAIR_STAGES_STG2 = 5; // pair mode, two staging tiles per warp (short K loops: the epilogue is the critical path)
constexpr int TG_PAIR_STAGE_BYTES = TG_A_BYTES + TG_WBOX_BYTES;  // 32 KB: own 128 time steps of A + own half of W
constexpr int TG_MAX_STAGES = 6;
static_assert(TG_PAIR_STAGES_STG2 * TG_PAIR_STAGE_BYTES + TG_PAIR_STG_BYTES == TG_RING_BYTES + TG_STG_BYTES &&
              TG_PAIR_STAGES * TG_PAIR_STAGE_BYTES == TG_RING_BYTES, "every mode uses the same amount of shared memory");
constexpr int TG_SMEM_BYTES = TG_RING_BYTES + TG_STG_BYTES + 256 + 1024;  // + barriers + alignment slack

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
  // AEWN_F_MERGE_NEXT: tile i and tile i+1 share ONE accumulator (one MMA of n_i + n_{i+1} columns over contiguous W rows,
  // one pass over the A operand); work items walk the head tiles only
  int n_heads;
  int head[AEWN_MAX_NTILES];
  int merged[AEWN_MAX_NTILES];
  int batch;
  int t_begin;
  int n_ttiles;
  int* err;
  uint32_t a_lbo, a_sbo;
  int cluster;   // 1, 2 or 4: CTAs of a cluster work on adjacent time tiles of the same (batch, n-tile) and share W
  int stg2;      // 1 (pair only): 5 ring stages + two epilogue staging tiles per warp instead of 6 + one
  int pair;      // 1 (cluster == 2 only): the pair issues cta_group::2 MMAs (M = 256), each CTA stages its own 128 time
                 // steps of A and HALF of the W rows in its own shared memory: no multicast, 2/3 of the smem traffic
  int n_tgroups; // ceil(n_ttiles / cluster)
  // Output maps for the TMA-store epilogue: o_map[tile][0] covers `out` (for GATE_BWD: the whole g_f;g_g tensor),
  // [1] `out2`, [2] `out3`.  o_tma[tile] != 0 when the tile's main outputs go through TMA (see tma_eligible()).
  CUtensorMap o_map[AEWN_MAX_NTILES][3];
  int o_tma[AEWN_MAX_NTILES];
  int gg_ch_off[AEWN_MAX_NTILES];  // GATE_BWD: channel offset of g_gate inside the g_f;g_g tensor
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
  it.ni = p.head[item % p.n_heads];
  int r = item / p.n_heads;
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
//
// Instruction economy matters here (the LINEAR / GATE_BWD tiles have little MMA work to hide behind): every n-tile field
// the inner loops need is copied into registers ONCE per item (the descriptor lives in the kernel parameter space and
// is selected by a runtime index, so touching it per element costs a constant-bank load + 64-bit address rebuild
// each time -- measured 35 instructions per stored element), row pointers advance by the channel stride, stores are
// predicated rather than branched, and fully valid 32-column chunks take a path with no per-column bounds test.

// ---- TMA-store staging ---------------------------------------------------------------------------------------------
// Each epilogue warp owns one 4 KB tile [32 channels][32 time steps].  Lane = time step, so element (j, lane) sits at
// j*128 + lane*4 bytes: bank-conflict free, and the offsets are compile-time immediates (one STS per element, no
// address arithmetic).  After a proxy fence one lane issues a TMA store (or reduce-add) of the box; rows outside the
// tensor extents (time >= t_hi, channel >= n_valid) are clipped by the hardware.  Rows of a partially active tile
// that lie below t_lo are written as zeros (stores) / add zero (reduce) -- by construction those positions are
// margins that hold zeros anyway.
struct StgOut {
  float* tile;       // this warp's staging tile(s)
  int ntiles;        // 1, or 2 (alternating: the box of store k is written while store k-1 is still being read)
  mutable int cur;   // tile of the most recent stg_acquire
  int t0;            // first time step of this warp's 32-row slab (tile origin + 32*q), output coordinates
  int b;
  int lane;
  bool slab_on;      // slab intersects [t_lo, t_hi)
};

// Returns this lane's column of the tile to fill next (element (j, lane) at [j * 32]).
__device__ __forceinline__ float* stg_acquire(const StgOut& so) {
  so.cur ^= so.ntiles - 1;
  if (elect_one()) {   // elect.sync names the same lane every time, i.e. the one that committed the bulk groups
    if (so.ntiles == 2) tma_store_wait_read1();   // all but the latest box have been read out: the older tile is free
    else tma_store_wait_read();
  }
  __syncwarp();
  return so.tile + so.cur * 1024 + so.lane;
}
__device__ __forceinline__ void stg_flush(const StgOut& so, const CUtensorMap* map, int c0, bool reduce) {
  fence_proxy_async_smem();                  // make this lane's generic-proxy writes visible to the async proxy
  __syncwarp();
  if (elect_one()) {
    if (reduce) tma_reduce_add_3d(map, so.tile + so.cur * 1024, so.t0, c0, so.b);
    else tma_store_3d(map, so.tile + so.cur * 1024, so.t0, c0, so.b);
    tma_store_commit();
  }
}

// ---- LINEAR -------------------------------------------------------------------------------------------------------
struct LinRegs {
  const CUtensorMap* omap;   // non-null: `out` goes through the TMA-store path
  bool tma_reduce;
  float* outp;         // (b, tile channel 0, tau + out_toff)
  float* dupp;         // dup store row or nullptr
  float* out3p;        // relu(pre) row or nullptr
  const float* srcp;   // prefetch source row (addend, mask source, or previous output) or nullptr
  const float* bias;
  long long out_cs, src_cs;
  int n, n_valid;
  float fill;
  bool in_range, live, src_ok, src_is_add, accum, relu, relu_first, maskpos, count;
};

__device__ __forceinline__ LinRegs lin_regs(const aewn_ntile& nt, int b, int tau, const CUtensorMap* omap) {
  LinRegs c;
  c.omap = omap;
  c.tma_reduce = omap && (nt.flags & AEWN_F_ACCUM);
  c.in_range = (tau >= nt.t_lo) && (tau < nt.t_hi);
  c.live = tau >= nt.t_zero_lo;
  c.out_cs = nt.out_cs;
  c.outp = nt.out + static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
  const int dup_t = tau + nt.dup_toff;
  const bool dup_ok = nt.out2 && c.in_range && dup_t >= 0 && dup_t < nt.dup_t_hi;
  c.dupp = dup_ok ? nt.out2 + static_cast<long long>(b) * nt.out_bs + dup_t : nullptr;
  c.relu_first = (nt.flags & AEWN_F_RELU_FIRST) != 0;
  c.out3p = (c.relu_first && nt.out3) ? nt.out3 + static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff) : nullptr;
  c.accum = (nt.flags & AEWN_F_ACCUM) != 0;
  c.relu = (nt.flags & AEWN_F_RELU) != 0;
  c.maskpos = (nt.flags & AEWN_F_MASKPOS) != 0;
  c.count = nt.zero_count != nullptr;
  c.bias = nt.bias;
  c.n = nt.n;
  c.n_valid = nt.n_valid;
  if (nt.add) {
    c.srcp = nt.add + static_cast<long long>(b) * nt.add_bs + (tau + nt.add_toff);
    c.src_cs = nt.add_cs;
    c.src_ok = c.in_range && c.live && tau >= nt.add_t_lo;
    c.src_is_add = true;
    c.fill = c.maskpos ? 1.0f : 0.0f;
  } else if (c.accum && !c.tma_reduce) {
    c.srcp = c.outp;
    c.src_cs = nt.out_cs;
    c.src_ok = c.in_range;
    c.src_is_add = false;
    c.fill = 0.0f;
  } else {
    c.srcp = nullptr;
    c.src_cs = 0;
    c.src_ok = false;
    c.src_is_add = false;
    c.fill = 0.0f;
  }
  return c;
}

// Issue the 32 loads of one column chunk of the prefetch source (select, not branch: all 32 go out back to back).
__device__ __forceinline__ void lin_issue(const LinRegs& c, int c0, float (&a)[32]) {
  const float* q = c.srcp + static_cast<long long>(c0) * c.src_cs;
  const int nrem = c.n_valid - c0;
  if (nrem >= 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      a[j] = c.src_ok ? __ldcg(q) : c.fill;
      q += c.src_cs;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      a[j] = (c.src_ok && j < nrem) ? __ldcg(q) : c.fill;
      q += c.src_cs;
    }
  }
}

template <bool FULL>
__device__ __forceinline__ void lin_chunk(const LinRegs& c, const StgOut& so, uint32_t taddr, int c0,
                                          const float (&buf)[32], unsigned int& zeros) {
  uint32_t v[32];
  tmem_ld32(taddr + c0, v);
  tmem_ld_wait();
  const int nrem = c.n_valid - c0;
  if (nrem <= 0) return;
  float r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
  if (c.bias) {
    const float* bp = c.bias + c0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (FULL || j < nrem) r[j] += __ldg(bp + j);
  }
  if (c.relu_first) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = fmaxf(r[j], 0.0f);
    if (c.out3p) {  // keep relu(pre): exact activation mask for the backward pass (wave_encoder.py:39)
      float* o3 = c.out3p + static_cast<long long>(c0) * c.out_cs;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (c.in_range && (FULL || j < nrem)) *o3 = c.live ? r[j] : 0.0f;
        o3 += c.out_cs;
      }
    }
  }
  if (c.srcp && c.src_is_add) {
    if (c.maskpos) {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = buf[j] > 0.0f ? r[j] : 0.0f;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] += buf[j];
    }
  }
  if (!c.live) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = 0.0f;
  }
  if (c.accum && !c.tma_reduce) {   // (an n-tile has either an addend or the accumulate flag, never both)
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] += buf[j];
  }
  if (c.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = fmaxf(r[j], 0.0f);
  }
  if (c.omap) {
    if (so.slab_on) {
      float* st = stg_acquire(so);
#pragma unroll
      for (int j = 0; j < 32; ++j) st[j * 32] = c.in_range ? r[j] : 0.0f;
      stg_flush(so, c.omap, c0, c.tma_reduce);
    }
  } else {
    float* o = c.outp + static_cast<long long>(c0) * c.out_cs;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (c.in_range && (FULL || j < nrem)) *o = r[j];
      o += c.out_cs;
    }
  }
  if (c.dupp) {
    float* d = c.dupp + static_cast<long long>(c0) * c.out_cs;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (FULL || j < nrem) *d = r[j];
      d += c.out_cs;
    }
  }
  if (c.count && c.in_range) {
#pragma unroll
    for (int j = 0; j < 32; ++j) zeros += ((FULL || j < nrem) && r[j] == 0.0f) ? 1u : 0u;
  }
}

__device__ __forceinline__ void lin_chunk_any(const LinRegs& c, const StgOut& so, uint32_t taddr, int c0,
                                              const float (&buf)[32], unsigned int& zeros) {
  if (c.n_valid - c0 >= 32) lin_chunk<true>(c, so, taddr, c0, buf, zeros);
  else lin_chunk<false>(c, so, taddr, c0, buf, zeros);
}

// LINEAR epilogue of one tile.  Column chunks c0 = 32*half + 64*i; chunk i uses buffer A (i even) or B (i odd), and
// as soon as a buffer is consumed the loads of chunk i+2 are issued into it: two chunks (2 x 32 x 128 B per warp) stay
// in flight; the first two are issued by the caller BEFORE it waits for the accumulator.
__device__ __forceinline__ void epi_linear(const LinRegs& c, const StgOut& so, uint32_t taddr, int half,
                                           float (&bufA)[32], float (&bufB)[32], unsigned long long* zero_count) {
  unsigned int zeros = 0;
  for (int c0 = half * 32; c0 < c.n; c0 += 128) {
    lin_chunk_any(c, so, taddr, c0, bufA, zeros);
    if (c.srcp && c0 + 128 < c.n_valid) lin_issue(c, c0 + 128, bufA);
    if (c0 + 64 < c.n) {
      lin_chunk_any(c, so, taddr, c0 + 64, bufB, zeros);
      if (c.srcp && c0 + 192 < c.n_valid) lin_issue(c, c0 + 192, bufB);
    }
  }
  if (c.count) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(zero_count, static_cast<unsigned long long>(zeros));
  }
}

// Second tile of a merged pair (AEWN_F_MERGE_NEXT): plain "accumulator -> TMA store / reduce-add", no addend, bias or
// ReLU (validated host-side).  Its columns start at an arbitrary offset inside the 256-column accumulator stage, so the
// last 32-column group is aligned to the tile's END (it must not read past the stage); the columns that group shares with
// the previous one are re-stored unchanged (plain store) or contribute zero (reduce-add).
__device__ __forceinline__ void epi_linear_tail(const aewn_ntile& nt, const StgOut& so, uint32_t taddr, int half, int tau,
                                                const CUtensorMap* omap) {
  const bool keep = (tau >= nt.t_lo) && (tau < nt.t_hi) && (tau >= nt.t_zero_lo);
  const bool reduce = (nt.flags & AEWN_F_ACCUM) != 0;
  for (int c0 = half * 32; c0 < nt.n; c0 += 64) {
    int c0e = c0, skip = 0;
    if (c0 + 32 > nt.n) {
      c0e = nt.n - 32;
      skip = c0 - c0e;
    }
    uint32_t v[32];
    tmem_ld32(taddr + c0e, v);
    tmem_ld_wait();
    if (so.slab_on) {
      float* st = stg_acquire(so);
#pragma unroll
      for (int j = 0; j < 32; ++j) st[j * 32] = (keep && (j >= skip || !reduce)) ? __uint_as_float(v[j]) : 0.0f;
      stg_flush(so, omap, c0e, reduce);
    }
  }
}

// ---- GATE_FWD -----------------------------------------------------------------------------------------------------
// wavenet.py:102  z = tanh(filt) * sigmoid(gate); columns [0,128) = filt, [128,256) = gate of the same channels.
__device__ __forceinline__ void epi_gate_fwd(const aewn_ntile& nt, const StgOut& so, const CUtensorMap* omaps,
                                             uint32_t taddr, int half, int b, int tau) {
  const bool in_range = (tau >= nt.t_lo) && (tau < nt.t_hi);
  const bool live = tau >= nt.t_zero_lo;
  const long long cs = nt.out_cs;
  const long long off = static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
  float* const th_base = nt.out ? nt.out + off : nullptr;
  float* const sg_base = nt.out2 ? nt.out2 + off : nullptr;
  float* const z_base = nt.out3 + off;
  const float* const bias = nt.bias;
  const int n_valid = nt.n_valid;
  // Fast path (warp-uniform decision): every row of this warp's slab is live and in range, full 128-channel block, no
  // separate bias, TMA stores.  It is the common case (all but the first tile of a sequence) and carries no per-element
  // selects: ~14 instructions per (filt, gate) pair instead of ~59 -- the gated epilogue was the kernel's critical path
  // once the MMA issue loop had been fixed (profiles/r2_*).
  const bool fast = omaps && !bias && n_valid == 128 && so.slab_on && __all_sync(0xffffffffu, in_range && live);
  for (int c0 = half * 32; c0 < 128; c0 += 64) {
    uint32_t vf[32], vg[32];
    tmem_ld32(taddr + c0, vf);
    tmem_ld32(taddr + 128 + c0, vg);
    tmem_ld_wait();
    if (fast) {
      float* st;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        vf[j] = __float_as_uint(fast_tanh(__uint_as_float(vf[j])));
        vg[j] = __float_as_uint(fast_sigmoid(__uint_as_float(vg[j])));
      }
      if (th_base) {
        st = stg_acquire(so);
#pragma unroll
        for (int j = 0; j < 32; ++j) st[j * 32] = __uint_as_float(vf[j]);
        stg_flush(so, &omaps[0], c0, false);
      }
      if (sg_base) {
        st = stg_acquire(so);
#pragma unroll
        for (int j = 0; j < 32; ++j) st[j * 32] = __uint_as_float(vg[j]);
        stg_flush(so, &omaps[1], c0, false);
      }
      st = stg_acquire(so);
#pragma unroll
      for (int j = 0; j < 32; ++j) st[j * 32] = __uint_as_float(vf[j]) * __uint_as_float(vg[j]);
      stg_flush(so, &omaps[2], c0, false);
      continue;
    }
    const int nrem = n_valid - c0;
    if (nrem <= 0) continue;
    float th[32], sg[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float f = __uint_as_float(vf[j]);
      float g = __uint_as_float(vg[j]);
      if (bias) {
        f += __ldg(bias + c0 + j);
        g += __ldg(bias + 128 + c0 + j);
      }
      th[j] = live ? fast_tanh(f) : 0.0f;
      sg[j] = live ? fast_sigmoid(g) : 0.0f;
    }
    if (omaps) {   // TMA-store path: tanh, sigmoid, z through the warp's staging tile, one box each
      if (so.slab_on) {
        float* st;
        if (th_base) {
          st = stg_acquire(so);
#pragma unroll
          for (int j = 0; j < 32; ++j) st[j * 32] = in_range ? th[j] : 0.0f;
          stg_flush(so, &omaps[0], c0, false);
        }
        if (sg_base) {
          st = stg_acquire(so);
#pragma unroll
          for (int j = 0; j < 32; ++j) st[j * 32] = in_range ? sg[j] : 0.0f;
          stg_flush(so, &omaps[1], c0, false);
        }
        st = stg_acquire(so);
#pragma unroll
        for (int j = 0; j < 32; ++j) st[j * 32] = in_range ? th[j] * sg[j] : 0.0f;
        stg_flush(so, &omaps[2], c0, false);
      }
      continue;
    }
    const long long o0 = static_cast<long long>(c0) * cs;
    if (th_base) {
      float* o = th_base + o0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (in_range && j < nrem) *o = th[j];
        o += cs;
      }
    }
    if (sg_base) {
      float* o = sg_base + o0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (in_range && j < nrem) *o = sg[j];
        o += cs;
      }
    }
    float* o = z_base + o0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (in_range && j < nrem) *o = th[j] * sg[j];
      o += cs;
    }
  }
}

// ---- GATE_BWD -----------------------------------------------------------------------------------------------------
// g_f = g_z * sg * (1 - th^2), g_g = g_z * th * sg * (1 - sg); acc columns = g_z of all D channels.
struct GateBwdCtx {
  const CUtensorMap* omap;   // non-null: g_f / g_g go through the TMA-store path (map over the whole g_f;g_g tensor)
  int gg_ch_off;
  const float* thp;   // (b, channel 0, tau + add_toff)
  const float* sgp;
  float* gfp;         // (b, channel 0, tau + out_toff)
  float* dupp;        // dup row of g_f or nullptr
  long long add_cs, out_cs, g_delta;
  int n, n_valid;
  bool in_range, live;
  bool ab16;          // AEWN_F_AB16: thp holds {fp16 a, fp16 b} words, the derivative factors themselves
  __nv_bfloat16* g16row;   // optional 16-bit channels-last copy of [g_f; g_g]: row of this lane's time step (or nullptr)
  int g16_goff;            // channel offset of g_gate in that row
  float g16_scale;         // 0: bf16 copy; else fp16(value * g16_scale), a power of two (aewn_ntile.out16_scale)
  bool no_out32;           // AEWN_F_NO_OUT32: the fp32 g_f / g_g stores are skipped
  int* err;                // device error word (fp16 range overflow of the scaled copy)
};

__device__ __forceinline__ GateBwdCtx gbwd_ctx(const aewn_ntile& nt, int b, int tau, const CUtensorMap* omap,
                                               int gg_ch_off) {
  GateBwdCtx cx;
  cx.omap = omap;
  cx.gg_ch_off = gg_ch_off;
  cx.in_range = (tau >= nt.t_lo) && (tau < nt.t_hi);
  cx.live = tau >= nt.t_zero_lo;
  const long long aoff = static_cast<long long>(b) * nt.add_bs + (tau + nt.add_toff);
  cx.thp = nt.add + aoff;
  cx.sgp = nt.add2 + aoff;
  cx.gfp = nt.out + static_cast<long long>(b) * nt.out_bs + (tau + nt.out_toff);
  const int dup_t = tau + nt.dup_toff;
  const bool dup_ok = nt.out3 && cx.in_range && dup_t >= 0 && dup_t < nt.dup_t_hi;
  cx.dupp = dup_ok ? nt.out3 + static_cast<long long>(b) * nt.out_bs + dup_t : nullptr;
  cx.g_delta = nt.out2 - nt.out;  // g_gate rows follow g_filt rows in the same tensor
  cx.add_cs = nt.add_cs;
  cx.out_cs = nt.out_cs;
  cx.n = nt.n;
  cx.n_valid = nt.n_valid;
  cx.ab16 = (nt.flags & AEWN_F_AB16) != 0;
  cx.g16row = (nt.out16 && cx.in_range)
                  ? reinterpret_cast<__nv_bfloat16*>(nt.out16) + static_cast<long long>(b) * nt.out16_bs +
                        static_cast<long long>(tau + nt.out_toff) * nt.out16_cp
                  : nullptr;
  cx.g16_goff = static_cast<int>((nt.out2 - nt.out) / nt.out_cs);
  cx.g16_scale = (nt.out16 && nt.out16_scale) ? __ldg(nt.out16_scale) : 0.0f;
  cx.err = nullptr;
  cx.no_out32 = (nt.flags & AEWN_F_NO_OUT32) != 0 && nt.out16 != nullptr;
  return cx;
}

__device__ __forceinline__ void gbwd_issue(const GateBwdCtx& cx, int c0, float (&th)[32], float (&sg)[32]) {
  const float* pt = cx.thp + static_cast<long long>(c0) * cx.add_cs;
  const float* ps = cx.sgp + static_cast<long long>(c0) * cx.add_cs;
  const int nrem = cx.n_valid - c0;
  const bool ok = cx.in_range && cx.live;
  if (cx.ab16) {     // one word per element: th[] receives a, sg[] receives b
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const unsigned int w = (ok && j < nrem) ? __ldcs(reinterpret_cast<const unsigned int*>(pt)) : 0u;
      const float2 ab = __half22float2(*reinterpret_cast<const __half2*>(&w));
      th[j] = ab.x;
      sg[j] = ab.y;
      pt += cx.add_cs;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    th[j] = (ok && j < nrem) ? __ldcg(pt) : 0.0f;
    sg[j] = (ok && j < nrem) ? __ldcg(ps) : 0.0f;
    pt += cx.add_cs;
    ps += cx.add_cs;
  }
}

__device__ __forceinline__ void gbwd_chunk(const GateBwdCtx& cx, const StgOut& so, uint32_t taddr, int c0,
                                           const float (&th)[32], const float (&sg)[32]) {
  uint32_t v[32];
  tmem_ld32(taddr + c0, v);
  tmem_ld_wait();
  const int nrem = cx.n_valid - c0;
  if (nrem <= 0) return;
  float gf[32], gg[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float gz = cx.live ? __uint_as_float(v[j]) : 0.0f;   // th, sg are 0 when !live
    if (cx.ab16) {
      gf[j] = cx.in_range ? gz * th[j] : 0.0f;
      gg[j] = cx.in_range ? gz * sg[j] : 0.0f;
    } else {
      const float gs = gz * sg[j];
      gf[j] = cx.in_range ? gs * (1.0f - th[j] * th[j]) : 0.0f;
      gg[j] = cx.in_range ? gs * th[j] * (1.0f - sg[j]) : 0.0f;
    }
  }
  if (cx.g16row && nrem >= 32) {
    // bf16 channels-last copy for the 16-bit data-gradient engine (aewn_grcc_dgrad): lane = time row, 32 channels =
    // 64 contiguous bytes per row for g_f and for g_g
    const float gsc = cx.g16_scale;
    float amax = 0.0f;
    auto pack = [&](float lo, float hi) {
      uint32_t r;
      if (gsc != 0.0f) {      // fp16 with a power-of-two scale: TF32-class mantissa, range checked
        lo *= gsc;
        hi *= gsc;
        amax = fmaxf(amax, fmaxf(fabsf(lo), fabsf(hi)));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
      } else {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
      }
      return r;
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 a, g;
      a.x = pack(gf[8 * i + 0], gf[8 * i + 1]); a.y = pack(gf[8 * i + 2], gf[8 * i + 3]);
      a.z = pack(gf[8 * i + 4], gf[8 * i + 5]); a.w = pack(gf[8 * i + 6], gf[8 * i + 7]);
      g.x = pack(gg[8 * i + 0], gg[8 * i + 1]); g.y = pack(gg[8 * i + 2], gg[8 * i + 3]);
      g.z = pack(gg[8 * i + 4], gg[8 * i + 5]); g.w = pack(gg[8 * i + 6], gg[8 * i + 7]);
      *reinterpret_cast<uint4*>(cx.g16row + c0 + 8 * i) = a;
      *reinterpret_cast<uint4*>(cx.g16row + cx.g16_goff + c0 + 8 * i) = g;
    }
    if (amax > 65504.0f && cx.err) atomicExch(cx.err, AEWN_ERR_RANGE);
  }
  if (cx.no_out32) {
    // every consumer (data gradient, weight gradients) reads the 16-bit copy: no fp32 [g_f; g_g] at all
  } else if (cx.omap) {   // TMA-store path (tile-uniform decision): g_f box, then g_g box, through the warp's staging tile
    if (so.slab_on) {
      float* st = stg_acquire(so);
#pragma unroll
      for (int j = 0; j < 32; ++j) st[j * 32] = gf[j];
      stg_flush(so, cx.omap, c0, false);
      st = stg_acquire(so);
#pragma unroll
      for (int j = 0; j < 32; ++j) st[j * 32] = gg[j];
      stg_flush(so, cx.omap, cx.gg_ch_off + c0, false);
    }
  } else {
    float* o = cx.gfp + static_cast<long long>(c0) * cx.out_cs;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (cx.in_range && j < nrem) {
        o[0] = gf[j];
        o[cx.g_delta] = gg[j];
      }
      o += cx.out_cs;
    }
  }
  if (cx.dupp) {   // shifted duplicate for a dilation-1/2 data-gradient tap: plain stores (unaligned time origin)
    float* d = cx.dupp + static_cast<long long>(c0) * cx.out_cs;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < nrem) {
        d[0] = gf[j];
        d[cx.g_delta] = gg[j];
      }
      d += cx.out_cs;
    }
  }
}

// tanh / sigmoid of the first chunk are loaded by the caller before it waits for the accumulator; later chunks load
// theirs right after the TMEM read is issued (a deeper software pipeline costs 64 more registers and spilled).
__device__ __forceinline__ void epi_gate_bwd(const GateBwdCtx& cx, const StgOut& so, uint32_t taddr, int half,
                                             float (&th)[32], float (&sg)[32]) {
  bool first = true;
  for (int c0 = half * 32; c0 < cx.n; c0 += 64) {
    if (!first) gbwd_issue(cx, c0, th, sg);
    first = false;
    gbwd_chunk(cx, so, taddr, c0, th, sg);
  }
}

// ------------------------------------------------------------------------------------------------ kernel
// MODE: epilogue mode shared by every n-tile of the launch (AEWN_EPI_*), or -1 = per-tile dispatch.  The three epilogues
// are fully unrolled; compiled into one kernel they make 230 KB of SASS and the epilogue warps spent 16 % of their
// samples on instruction fetch (stall_no_inst, profiles/r2b_*), so each mode gets its own instantiation.
template <bool PAIR, int MODE>
__global__ void __launch_bounds__(TG_THREADS, 1) tgemm_kernel(const __grid_constant__ TgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  float* stg_base = reinterpret_cast<float*>(
      smem + ((PAIR && p.stg2) ? TG_PAIR_STAGES_STG2 * TG_PAIR_STAGE_BYTES : TG_RING_BYTES));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TG_RING_BYTES + TG_STG_BYTES);
  uint64_t* empty_bar = full_bar + TG_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + TG_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool pair = PAIR;   // compile-time: the two modes are separate instantiations (no spills at 40 registers)
  const bool stg2 = pair && p.stg2 != 0;
  const uint32_t n_stages = pair ? (stg2 ? TG_PAIR_STAGES_STG2 : TG_PAIR_STAGES) : TG_STAGES;
  const uint32_t stage_bytes = pair ? TG_PAIR_STAGE_BYTES : TG_STAGE_BYTES;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int i = 0; i < TG_MAX_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      // multicast mode: one tcgen05.commit arrival from every CTA that reads the stage's W; pair mode: the leader's
      // cta_group::2 commit alone releases the stage in both CTAs
      mbar_init(&empty_bar[i], pair ? 1 : p.cluster);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      // pair mode: the leader's MMA warp may only overwrite an accumulator stage once the epilogue warps of BOTH CTAs
      // have drained it (the peer's warps arrive remotely on the leader's barrier)
      mbar_init(&tempty_bar[i], pair ? 2 * TG_EPI_WARPS : TG_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (pair) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
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
  // (the item loops below rebuild their bounds from the kernel parameters: a value computed HERE, before the register
  // split, is spilled to local memory for all three roles, and a local load in the epilogue's tile loop queues behind the
  // warp's own stores in the LSU)
  const uint16_t cmask = static_cast<uint16_t>((1u << p.cluster) - 1u);

  // Register reallocation: warps 0-3 (TMA / MMA / TMEM-alloc roles, one warpgroup) shrink to 88 registers and the 8
  // epilogue warps grow to 208, so 32-wide column chunks + prefetch buffers stay in registers
  // (128*88 + 256*208 = 64512 <= 65536).  Each setmaxnreg dominates its role code (no merge of limits).
  if (warp < 4) {
  reg_dealloc<40>();   // 128*40 + 256*232 = 64512 = 384*168: the CTA pool is what the launch allocated, NOT the SM file
  // Both single-issuer roles run WARP-CONVERGENT: all 32 lanes walk the loops (so the pipeline state, tile coordinates
  // and descriptors are warp-uniform and live in uniform registers) and one elected lane issues the TMA / MMA /
  // commit instructions.  A loop walked by `if (lane == 0)` alone compiles to vector-register state and a
  // waterfall (ELECT + 7 x R2UR + BRA.U.ANY) around EVERY tcgen05 / TMA instruction: ncu showed the MMA thread spending
  // ~1000 cycles per K block on issue overhead against 512 cycles of tensor work (profiles/r2_*).
  if (warp == 0) {
    // ===================================================== TMA producer
    {
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      const uint32_t lead_full = pair ? mapa_u32(&full_bar[0], 0) : 0u;
      for (int item = blockIdx.x / p.cluster; item < p.batch * p.n_tgroups * p.n_heads && ok; item += gridDim.x / p.cluster) {
        const TgItem it = tg_decode(p, item, crank);
        if (!it.active) continue;
        const aewn_ntile& nt = p.nt[it.ni];
        const int n_mma = nt.n + (p.merged[it.ni] ? p.nt[it.ni + 1].n : 0);   // columns of the (merged) accumulator
        // W rows are split into `cluster` slices of wrows each; CTA r loads slice r and multicasts it to all peers
        const int wrows = 256 / p.cluster;
        const int wslices = (nt.n + wrows - 1) / wrows;
        for (int s = 0; s < p.n_segs && ok; ++s) {
          if (!((nt.seg_mask >> s) & 1)) continue;
          const TgSeg sg = p.seg[s];
          for (int kb = 0; kb < sg.kblocks; ++kb) {
            if (!mbar_wait_warp(&empty_bar[stage], phase ^ 1u, abort_flag)) { ok = false; break; }
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sw = sa + TG_A_BYTES;
            if (pair) {
              // Both CTAs signal the LEADER's full barrier (cta_group::2 loads); the leader expects the bytes of both.
              // CTA r stages W rows [r * n/2, (r+1) * n/2) of the tile (a 128-row box; the MMA reads n/2 of them).
              if (elect_one()) {
                if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * TG_PAIR_STAGE_BYTES);
                const uint32_t fb = lead_full + stage * 8u;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  tma_load_3d_pair(sa + i * 4096, &p.a_map[sg.map], fb, it.tau0 + sg.shift + 32 * i, kb * TG_BK, it.b);
                tma_load_2d_pair(sw, &p.w_map, fb, sg.w_koff + kb * TG_BK, nt.w_row + crank * (n_mma >> 1));
              }
              __syncwarp();
              if (++stage == n_stages) { stage = 0; phase ^= 1u; }
              continue;
            }
            const int wboxes = (nt.n + 127) >> 7;
            if (elect_one()) {
              mbar_expect_tx(&full_bar[stage],
                             TG_A_BYTES + (p.cluster == 1 ? wboxes * TG_WBOX_BYTES : wslices * wrows * 128));
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
            }
            __syncwarp();
            if (++stage == n_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (pair mode: the leader CTA issues for both)
    if (!pair || crank == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ok = true;
      // A: MN-major, 128B swizzle with 32B atoms: 4-row groups 512 B apart (SBO), 32-step chunks 4 KB apart (LBO);
      // W: K-major, 128B swizzle: 8-row groups 1 KB apart (SBO).  The descriptors of a stage differ from these only in
      // the 14-bit start-address field (bytes >> 4), so they are built once and the address is ADDED per MMA.
      const uint64_t adesc0 = make_smem_desc(0, p.a_lbo, p.a_sbo, kLayoutSW128Base32);
      const uint64_t bdesc0 = make_smem_desc(0, 16, 1024, kLayoutSW128);
      const uint32_t ring = smem_u32(smem);
      for (int item = blockIdx.x / p.cluster; item < p.batch * p.n_tgroups * p.n_heads && ok; item += gridDim.x / p.cluster) {
        const TgItem it = tg_decode(p, item, crank);
        if (!it.active) continue;
        const aewn_ntile& nt = p.nt[it.ni];
        if (!mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1u, abort_flag)) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        const int n_mma = nt.n + (p.merged[it.ni] ? p.nt[it.ni + 1].n : 0);
        const uint32_t idesc = make_idesc_tf32(pair ? 2 * TG_BM : TG_BM, n_mma, /*a_mn=*/1, /*b_mn=*/0);
        uint32_t kiter = 0;
        for (int s = 0; s < p.n_segs && ok; ++s) {
          if (!((nt.seg_mask >> s) & 1)) continue;
          const int kblocks = p.seg[s].kblocks;
          for (int kb = 0; kb < kblocks; ++kb) {
            if (!mbar_wait_warp(&full_bar[stage], phase, abort_flag)) { ok = false; break; }
            tc_fence_after();
            if (elect_one()) {
              // start-address fields (16-byte units).  The mask matters: in CTA rank 1 of a cluster the shared::cta
              // window address carries the rank in bit 24, which would otherwise spill into the LBO field.
              const uint32_t a16 = ((ring + stage * stage_bytes) >> 4) & 0x3FFFu;
              const uint32_t w16 = a16 + (TG_A_BYTES >> 4);
#pragma unroll
              for (int ks = 0; ks < TG_BK / 8; ++ks) {
                // A advances 8 K rows = 1 KB per MMA; W advances 8 K columns = 32 B inside the swizzle row
                const uint64_t adesc = adesc0 + (a16 + ks * 64);
                const uint64_t bdesc = bdesc0 + (w16 + ks * 2);
                if (pair) umma_tf32_ss_pair(d_tmem, adesc, bdesc, idesc, (kiter | ks) != 0u);
                else umma_tf32_ss(d_tmem, adesc, bdesc, idesc, (kiter | ks) != 0u);
              }
              if (pair) umma_commit_pair(&empty_bar[stage], 0x3);
              else if (p.cluster == 1) umma_commit(&empty_bar[stage]);
              else umma_commit_mcast(&empty_bar[stage], cmask);
            }
            __syncwarp();
            ++kiter;
            if (++stage == n_stages) { stage = 0; phase ^= 1u; }
          }
        }
        if (!ok) break;
        if (elect_one()) {
          if (pair) umma_commit_pair(&tfull_bar[acc], 0x3);   // accumulator ready: wake the epilogues of both CTAs
          else umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
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
    int stg_cur = 0;
    for (int item = blockIdx.x / p.cluster; item < p.batch * p.n_tgroups * p.n_heads; item += gridDim.x / p.cluster) {
      const TgItem it = tg_decode(p, item, crank);
      if (!it.active) continue;
      const aewn_ntile& nt = p.nt[it.ni];
      const int tau = it.tau0 + q * 32 + lane;
      // Each mode keeps its own register context (exclusive branches, so the contexts can share registers); the first
      // chunks' loads (addend / previous output / tanh+sigmoid) are issued BEFORE waiting for the accumulator.
      const uint32_t taddr = tmem_base + acc * 256u + (static_cast<uint32_t>(q * 32) << 16);
      const int mode = MODE >= 0 ? MODE : nt.mode;
      const bool use_tma = p.o_tma[it.ni] != 0;
      StgOut so;
      so.ntiles = stg2 ? 2 : 1;
      so.cur = stg_cur;     // the alternation continues across items: the last box of the previous item may still be read
      so.tile = stg_base + (warp - 4) * 1024 * so.ntiles;
      so.lane = lane;
      so.b = it.b;
      so.t0 = it.tau0 + q * 32 + nt.out_toff;
      so.slab_on = (it.tau0 + q * 32 + 32 > nt.t_lo) && (it.tau0 + q * 32 < nt.t_hi);
      bool ok;
      if (mode == AEWN_EPI_LINEAR) {
        const LinRegs lc = lin_regs(nt, it.b, tau, use_tma ? &p.o_map[it.ni][0] : nullptr);
        float bufA[32], bufB[32];
        if (lc.srcp && half * 32 < lc.n_valid) lin_issue(lc, half * 32, bufA);
        if (lc.srcp && half * 32 + 64 < lc.n_valid) lin_issue(lc, half * 32 + 64, bufB);
        ok = mbar_wait(&tfull_bar[acc], acc_phase, abort_flag);
        if (ok) {
          tc_fence_after();
          epi_linear(lc, so, taddr, half, bufA, bufB, nt.zero_count);
          if (p.merged[it.ni]) {   // the partner tile sits in the columns after this tile's n
            const aewn_ntile& nt2 = p.nt[it.ni + 1];
            StgOut so2 = so;
            so2.t0 = it.tau0 + q * 32 + nt2.out_toff;
            so2.slab_on = (it.tau0 + q * 32 + 32 > nt2.t_lo) && (it.tau0 + q * 32 < nt2.t_hi);
            epi_linear_tail(nt2, so2, taddr + static_cast<uint32_t>(nt.n), half, tau, &p.o_map[it.ni + 1][0]);
            so.cur = so2.cur;
          }
        }
      } else if (mode == AEWN_EPI_GATE_BWD) {
        GateBwdCtx gcx = gbwd_ctx(nt, it.b, tau, use_tma ? &p.o_map[it.ni][0] : nullptr, p.gg_ch_off[it.ni]);
        gcx.err = p.err;
        float thA[32], sgA[32];
        gbwd_issue(gcx, half * 32, thA, sgA);
        ok = mbar_wait(&tfull_bar[acc], acc_phase, abort_flag);
        if (ok) {
          tc_fence_after();
          epi_gate_bwd(gcx, so, taddr, half, thA, sgA);
        }
      } else {
        ok = mbar_wait(&tfull_bar[acc], acc_phase, abort_flag);
        if (ok) {
          tc_fence_after();
          epi_gate_fwd(nt, so, use_tma ? &p.o_map[it.ni][0] : nullptr, taddr, half, it.b, tau);
        }
      }
      if (!ok) break;
      stg_cur = so.cur;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pair && crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));   // the leader's barrier
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (elect_one()) tma_store_wait_all();   // bulk stores must have completed before the CTA's smem goes away
    __syncwarp();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();   // no CTA may exit while a peer can still multicast into its smem
  if (threadIdx.x == 0 && *abort_flag && p.err) atomicExch(p.err, AEWN_ERR_TIMEOUT);
  if (warp == 2) {
    tc_fence_after();
    if (pair) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
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
    // (the attribute is set on the selected instantiation right before the launch)
  }

  TgParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < d->n_acts; ++i) {
    int rc = encode_act_map(&p.a_map[i], d->acts[i], TG_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (d->acts[i].batch < d->batch) return set_err(AEWN_ERR_INVALID, "tgemm: act %d batch smaller than problem batch", i);
  }
  for (int i = d->n_acts; i < AEWN_MAX_ACTS; ++i) p.a_map[i] = p.a_map[0];
  int cluster = d->cluster > 0 ? d->cluster : AEWN_TG_DEFAULT_CLUSTER;
  if (cluster == AEWN_CLUSTER_PAIR_MMA) {
    p.pair = 1;
    cluster = 2;
  }
  {
    // Short K loops (res+skip: 8 K blocks, gate derivative: 20) are bound by the epilogue's store chain: give them the
    // second staging tile.  Long ones (conv+gate: 29, data gradient: 32) need the sixth ring stage to cover HBM
    // latency (measured: 4.36 vs 4.72 ms / 20 layers for conv+gate, 3.96 vs 3.79 for res+skip).  AEWN_TG_STG2=0|1 forces.
    int kb_total = 0;
    for (int s2 = 0; s2 < d->n_segs; ++s2) kb_total += (d->segs[s2].channels + TG_BK - 1) / TG_BK;
    static const int forced = []() { const char* e = getenv("AEWN_TG_STG2"); return e ? atoi(e) : -1; }();
    p.stg2 = p.pair && (forced >= 0 ? forced : (kb_total <= 24));
  }
  if (cluster != 1 && cluster != 2 && cluster != 4)
    return set_err(AEWN_ERR_INVALID, "tgemm: cluster must be 1, 2, 4 or AEWN_CLUSTER_PAIR_MMA");
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
      if (nt.add && (nt.flags & AEWN_F_ACCUM))
        return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d: an addend and AEWN_F_ACCUM are mutually exclusive", i);
    } else if (nt.mode == AEWN_EPI_GATE_FWD) {
      if (nt.n != 256 || !nt.out3 || nt.n_valid > 128)
        return set_err(AEWN_ERR_INVALID, "tgemm: GATE_FWD tile %d needs n=256, n_valid<=128 and out3", i);
    } else if (nt.mode == AEWN_EPI_GATE_BWD) {
      if (!nt.out || !nt.out2 || !nt.add || (!nt.add2 && !(nt.flags & AEWN_F_AB16)))
        return set_err(AEWN_ERR_INVALID, "tgemm: GATE_BWD tile %d needs out, out2, add, add2", i);
      if (nt.out16 && ((nt.n_valid & 31) || (nt.out16_cp & 7) || (nt.out16_bs & 7) || (reinterpret_cast<uintptr_t>(nt.out16) & 15u)))
        return set_err(AEWN_ERR_INVALID, "tgemm: GATE_BWD tile %d: out16 needs n_valid %% 32 == 0 and 16-byte aligned rows", i);
    } else {
      return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d unknown mode %d", i, nt.mode);
    }
    if (nt.w_row < 0 || nt.w_row >= d->w_rows) return set_err(AEWN_ERR_INVALID, "tgemm: n-tile %d w_row out of range", i);
    if (nt.t_lo < d->t_begin) nt.t_lo = d->t_begin;
    if (nt.t_hi > d->t_end) nt.t_hi = d->t_end;
    p.nt[i] = nt;

    // ---- TMA-store eligibility (otherwise the tile keeps the st.global epilogue)
    auto aligned = [&](const float* q) { return q && (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool geo_ok = !d->no_tma_store && (nt.out_toff & 3) == 0 && (nt.out_cs & 3) == 0 && (nt.out_bs & 3) == 0 &&
                        nt.t_hi + nt.out_toff > 0;
    const int t_ext = nt.t_hi + nt.out_toff;
    p.o_tma[i] = 0;
    p.gg_ch_off[i] = 0;
    if (geo_ok && nt.mode == AEWN_EPI_LINEAR && aligned(nt.out) &&
        !((nt.flags & AEWN_F_ACCUM) && (nt.flags & AEWN_F_RELU))) {
      int rc2 = encode_out_map(&p.o_map[i][0], nt.out, t_ext, nt.n_valid, d->batch, nt.out_cs, nt.out_bs);
      if (rc2) return rc2;
      p.o_tma[i] = 1;
    } else if (geo_ok && nt.mode == AEWN_EPI_GATE_FWD && aligned(nt.out3) && (!nt.out || aligned(nt.out)) &&
               (!nt.out2 || aligned(nt.out2))) {
      float* ptrs[3] = {nt.out, nt.out2, nt.out3};
      for (int k = 0; k < 3; ++k) {
        if (!ptrs[k]) continue;
        int rc2 = encode_out_map(&p.o_map[i][k], ptrs[k], t_ext, nt.n_valid, d->batch, nt.out_cs, nt.out_bs);
        if (rc2) return rc2;
      }
      p.o_tma[i] = 1;
    } else if (geo_ok && nt.mode == AEWN_EPI_GATE_BWD && aligned(nt.out) && (nt.n_valid & 31) == 0 &&
               nt.out2 > nt.out && ((nt.out2 - nt.out) % nt.out_cs) == 0) {
      const int off = static_cast<int>((nt.out2 - nt.out) / nt.out_cs);
      int rc2 = encode_out_map(&p.o_map[i][0], nt.out, t_ext, off + nt.n_valid, d->batch, nt.out_cs, nt.out_bs);
      if (rc2) return rc2;
      p.o_tma[i] = 1;
      p.gg_ch_off[i] = off;
    }
  }
  p.n_ntiles = d->n_ntiles;
  p.n_heads = 0;
  for (int i = 0; i < d->n_ntiles; ++i) {
    p.head[p.n_heads++] = i;
    if (!(p.nt[i].flags & AEWN_F_MERGE_NEXT)) continue;
    if (i + 1 >= d->n_ntiles) return set_err(AEWN_ERR_INVALID, "tgemm: AEWN_F_MERGE_NEXT on the last n-tile");
    const aewn_ntile &a = p.nt[i], &b = p.nt[i + 1];
    const bool ok = p.pair && a.mode == AEWN_EPI_LINEAR && b.mode == AEWN_EPI_LINEAR && a.n == a.n_valid &&
                    b.w_row == a.w_row + a.n && a.n + b.n <= 256 && b.n >= 32 && !b.add && !b.bias && !b.out2 && !b.out3 &&
                    !b.zero_count && (b.flags & ~AEWN_F_ACCUM) == 0 && p.o_tma[i + 1] &&
                    ((b.seg_mask ^ a.seg_mask) & all_mask & b.seg_mask) == 0;
    if (!ok)
      return set_err(AEWN_ERR_INVALID,
                     "tgemm: n-tiles %d/%d cannot share an accumulator (need pair mode, LINEAR, contiguous W rows, "
                     "n <= 256 in total, a plain TMA-stored partner whose segments are a subset of the head's)", i, i + 1);
    p.merged[i] = 1;
    ++i;   // the partner is not a work item of its own
  }
  p.batch = d->batch;
  p.t_begin = d->t_begin;
  p.n_ttiles = (d->t_end - d->t_begin + TG_BM - 1) / TG_BM;
  p.err = d->err;
  p.a_lbo = d->dbg_lbo > 0 ? d->dbg_lbo : 4096;
  p.a_sbo = d->dbg_sbo > 0 ? d->dbg_sbo : 512;

  p.cluster = cluster;
  p.n_tgroups = (p.n_ttiles + cluster - 1) / cluster;
  const long long total = static_cast<long long>(p.batch) * p.n_tgroups * p.n_heads;   // items per cluster walk
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
  int mode = d->ntiles[0].mode;
  for (int i = 1; i < d->n_ntiles; ++i)
    if (d->ntiles[i].mode != mode) mode = -1;
  using KernelFn = void (*)(TgParams);
  KernelFn fn;
  if (p.pair) {
    fn = mode == AEWN_EPI_LINEAR ? tgemm_kernel<true, AEWN_EPI_LINEAR>
         : mode == AEWN_EPI_GATE_FWD ? tgemm_kernel<true, AEWN_EPI_GATE_FWD>
         : mode == AEWN_EPI_GATE_BWD ? tgemm_kernel<true, AEWN_EPI_GATE_BWD> : tgemm_kernel<true, -1>;
  } else {
    fn = mode == AEWN_EPI_LINEAR ? tgemm_kernel<false, AEWN_EPI_LINEAR>
         : mode == AEWN_EPI_GATE_FWD ? tgemm_kernel<false, AEWN_EPI_GATE_FWD>
         : mode == AEWN_EPI_GATE_BWD ? tgemm_kernel<false, AEWN_EPI_GATE_BWD> : tgemm_kernel<false, -1>;
  }
  cudaError_t ae = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES);
  if (ae != cudaSuccess) return cuda_err(ae, "tgemm: cudaFuncSetAttribute");
  cudaError_t le = cudaLaunchKernelEx(&cfg, fn, p);
  count_launch();
  if (le != cudaSuccess) return cuda_err(le, "tgemm launch");
  return cuda_err(cudaGetLastError(), "tgemm launch");
}
