// wgradw: weight-gradient GEMM with WIDE units on CTA pairs (tcgen05 cta_group::2, TF32 in, FP32 accumulate in TMEM).
//
//   out_c[m * rs_c + n * cs_c] += sum_b sum_u G[b, g_row + m, u] * X_c[b, x_row_c + n, u + shift_c]      c = chunk
//
// Same contraction as wgrad.cu (the dW lines of wavenet.py:25-34), different tiling.  ncu of wgrad.cu
// showed the weight gradients bound by operand delivery (L2 -> shared memory), not by the tensor pipe: a 128 x 256
// accumulator re-reads the G tile once per 256 columns of X.  Here one unit is M = 256 rows of G (128 per CTA of the
// pair) against up to THREE column chunks of X (<= 512 columns in total = the whole TMEM of each CTA), so G is staged
// once per ~512 output columns and every X row is staged by exactly one CTA of the pair (half-boxes of 64 or 128
// rows).  Per K block of 32 time steps a CTA receives <= 48 KB for 128 x 512 x 32 MACs (42.7 MAC/B, was 32 and less
// for the narrow tail chunks).  The price: one accumulator stage, i.e. the epilogue (fp32 red.global.add of the
// partial tile) does not overlap the next unit's MMAs -- split-K units are >= 48 K blocks long, so that is a few %.
//
// Warp roles as in tgemm.cu / wgrad.cu; single-issuer roles run warp-convergent with an elected issuing lane.
#include "host_util.h"
#include "ptx.cuh"

namespace aewn {

constexpr int WW_BK = 32;
constexpr int WW_STAGES = 4;
constexpr int WW_BOX_BYTES = 128 * WW_BK * 4;         // 16 KB: 128 rows x 32 time steps
constexpr int WW_STAGE_BYTES = 3 * WW_BOX_BYTES;      // G box + 32 KB of X half-boxes (<= 256 rows per CTA)
constexpr int WW_THREADS = 384;
constexpr int WW_EPI_WARPS = 8;
constexpr int WW_XPOSE_BYTES = WW_EPI_WARPS * 4096;   // one 32 x 32 fp32 transpose tile per epilogue warp
constexpr int WW_SMEM_BYTES = WW_STAGES * WW_STAGE_BYTES + WW_XPOSE_BYTES + 256 + 1024;

struct WwParams {
  CUtensorMap map128[AEWN_WGRAD_MAX_ACTS];   // box {32 t, 128 rows}
  CUtensorMap map64[AEWN_WGRAD_MAX_ACTS];    // box {32 t, 64 rows}
  aewn_wgw_unit units[AEWN_WGW_MAX_UNITS];
  int x_off[AEWN_WGW_MAX_UNITS][AEWN_WGW_MAX_CHUNKS];    // byte offset of the chunk's half-box inside the X region
  int tm_col[AEWN_WGW_MAX_UNITS][AEWN_WGW_MAX_CHUNKS];   // first TMEM column of the chunk
  int tx_bytes[AEWN_WGW_MAX_UNITS];                      // bytes both CTAs deliver per K block
  int work_begin[AEWN_WGW_MAX_UNITS + 1];                // prefix sum of n_split: work item -> (unit, split)
  int n_units;
  int batch;
  int* err;
};

struct WwWork {
  int unit;
  int kb_begin, kb_end;   // flattened (batch, time-block) range
  int blocks_per_b;
};

__device__ __forceinline__ WwWork ww_decode(const WwParams& p, int work) {
  WwWork w;
  int u = 0;
  while (u + 1 < p.n_units && work >= p.work_begin[u + 1]) ++u;
  const int split = work - p.work_begin[u];
  const aewn_wgw_unit& un = p.units[u];
  w.unit = u;
  w.blocks_per_b = (un.t_hi - un.t_lo + WW_BK - 1) / WW_BK;
  const long long total = static_cast<long long>(w.blocks_per_b) * p.batch;
  w.kb_begin = static_cast<int>(total * split / un.n_split);
  w.kb_end = static_cast<int>(total * (split + 1) / un.n_split);
  return w;
}

__global__ void __launch_bounds__(WW_THREADS, 1) wgradw_kernel(const __grid_constant__ WwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  float* xpose = reinterpret_cast<float*>(smem + WW_STAGES * WW_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WW_STAGES * WW_STAGE_BYTES + WW_XPOSE_BYTES);
  uint64_t* empty_bar = full_bar + WW_STAGES;
  uint64_t* tfull_bar = empty_bar + WW_STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int i = 0; i < WW_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);    // leader: its own arrive.expect_tx; the bytes of BOTH CTAs complete on it
      mbar_init(&empty_bar[i], 1);   // the leader's cta_group::2 commit releases the stage in both CTAs
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * WW_EPI_WARPS);   // on the leader: the epilogue warps of both CTAs
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int crank = static_cast<int>(cluster_ctarank());
  const int cid = blockIdx.x >> 1;
  const int n_cl = gridDim.x >> 1;
  const int total_work = p.work_begin[p.n_units];

  if (warp < 4) {
    reg_dealloc<88>();
    if (warp == 0) {
      // ===================================================== TMA producer (both CTAs)
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      const uint32_t lead_full = mapa_u32(&full_bar[0], 0);
      for (int work = cid; work < total_work && ok; work += n_cl) {
        const WwWork w = ww_decode(p, work);
        if (w.kb_end <= w.kb_begin) continue;
        const aewn_wgw_unit& un = p.units[w.unit];
        int b_next = w.kb_begin / w.blocks_per_b;
        int tb = w.kb_begin - b_next * w.blocks_per_b;
        for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
          const int b = b_next;
          const int t = un.t_lo + tb * WW_BK;
          if (++tb == w.blocks_per_b) { tb = 0; ++b_next; }
          if (!mbar_wait_warp(&empty_bar[stage], phase ^ 1u, abort_flag)) { ok = false; break; }
          if (elect_one()) {
            uint8_t* sg = smem + stage * WW_STAGE_BYTES;
            uint8_t* sx = sg + WW_BOX_BYTES;
            if (crank == 0) mbar_expect_tx(&full_bar[stage], p.tx_bytes[w.unit]);
            const uint32_t fb = lead_full + stage * 8u;
            tma_load_3d_pair(sg, &p.map128[un.g_act], fb, t, un.g_row + crank * 128, b);
            for (int c = 0; c < un.n_chunks; ++c) {
              const aewn_wgw_chunk& ch = un.chunk[c];
              // CTA r stages X rows [r * n/2, (r + 1) * n/2) of the chunk: a 64- or 128-row box, the MMA reads n/2
              const CUtensorMap* m = ch.n > 128 ? &p.map128[ch.x_act] : &p.map64[ch.x_act];
              tma_load_3d_pair(sx + p.x_off[w.unit][c], m, fb, t + ch.shift, ch.x_row + crank * (ch.n >> 1), b);
            }
          }
          __syncwarp();
          if (++stage == WW_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // ===================================================== MMA issuer (the pair's leader CTA)
      if (crank == 0) {
        uint32_t stage = 0, phase = 0, acc_phase = 0;
        bool ok = true;
        const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSW128);   // K-major, 128B swizzle, 8-row groups 1 KB apart
        const uint32_t ring = smem_u32(smem);
        for (int work = cid; work < total_work && ok; work += n_cl) {
          const WwWork w = ww_decode(p, work);
          if (w.kb_end <= w.kb_begin) continue;
          const aewn_wgw_unit& un = p.units[w.unit];
          if (!mbar_wait_warp(tempty_bar, acc_phase ^ 1u, abort_flag)) break;
          tc_fence_after();
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            if (!mbar_wait_warp(&full_bar[stage], phase, abort_flag)) { ok = false; break; }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t g16 = ((ring + stage * WW_STAGE_BYTES) >> 4) & 0x3FFFu;
              const uint32_t x16 = g16 + (WW_BOX_BYTES >> 4);
              const uint32_t accum = kb > w.kb_begin;
              for (int c = 0; c < un.n_chunks; ++c) {
                const uint32_t idesc = make_idesc_tf32(256, un.chunk[c].n, 0, 0);
                const uint32_t xc16 = x16 + (static_cast<uint32_t>(p.x_off[w.unit][c]) >> 4);
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(p.tm_col[w.unit][c]);
#pragma unroll
                for (int ks = 0; ks < WW_BK / 8; ++ks)   // K advances 32 B inside the swizzle row
                  umma_tf32_ss_pair(d_tmem, desc0 + (g16 + ks * 2), desc0 + (xc16 + ks * 2), idesc, accum | (ks > 0));
              }
              umma_commit_pair(&empty_bar[stage], 0x3);
            }
            __syncwarp();
            if (++stage == WW_STAGES) { stage = 0; phase ^= 1u; }
          }
          if (!ok) break;
          if (elect_one()) umma_commit_pair(tfull_bar, 0x3);
          __syncwarp();
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    reg_alloc<208>();
    // ===================================================== epilogue (both CTAs): partial tile -> red.global.add
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t acc_phase = 0;
    for (int work = cid; work < total_work; work += n_cl) {
      const WwWork w = ww_decode(p, work);
      if (w.kb_end <= w.kb_begin) continue;
      const aewn_wgw_unit& un = p.units[w.unit];
      if (!mbar_wait(tfull_bar, acc_phase, abort_flag)) break;
      tc_fence_after();
      const int m0 = crank * 128 + q * 32;          // first output row of this warp's 32 TMEM lanes
      const int m = m0 + lane;
      float* tile = xpose + (warp - 4) * 1024;
      for (int c = 0; c < un.n_chunks; ++c) {
        const aewn_wgw_chunk& ch = un.chunk[c];
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(p.tm_col[w.unit][c]) + (static_cast<uint32_t>(q * 32) << 16);
        // Lane = output row m.  When the rows are contiguous in memory (out_rs == 1: the transposed outputs) a warp-wide
        // red already hits one 128-byte line.  Otherwise (conv weights: rows 2R floats apart) every lane would touch
        // its own line -- 8 M single-lane reds per launch were a third of all L2 tag requests in ncu -- so the 32 x 32
        // block goes through a swizzled shared-memory tile and is reduced with lane = column instead.
        const bool direct = ch.out_rs == 1;
        float* orow = ch.out + static_cast<long long>(m) * ch.out_rs;
        for (int c0 = half * 32; c0 < ch.n; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (direct) {
            if (m < un.m_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < ch.n_valid) atomicAdd(orow + static_cast<long long>(c0 + j) * ch.out_cs, __uint_as_float(v[j]));
            }
          } else {
            __syncwarp();                            // the previous block has been read out of the tile
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 32 + (j ^ lane)] = __uint_as_float(v[j]);   // (row lane, col j)
            __syncwarp();
            if (c0 + lane < ch.n_valid) {
              float* ocol = ch.out + static_cast<long long>(m0) * ch.out_rs + static_cast<long long>(c0 + lane) * ch.out_cs;
              const int rows = un.m_valid - m0 < 32 ? un.m_valid - m0 : 32;
#pragma unroll 8
              for (int r = 0; r < rows; ++r)         // (row r, col lane) sits at physical column lane ^ r
                atomicAdd(ocol + static_cast<long long>(r) * ch.out_rs, tile[r * 32 + (lane ^ r)]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (crank != 0) mbar_arrive_cluster(mapa_u32(tempty_bar, 0));   // the leader's barrier
        else mbar_arrive(tempty_bar);
      }
      acc_phase ^= 1u;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA may exit while its peer can still signal it / read its shared memory
  if (threadIdx.x == 0 && *abort_flag && p.err) atomicExch(p.err, AEWN_ERR_TIMEOUT);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace aewn

using namespace aewn;

extern "C" int aewn_wgradw(const aewn_wgradw_desc* d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_err(AEWN_ERR_INVALID, "wgradw: null descriptor");
  if (d->n_acts < 1 || d->n_acts > AEWN_WGRAD_MAX_ACTS || d->n_units < 1 || d->n_units > AEWN_WGW_MAX_UNITS ||
      d->batch <= 0)
    return set_err(AEWN_ERR_INVALID, "wgradw: n_acts/n_units/batch out of range (%d/%d/%d)", d->n_acts, d->n_units,
                   d->batch);
  cudaError_t e = cudaFuncSetAttribute(wgradw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WW_SMEM_BYTES);
  if (e != cudaSuccess) return cuda_err(e, "wgradw: cudaFuncSetAttribute");

  static_assert(sizeof(WwParams) <= 4000, "WwParams must fit the kernel parameter space");
  WwParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < d->n_acts; ++i) {
    int rc = encode_act_map(&p.map128[i], d->acts[i], 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_act_map(&p.map64[i], d->acts[i], 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (d->acts[i].batch < d->batch) return set_err(AEWN_ERR_INVALID, "wgradw: act %d batch smaller than problem batch", i);
  }
  int work = 0;
  for (int u = 0; u < d->n_units; ++u) {
    const aewn_wgw_unit& un = d->units[u];
    if (un.g_act < 0 || un.g_act >= d->n_acts || un.m_valid < 1 || un.m_valid > 256 || un.n_chunks < 1 ||
        un.n_chunks > AEWN_WGW_MAX_CHUNKS || un.n_split < 1)
      return set_err(AEWN_ERR_INVALID, "wgradw: unit %d invalid (g_act=%d m_valid=%d n_chunks=%d n_split=%d)", u, un.g_act,
                     un.m_valid, un.n_chunks, un.n_split);
    if (un.t_hi <= un.t_lo || (un.t_lo & 3))
      return set_err(AEWN_ERR_INVALID, "wgradw: unit %d needs t_hi > t_lo and t_lo a multiple of 4 (TMA 16-byte origin rule)",
                     u);
    int cols = 0, rows = 0;
    for (int c = 0; c < un.n_chunks; ++c) {
      const aewn_wgw_chunk& ch = un.chunk[c];
      if (ch.x_act < 0 || ch.x_act >= d->n_acts || ch.n < 16 || ch.n > 256 || (ch.n & 15) || ch.n_valid < 1 ||
          ch.n_valid > ch.n || !ch.out)
        return set_err(AEWN_ERR_INVALID, "wgradw: unit %d chunk %d invalid (n=%d n_valid=%d)", u, c, ch.n, ch.n_valid);
      if (ch.shift & 3)
        return set_err(AEWN_ERR_INVALID, "wgradw: unit %d chunk %d shift %d is not a multiple of 4 (TMA 16-byte origin rule)",
                       u, c, ch.shift);
      p.tm_col[u][c] = cols;
      p.x_off[u][c] = rows * 128;
      cols += (ch.n + 31) & ~31;            // the epilogue reads 32-column groups: keep every chunk 32-aligned
      rows += ch.n > 128 ? 128 : 64;
    }
    if (cols > 512 || rows > 256)
      return set_err(AEWN_ERR_INVALID, "wgradw: unit %d needs %d TMEM columns / %d staged rows per CTA (limits 512 / 256)", u,
                     cols, rows);
    p.tx_bytes[u] = 2 * (WW_BOX_BYTES + rows * 128);
    p.units[u] = un;
    p.work_begin[u] = work;
    work += un.n_split;
  }
  p.work_begin[d->n_units] = work;
  p.n_units = d->n_units;
  p.batch = d->batch;
  p.err = d->err;

  int ctas = (d->max_ctas > 0 ? d->max_ctas : sm_count()) & ~1;
  if (ctas > 2 * work) ctas = 2 * work;
  if (ctas < 2) ctas = 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(WW_THREADS);
  cfg.dynamicSmemBytes = WW_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, wgradw_kernel, p);
  count_launch();
  if (le != cudaSuccess) return cuda_err(le, "wgradw launch");
  return cuda_err(cudaGetLastError(), "wgradw launch");
}
