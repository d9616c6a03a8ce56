// wgradh: the wide-unit weight-gradient GEMM of wgradw.cu on 16-bit CHANNELS-LAST operands (tcgen05 cta_group::2,
// kind::f16, FP32 accumulate in TMEM).
//
//   out_c[m * rs_c + n * cs_c] += inv_scale * sum_b sum_u G[b, u, g_row + m] * X_c[b, u + shift_c, x_row_c + n]
//
// The contraction runs over TIME, and the fp16 copies the fused layer kernels keep (x16, cond16, the scaled [g_f; g_g]
// copy of the gate-derivative launch) are channels-last: a box of {64 channels, 64 time steps} is a
// tile whose rows are K and whose 128-byte lines run along M / N -- an MN-MAJOR operand, which kind::f16 accepts for both
// A and B (canonical layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) under Swizzle<3,4,3>: LBO = the next 64-channel box,
// SBO = 1024 bytes = the next 8 time steps).  A tap shifted by the dilation is a shift of the box's ROW coordinate: no
// 16-byte origin rule, no pre-shifted duplicates.  Per K block of 64 time steps a CTA receives the same <= 48 KB as
// wgradw.cu does for 32, at twice the MMA rate: half the operand bytes per MAC.
// Unit / chunk description, split-K scheduling, warp roles and the red.global.add epilogue are those of wgradw.cu.
#include "host_util.h"
#include "ptx.cuh"

namespace aewn {

constexpr int WH_BK = 64;
constexpr int WH_STAGES = 4;
constexpr int WH_BOX_BYTES = 64 * 128;             // 8 KB: 64 time steps x 64 channels
constexpr int WH_G_BYTES = 2 * WH_BOX_BYTES;       // this CTA's 128 rows of G
constexpr int WH_X_BOXES = 4;                      // <= 256 X channels per CTA
constexpr int WH_STAGE_BYTES = WH_G_BYTES + WH_X_BOXES * WH_BOX_BYTES;   // 48 KB
constexpr int WH_THREADS = 384;
constexpr int WH_EPI_WARPS = 8;
constexpr int WH_XPOSE_BYTES = WH_EPI_WARPS * 4096;
constexpr int WH_SMEM_BYTES = WH_STAGES * WH_STAGE_BYTES + WH_XPOSE_BYTES + 256 + 1024;

struct WhParams {
  CUtensorMap map[AEWN_WGRAD_MAX_ACTS];                  // box {64 channels, 64 t, 1}
  aewn_wgw_unit units[AEWN_WGW_MAX_UNITS];
  int x_box[AEWN_WGW_MAX_UNITS][AEWN_WGW_MAX_CHUNKS];    // first X box of the chunk inside the stage
  int x_nbox[AEWN_WGW_MAX_UNITS][AEWN_WGW_MAX_CHUNKS];   // boxes this CTA stages for the chunk (n/2 channels)
  int tm_col[AEWN_WGW_MAX_UNITS][AEWN_WGW_MAX_CHUNKS];
  int tx_bytes[AEWN_WGW_MAX_UNITS];
  int work_begin[AEWN_WGW_MAX_UNITS + 1];
  int n_units;
  int batch;
  const float* inv_scale;
  int* err;
};

struct WhWork {
  int unit;
  int kb_begin, kb_end;
  int blocks_per_b;
};

__device__ __forceinline__ WhWork wh_decode(const WhParams& p, int work) {
  WhWork w;
  int u = 0;
  while (u + 1 < p.n_units && work >= p.work_begin[u + 1]) ++u;
  const int split = work - p.work_begin[u];
  const aewn_wgw_unit& un = p.units[u];
  w.unit = u;
  w.blocks_per_b = (un.t_hi - un.t_lo + WH_BK - 1) / WH_BK;
  const long long total = static_cast<long long>(w.blocks_per_b) * p.batch;
  w.kb_begin = static_cast<int>(total * split / un.n_split);
  w.kb_end = static_cast<int>(total * (split + 1) / un.n_split);
  return w;
}

__device__ __forceinline__ void umma_h_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::f16, A = B = F16, BOTH MN-major (bits 15 / 16), FP32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int M, int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__global__ void __launch_bounds__(WH_THREADS, 1) wgradh_kernel(const __grid_constant__ WhParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  float* xpose = reinterpret_cast<float*>(smem + WH_STAGES * WH_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WH_STAGES * WH_STAGE_BYTES + WH_XPOSE_BYTES);
  uint64_t* empty_bar = full_bar + WH_STAGES;
  uint64_t* tfull_bar = empty_bar + WH_STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int i = 0; i < WH_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * WH_EPI_WARPS);
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

  // (loop bounds are rebuilt per role from the kernel parameters: a value computed here would be spilled for all roles)
  if (warp < 4) {
    reg_dealloc<88>();
    if (warp == 0) {
      // ===================================================== TMA producer (both CTAs)
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      const uint32_t lead_full = mapa_u32(&full_bar[0], 0);
      for (int work = blockIdx.x >> 1; work < p.work_begin[p.n_units] && ok; work += gridDim.x >> 1) {
        const WhWork w = wh_decode(p, work);
        if (w.kb_end <= w.kb_begin) continue;
        const aewn_wgw_unit& un = p.units[w.unit];
        int b_next = w.kb_begin / w.blocks_per_b;
        int tb = w.kb_begin - b_next * w.blocks_per_b;
        for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
          const int b = b_next;
          const int t = un.t_lo + tb * WH_BK;
          if (++tb == w.blocks_per_b) { tb = 0; ++b_next; }
          if (!mbar_wait_warp(&empty_bar[stage], phase ^ 1u, abort_flag)) { ok = false; break; }
          if (elect_one()) {
            uint8_t* sg = smem + stage * WH_STAGE_BYTES;
            uint8_t* sx = sg + WH_G_BYTES;
            if (crank == 0) mbar_expect_tx(&full_bar[stage], p.tx_bytes[w.unit]);
            const uint32_t fb = lead_full + stage * 8u;
            // rows of the last K block beyond t_hi: the callers' tensors are zero there (or beyond the map: zero fill)
            tma_load_3d_pair(sg, &p.map[un.g_act], fb, un.g_row + crank * 128, t, b);
            tma_load_3d_pair(sg + WH_BOX_BYTES, &p.map[un.g_act], fb, un.g_row + crank * 128 + 64, t, b);
            for (int c = 0; c < un.n_chunks; ++c) {
              const aewn_wgw_chunk& ch = un.chunk[c];
              // CTA r stages X channels [r * n/2, (r + 1) * n/2) of the chunk in one or two 64-channel boxes
              for (int j = 0; j < p.x_nbox[w.unit][c]; ++j)
                tma_load_3d_pair(sx + (p.x_box[w.unit][c] + j) * WH_BOX_BYTES, &p.map[ch.x_act], fb,
                                 ch.x_row + crank * (ch.n >> 1) + 64 * j, t + ch.shift, b);
            }
          }
          __syncwarp();
          if (++stage == WH_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // ===================================================== MMA issuer (the pair's leader CTA)
      if (crank == 0) {
        uint32_t stage = 0, phase = 0, acc_phase = 0;
        bool ok = true;
        // MN-major, 128B swizzle: 64-channel boxes 8 KB apart (LBO), 8-time-step groups 1 KB apart (SBO)
        const uint64_t desc0 = make_smem_desc(0, WH_BOX_BYTES, 1024, kLayoutSW128);
        const uint32_t ring = smem_u32(smem);
        for (int work = blockIdx.x >> 1; work < p.work_begin[p.n_units] && ok; work += gridDim.x >> 1) {
          const WhWork w = wh_decode(p, work);
          if (w.kb_end <= w.kb_begin) continue;
          const aewn_wgw_unit& un = p.units[w.unit];
          if (!mbar_wait_warp(tempty_bar, acc_phase ^ 1u, abort_flag)) break;
          tc_fence_after();
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            if (!mbar_wait_warp(&full_bar[stage], phase, abort_flag)) { ok = false; break; }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t g16 = ((ring + stage * WH_STAGE_BYTES) >> 4) & 0x3FFFu;
              const uint32_t x16 = g16 + (WH_G_BYTES >> 4);
              const uint32_t accum = kb > w.kb_begin;
              for (int c = 0; c < un.n_chunks; ++c) {
                const uint32_t idesc = make_idesc_f16_mn(256, un.chunk[c].n);
                const uint32_t xc16 = x16 + (static_cast<uint32_t>(p.x_box[w.unit][c] * WH_BOX_BYTES) >> 4);
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(p.tm_col[w.unit][c]);
#pragma unroll
                for (int ks = 0; ks < WH_BK / 16; ++ks)   // K advances 16 time steps = 2 KB
                  umma_h_ss_pair(d_tmem, desc0 + (g16 + ks * 128), desc0 + (xc16 + ks * 128), idesc, accum | (ks > 0));
              }
              umma_commit_pair(&empty_bar[stage], 0x3);
            }
            __syncwarp();
            if (++stage == WH_STAGES) { stage = 0; phase ^= 1u; }
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
    const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.0f;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x >> 1; work < p.work_begin[p.n_units]; work += gridDim.x >> 1) {
      const WhWork w = wh_decode(p, work);
      if (w.kb_end <= w.kb_begin) continue;
      const aewn_wgw_unit& un = p.units[w.unit];
      if (!mbar_wait(tfull_bar, acc_phase, abort_flag)) break;
      tc_fence_after();
      const int m0 = crank * 128 + q * 32;
      const int m = m0 + lane;
      float* tile = xpose + (warp - 4) * 1024;
      for (int c = 0; c < un.n_chunks; ++c) {
        const aewn_wgw_chunk& ch = un.chunk[c];
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(p.tm_col[w.unit][c]) + (static_cast<uint32_t>(q * 32) << 16);
        const bool direct = ch.out_rs == 1;          // see wgradw.cu: rows contiguous -> lane = row; else transpose
        float* orow = ch.out + static_cast<long long>(m) * ch.out_rs;
        for (int c0 = half * 32; c0 < ch.n; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (direct) {
            if (m < un.m_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < ch.n_valid)
                  atomicAdd(orow + static_cast<long long>(c0 + j) * ch.out_cs, __uint_as_float(v[j]) * inv);
            }
          } else {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 32 + (j ^ lane)] = __uint_as_float(v[j]) * inv;
            __syncwarp();
            if (c0 + lane < ch.n_valid) {
              float* ocol = ch.out + static_cast<long long>(m0) * ch.out_rs + static_cast<long long>(c0 + lane) * ch.out_cs;
              const int rows = un.m_valid - m0 < 32 ? un.m_valid - m0 : 32;
#pragma unroll 8
              for (int r = 0; r < rows; ++r) atomicAdd(ocol + static_cast<long long>(r) * ch.out_rs, tile[r * 32 + (lane ^ r)]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (crank != 0) mbar_arrive_cluster(mapa_u32(tempty_bar, 0));
        else mbar_arrive(tempty_bar);
      }
      acc_phase ^= 1u;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0 && *abort_flag && p.err) atomicExch(p.err, AEWN_ERR_TIMEOUT);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace aewn

using namespace aewn;

extern "C" int aewn_wgradh(const aewn_wgradh_desc* d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_err(AEWN_ERR_INVALID, "wgradh: null descriptor");
  if (d->n_acts < 1 || d->n_acts > AEWN_WGRAD_MAX_ACTS || d->n_units < 1 || d->n_units > AEWN_WGW_MAX_UNITS || d->batch <= 0)
    return set_err(AEWN_ERR_INVALID, "wgradh: n_acts/n_units/batch out of range (%d/%d/%d)", d->n_acts, d->n_units, d->batch);
  cudaError_t e = cudaFuncSetAttribute(wgradh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM_BYTES);
  if (e != cudaSuccess) return cuda_err(e, "wgradh: cudaFuncSetAttribute");

  static_assert(sizeof(WhParams) <= 4000, "WhParams must fit the kernel parameter space");
  WhParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < d->n_acts; ++i) {
    const aewn_act16& a = d->acts[i];
    if (!a.ptr || a.channels < 8 || (a.channels & 7) || a.row_pitch < a.channels || (a.row_pitch & 7) || (a.batch_stride & 7) ||
        a.t_rows < 1 || a.batch < d->batch)
      return set_err(AEWN_ERR_INVALID, "wgradh: act %d invalid (channels %% 8, 16-byte aligned rows, batch)", i);
    cuuint64_t dims[3] = {(cuuint64_t)a.channels, (cuuint64_t)a.t_rows, (cuuint64_t)a.batch};
    cuuint64_t str[2] = {(cuuint64_t)a.row_pitch * 2u, (cuuint64_t)a.batch_stride * 2u};
    cuuint32_t box[3] = {64u, 64u, 1u};
    if (int rc = encode_f16_map(&p.map[i], a.ptr, 3, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "wgradh act")) return rc;
  }
  int work = 0;
  for (int u = 0; u < d->n_units; ++u) {
    const aewn_wgw_unit& un = d->units[u];
    if (un.g_act < 0 || un.g_act >= d->n_acts || un.m_valid < 1 || un.m_valid > 256 || un.n_chunks < 1 ||
        un.n_chunks > AEWN_WGW_MAX_CHUNKS || un.n_split < 1 || (un.g_row & 7))
      return set_err(AEWN_ERR_INVALID, "wgradh: unit %d invalid (g_act=%d g_row=%d m_valid=%d n_chunks=%d n_split=%d)", u,
                     un.g_act, un.g_row, un.m_valid, un.n_chunks, un.n_split);
    if (un.t_hi <= un.t_lo) return set_err(AEWN_ERR_INVALID, "wgradh: unit %d needs t_hi > t_lo", u);
    int cols = 0, boxes = 0;
    for (int c = 0; c < un.n_chunks; ++c) {
      const aewn_wgw_chunk& ch = un.chunk[c];
      if (ch.x_act < 0 || ch.x_act >= d->n_acts || ch.n < 16 || ch.n > 256 || (ch.n & 15) || ch.n_valid < 1 ||
          ch.n_valid > ch.n || !ch.out || (ch.x_row & 7))
        return set_err(AEWN_ERR_INVALID, "wgradh: unit %d chunk %d invalid (n=%d n_valid=%d x_row=%d)", u, c, ch.n, ch.n_valid,
                       ch.x_row);
      p.tm_col[u][c] = cols;
      p.x_box[u][c] = boxes;
      p.x_nbox[u][c] = ((ch.n >> 1) + 63) / 64;
      cols += (ch.n + 31) & ~31;
      boxes += p.x_nbox[u][c];
    }
    if (cols > 512 || boxes > WH_X_BOXES)
      return set_err(AEWN_ERR_INVALID, "wgradh: unit %d needs %d TMEM columns / %d staged boxes per CTA (limits 512 / %d)", u,
                     cols, boxes, WH_X_BOXES);
    p.tx_bytes[u] = 2 * (WH_G_BYTES + boxes * WH_BOX_BYTES);
    p.units[u] = un;
    p.work_begin[u] = work;
    work += un.n_split;
  }
  p.work_begin[d->n_units] = work;
  p.n_units = d->n_units;
  p.batch = d->batch;
  p.inv_scale = d->inv_scale;
  p.err = d->err;

  int ctas = (d->max_ctas > 0 ? d->max_ctas : sm_count()) & ~1;
  if (ctas > 2 * work) ctas = 2 * work;
  if (ctas < 2) ctas = 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(WH_THREADS);
  cfg.dynamicSmemBytes = WH_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, wgradh_kernel, p);
  count_launch();
  if (le != cudaSuccess) return cuda_err(le, "wgradh launch");
  return cuda_err(cudaGetLastError(), "wgradh launch");
}
