// wgrad: weight-gradient GEMM, contraction over (batch, time), on tcgen05 (TF32 in, FP32 accumulate in TMEM).
//
//   out[m, n] += sum_b sum_u G[b, g_row + m, u] * X[b, x_row + n, u + shift]          (dW lines)
//
// Both operands are K-major (time contiguous in the NCT tensors), so each K block is one TMA box {32 t x 128 rows}
// for G and up to two for X, all SWIZZLE_128B.  Work unit = (item, split); a unit accumulates its slice of the
// (batch x time) axis in TMEM and adds the 128 x n partial tile to `out` with red.global.add.f32.
// Warp roles and pipelines are the same as tgemm.cu.
#include "host_util.h"
#include "ptx.cuh"

namespace aewn {

constexpr int WG_BK = 32;
constexpr int WG_STAGES = 4;
constexpr int WG_BOX_BYTES = 128 * WG_BK * 4;        // 16 KB
constexpr int WG_STAGE_BYTES = 3 * WG_BOX_BYTES;     // G box + 2 X boxes (4 stages); "wide" launches: G + 3 X boxes, 3 stages
constexpr int WG_THREADS = 384;
constexpr int WG_EPI_WARPS = 8;
constexpr int WG_MAX_STAGES = 6;                     // pair-MMA mode: 6 stages of (G box + own half of the X rows)
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 256 + 1024;
static_assert(WG_MAX_STAGES * 2 * WG_BOX_BYTES == WG_STAGES * WG_STAGE_BYTES, "same ring in every mode");

struct WgParams {
  CUtensorMap map[AEWN_WGRAD_MAX_ACTS];
  aewn_wgrad_item items[AEWN_WGRAD_MAX_ITEMS];
  int unit_begin[AEWN_WGRAD_MAX_ITEMS + 1];  // prefix sum of n_split
  int n_items;
  int batch;
  int* err;
  int uniform_split;  // > 0: every item has this split count and units are ordered split-major (see wg_decode)
  int pair;           // 1: 2-CTA clusters; CTA r of cluster c takes item 2*pair_index + r; the pair shares its X tile
  int pair_mma;       // 1 (with pair): ONE cta_group::2 MMA stream per pair (M = 256 = the G rows of both items); each
                      // CTA stages its own G box and HALF of the X rows in its own shared memory (no multicast)
  int wide;           // 1: some item has 256 < n <= 384: 3 stages of (G + 3 X boxes), one 512-column accumulator
};

struct WgUnit {
  int item;
  int kb_begin, kb_end;  // flattened (b, time-block) range
  int blocks_per_b;
};

__device__ __forceinline__ WgUnit wg_decode(const WgParams& p, int unit, int crank) {
  // Split-major order when possible: the units of one split (all items) are adjacent, so the CTAs that run together
  // stream the SAME (batch, time) range of G and X -- the m-tiles of an X tile and the n-tiles of a G tile then hit in L2
  // instead of re-reading HBM (ncu, item-major order: 1.83 GB DRAM reads for 0.6 GB of unique operands).
  WgUnit u;
  int it, split;
  if (p.pair) {           // unit indexes (split, item PAIR); requires uniform_split (validated host-side)
    const int n_pairs = p.n_items >> 1;
    split = unit / n_pairs;
    it = 2 * (unit - split * n_pairs) + crank;
  } else if (p.uniform_split > 0) {
    split = unit / p.n_items;
    it = unit - split * p.n_items;
  } else {
    it = 0;
    while (it + 1 < p.n_items && unit >= p.unit_begin[it + 1]) ++it;
    split = unit - p.unit_begin[it];
  }
  u.item = it;
  const aewn_wgrad_item& im = p.items[it];
  u.blocks_per_b = (im.t_hi - im.t_lo + WG_BK - 1) / WG_BK;
  const long long total = static_cast<long long>(u.blocks_per_b) * p.batch;
  u.kb_begin = static_cast<int>(total * split / im.n_split);
  u.kb_end = static_cast<int>(total * (split + 1) / im.n_split);
  return u;
}

template <bool PMMA>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + WG_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Wide launches trade pipeline depth and the second accumulator for N up to 384 per unit: the G tile is then loaded
  // once for a whole 368-row X tile instead of once for 256 + once for 112 rows (the 112-wide units were L2-bound).
  const uint32_t n_stages = PMMA ? WG_MAX_STAGES : (p.wide ? 3u : 4u);
  const uint32_t stage_bytes = PMMA ? 2u * WG_BOX_BYTES : (p.wide ? 4u * WG_BOX_BYTES : 3u * WG_BOX_BYTES);
  const uint32_t acc_stages = (!PMMA && p.wide) ? 1u : 2u;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int i = 0; i < WG_MAX_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], (p.pair && !PMMA) ? 2 : 1);   // pair-MMA: the leader's commit releases both CTAs' stages
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PMMA ? 2 * WG_EPI_WARPS : WG_EPI_WARPS);   // pair-MMA: both CTAs' epilogues, on the leader
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PMMA) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.pair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int crank = p.pair ? static_cast<int>(cluster_ctarank()) : 0;
  const int cid = p.pair ? blockIdx.x >> 1 : blockIdx.x;
  const int n_cl = p.pair ? gridDim.x >> 1 : gridDim.x;
  const int total_units = p.pair ? p.uniform_split * (p.n_items >> 1) : p.unit_begin[p.n_items];

  if (warp < 4) {
  reg_dealloc<88>();
  // Single-issuer roles run warp-convergent with one elected issuing lane (see tgemm.cu: keeps the loop state in
  // uniform registers and removes the per-instruction waterfall loops).
  if (warp == 0) {
    {
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      const uint32_t lead_full = PMMA ? mapa_u32(&full_bar[0], 0) : 0u;
      for (int unit = cid; unit < total_units && ok; unit += n_cl) {
        const WgUnit u = wg_decode(p, unit, crank);
        if (u.kb_end <= u.kb_begin) continue;
        const aewn_wgrad_item& im = p.items[u.item];
        const int xboxes = (im.n + 127) >> 7;
        int b_next = u.kb_begin / u.blocks_per_b;                  // (batch, time block) of the unit's first K block;
        int tb = u.kb_begin - b_next * u.blocks_per_b;             // advanced incrementally (no division per block)
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          const int b = b_next;
          const int t = im.t_lo + tb * WG_BK;
          if (++tb == u.blocks_per_b) { tb = 0; ++b_next; }
          if (!mbar_wait_warp(&empty_bar[stage], phase ^ 1u, abort_flag)) { ok = false; break; }
          uint8_t* sg = smem + stage * stage_bytes;
          uint8_t* sx = sg + WG_BOX_BYTES;
          if (PMMA) {
            // both CTAs complete their bytes on the LEADER's full barrier; CTA r stages X rows [r * n/2, +n/2) (one
            // 128-row box, of which the MMA reads n/2)
            if (elect_one()) {
              if (crank == 0) mbar_expect_tx(&full_bar[stage], 4 * WG_BOX_BYTES);
              const uint32_t fb = lead_full + stage * 8u;
              tma_load_3d_pair(sg, &p.map[im.g_act], fb, t, im.g_row, b);
              tma_load_3d_pair(sx, &p.map[im.x_act], fb, t + im.shift, im.x_row + crank * (im.n >> 1), b);
            }
            __syncwarp();
            if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], (1 + xboxes) * WG_BOX_BYTES);
            tma_load_3d(sg, &p.map[im.g_act], &full_bar[stage], t, im.g_row, b);
            if (!p.pair) {
              for (int j = 0; j < xboxes; ++j)
                tma_load_3d(sx + j * WG_BOX_BYTES, &p.map[im.x_act], &full_bar[stage], t + im.shift, im.x_row + j * 128,
                            b);
            } else {   // CTA r fetches X boxes r, r+2 and multicasts them to both CTAs of the pair
              for (int j = crank; j < xboxes; j += 2)
                tma_load_3d_mcast(sx + j * WG_BOX_BYTES, &p.map[im.x_act], &full_bar[stage], t + im.shift,
                                  im.x_row + j * 128, b, 0x3);
            }
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (!PMMA || crank == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ok = true;
      // both operands K-major, 128B swizzle, 8-row groups 1 KB apart; per stage only the start-address field changes
      const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSW128);
      const uint32_t ring = smem_u32(smem);
      for (int unit = cid; unit < total_units && ok; unit += n_cl) {
        const WgUnit u = wg_decode(p, unit, crank);
        if (u.kb_end <= u.kb_begin) continue;
        const aewn_wgrad_item& im = p.items[u.item];
        if (!mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1u, abort_flag)) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        const int n0 = im.n > 256 ? 256 : im.n;
        const int n1 = im.n - n0;                       // second MMA part (wide items): X rows 256.., TMEM columns 256..
        const uint32_t idesc = make_idesc_tf32(PMMA ? 256 : 128, n0, 0, 0);
        const uint32_t idesc1 = n1 > 0 ? make_idesc_tf32(128, n1, 0, 0) : 0u;
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          if (!mbar_wait_warp(&full_bar[stage], phase, abort_flag)) { ok = false; break; }
          tc_fence_after();
          if (elect_one()) {
            // start-address fields (16-byte units), masked: the window address of CTA rank 1 carries the rank above
            const uint32_t g16 = ((ring + stage * stage_bytes) >> 4) & 0x3FFFu;
            const uint32_t x16 = g16 + (WG_BOX_BYTES >> 4);
            const uint32_t accum = kb > u.kb_begin;
#pragma unroll
            for (int ks = 0; ks < WG_BK / 8; ++ks) {
              const uint64_t adesc = desc0 + (g16 + ks * 2);            // K advances 32 B inside the swizzle row
              const uint64_t bdesc = desc0 + (x16 + ks * 2);
              if (PMMA) {
                umma_tf32_ss_pair(d_tmem, adesc, bdesc, idesc, accum | (ks > 0));
                continue;
              }
              umma_tf32_ss(d_tmem, adesc, bdesc, idesc, accum | (ks > 0));
              if (n1 > 0) {
                const uint64_t bdesc1 = bdesc + ((2 * WG_BOX_BYTES) >> 4);
                umma_tf32_ss(d_tmem + 256u, adesc, bdesc1, idesc1, accum | (ks > 0));
              }
            }
            if (PMMA) umma_commit_pair(&empty_bar[stage], 0x3);
            else if (p.pair) umma_commit_mcast(&empty_bar[stage], 0x3);
            else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        if (!ok) break;
        if (elect_one()) {
          if (PMMA) umma_commit_pair(&tfull_bar[acc], 0x3);
          else umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++acc == acc_stages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }
  } else {
    reg_alloc<208>();
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t acc = 0, acc_phase = 0;
    for (int unit = cid; unit < total_units; unit += n_cl) {
      const WgUnit u = wg_decode(p, unit, crank);
      if (u.kb_end <= u.kb_begin) continue;
      const aewn_wgrad_item& im = p.items[u.item];
      if (!mbar_wait(&tfull_bar[acc], acc_phase, abort_flag)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256u + (static_cast<uint32_t>(q * 32) << 16);
      const int m = q * 32 + lane;
      float* orow = im.out + static_cast<long long>(m) * im.out_rs;
      for (int c0 = half * 32; c0 < im.n; c0 += 64) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        if (m < im.m_valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < im.n_valid) atomicAdd(orow + static_cast<long long>(c0 + j) * im.out_cs, __uint_as_float(v[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PMMA && crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));   // the leader's barrier
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == acc_stages) { acc = 0; acc_phase ^= 1u; }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (p.pair) cluster_sync_all();
  if (threadIdx.x == 0 && *abort_flag && p.err) atomicExch(p.err, AEWN_ERR_TIMEOUT);
  if (warp == 2) {
    tc_fence_after();
    if (PMMA) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace aewn

using namespace aewn;

extern "C" int aewn_wgrad(const aewn_wgrad_desc* d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_err(AEWN_ERR_INVALID, "wgrad: null descriptor");
  if (d->n_acts < 1 || d->n_acts > AEWN_WGRAD_MAX_ACTS || d->n_items < 1 || d->n_items > AEWN_WGRAD_MAX_ITEMS ||
      d->batch <= 0)
    return set_err(AEWN_ERR_INVALID, "wgrad: n_acts/n_items/batch out of range (%d/%d/%d)", d->n_acts, d->n_items,
                   d->batch);
  cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
  if (e != cudaSuccess) return cuda_err(e, "wgrad: cudaFuncSetAttribute");

  // kernel parameters are limited to 4 KB: the item table rides in the parameter block
  static_assert(sizeof(WgParams) <= 4000, "WgParams must fit the kernel parameter space");
  WgParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < d->n_acts; ++i) {
    int rc = encode_act_map(&p.map[i], d->acts[i], 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (d->acts[i].batch < d->batch) return set_err(AEWN_ERR_INVALID, "wgrad: act %d batch smaller than problem batch", i);
  }
  int units = 0;
  for (int i = 0; i < d->n_items; ++i) {
    const aewn_wgrad_item& im = d->items[i];
    if (im.g_act < 0 || im.g_act >= d->n_acts || im.x_act < 0 || im.x_act >= d->n_acts)
      return set_err(AEWN_ERR_INVALID, "wgrad: item %d operand index out of range", i);
    if (im.n < 16 || im.n > 384 || (im.n & 15) || im.n_valid < 1 || im.n_valid > im.n || im.m_valid < 0 ||
        im.m_valid > 128)
      return set_err(AEWN_ERR_INVALID, "wgrad: item %d tile shape invalid (m_valid=%d n=%d n_valid=%d)", i, im.m_valid,
                     im.n, im.n_valid);
    if (im.t_hi <= im.t_lo || im.n_split < 1 || !im.out)
      return set_err(AEWN_ERR_INVALID, "wgrad: item %d needs t_hi>t_lo, n_split>=1, out", i);
    if ((im.t_lo & 3) || (im.shift & 3))
      return set_err(AEWN_ERR_INVALID, "wgrad: item %d t_lo=%d / shift=%d must be multiples of 4 (TMA 16-byte origin rule)",
                     i, im.t_lo, im.shift);
    p.items[i] = im;
    p.unit_begin[i] = units;
    units += im.n_split;
  }
  p.unit_begin[d->n_items] = units;
  p.n_items = d->n_items;
  p.wide = 0;
  for (int i = 0; i < d->n_items; ++i)
    if (d->items[i].n > 256) p.wide = 1;
  p.uniform_split = d->items[0].n_split;
  for (int i = 1; i < d->n_items; ++i)
    if (d->items[i].n_split != d->items[0].n_split) p.uniform_split = 0;
  p.batch = d->batch;
  p.err = d->err;

  p.pair = 0;
  if (d->pair_x) {
    // items (2i, 2i+1) must read the same X tile over the same (batch, time) range with the same split
    if ((d->n_items & 1) || p.uniform_split <= 0)
      return set_err(AEWN_ERR_INVALID, "wgrad: pair_x needs an even item count and a uniform n_split");
    for (int i = 0; i < d->n_items; i += 2) {
      const aewn_wgrad_item &a = d->items[i], &b = d->items[i + 1];
      if (a.x_act != b.x_act || a.x_row != b.x_row || a.n != b.n || a.shift != b.shift || a.t_lo != b.t_lo ||
          a.t_hi != b.t_hi || a.g_act != b.g_act)
        return set_err(AEWN_ERR_INVALID, "wgrad: pair_x items %d/%d do not share their X tile", i, i + 1);
    }
    p.pair = 1;
    if (d->pair_x == 2) {
      if (p.wide) return set_err(AEWN_ERR_INVALID, "wgrad: pair_x = 2 (cta_group::2 MMAs) needs n <= 256 for every item");
      p.pair_mma = 1;
    }
  }
  int ctas = d->max_ctas > 0 ? d->max_ctas : sm_count();
  const int work = p.pair ? 2 * p.uniform_split * (d->n_items / 2) : units;
  if (ctas > work) ctas = work;
  if (p.pair) ctas &= ~1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = WG_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = p.pair_mma ? cudaLaunchKernelEx(&cfg, wgrad_kernel<true>, p) : cudaLaunchKernelEx(&cfg, wgrad_kernel<false>, p);
  count_launch();
  if (le != cudaSuccess) return cuda_err(le, "wgrad launch");
  return cuda_err(cudaGetLastError(), "wgrad launch");
}
