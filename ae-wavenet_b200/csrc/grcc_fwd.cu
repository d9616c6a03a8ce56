// grcc_fwd: ONE launch per dilation layer (wavenet.py:91-111): filter conv + gate conv + conditioning projections + bias,
// tanh * sigmoid, 1x1 residual (+ x) and 1x1 skip (+= skip sum), on tcgen05 CTA pairs.
//
// Why a second engine next to tgemm.cu: ncu / bench arithmetic of the two-launch TF32 path showed the conv+gate launch
// bound by operand DELIVERY -- every CTA needs 64 B/clk of operands from L2 (32 KB per 512-cycle K block) against ~40
// B/clk/SM the chip delivers with all SMs pulling -- and the res+skip launch bound by HBM while the tensor pipe idles.
// This kernel (a) feeds the tensor cores FP16 operands (10-bit mantissa = TF32's, round-to-nearest instead of TF32's
// truncation; FP32 accumulation in TMEM; the residual stream itself stays FP32): half the bytes per MAC, twice the MMA
// rate; (b) reads the activations from CHANNELS-LAST fp16 copies (B, T, C): both operands are plain K-major
// SWIZZLE_128B tiles, one contiguous 16 KB box per K block, and a dilated tap is a shift of the box's ROW coordinate, so
// the 16-byte TMA origin rule no longer forces pre-shifted duplicates; (c) keeps z = tanh*sigmoid in shared memory as
// the A operand of the residual/skip GEMM (in inference mode it never travels to HBM).
//
// Per CTA pair and 256 time steps (128 per CTA), jobs run in a fixed order through two 256-column TMEM regions:
//   G1.j (j < D/128): acc[t, 0:128 | 128:256] = filt | gate pre-activations of channels [128j, 128j+128)
//                     K = [x16(t-d) | x16(t) | cond16(t), 1], 64 channels per ring stage (4 x 32 KB ring)
//   epilogue G1.j   : tanh, sigmoid (-> packed fp16 derivative factors for the backward pass), z -> fp16 -> zbuf (K-major,
//                     manual 128B swizzle)
//   SKP.c           : acc = Ws[c] . z   -> skip sum (store / TMA reduce-add / relu(old + acc))          (A operand = zbuf)
//   RES.c           : acc = Wr[c] . z   -> x32_next = acc + x32 ; x16_next = fp16(x32_next)
// The residual / skip jobs come smallest-epilogue first: the next tile's first gate job may start as soon as the region
// of the second-to-last job is drained, so the 256-channel residual chunk (the longest epilogue) goes last and overlaps
// the next tile's gate MMAs.
// Warps: 0 TMA producer, 1 MMA issuer (leader CTA), 2 TMEM allocator, 3 idle, 4-19 epilogue (16 warps, 4 per TMEM lane
// quadrant, 16-column chunks; registers per SM sub-partition: one control warp at 64 + four epilogue warps at 104).  The
// saved activations / z / x_next leave the SM with plain coalesced stores (lane = time step: one 128-byte line per
// instruction, st.global.cs), the fp16 copy with one 32-byte st.global.v8 per lane; a staged TMA-store path measured
// slower here.  The running skip sum keeps the TMA reduce-add (one 2 KB box per 16 channels through a per-warp staging
// tile): red.global.add per element measured 2x slower (profiles/r3_gf_*).
// NO local memory in the loops: a spill is a global-memory access that queues behind the warp's own streaming stores
// (thousands of cycles each in this kernel) -- check `-Xptxas -v` / STL, LDL in the SASS after every edit.
// All waits are bounded.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "host_util.h"
#include "ptx.cuh"

namespace aewn {

constexpr int GF_BM = 128;
constexpr int GF_KB = 64;                         // K elements per ring stage: 64 fp16 = one 128-byte swizzle row
constexpr int GF_MAX_STAGES = 4;                  // operand ring: 4 x 32 KB
constexpr int GF_A_BYTES = GF_BM * 128;           // 16 KB: 128 time rows x 128 B
constexpr int GF_W_BYTES = 128 * 128;             // 16 KB: 128 weight rows x 128 B (this CTA's half of the N rows)
constexpr int GF_STAGE_BYTES = GF_A_BYTES + GF_W_BYTES;
constexpr int GF_MAX_D = 256;
constexpr int GF_ZBUF_BYTES = (GF_MAX_D / GF_KB) * GF_A_BYTES;   // 64 KB: z of this CTA's 128 time rows, all D channels
constexpr int GF_CTRL_WARPS = 4;                 // 0: TMA producer, 1: MMA issuer, 2: TMEM allocator, 3: idle (one per SM sub-partition)
constexpr int GF_THREADS = 640;                  // 4 control warps + 16 epilogue warps (see the register split below)
constexpr int GF_EPI_WARPS = 16;
constexpr int GF_STG_BYTES = GF_EPI_WARPS * 2048;               // one [16 ch][32 t] fp32 tile per epilogue warp: the skip
                                                                // sum's TMA reduce-add (red.global.add measured 2x slower)
constexpr int GF_POOL_BYTES = GF_MAX_STAGES * GF_STAGE_BYTES + GF_STG_BYTES;   // 160 KB
constexpr int GF_SMEM_BYTES = GF_POOL_BYTES + GF_ZBUF_BYTES + 256 + 640 + 1024;   // + barriers + GfHot + alignment slack
constexpr int GF_MAX_JOBS = 8;

enum { GF_GATE = 0, GF_RES = 1, GF_SKP = 2, GF_GZ = 3 };

struct GfJob {
  int kind;
  int w_row;     // first weight row (w1 for GATE, w2 otherwise)
  int n;         // accumulator columns (MMA N)
  int n_valid;   // output channels of this job
  int ch0;       // first output channel (GATE: first z channel)
  int a_ring;    // RES / SKP jobs: 1 = the A operand comes through the ring (segments below) instead of the z buffer
  int split;     // RES jobs: accumulator columns >= split are reduce-added into the `skp` tensor (channel = column - split)
};

struct GfSeg {
  int map;       // 0: xa, 1: ca
  int shift;     // row (time) shift of the box
  int kb;        // ring stages (64 channels each)
};

// Scalars and the job table: copied to shared memory at kernel start.  Read from the kernel parameter space they cost an
// LDC per use (the job table is indexed at run time), and ncu attributed ~17 % of the epilogue warps' samples to
// instructions waiting for those constant loads.
struct GfHot {
  GfJob job[GF_MAX_JOBS];
  int n_jobs, n_gate;
  int kb_z;                               // ring stages of a z-operand job
  GfSeg seg[3];                           // K segments of a ring-operand job, in W column order
  int n_segs, ring_stages;
  int ab_bf16;                            // operands are bf16 (data-gradient variant) instead of fp16
  int add_t_lo;                           // RES epilogue: the addend applies for t >= add_t_lo
  const float* x32;                        // residual source (B, R, Tp)
  float* xo32;                             // x_next (B, R, Tp), same strides
  long long x_bs, x_cs;
  float* dup;                              // optional pre-shifted fp32 duplicate of x_next (backward wgrad tap, d_next % 4 != 0)
  int dup_toff, dup_t_hi;
  __half* xo16;                            // (B, Tp, x16_cp) channels-last fp16 copy of x_next
  long long x16_bs;
  int x16_cp;
  float* skp;
  long long s_bs, s_cs;
  int save, z_out, skp_mode;               // skp_mode: 0 store, 1 reduce-add, 2 relu(old + acc), 3 relu(acc)
  float *th, *sg, *z;                      // (B, D, Tp) saved activations (save / z_out)
  long long a_bs, a_cs;
  int batch, t_begin, n_tgroups, n_res;
  int t_lo, t_zero_lo, t_hi, skp_t_lo, skp_zero_lo;
  int prefetch;
  int dbg;                                 // AEWN_GF_DBG bit mask: skip parts of the epilogues (timing experiments only)
  int* err;
  long long* dbg_clock;   // optional: cluster 0 / CTA 0 stamps its first items (profiles/gf_phase_clock.py)
  const float* out_scale; // optional (data-gradient variant with scaled fp16 operands): accumulators are multiplied by *out_scale
  __half* z16;            // optional fp16 channels-last copy of z (B, Tp, z16_cp)
  long long z16_bs;
  int z16_cp, pad0;
};
static_assert(sizeof(GfHot) % 8 == 0 && sizeof(GfHot) <= 640, "GfHot is copied to shared memory as 64-bit words");

struct GfParams {
  CUtensorMap xa, ca, w1, w2;             // operand loads (fp16)
  CUtensorMap xr_m;                          // fp32 residual source, box {128 t, 32 ch, 1}: L2 prefetch only
  CUtensorMap skp_m;                         // skip sum (t, ch, b), box {32 t, 16 ch, 1}: TMA reduce-add
  GfHot hot;
};

__device__ __forceinline__ void umma_f16_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::f16, A and B = F16 (format 0), both K-major, FP32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int bf16 = 0) {
  return (1u << 4) | (static_cast<uint32_t>(bf16) << 7) | (static_cast<uint32_t>(bf16) << 10) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// TMEM -> registers: 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// cta_group::2 loads with an L2 cache policy (createpolicy): operands that are read again soon -- the weights by every
// CTA pair, the activation tile by the second gate job and, dil steps later, as the shifted tap -- are kept with
// evict_last, so that the kernel's ~0.9 GB of streaming output per launch does not push them out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                      int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair_hint(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                      int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}

// Pull one box of a tensor into L2 (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

struct GfItem {
  int b, g0, tau0;
  bool do_skp;
};

__device__ __forceinline__ GfItem gf_decode(const GfHot& p, int item, int crank) {
  GfItem it;
  const int tg = item % p.n_tgroups;
  it.b = item / p.n_tgroups;
  it.g0 = p.t_begin + tg * 2 * GF_BM;
  it.tau0 = it.g0 + crank * GF_BM;
  it.do_skp = it.g0 + 2 * GF_BM > p.skp_t_lo;   // decided per GROUP: both CTAs walk the same job sequence
  return it;
}

// read-once global load: no L1 allocation (the L1 is ~20 KB next to 225 KB of shared memory)
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
#ifdef AEWN_GF_LDCS
  v = __ldcs(p);
#else
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
#endif
  return v;
}

// KINDS: bit mask of the job kinds (1 << GF_*) this instantiation can run.  The epilogues of all kinds in ONE function
// strain the 104-register budget of the epilogue warps (adding the gate-derivative epilogue made ptxas spill in the others);
// the forward layer, the data gradient and the gate derivative each get their own instantiation.
constexpr int GF_KINDS_FWD = (1 << GF_GATE) | (1 << GF_RES) | (1 << GF_SKP);
constexpr int GF_KINDS_DGRAD = 1 << GF_RES;
constexpr int GF_KINDS_GZ = 1 << GF_GZ;

template <int STAGES, int KINDS>
__global__ void __launch_bounds__(GF_THREADS, 1) grcc_fwd_kernel(const __grid_constant__ GfParams p) {
  constexpr int RING_BYTES = GF_MAX_STAGES * GF_STAGE_BYTES;
  static_assert(STAGES <= GF_MAX_STAGES, "ring depth");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* zbuf = smem + RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES + GF_ZBUF_BYTES + GF_STG_BYTES);
  uint64_t* empty_bar = full_bar + GF_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + GF_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* zready_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zready_bar + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  GfHot* hot_s = reinterpret_cast<GfHot*>(smem + RING_BYTES + GF_ZBUF_BYTES + GF_STG_BYTES + 256);
  if (threadIdx.x < sizeof(GfHot) / 8)
    reinterpret_cast<unsigned long long*>(hot_s)[threadIdx.x] = reinterpret_cast<const unsigned long long*>(&p.hot)[threadIdx.x];
  const GfHot& hp = *hot_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // accumulator scale / copy scale slots (the pointer is rebuilt where it is used: a value computed before the register
  // split is spilled for every role)
#define GF_OSC_S(smem_) reinterpret_cast<float*>((smem_) + RING_BYTES + GF_ZBUF_BYTES + GF_STG_BYTES + 116)
  if (threadIdx.x == 0) {
    float* osc_s = GF_OSC_S(smem);
    osc_s[0] = p.hot.out_scale ? __ldg(p.hot.out_scale) : 1.0f;
    osc_s[1] = 1.0f / osc_s[0];                      // powers of two: exact (the 16-bit copy of the output is scaled back up)
    *abort_flag = 0;
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);    // leader: its own arrive.expect_tx; the bytes of BOTH CTAs complete on it
      mbar_init(&empty_bar[i], 1);   // the leader's cta_group::2 commit releases the stage in both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * GF_EPI_WARPS);   // on the leader: the epilogue warps of both CTAs
    }
    mbar_init(zready_bar, 2 * GF_EPI_WARPS * (p.hot.n_gate > 0 ? p.hot.n_gate : 1));   // every epilogue warp of the pair, once per gate job
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.xa);
    tma_prefetch_desc(&p.ca);
    tma_prefetch_desc(&p.w1);
    tma_prefetch_desc(&p.w2);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int crank = static_cast<int>(cluster_ctarank());
  const int n_cl = gridDim.x >> 1;
  const int total = hp.batch * hp.n_tgroups;

  if (warp < GF_CTRL_WARPS) {
    // Registers are a per-sub-partition resource (16 K each; warp w lives on sub-partition w % 4): one control warp and
    // four epilogue warps each, 64 + 4 x 104 = 480 = 5 x 96 (the launch allocation).  The epilogue code must not
    // spill: a local-memory load queues behind the warp's own streaming stores in the LSU and costs 2-6 k cycles in this
    // kernel (phase clock, profiles/r3_gf_spill_*).
    reg_dealloc<64>();
    if (warp == 0) {
      // ===================================================== TMA producer (both CTAs)
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      const uint32_t lead_full = mapa_u32(&full_bar[0], 0);
      const uint64_t pol_keep = l2_policy_evict_last();
      const uint64_t pol_drop = (hp.prefetch & 16) ? l2_policy_evict_first() : pol_keep;
      for (int item = blockIdx.x >> 1; item < total && ok; item += n_cl) {
        const GfItem it = gf_decode(hp, item, crank);
        // the residual rows this tile's RES epilogues will add (128 time steps x R channels of x32): pull them into L2
        // now, ~20 k cycles before the epilogue warps load them
        if ((hp.prefetch & 1) && elect_one()) {
          for (int c = 0; c < hp.n_res; c += 32) tma_prefetch_l2_3d(&p.xr_m, it.tau0, c, it.b);
        }
        __syncwarp();
        for (int jb = 0; jb < hp.n_jobs && ok; ++jb) {
          const GfJob jd = hp.job[jb];
          if (jd.kind == GF_SKP && !it.do_skp) continue;
          const bool use_ring = jd.kind == GF_GATE || jd.a_ring;
          const int nst = use_ring ? hp.ring_stages : hp.kb_z;
          // After the last gate job of this tile: pull the NEXT tile's HBM-fresh activation operand (fp16 x and conditioning
          // rows at shift 0) into L2, so that its first gate job -- which otherwise waits for HBM behind a 4-stage ring
          // (phase clock: ~1000 cycles per stage against ~570 from L2) -- finds it there.  (bit 1 of AEWN_GF_PREFETCH)
          if ((hp.prefetch & 2) && jb == hp.n_gate && hp.n_gate > 0 && item + n_cl < total && elect_one()) {
            const GfItem nx = gf_decode(hp, item + n_cl, crank);
            for (int si = 0; si < hp.n_segs; ++si)
              if (hp.seg[si].shift == 0)
                for (int k2 = 0; k2 < hp.seg[si].kb; ++k2)
                  tma_prefetch_l2_3d(hp.seg[si].map ? &p.ca : &p.xa, k2 * GF_KB, nx.tau0, nx.b);
          }
          __syncwarp();
          int sgi = 0, kb = 0;                                   // current K segment / block inside it (ring-operand jobs)
          for (int s = 0; s < nst; ++s) {
            if (!mbar_wait_warp(&empty_bar[stage], phase ^ 1u, abort_flag)) { ok = false; break; }
            if (elect_one()) {
              uint8_t* sa = smem + stage * GF_STAGE_BYTES;
              uint8_t* sw = sa + GF_A_BYTES;
              const uint32_t fb = lead_full + stage * 8u;
              if (use_ring) {
                if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * GF_STAGE_BYTES);
                const GfSeg sg = hp.seg[sgi];
                // the LAST read of an activation row in this layer (last gate job, unshifted tap) may leave L2 early
                const bool last_use = jd.kind == GF_GATE && jb == hp.n_gate - 1 && sg.map == 0 && sg.shift == 0;
                tma_load_3d_pair_hint(sa, sg.map ? &p.ca : &p.xa, fb, kb * GF_KB, it.tau0 + sg.shift, it.b,
                                      last_use ? pol_drop : pol_keep);
                tma_load_2d_pair_hint(sw, &p.w1, fb, s * GF_KB, jd.w_row + crank * (jd.n >> 1), pol_keep);
              } else {
                // CTA r stages W2 rows [r * n/2, (r+1) * n/2) of the job (a 128-row box; the MMA reads n/2 of them)
                if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * GF_W_BYTES);
                tma_load_2d_pair_hint(sw, &p.w2, fb, s * GF_KB, jd.w_row + crank * (jd.n >> 1), pol_keep);
              }
            }
            __syncwarp();
            if (use_ring && ++kb == hp.seg[sgi].kb) { kb = 0; ++sgi; }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1) {
      // ===================================================== MMA issuer (the pair's leader CTA)
      if (crank == 0) {
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, zphase = 0;
        bool ok = true;
        const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSW128);   // K-major, 128B swizzle, 8-row groups 1 KB apart
        const uint32_t ring = smem_u32(smem);
        const uint32_t z16_0 = (smem_u32(zbuf) >> 4) & 0x3FFFu;
        // whole-kernel stamps of cluster 0's issuer (slot [tile 3][job 6]): cycles and nanoseconds -> effective SM clock
        long long dbg_c0 = 0, dbg_n0 = 0;
        int dbg_items = 0;
        if (hp.dbg_clock && blockIdx.x == 0 && lane == 0) {
          dbg_c0 = clock64();
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_n0));
        }
        for (int item = blockIdx.x >> 1; item < total && ok; item += n_cl) {
          const GfItem it = gf_decode(hp, item, crank);
          bool z_waited = false;
          ++dbg_items;
          if (hp.dbg_clock && blockIdx.x == 0 && item < 4 * n_cl && lane == 0)
            hp.dbg_clock[((item / n_cl) * GF_MAX_JOBS + 7) * 6] = clock64();      // top of the tile (slot of the unused job 7)
          for (int jb = 0; jb < hp.n_jobs && ok; ++jb) {
            const GfJob jd = hp.job[jb];
            if (jd.kind == GF_SKP && !it.do_skp) continue;
            const bool stamp = hp.dbg_clock && blockIdx.x == 0 && item < 4 * n_cl && lane == 0;
            long long* ck = hp.dbg_clock + ((item / n_cl) * GF_MAX_JOBS + jb) * 6;
            unsigned int w_full = 0;
            if (stamp) ck[0] = clock64();
            if (!mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1u, abort_flag)) { ok = false; break; }
            const bool use_ring = jd.kind == GF_GATE || jd.a_ring;
            if (!use_ring && !z_waited) {
              // every epilogue warp of the pair has written its z columns of this tile (generic proxy -> fence -> arrive)
              const unsigned int tz = clock();
              if (!mbar_wait_warp(zready_bar, zphase, abort_flag)) { ok = false; break; }
              if (stamp) ck[5] = clock() - tz;
              zphase ^= 1u;
              z_waited = true;
            }
            tc_fence_after();
            if (stamp) ck[1] = clock64();
            const uint32_t d_tmem = tmem_base + acc * 256u;
            const uint32_t idesc = make_idesc_f16(2 * GF_BM, jd.n, hp.ab_bf16);
            const int nst = use_ring ? hp.ring_stages : hp.kb_z;
            for (int s = 0; s < nst; ++s) {
              const unsigned int tf = clock();
              if (!mbar_wait_warp(&full_bar[stage], phase, abort_flag)) { ok = false; break; }
              w_full += clock() - tf;
              tc_fence_after();
              if (elect_one()) {
                const uint32_t s16 = ((ring + stage * GF_STAGE_BYTES) >> 4) & 0x3FFFu;
                const uint32_t w16 = s16 + (GF_A_BYTES >> 4);
                const uint32_t a16 = use_ring ? s16 : z16_0 + static_cast<uint32_t>(s) * (GF_A_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < GF_KB / 16; ++ks)   // K advances 16 fp16 = 32 B inside the swizzle row
                  umma_f16_ss_pair(d_tmem, desc0 + (a16 + ks * 2), desc0 + (w16 + ks * 2), idesc,
                                   static_cast<uint32_t>((s | ks) != 0));
                umma_commit_pair(&empty_bar[stage], 0x3);
              }
              __syncwarp();
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (!ok) break;
            if (elect_one()) umma_commit_pair(&tfull_bar[acc], 0x3);
            __syncwarp();
            if (stamp) { ck[2] = clock64(); ck[3] = w_full; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          }
          // a tile whose skip job was skipped still has to consume the z phase (the epilogues always arrive)
          if (ok && !z_waited && hp.n_gate > 0) {
            if (!mbar_wait_warp(zready_bar, zphase, abort_flag)) { ok = false; break; }
            zphase ^= 1u;
          }
        }
        if (hp.dbg_clock && blockIdx.x == 0 && lane == 0) {
          long long n1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
          long long* ck = hp.dbg_clock + (3 * GF_MAX_JOBS + 6) * 6;
          ck[0] = 1;                    // marks the slot as used
          ck[1] = clock64() - dbg_c0;   // cycles the issuer spent on its `dbg_items` tiles
          ck[2] = n1 - dbg_n0;          // the same span in nanoseconds
          ck[3] = dbg_items;
        }
      }
    }
  } else {
    reg_alloc<104>();
    // ===================================================== epilogue (both CTAs): 16 warps, 4 per TMEM lane quadrant
    // Warp (q, h): TMEM lanes [32q, 32q+32) = time rows, a quarter of every job's columns, in chunks of 16 columns
    // (104 registers per thread: 640 threads share the 64 K register file).  fp32 outputs go through two 2 KB half-tiles
    // per warp ([16 channels][32 time steps], alternating: the box of store k is written while the TMA engine still reads
    // store k-1) and leave as TMA stores / reduce-adds of {32 t, 16 ch} boxes.
    const int q = warp & 3;
    const int h = (warp - GF_CTRL_WARPS) >> 2;
    const int row = q * 32 + lane;
#ifdef AEWN_GF_EXPERIMENTS
    const int dbg = hp.dbg;      // timing experiments (profiles/r3_gf_epilogue_experiments.txt): epilogue parts switched off
#else
    constexpr int dbg = 0;
#endif
    float* const stg = reinterpret_cast<float*>(smem + RING_BYTES + GF_ZBUF_BYTES) + (warp - GF_CTRL_WARPS) * 512;
    uint32_t acc = 0, acc_phase = 0;
    float xmax = 0.0f;
    bool ok = true;
    // The two loop counters live in shared memory, not in registers: under the 104-register budget ptxas spilled exactly
    // these two to LOCAL memory, and a local load at a loop back-edge queues behind the warp's own streaming stores in the
    // L1 pipeline (2-6 k cycles per job boundary in the phase clock).  Shared-memory loads do not take that path.
    volatile int* const wst = reinterpret_cast<volatile int*>(smem + RING_BYTES + GF_ZBUF_BYTES + GF_STG_BYTES + 128) +
                              (warp - GF_CTRL_WARPS) * 2;
#ifdef AEWN_GF_PHASE_CLOCK
    if (warp == 4 && lane == 0) *reinterpret_cast<volatile int*>(smem + RING_BYTES + GF_ZBUF_BYTES + GF_STG_BYTES + 112) = -1;
#endif
    for (wst[0] = blockIdx.x >> 1; wst[0] < hp.batch * hp.n_tgroups && ok; wst[0] = wst[0] + (gridDim.x >> 1)) {
      const int item = wst[0];
#ifdef AEWN_GF_PHASE_CLOCK
      if (warp == 4 && lane == 0) {
        volatile int* tile_no = reinterpret_cast<volatile int*>(smem + RING_BYTES + GF_ZBUF_BYTES + GF_STG_BYTES + 112);
        *tile_no = *tile_no + 1;
      }
#endif
      const GfItem it = gf_decode(hp, item, crank);
      const int tau = it.tau0 + row;
      const int slab0 = it.tau0 + q * 32;
      const bool in_range = tau >= hp.t_lo && tau < hp.t_hi;
      const bool keep = in_range && tau >= hp.t_zero_lo;
      // Residual values (x32) of the chunk the RES epilogue will process next.  The loads run one chunk ahead, ACROSS job
      // boundaries: the first chunk of a RES job is requested before the skip epilogue / at the last chunk of the RES job
      // in front of it (phase clock: 5.5-8 k cycles from "job seen" to the first values when requested at the job's top).
      float buf[16];
      int buf_job = -1;                 // job whose first chunk `buf` holds
      auto res_load = [&](int j2, int c0) {
        const int nv2 = hp.job[j2].n_valid;
        const float* sp = hp.x32 + static_cast<long long>(it.b) * hp.x_bs +
                          static_cast<long long>(hp.job[j2].ch0 + c0) * hp.x_cs + tau;
        const bool on = keep && hp.x32 != nullptr && tau >= hp.add_t_lo && !(dbg & 1);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          buf[j] = (on && c0 + j < nv2) ? ld_stream(sp) : 0.0f;
          sp += hp.x_cs;
        }
      };
      auto res_first = [&](int j2) {    // first chunk of job j2 (a z-operand RES job), if this warp has columns in it
        const int n2 = hp.job[j2].n;
        const int cb2 = h * (((n2 + 63) >> 6) << 4);
        if (cb2 < n2) res_load(j2, cb2);
        buf_job = j2;
      };
      for (wst[1] = 0; wst[1] < hp.n_jobs && ok; wst[1] = wst[1] + 1) {
        const int jb = wst[1];
        const GfJob jd = hp.job[jb];
        if (jd.kind == GF_SKP && !it.do_skp) continue;
        // At the tile's LAST gate job: pull the residual rows this warp's RES epilogues will add (x32[b, ch, slab]) from
        // HBM into L2, one 128-byte line per LANE and instruction (lane = channel), 15-30 k cycles before they are
        // loaded.  Unprefetched, each 16-channel chunk of a RES epilogue waits ~4 k cycles for HBM behind the write
        // stream.  The moment matters: the operand ring is idle now (the residual / skip GEMMs only stream weights);
        // issued at the tile's first job the prefetch competed with the second gate job's operand feed (+7 k cycles).
        if (jb == hp.n_gate - 1 && (hp.prefetch & 4) && hp.x32) {
#pragma unroll 1
          for (int j2 = 0; j2 < hp.n_jobs; ++j2) {
            if (hp.job[j2].kind != GF_RES || hp.job[j2].a_ring) continue;
            const int n2 = hp.job[j2].n, nv2 = min(hp.job[j2].n_valid, hp.job[j2].split);
            const int span2 = ((n2 + 63) >> 6) << 4;
#pragma unroll 1
            for (int c = h * span2 + lane; c < min(h * span2 + span2, nv2); c += 32) {
              const float* pl = hp.x32 + static_cast<long long>(it.b) * hp.x_bs +
                                static_cast<long long>(hp.job[j2].ch0 + c) * hp.x_cs + slab0;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pl));
            }
          }
        }
        const uint32_t taddr = tmem_base + acc * 256u + (static_cast<uint32_t>(q * 32) << 16);
        // phase clock (profiles/gf_phase_clock.py): nothing of it stays live in registers between the stamps
        auto stamp_at = [&](int k) {
#ifdef AEWN_GF_PHASE_CLOCK
          // (build with -DAEWN_GF_PHASE_CLOCK; warp 4 = (q 0, h 0) of CTA 0; the tile number lives in shared memory too)
          volatile int* tile_no = reinterpret_cast<volatile int*>(smem + RING_BYTES + GF_ZBUF_BYTES + GF_STG_BYTES + 112);
          if (hp.dbg_clock && blockIdx.x == 0 && warp == 4 && lane == 0 && *tile_no < 4)
            hp.dbg_clock[4 * GF_MAX_JOBS * 6 + (*tile_no * GF_MAX_JOBS + wst[1]) * 6 + k] = clock64();
#endif
        };
        stamp_at(0);
        if ((KINDS & (1 << GF_GATE)) && jd.kind == GF_GATE) {
          if (!mbar_wait_warp(&tfull_bar[acc], acc_phase, abort_flag)) { ok = false; break; }
          tc_fence_after();
          stamp_at(1);
#pragma unroll 1
          for (int i = 0; i < 2; ++i) {
            const int c0 = 32 * h + 16 * i;            // z channels [c0, c0 + 16) of this 128-channel block
            uint32_t vf[16], vg[16];
            tmem_ld16(taddr + c0, vf);
            tmem_ld16(taddr + 128 + c0, vg);
            tmem_ld_wait();
            // tanh(f) = (1 - a) / (1 + a), a = e^(-2f); sigmoid(g) = 1 / (1 + b), b = e^(-g): ONE reciprocal of
            // (1 + a)(1 + b) serves both (3 MUFU ops per element instead of 4; the exponents are clamped so that the
            // product stays finite: 2^60 * 2^60)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float ea = ex2_ftz(fminf(-2.8853900817779268f * __uint_as_float(vf[j]), 60.0f));
              const float eb = ex2_ftz(fminf(-1.4426950408889634f * __uint_as_float(vg[j]), 60.0f));
              const float pa = 1.0f + ea, pb = 1.0f + eb;
              const float inv = rcp_ftz(pa * pb);
              vf[j] = __float_as_uint(keep ? (1.0f - ea) * pb * inv : 0.0f);
              vg[j] = __float_as_uint(keep ? pa * inv : 0.0f);
            }
            // z -> fp16 -> zbuf: K-major rows of 128 B (64 channels), 16-byte chunk index XOR (row & 7) = SWIZZLE_128B
            const int ch = jd.ch0 + c0;
            uint32_t zw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              zw[k] = pack_f16x2(__uint_as_float(vf[2 * k]) * __uint_as_float(vg[2 * k]),
                                 __uint_as_float(vf[2 * k + 1]) * __uint_as_float(vg[2 * k + 1]));
            {
              uint8_t* zrow = zbuf + (ch >> 6) * GF_A_BYTES + row * 128;
              const int lc0 = (ch & 63) >> 3;
              *reinterpret_cast<uint4*>(zrow + (((lc0 + 0) ^ (row & 7)) << 4)) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
              *reinterpret_cast<uint4*>(zrow + (((lc0 + 1) ^ (row & 7)) << 4)) = make_uint4(zw[4], zw[5], zw[6], zw[7]);
            }
            if (i == 1) {
              // Both chunks' z rows are in shared memory and the accumulator has been read: release the TMEM region and
              // signal z NOW, ahead of this chunk's global stores -- the arrivals (remote ones for the pair's second CTA)
              // otherwise queue behind the store burst, and the MMA issuer saw its regions thousands of cycles late.
              fence_proxy_async_smem();   // the z rows written above are read by tcgen05.mma (async proxy)
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive_cluster(mapa_u32(zready_bar, 0));
                if (crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
                else mbar_arrive(&tempty_bar[acc]);
              }
            }
            if (hp.z16 && in_range) {      // fp16 channels-last copy of z for the weight gradients: one 32-byte store per lane
              __half* zr = hp.z16 + static_cast<long long>(it.b) * hp.z16_bs + static_cast<long long>(tau) * hp.z16_cp + ch;
              asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(zr), "r"(zw[0]), "r"(zw[1]),
                           "r"(zw[2]), "r"(zw[3]), "r"(zw[4]), "r"(zw[5]), "r"(zw[6]), "r"(zw[7])
                           : "memory");
            }
            // tanh / sigmoid (/ z) for the backward pass: plain coalesced stores (lane = time step: every store
            // instruction writes one full 128-byte line).  Through the staging tiles these were 3 TMA stores per chunk,
            // each with its acquire / proxy fence / issue latency (~800 cycles, phase clock) in the warp's serial chain.
            if (in_range && !(dbg & 8)) {
              const long long o0 = static_cast<long long>(it.b) * hp.a_bs + static_cast<long long>(ch) * hp.a_cs + tau;
              if (hp.save == 2) {
                // what the backward pass needs from tanh / sigmoid are the two derivative factors
                //   a = d z / d filt = sg (1 - th^2),   b = d z / d gate = th sg (1 - sg)          
                // kept as ONE 32-bit word {fp16 a, fp16 b} per element (relative precision 2^-11, the operand precision
                // of the backward engines): half the saved-activation bytes of fp32 tanh + sigmoid, in both passes
                uint32_t* ap = reinterpret_cast<uint32_t*>(hp.th) + o0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float th = __uint_as_float(vf[j]), sg = __uint_as_float(vg[j]);
                  __stcs(ap, pack_f16x2(sg * fmaf(-th, th, 1.0f), th * sg * (1.0f - sg)));
                  ap += hp.a_cs;
                }
              } else if (hp.save) {
                float* tp = hp.th + o0;
                float* sp = hp.sg + o0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  __stcs(tp, __uint_as_float(vf[j]));
                  __stcs(sp, __uint_as_float(vg[j]));
                  tp += hp.a_cs;
                  sp += hp.a_cs;
                }
              }
              if (hp.z_out) {
                float* zp = hp.z + o0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  __stcs(zp, __uint_as_float(vf[j]) * __uint_as_float(vg[j]));
                  zp += hp.a_cs;
                }
              }
            }
          }
        } else if ((KINDS & (1 << GF_GZ)) && jd.kind == GF_GZ) {
          // Gate derivative (autograd of wavenet.py:100-103): the accumulator is g_z * s (its operands are
          // the scaled fp16 copies of g_x and g_skp); with the saved word {fp16 a, fp16 b} of the forward pass
          // g_f = g_z a, g_g = g_z b leave as the scaled fp16 channels-last copy [g_f | g_g] every other backward engine reads.
          const int span = ((jd.n + 63) >> 6) << 4;
          const int cb = h * span;
          const int ce = min(cb + span, jd.n);
          const uint32_t* ap = reinterpret_cast<const uint32_t*>(hp.th) + static_cast<long long>(it.b) * hp.a_bs + tau;
          uint32_t abw[16];
          auto ld_ab = [&](int c0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              abw[j] = (keep && c0 + j < jd.n_valid) ? __ldcs(ap + static_cast<long long>(c0 + j) * hp.a_cs) : 0u;
          };
          if (cb < ce) ld_ab(cb);
          if (!mbar_wait_warp(&tfull_bar[acc], acc_phase, abort_flag)) { ok = false; break; }
          tc_fence_after();
          stamp_at(1);
#pragma unroll 1
          for (int c0 = cb; c0 < ce; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
            if (c0 + 16 >= ce) {      // last chunk of this warp: release the region
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
                else mbar_arrive(&tempty_bar[acc]);
              }
            }
            uint32_t gfw[8], ggw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t w0 = abw[2 * k], w1 = abw[2 * k + 1];        // {fp16 a (low half), fp16 b (high half)}
              const float a0 = __half2float(__ushort_as_half(static_cast<unsigned short>(w0 & 0xffffu)));
              const float b0 = __half2float(__ushort_as_half(static_cast<unsigned short>(w0 >> 16)));
              const float a1 = __half2float(__ushort_as_half(static_cast<unsigned short>(w1 & 0xffffu)));
              const float b1 = __half2float(__ushort_as_half(static_cast<unsigned short>(w1 >> 16)));
              const float z0 = __uint_as_float(v[2 * k]), z1 = __uint_as_float(v[2 * k + 1]);
              const float f0 = z0 * a0, f1 = z1 * a1, g0 = z0 * b0, g1 = z1 * b1;
              // (NaN-aware: an inf * 0 product must not slip through a fmaxf chain)
              if (!(fabsf(f0) <= 65504.0f) || !(fabsf(f1) <= 65504.0f) || !(fabsf(g0) <= 65504.0f) || !(fabsf(g1) <= 65504.0f))
                xmax = INFINITY;
              gfw[k] = pack_f16x2(f0, f1);
              ggw[k] = pack_f16x2(g0, g1);
            }
            if (c0 + 16 < ce) ld_ab(c0 + 16);            // ahead of this chunk's stores
            if (in_range) {
              __half* gr = hp.xo16 + static_cast<long long>(it.b) * hp.x16_bs + static_cast<long long>(tau) * hp.x16_cp + c0;
              asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gr), "r"(gfw[0]), "r"(gfw[1]),
                           "r"(gfw[2]), "r"(gfw[3]), "r"(gfw[4]), "r"(gfw[5]), "r"(gfw[6]), "r"(gfw[7])
                           : "memory");
              asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gr + hp.dup_toff), "r"(ggw[0]),
                           "r"(ggw[1]), "r"(ggw[2]), "r"(ggw[3]), "r"(ggw[4]), "r"(ggw[5]), "r"(ggw[6]), "r"(ggw[7])
                           : "memory");
            }
          }
          if (cb >= ce) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
              else mbar_arrive(&tempty_bar[acc]);
            }
          }
        } else if ((KINDS & (1 << GF_RES)) && jd.kind == GF_RES) {
          // x_next = acc + x (wavenet.py:108).  fp32 through the staging half-tiles + TMA stores; fp16 channels-last copy
          // with 16-byte stores (lane = time row: 32 contiguous bytes per 16 channels); optional shifted fp32 duplicate.
          // The residual rows were pulled into L2 by the producer warp's bulk prefetch at the start of the tile.
          const int span = ((jd.n + 63) >> 6) << 4;          // columns per warp: 64 for n = 256, 32 for n = 112
          const int cb = h * span;
          const int ce = min(cb + span, jd.n);
          const int nv = jd.n_valid;
          // ONE element offset serves x32 / xo32 / dup (same strides); the pointers are rebuilt from the shared-memory copy
          // of the parameters where they are used.  (With four live 64-bit pointers ptxas spilled five of the first
          // sixteen prefetched residual values right after their loads: every spill store waits for its load, ~8 k
          // cycles of serial HBM latency per job in the phase clock.)
          const long long eoff = static_cast<long long>(it.b) * hp.x_bs + static_cast<long long>(jd.ch0) * hp.x_cs + tau;
          const bool nxt_res = jb + 1 < hp.n_jobs && hp.job[jb + 1].kind == GF_RES && !hp.job[jb + 1].a_ring && !jd.a_ring;
          if (buf_job != jb) res_first(jb);
          stamp_at(3);
          if (!mbar_wait_warp(&tfull_bar[acc], acc_phase, abort_flag)) { ok = false; break; }
          tc_fence_after();
          stamp_at(1);
          auto chunk = [&](int c0) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
            if (c0 + 16 >= ce) {      // last chunk of this warp: the accumulator is in registers, release the region now
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
                else mbar_arrive(&tempty_bar[acc]);
              }
            }
            if (c0 >= jd.split) {
              // the reduce-add columns of a merged job (data-gradient variant: g_cond += P^T [g_f; g_g])
              if ((slab0 + 32 > hp.skp_t_lo) && (slab0 < hp.t_hi)) {
                if (elect_one()) tma_store_wait_read();
                __syncwarp();
                const bool s_keep = tau >= hp.skp_t_lo && tau < hp.t_hi && tau >= hp.skp_zero_lo;
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[j * 32 + lane] = s_keep ? __uint_as_float(v[j]) * GF_OSC_S(smem)[0] : 0.0f;
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                  tma_reduce_add_3d(&p.skp_m, stg, slab0, c0 - jd.split, it.b);
                  tma_store_commit();
                }
              }
              return;
            }
            float r[16];
            const float osc = GF_OSC_S(smem)[0];   // 1.0f except in the scaled-fp16 data-gradient variant
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              r[j] = (keep && c0 + j < nv) ? fmaf(__uint_as_float(v[j]), osc, buf[j]) : 0.0f;
              xmax = fmaxf(xmax, fabsf(r[j]));
            }
            if (c0 == cb) stamp_at(4);             // first chunk: accumulator and residual values are in registers
            if (c0 + 16 < ce) res_load(jb, c0 + 16);      // ahead of this chunk's stores in the LSU queue
            else if (nxt_res) res_first(jb + 1);
            if (in_range && !(dbg & 2)) {      // x_next fp32: plain coalesced stores (lane = time step), like tanh / sigmoid above
              float* xo = hp.xo32 + eoff + static_cast<long long>(c0) * hp.x_cs;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < nv) __stcs(xo, r[j]);
                xo += hp.x_cs;
              }
            }
            const int dup_t = tau + hp.dup_toff;
            if (hp.dup && in_range && dup_t >= 0 && dup_t < hp.dup_t_hi) {
              float* dd = hp.dup + eoff + hp.dup_toff + static_cast<long long>(c0) * hp.x_cs;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < nv) __stcs(dd, r[j]);
                dd += hp.x_cs;
              }
            }
            if (in_range && hp.xo16 && !(dbg & 4)) {
              __half* x16row = hp.xo16 + static_cast<long long>(it.b) * hp.x16_bs + static_cast<long long>(tau) * hp.x16_cp + jd.ch0;
              if (c0 < nv) {
                // ONE 32-byte store per lane (st.global.v8: a full sector; two 16-byte stores are two half-filled
                // sectors on the SM -> L2 write port).  Channels in [nv, c0 + 16) are zeros and land in the padding.
                uint32_t hw[8];
                const float csc = GF_OSC_S(smem)[1];   // scale of the 16-bit copy (1.0f in the forward kernel)
#pragma unroll
                for (int k = 0; k < 8; ++k) hw[k] = pack_f16x2(r[2 * k] * csc, r[2 * k + 1] * csc);
                if (hp.prefetch & 8)     // keep the next layer's operand in L2 (experiment)
                  asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(x16row + c0),
                               "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]), "r"(hw[4]), "r"(hw[5]), "r"(hw[6]), "r"(hw[7])
                               : "memory");
                else
                  asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(x16row + c0), "r"(hw[0]),
                               "r"(hw[1]), "r"(hw[2]), "r"(hw[3]), "r"(hw[4]), "r"(hw[5]), "r"(hw[6]), "r"(hw[7])
                               : "memory");
              }
            }
          };
#pragma unroll 1
          for (int c0 = cb; c0 < ce; c0 += 16) chunk(c0);
          if (cb >= ce && nxt_res) res_first(jb + 1);
          if (cb >= ce) {             // a warp without columns in this job still owes its arrival
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
              else mbar_arrive(&tempty_bar[acc]);
            }
          }
        } else if (KINDS & (1 << GF_SKP)) {
          // skip sum (wavenet.py:104,110-111 + the caller's running sum): store (first layer), reduce-add in L2, or
          // relu(old + acc) for the last layer (wavenet.py:359)
          const int span = ((jd.n + 63) >> 6) << 4;
          const int cb = h * span;
          const int ce = min(cb + span, jd.n);
          const bool s_in = tau >= hp.skp_t_lo && tau < hp.t_hi;
          const bool s_keep = s_in && tau >= hp.skp_zero_lo;
          const float* old = hp.skp + static_cast<long long>(it.b) * hp.s_bs + static_cast<long long>(jd.ch0) * hp.s_cs + tau;
          if (jb + 1 < hp.n_jobs && hp.job[jb + 1].kind == GF_RES && !hp.job[jb + 1].a_ring) res_first(jb + 1);
          if (!mbar_wait_warp(&tfull_bar[acc], acc_phase, abort_flag)) { ok = false; break; }
          tc_fence_after();
          stamp_at(1);
#pragma unroll 1
          for (int c0 = cb; c0 < ce; c0 += 16) {
            uint32_t v[16];
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = 0.0f;
            if (hp.skp_mode == 2) {
              const float* sp = old + static_cast<long long>(c0) * hp.s_cs;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                o[j] = (s_keep && c0 + j < jd.n_valid) ? __ldcg(sp) : 0.0f;
                sp += hp.s_cs;
              }
            }
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
            if (c0 + 16 >= ce) {      // last chunk: release the region before the stores
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
                else mbar_arrive(&tempty_bar[acc]);
              }
            }
            if (hp.skp_mode == 1) {
              // running skip sum: TMA reduce-add of a {32 t, 16 ch} box (the read-modify-write happens in L2)
              if ((slab0 + 32 > hp.skp_t_lo) && (slab0 < hp.t_hi)) {
                if (elect_one()) tma_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[j * 32 + lane] = s_keep ? __uint_as_float(v[j]) : 0.0f;
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                  tma_reduce_add_3d(&p.skp_m, stg, slab0, jd.ch0 + c0, it.b);
                  tma_store_commit();
                }
              }
            } else if (s_in) {
              float* op = hp.skp + static_cast<long long>(it.b) * hp.s_bs + static_cast<long long>(jd.ch0 + c0) * hp.s_cs + tau;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < jd.n_valid) {
                  const float a = s_keep ? __uint_as_float(v[j]) : 0.0f;
                  *op = hp.skp_mode >= 2 ? (s_keep ? fmaxf(a + o[j], 0.0f) : 0.0f) : a;
                }
                op += hp.s_cs;
              }
            }
          }
          if (cb >= ce) {             // a warp without columns in this job still owes its arrival
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
              else mbar_arrive(&tempty_bar[acc]);
            }
          }
        }
        stamp_at(2);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    if (xmax * GF_OSC_S(smem)[1] > 65504.0f && hp.err && hp.xo16) atomicExch(hp.err, AEWN_ERR_RANGE);   // fp16 copy saturated
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0 && *abort_flag && hp.err) atomicExch(hp.err, AEWN_ERR_TIMEOUT);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (B, C, T) fp32  ->  (B, Tp, Cp) fp16 channels-last operand copy; channel `ones_ch` (if >= 0) is set to 1.0 (the bias
// rides on it), channels in [C, Cp) other than that are written as 0.  64 channels x 32 time steps per CTA through a
// padded shared-memory tile: reads coalesced along time, writes 128 contiguous bytes per time row.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cvt_f16_cl_kernel(const float* __restrict__ src, long long s_bs, long long s_cs,
                                                         __half* __restrict__ dst, long long d_bs, int Cp, int C, int T,
                                                         int ones_ch, int* err, const float* __restrict__ scale = nullptr) {
  __shared__ float tl[64][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 64;
  const int t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows of 32
  const float sc = scale ? __ldg(scale) : 1.0f;
  float m = 0.0f;
  for (int cc = ty; cc < 64; cc += 8) {
    const int c = c0 + cc, t = t0 + tx;
    float v = 0.0f;
    if (c < C && t < T) v = src[static_cast<long long>(b) * s_bs + static_cast<long long>(c) * s_cs + t] * sc;
    if (c == ones_ch) v = 1.0f;
    m = fmaxf(m, fabsf(v));
    tl[cc][tx] = v;
  }
  __syncthreads();
  // thread -> (time row, 8-channel group): 32 rows x 8 groups = 256 threads, one 16-byte store each
  const int tr = threadIdx.x >> 3, g = threadIdx.x & 7;
  const int t = t0 + tr;
  if (t < T && c0 + 8 * g < Cp) {
    uint4 h;
    h.x = pack_f16x2(tl[8 * g + 0][tr], tl[8 * g + 1][tr]);
    h.y = pack_f16x2(tl[8 * g + 2][tr], tl[8 * g + 3][tr]);
    h.z = pack_f16x2(tl[8 * g + 4][tr], tl[8 * g + 5][tr]);
    h.w = pack_f16x2(tl[8 * g + 6][tr], tl[8 * g + 7][tr]);
    *reinterpret_cast<uint4*>(dst + static_cast<long long>(b) * d_bs + static_cast<long long>(t) * Cp + c0 + 8 * g) = h;
  }
  if (m > 65504.0f && err) atomicExch(err, AEWN_ERR_RANGE);
}

// Weight repacking into fp16 operand matrices: the aewn_copy_block table of aewn_pack_blocks with `dst` counted in
// halves:  dst16[i*di + j] = fp16(src[i*si + j*sj]).
__global__ void pack_blocks_f16_kernel(const aewn_copy_block* __restrict__ blocks, int n_blocks) {
  for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
    const aewn_copy_block b = blocks[bi];
    __half* d = reinterpret_cast<__half*>(b.dst);
    const long long total = static_cast<long long>(b.ni) * b.nj;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
      const long long i = e / b.nj, j = e - i * b.nj;
      d[i * b.di + j] = __float2half_rn(b.src[i * b.si + j * b.sj]);
    }
  }
}

__global__ void pack_blocks_bf16_kernel(const aewn_copy_block* __restrict__ blocks, int n_blocks) {
  for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
    const aewn_copy_block b = blocks[bi];
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(b.dst);
    const long long total = static_cast<long long>(b.ni) * b.nj;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
      const long long i = e / b.nj, j = e - i * b.nj;
      d[i * b.di + j] = __float2bfloat16_rn(b.src[i * b.si + j * b.sj]);
    }
  }
}

int encode_f16_map(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                   const cuuint32_t* box, CUtensorMapL2promotion promo, const char* what, bool bf16) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  if (!fn) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if (!ptr || (reinterpret_cast<uintptr_t>(ptr) & 15u))
    return set_err(AEWN_ERR_INVALID, "grcc_fwd: %s pointer null or not 16-byte aligned", what);
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), dims,
                  strides_b, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled(%s) failed: CUresult %d", what, (int)r);
  return AEWN_OK;
}

}  // namespace aewn

using namespace aewn;

static int gf_launch(GfParams& p, int max_ctas, cudaStream_t stream, const char* what, int kinds = GF_KINDS_FWD) {
  const long long total = static_cast<long long>(p.hot.batch) * p.hot.n_tgroups;
  int clusters = (max_ctas > 0 ? max_ctas : sm_count()) / 2;
  if (clusters > total) clusters = static_cast<int>(total);
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(GF_THREADS);
  cfg.dynamicSmemBytes = GF_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // ring depth: 4 stages (default) or 3 (AEWN_GF_RING=3, A/B runs)
  static const int ring = []() { const char* e = getenv("AEWN_GF_RING"); return e && atoi(e) == 3 ? 3 : 4; }();
  using KernelFn = void (*)(GfParams);
  KernelFn fn = kinds == GF_KINDS_GZ      ? grcc_fwd_kernel<4, GF_KINDS_GZ>
                : kinds == GF_KINDS_DGRAD ? grcc_fwd_kernel<4, GF_KINDS_DGRAD>
                : ring == 3               ? grcc_fwd_kernel<3, GF_KINDS_FWD>
                                          : grcc_fwd_kernel<4, GF_KINDS_FWD>;
  cudaError_t ae = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM_BYTES);
  if (ae != cudaSuccess) return cuda_err(ae, what);
  cudaError_t le = cudaLaunchKernelEx(&cfg, fn, p);
  count_launch();
  if (le != cudaSuccess) return cuda_err(le, what);
  return cuda_err(cudaGetLastError(), what);
}

extern "C" int aewn_grcc_fwd(const aewn_grcc_fwd_desc* d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_err(AEWN_ERR_INVALID, "grcc_fwd: null descriptor");
  const int R = d->R, D = d->D, S = d->S, C1 = d->n_cond1;
  if (D != 128 && D != 256) return set_err(AEWN_ERR_INVALID, "grcc_fwd: n_dil must be 128 or 256 (got %d)", D);
  if (R < 8 || R > 1024 || (R & 7) || S < 32 || S > 1024 || (S & 31) || C1 < 1 || C1 > 1024)
    return set_err(AEWN_ERR_INVALID, "grcc_fwd: need R %% 8 == 0, S %% 32 == 0 (R=%d S=%d C+1=%d)", R, S, C1);
  if (d->batch <= 0 || d->t_hi <= d->t_lo || d->dil <= 0 || d->t_rows < d->t_hi)
    return set_err(AEWN_ERR_INVALID, "grcc_fwd: bad batch / time range");
  const int KR = (R + 63) & ~63, KC = (C1 + 63) & ~63;
  if (d->x16_cp != KR || d->c16_cp != KC || d->w1_k != 2 * KR + KC)
    return set_err(AEWN_ERR_INVALID, "grcc_fwd: operand pitches must be the 64-rounded channel counts (x %d/%d c %d/%d w1 %d/%d)",
                   d->x16_cp, KR, d->c16_cp, KC, d->w1_k, 2 * KR + KC);
  if (!d->x16 || !d->c16 || !d->w1h || !d->w2h || !d->skp || (!d->final_layer && (!d->x32 || !d->xo32 || !d->xo16)))
    return set_err(AEWN_ERR_INVALID, "grcc_fwd: null operand / output pointer");
  if ((d->x_cs & 3) || (d->x_bs & 3) || (d->a_cs & 3) || (d->a_bs & 3) || (d->s_cs & 3) || (d->s_bs & 3))
    return set_err(AEWN_ERR_INVALID, "grcc_fwd: fp32 strides must be multiples of 4 elements");
  if (d->save && (!d->th || (d->save != 2 && !d->sg))) return set_err(AEWN_ERR_INVALID, "grcc_fwd: save needs th (and sg)");

  GfParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)KR, (cuuint64_t)d->t_rows, (cuuint64_t)d->batch};
    cuuint64_t str[2] = {(cuuint64_t)KR * 2u, (cuuint64_t)d->x16_bs * 2u};
    cuuint32_t box[3] = {64u, 128u, 1u};
    if ((rc = encode_f16_map(&p.xa, d->x16, 3, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "x16"))) return rc;
    cuuint64_t dimc[3] = {(cuuint64_t)KC, (cuuint64_t)d->t_rows, (cuuint64_t)d->batch};
    cuuint64_t strc[2] = {(cuuint64_t)KC * 2u, (cuuint64_t)d->c16_bs * 2u};
    if ((rc = encode_f16_map(&p.ca, d->c16, 3, dimc, strc, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "c16"))) return rc;
    cuuint64_t dw1[2] = {(cuuint64_t)d->w1_k, (cuuint64_t)(2 * D)};
    cuuint64_t sw1[1] = {(cuuint64_t)d->w1_k * 2u};
    cuuint32_t bw[2] = {64u, 128u};
    if ((rc = encode_f16_map(&p.w1, d->w1h, 2, dw1, sw1, bw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w1h"))) return rc;
    const int rows2 = (d->final_layer ? 0 : R) + S;
    cuuint64_t dw2[2] = {(cuuint64_t)D, (cuuint64_t)rows2};
    cuuint64_t sw2[1] = {(cuuint64_t)D * 2u};
    if ((rc = encode_f16_map(&p.w2, d->w2h, 2, dw2, sw2, bw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w2h"))) return rc;
  }
  if (!d->final_layer) {
    if ((rc = encode_out_map(&p.xr_m, const_cast<float*>(d->x32), d->t_hi, R, d->batch, d->x_cs, d->x_bs, 32, 128))) return rc;
    // bit 0: residual rows by TMA prefetch at the top of the tile, bit 1: next tile's operands (both measured SLOWER:
    // 356 -> 379 / 377 / 394 us per layer), bit 2: residual rows by per-lane prefetch.global.L2 from the epilogue warps
    static const int pf = []() { const char* e = getenv("AEWN_GF_PREFETCH"); return e ? atoi(e) : 0; }();
    p.hot.prefetch = pf;
    static const int dbgf = []() { const char* e = getenv("AEWN_GF_DBG"); return e ? atoi(e) : 0; }();
    p.hot.dbg = dbgf;
    p.hot.n_res = R;
  }

  if ((rc = encode_out_map(&p.skp_m, d->skp, d->t_hi, S, d->batch, d->s_cs, d->s_bs, 16))) return rc;

  int nj = 0;
  for (int j = 0; j < D / 128; ++j) p.hot.job[nj++] = GfJob{GF_GATE, 256 * j, 256, 128, 128 * j, 1, 256};
  p.hot.n_gate = nj;
  // Order of the residual / skip jobs: their epilogues are long (HBM stores) and their MMAs short, and the next tile's
  // first gate job may start as soon as the region of the SECOND-to-last job has been drained -- so the job with the
  // longest epilogue (a full 256-channel residual chunk) goes last, where it overlaps the next tile's gate MMAs
  // (measured with the phase clock: 85 k -> cycles per tile with the residual chunks first).
  const int row_s = d->final_layer ? 0 : R;
  for (int c0 = 0; c0 < S; c0 += 256) {
    const int nv = S - c0 < 256 ? S - c0 : 256;
    p.hot.job[nj++] = GfJob{GF_SKP, row_s + c0, (nv + 15) & ~15, nv, c0, 0, 256};
  }
  if (!d->final_layer)
    for (int c0 = ((R - 1) / 256) * 256; c0 >= 0; c0 -= 256) {      // the partial chunk (if any) first
      const int nv = R - c0 < 256 ? R - c0 : 256;
      p.hot.job[nj++] = GfJob{GF_RES, c0, (nv + 15) & ~15, nv, c0, 0, 256};
    }
  if (nj > GF_MAX_JOBS) return set_err(AEWN_ERR_INVALID, "grcc_fwd: too many jobs (%d)", nj);
  p.hot.n_jobs = nj;
  p.hot.seg[0] = GfSeg{0, -d->dil, KR / 64};      // x(t - dil)
  p.hot.seg[1] = GfSeg{0, 0, KR / 64};            // x(t)
  p.hot.seg[2] = GfSeg{1, 0, KC / 64};            // cond(t), 1
  p.hot.n_segs = 3;
  p.hot.ring_stages = 2 * (KR / 64) + KC / 64;
  p.hot.kb_z = D / 64;
  p.hot.x32 = d->x32;
  p.hot.xo32 = d->xo32;
  p.hot.x_bs = d->x_bs;
  p.hot.x_cs = d->x_cs;
  p.hot.dup = d->dup;
  p.hot.dup_toff = d->dup_toff;
  p.hot.dup_t_hi = d->dup_t_hi;
  p.hot.xo16 = reinterpret_cast<__half*>(d->xo16);
  p.hot.x16_bs = d->x16_bs;
  p.hot.x16_cp = d->x16_cp;
  p.hot.skp = d->skp;
  p.hot.s_bs = d->s_bs;
  p.hot.s_cs = d->s_cs;
  p.hot.th = d->th;
  p.hot.sg = d->sg;
  p.hot.z = d->z;
  p.hot.a_bs = d->a_bs;
  p.hot.a_cs = d->a_cs;
  p.hot.save = d->save;
  p.hot.z_out = d->z != nullptr;
  if (d->z16) {
    if ((d->z16_cp & 15) || d->z16_cp < D || (d->z16_bs & 7) || (reinterpret_cast<uintptr_t>(d->z16) & 31u))
      return set_err(AEWN_ERR_INVALID, "grcc_fwd: z16 needs z16_cp %% 16 == 0, >= D, 32-byte aligned rows");
    p.hot.z16 = reinterpret_cast<__half*>(d->z16);
    p.hot.z16_bs = d->z16_bs;
    p.hot.z16_cp = d->z16_cp;
  }
  p.hot.skp_mode = d->skp_mode;
  p.hot.batch = d->batch;
  p.hot.t_begin = d->t_lo & ~31;
  p.hot.n_tgroups = (d->t_hi - p.hot.t_begin + 2 * GF_BM - 1) / (2 * GF_BM);
  p.hot.t_lo = d->t_lo;
  p.hot.t_zero_lo = d->t_zero_lo;
  p.hot.t_hi = d->t_hi;
  p.hot.skp_t_lo = d->skp_t_lo;
  p.hot.skp_zero_lo = d->skp_zero_lo;
  p.hot.err = d->err;
  p.hot.dbg_clock = d->dbg_clock;
  if ((p.hot.t_lo & 3) || (p.hot.skp_t_lo & 3)) return set_err(AEWN_ERR_INVALID, "grcc_fwd: t_lo / skp_t_lo must be multiples of 4");

  return gf_launch(p, d->max_ctas, stream, "grcc_fwd launch");
}

// ---------------------------------------------------------------------------------------------------------------
// Data gradient of a dilation layer on the same engine:
//   g_x[t] = tap1^T gfg[t] + tap0^T gfg[t + dil] (+ g_sig[t], t >= add_t_lo);     g_cond[t] += P^T gfg[t]
// gfg = [g_f; g_g] arrives as a bf16 channels-last copy (written by the gate-derivative launch), the transposed weights
// as a bf16 K-major matrix [R + C][2 K2] (tap0^T | tap1^T; the conditioning rows hold zeros under the shifted block).
// Jobs: 256-channel chunks of g_x; the last, partial chunk shares its accumulator with the conditioning columns
// (columns >= split are reduce-added into g_cond), so the operand tile is streamed twice instead of three times.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int aewn_grcc_dgrad(const aewn_grcc_dgrad_desc* d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_err(AEWN_ERR_INVALID, "grcc_dgrad: null descriptor");
  const int R = d->R, C = d->n_cond, K2 = d->g16_cp;
  if (R < 16 || R > 1024 || (R & 15) || C < 1 || C > 240 || K2 < 64 || (K2 & 63) || K2 > 1024)
    return set_err(AEWN_ERR_INVALID, "grcc_dgrad: need R %% 16 == 0, 1 <= C <= 240, g16_cp %% 64 == 0 (R=%d C=%d K2=%d)", R, C, K2);
  if (!d->g16 || !d->w1t16 || !d->gx || !d->g_cond || d->batch <= 0 || d->t_hi <= d->t_lo || d->t_rows < d->t_hi ||
      d->w_k != 2 * K2 || (d->t_lo & 3) || (d->cond_t_lo & 3) || (d->x_cs & 3) || (d->x_bs & 3) || (d->c_cs & 3) || (d->c_bs & 3))
    return set_err(AEWN_ERR_INVALID, "grcc_dgrad: bad pointer / range / stride");
  GfParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)K2, (cuuint64_t)d->t_rows, (cuuint64_t)d->batch};
    cuuint64_t str[2] = {(cuuint64_t)K2 * 2u, (cuuint64_t)d->g16_bs * 2u};
    cuuint32_t box[3] = {64u, 128u, 1u};
    const bool bf16 = d->g_inv_scale == nullptr;
    if ((rc = encode_f16_map(&p.xa, d->g16, 3, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "g16", bf16))) return rc;
    p.ca = p.xa;
    cuuint64_t dw[2] = {(cuuint64_t)d->w_k, (cuuint64_t)(R + C)};
    cuuint64_t sw[1] = {(cuuint64_t)d->w_k * 2u};
    cuuint32_t bw[2] = {64u, 128u};
    if ((rc = encode_f16_map(&p.w1, d->w1t16, 2, dw, sw, bw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w1t16", bf16))) return rc;
    p.w2 = p.w1;
  }
  if ((rc = encode_out_map(&p.skp_m, d->g_cond, d->t_hi, C, d->batch, d->c_cs, d->c_bs, 16))) return rc;
  const int Cp = (C + 15) & ~15;
  int nj = 0;
  bool cond_done = false;
  for (int c0 = 0; c0 < R; c0 += 256) {
    const int nv = R - c0 < 256 ? R - c0 : 256;
    if (nv + Cp <= 256) {      // the partial chunk takes the conditioning columns along (their W rows follow directly)
      p.hot.job[nj++] = GfJob{GF_RES, c0, nv + Cp, nv, c0, 1, nv};
      cond_done = true;
    } else {
      p.hot.job[nj++] = GfJob{GF_RES, c0, nv, nv, c0, 1, 256};
    }
  }
  if (!cond_done) p.hot.job[nj++] = GfJob{GF_RES, R, Cp, 0, 0, 1, 0};
  if (nj > GF_MAX_JOBS) return set_err(AEWN_ERR_INVALID, "grcc_dgrad: too many jobs (%d)", nj);
  p.hot.n_jobs = nj;
  p.hot.n_gate = 0;
  p.hot.seg[0] = GfSeg{0, d->dil, K2 / 64};       // tap0^T block: gfg[t + dil]
  p.hot.seg[1] = GfSeg{0, 0, K2 / 64};            // tap1^T block: gfg[t]
  p.hot.n_segs = 2;
  p.hot.ring_stages = 2 * (K2 / 64);
  p.hot.kb_z = 0;
  p.hot.ab_bf16 = d->g_inv_scale ? 0 : 1;
  p.hot.out_scale = d->g_inv_scale;
  if (d->gx16) {
    if (!d->g_inv_scale || (d->gx16_cp & 15) || d->gx16_cp < R || (d->gx16_bs & 7) || (reinterpret_cast<uintptr_t>(d->gx16) & 31u))
      return set_err(AEWN_ERR_INVALID, "grcc_dgrad: gx16 needs the scaled variant, gx16_cp %% 16 == 0, >= R, 32-byte aligned rows");
    p.hot.xo16 = reinterpret_cast<__half*>(d->gx16);
    p.hot.x16_bs = d->gx16_bs;
    p.hot.x16_cp = d->gx16_cp;
  }
  p.hot.add_t_lo = d->add_t_lo;
  p.hot.x32 = d->g_sig;
  p.hot.xo32 = d->gx;
  p.hot.x_bs = d->x_bs;
  p.hot.x_cs = d->x_cs;
  p.hot.skp = d->g_cond;
  p.hot.s_bs = d->c_bs;
  p.hot.s_cs = d->c_cs;
  p.hot.skp_mode = 1;
  p.hot.batch = d->batch;
  p.hot.t_begin = d->t_lo & ~31;
  p.hot.n_tgroups = (d->t_hi - p.hot.t_begin + 2 * GF_BM - 1) / (2 * GF_BM);
  p.hot.n_res = R;
  p.hot.t_lo = d->t_lo;
  p.hot.t_zero_lo = d->t_zero_lo;
  p.hot.t_hi = d->t_hi;
  p.hot.skp_t_lo = d->cond_t_lo;
  p.hot.skp_zero_lo = d->cond_zero_lo;
  p.hot.err = d->err;
  return gf_launch(p, d->max_ctas, stream, "grcc_dgrad launch", GF_KINDS_DGRAD);
}

extern "C" int aewn_pack_blocks_bf16(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!blocks_dev || n_blocks <= 0) return set_err(AEWN_ERR_INVALID, "pack_blocks_bf16: bad arguments");
  int grid = n_blocks < 148 * 8 ? n_blocks : 148 * 8;
  pack_blocks_bf16_kernel<<<grid, 256, 0, stream>>>(blocks_dev, n_blocks);
  count_launch();
  return cuda_err(cudaGetLastError(), "pack_blocks_bf16 launch");
}

extern "C" int aewn_cvt_f16_cl(const float* src, long long s_bs, long long s_cs, void* dst, long long d_bs, int Cp, int C,
                               int T, int batch, int ones_ch, int* err, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst || batch <= 0 || C <= 0 || T <= 0 || Cp < C || (Cp & 7) || (reinterpret_cast<uintptr_t>(dst) & 15u) ||
      (d_bs & 7))
    return set_err(AEWN_ERR_INVALID, "cvt_f16_cl: bad arguments");
  dim3 grid((T + 31) / 32, (Cp + 63) / 64, batch);
  cvt_f16_cl_kernel<<<grid, 256, 0, stream>>>(src, s_bs, s_cs, reinterpret_cast<__half*>(dst), d_bs, Cp, C, T, ones_ch, err);
  count_launch();
  return cuda_err(cudaGetLastError(), "cvt_f16_cl launch");
}

extern "C" int aewn_pack_blocks_f16(const aewn_copy_block* blocks_dev, int n_blocks, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!blocks_dev || n_blocks <= 0) return set_err(AEWN_ERR_INVALID, "pack_blocks_f16: bad arguments");
  int grid = n_blocks < 148 * 8 ? n_blocks : 148 * 8;
  pack_blocks_f16_kernel<<<grid, 256, 0, stream>>>(blocks_dev, n_blocks);
  count_launch();
  return cuda_err(cudaGetLastError(), "pack_blocks_f16 launch");
}

extern "C" int aewn_cvt_f16_cl_scaled(const float* src, long long s_bs, long long s_cs, void* dst, long long d_bs, int Cp, int C,
                                      int T, int batch, const float* scale, int* err, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst || !scale || batch <= 0 || C <= 0 || T <= 0 || Cp < C || (Cp & 7) || (reinterpret_cast<uintptr_t>(dst) & 15u) ||
      (d_bs & 7))
    return set_err(AEWN_ERR_INVALID, "cvt_f16_cl_scaled: bad arguments");
  dim3 grid((T + 31) / 32, (Cp + 63) / 64, batch);
  cvt_f16_cl_kernel<<<grid, 256, 0, stream>>>(src, s_bs, s_cs, reinterpret_cast<__half*>(dst), d_bs, Cp, C, T, -1, err, scale);
  count_launch();
  return cuda_err(cudaGetLastError(), "cvt_f16_cl_scaled launch");
}

// ---------------------------------------------------------------------------------------------------------------
// Gate derivative on the fused-layer engine: g_z = Wr^T g_x + Ws^T g_skp from the scaled fp16 channels-last copies of
// g_x and g_skp, then [g_f | g_g] = g_z {a, b} (saved fp16 words of the forward) stored as the scaled fp16 copy.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int aewn_grcc_gz(const aewn_grcc_gz_desc* d, aewn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_err(AEWN_ERR_INVALID, "grcc_gz: null descriptor");
  const int D = d->D, KR = d->gx16 ? d->gx16_cp : 0, KS = d->gs16_cp;
  if ((D != 128 && D != 256) || !d->gs16 || !d->w2t16 || !d->ab || !d->g16 || d->batch <= 0 || d->t_hi <= d->t_lo ||
      d->t_rows < d->t_hi || (KR & 63) || (KS & 63) || KS < 64 || d->w_k < KR + KS || (d->w_k & 7) || (d->g16_cp & 15) ||
      d->gg_off < D || d->gg_off + D > d->g16_cp || (d->gg_off & 15) || (reinterpret_cast<uintptr_t>(d->g16) & 31u) ||
      (d->g16_bs & 7))
    return set_err(AEWN_ERR_INVALID, "grcc_gz: bad pointer / range / stride (D=%d KR=%d KS=%d w_k=%d)", D, KR, KS, d->w_k);
  GfParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  cuuint32_t box[3] = {64u, 128u, 1u};
  if (d->gx16) {
    cuuint64_t dims[3] = {(cuuint64_t)KR, (cuuint64_t)d->t_rows, (cuuint64_t)d->batch};
    cuuint64_t str[2] = {(cuuint64_t)KR * 2u, (cuuint64_t)d->gx16_bs * 2u};
    if ((rc = encode_f16_map(&p.xa, d->gx16, 3, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "gx16"))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)KS, (cuuint64_t)d->t_rows, (cuuint64_t)d->batch};
    cuuint64_t str[2] = {(cuuint64_t)KS * 2u, (cuuint64_t)d->gs16_bs * 2u};
    if ((rc = encode_f16_map(&p.ca, d->gs16, 3, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "gs16"))) return rc;
    if (!d->gx16) p.xa = p.ca;
    // W2^T [D][KR | KS]: without the g_x segment the columns start at the skip block
    const __half* w = reinterpret_cast<const __half*>(d->w2t16) + (d->gx16 ? 0 : d->w_koff_skp);
    cuuint64_t dw[2] = {(cuuint64_t)(KR + KS), (cuuint64_t)D};
    cuuint64_t sw[1] = {(cuuint64_t)d->w_k * 2u};
    cuuint32_t bw[2] = {64u, 128u};
    if ((rc = encode_f16_map(&p.w1, w, 2, dw, sw, bw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w2t16"))) return rc;
    p.w2 = p.w1;
  }
  p.hot.job[0] = GfJob{GF_GZ, 0, D == 128 ? 128 : 256, D, 0, 1, 1 << 30};
  p.hot.n_jobs = 1;
  p.hot.n_gate = 0;
  int ns = 0;
  if (d->gx16) p.hot.seg[ns++] = GfSeg{0, 0, KR / 64};
  p.hot.seg[ns++] = GfSeg{1, 0, KS / 64};
  p.hot.n_segs = ns;
  p.hot.ring_stages = (KR + KS) / 64;
  p.hot.th = const_cast<float*>(reinterpret_cast<const float*>(d->ab));
  p.hot.a_bs = d->a_bs;
  p.hot.a_cs = d->a_cs;
  p.hot.xo16 = reinterpret_cast<__half*>(d->g16);
  p.hot.x16_bs = d->g16_bs;
  p.hot.x16_cp = d->g16_cp;
  p.hot.dup_toff = d->gg_off;                  // (reused field) channel offset of g_g inside a g16 row
  p.hot.batch = d->batch;
  p.hot.t_begin = d->t_lo & ~31;
  p.hot.n_tgroups = (d->t_hi - p.hot.t_begin + 2 * GF_BM - 1) / (2 * GF_BM);
  p.hot.t_lo = d->t_lo;
  p.hot.t_zero_lo = d->t_zero_lo;
  p.hot.t_hi = d->t_hi;
  p.hot.skp_t_lo = d->t_hi;
  p.hot.skp_zero_lo = d->t_hi;
  p.hot.err = d->err;
  return gf_launch(p, d->max_ctas, stream, "grcc_gz launch", GF_KINDS_GZ);
}
