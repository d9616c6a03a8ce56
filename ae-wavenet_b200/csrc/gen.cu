// Persistent incremental sampler (wavenet.py:367-531 WaveNet.forward_test) -- see include/aewn.h for the contract.
//
// One thread-block cluster generates for n_rep replicas.  Every CTA owns 1/cluster of the rows of every matrix of the
// decoder and streams them (fp32, consumption order, prepared once by the host) from L2 into a shared-memory ring with
// 1-D TMA bulk copies issued by a dedicated producer warp, so the ~50 MB of weights an arch.basic step touches flow at
// the aggregate L2 bandwidth of the cluster while the eight compute warps do the mat-vec work out of shared memory.
//
// The vectors that cross CTAs (z: D floats, x: R floats, h: S/P floats, logits: Q floats per replica) are exchanged
// all-to-all through distributed shared memory: every CTA writes its slice locally, then one thread per peer issues a
// shared->remote-shared bulk copy (cp.async.bulk.shared::cluster.shared::cta) that completes transaction bytes on the
// PEER's mbarrier -- data movement and synchronisation are one operation, there is no separate arrive and no scalar
// remote store (measured on B200: a scalar st.shared::cluster costs ~50 issue cycles; the first version of this kernel
// spent 7000 cycles per layer pushing 368 of them one by one).  Vector layouts are padded per CTA slice to 16 bytes
// (the bulk-copy granule); the host permutes weight columns to match.  Two barriers alternate so that a fast peer's
// next-phase bytes can never complete the current phase.  Per generated sample: 2*L + 3 exchanges, zero launches.
#include "host_util.h"
#include "ptx.cuh"

namespace aewn {
namespace {

constexpr int kGenComputeWarps = 8;
constexpr int kCT = kGenComputeWarps * 32;  // compute threads
constexpr int kGenThreads = kCT + 32;       // + producer warp
constexpr int kGateChunkRows = 8;           // rows per ring stage in a gate block (4 channel pairs)

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

// Shared-memory carve-up and slice geometry, identical on host (size query, validation) and device.
struct GenLayout {
  int stages, xbuf, zbuf, h0, h1, lg, sacc, part, codes, bars, total;
  int pairs, pairs_p, nres, nres_p, nskp, nskp_p, np1, np1_p, np2, np2_p;  // rows per CTA and their padded slices
  int Rp, Dp, Sp, Pp, Qp;  // padded vector lengths (cluster * slice)
  int Dz, Sz, Pz;          // vector strides incl. the bias slot
  int KA, n_gate_chunks, n_gate_groups;
};

__host__ __device__ inline GenLayout gen_layout(const aewn_gen_desc& p) {
  GenLayout L;
  const int nrep = p.n_rep, cl = p.cluster;
  L.pairs = p.D / cl, L.pairs_p = round4(L.pairs);
  L.nres = p.R / cl, L.nres_p = round4(L.nres);
  L.nskp = p.S / cl, L.nskp_p = round4(L.nskp);
  L.np1 = p.P / cl, L.np1_p = round4(L.np1);
  L.np2 = p.Q / cl, L.np2_p = round4(L.np2);
  L.Rp = cl * L.nres_p, L.Dp = cl * L.pairs_p, L.Sp = cl * L.nskp_p, L.Pp = cl * L.np1_p, L.Qp = cl * L.np2_p;
  L.Dz = L.Dp + 4, L.Sz = L.Sp + 4, L.Pz = L.Pp + 4;
  L.KA = 2 * L.Rp + p.cond_pitch;
  L.n_gate_chunks = (2 * L.pairs + kGateChunkRows - 1) / kGateChunkRows;
  const int group = 32 / (kGateChunkRows * nrep);  // chunks reduced together (32 partial sums per lane)
  L.n_gate_groups = (L.n_gate_chunks + group - 1) / group;
  int o = 0;
  auto take = [&](int bytes) {
    int r = o;
    o += (bytes + 127) & ~127;
    return r;
  };
  L.stages = take(p.n_stages * p.stage_bytes);
  L.xbuf = take(2 * nrep * L.Rp * 4);
  L.zbuf = take(nrep * L.Dz * 4);
  L.h0 = take(nrep * L.Sz * 4);
  L.h1 = take(nrep * L.Pz * 4);
  L.lg = take(nrep * L.Qp * 4);
  L.sacc = take(nrep * L.nskp * 4);
  L.part = take(kGenComputeWarps * L.n_gate_groups * 32 * 4);
  L.codes = take(64 + 4 * AEWN_GEN_MAX_LAYERS);  // codes[4], flags, per-layer ring slots
  L.bars = take((2 * p.n_stages + 4) * 8);
  L.total = o + 128;  // slack for aligning the dynamic base to 128 B
  return L;
}

// ------------------------------------------------------------------ small PTX helpers local to this kernel
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 16 bytes -> a peer CTA's shared memory, completing 16 transaction bytes on the PEER's mbarrier (SASS: STAS.128)
__device__ __forceinline__ void st_async_v4(uint32_t dst_remote, const float4& v, uint32_t bar_remote) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   dst_remote),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(bar_remote)
               : "memory");
}
// Explicit shared-space 128-bit load.  The ring / vector pointers travel through structs and selects, where the compiler
// loses the address space and falls back to generic LD.E.128 issued one at a time (measured: 2400 cycles for a 5-row
// mat-vec pass); volatile keeps the loads in program order after the mbarrier wait, but lets them be issued together.
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

// Warp reduction of V values per lane by recursive halving: 31 shuffles reduce 32 values (a butterfly per value would
// take 160).  Afterwards lane L holds in v[0] the warp total of value index  L >> (5 - log2 V).
template <int N, int S>
struct MultiReduce {
  template <int V>
  static __device__ __forceinline__ void run(float (&v)[V], int lane) {
    if constexpr (S >= 1) {
      if constexpr (N > 1) {
        constexpr int H = N / 2;
        const bool up = (lane & S) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float mine = up ? v[i + H] : v[i];
          const float other = up ? v[i] : v[i + H];
          v[i] = mine + __shfl_xor_sync(0xffffffffu, other, S);
        }
        MultiReduce<H, S / 2>::run(v, lane);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], S);
        MultiReduce<1, S / 2>::run(v, lane);
      }
    }
  }
};
template <int V>
__host__ __device__ constexpr int multi_shift() {
  int s = 5;
  for (int v = V; v > 1; v >>= 1) --s;
  return s;
}

__device__ __forceinline__ int chunk_cap(int kind, int rowf, int stage_bytes) {
  if (kind == 0) return kGateChunkRows;
  int c = stage_bytes / (rowf * 4);
  return c < 32 ? c : 32;
}

struct GenCtx {
  uint32_t stages_a;  // shared-space address of the weight ring
  uint64_t* full;
  uint64_t* empty;
  uint64_t* xbar;     // [2] exchange barriers (count 1 + transaction bytes)
  uint64_t* stepbar;  // arrive-counted barrier (count = cluster), once per step
  volatile int* abort;
  uint32_t st, ph;    // ring stage / phase of the next chunk to consume
  uint32_t xcnt;      // exchanges done
  uint32_t stepph;
  int stage_floats, n_stages;
  int cl, cl_log2, nrep;
  uint32_t rank;
};

__device__ __forceinline__ void bounded_wait_cluster(GenCtx& g, uint64_t* bar, uint32_t parity) {
  bool ok = false;
  for (uint32_t i = 0; i < kSpinLimit; ++i) {
    if (mbar_try_wait_cluster(bar, parity)) {
      ok = true;
      break;
    }
    if ((i & 255u) == 255u && *g.abort) break;
  }
  if (!ok) *g.abort = 1;
}

// All-to-all exchange of one vector.  Every compute thread has finished writing this CTA's slice (n_pad floats per
// replica, replica stride rep_stride) into its own copy of the vector; on return every CTA's slice is in place.
// Each 16-byte piece travels as one st.async: the store and its "16 bytes arrived" signal on the peer's mbarrier are
// a single instruction, so a peer wakes up as soon as the last piece of the last slice has landed.
__device__ __forceinline__ void exchange_begin(GenCtx& g, int tid, const float* slice0, int rep_stride, int n_pad,
                                               long long* dbg = nullptr) {
  if (dbg) dbg[0] = clock64();
  compute_bar();
  if (dbg) dbg[1] = clock64();
  uint64_t* bar = &g.xbar[g.xcnt & 1u];
  const int nq = n_pad >> 2;
  const int per_dst = g.nrep * nq;
  if (tid == 0) mbar_expect_tx(bar, static_cast<uint32_t>(g.cl - 1) * per_dst * 16u);
  const uint32_t bar_a = smem_u32(bar);
  // item i: destination CTA = i mod cluster (a power of two), piece j = i / cluster = (replica q, 16-byte piece k)
  for (int i = tid; i < (per_dst << g.cl_log2); i += kCT) {
    const uint32_t c = static_cast<uint32_t>(i) & static_cast<uint32_t>(g.cl - 1);
    const int j = i >> g.cl_log2;
    const int q = (j >= nq) + (j >= 2 * nq) + (j >= 3 * nq);  // n_rep <= 4
    const int k = j - q * nq;
    if (c != g.rank) {
      const float* src = slice0 + q * rep_stride + 4 * k;
      st_async_v4(mapa_u32(smem_u32(src), c), *reinterpret_cast<const float4*>(src), mapa_u32(bar_a, c));
    }
  }
  if (dbg) dbg[2] = clock64();
}
__device__ __forceinline__ void exchange_wait(GenCtx& g) {
  bounded_wait_cluster(g, &g.xbar[g.xcnt & 1u], (g.xcnt >> 1) & 1u);
  ++g.xcnt;
}

// Release/acquire barrier over the cluster (remote arrives): orders the GLOBAL-memory history-ring stores of one step
// before the ring loads of later steps.  Once per step; the per-phase exchanges above carry no release semantics for
// generic-proxy global stores.
__device__ __forceinline__ void step_barrier(GenCtx& g, int tid) {
  compute_bar();
  if (tid < g.cl) {
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    mbar_arrive_remote(mapa_u32(smem_u32(g.stepbar), static_cast<uint32_t>(tid)));
  }
  bounded_wait_cluster(g, g.stepbar, g.stepph);
  g.stepph ^= 1u;
}

// Returns false once the CTA is aborting (the caller skips the math but keeps the control flow).
__device__ __forceinline__ bool acquire_chunk(GenCtx& g, uint32_t& addr) {
  addr = g.stages_a + g.st * static_cast<uint32_t>(g.stage_floats) * 4u;
  return mbar_wait(&g.full[g.st], g.ph, g.abort);
}
__device__ __forceinline__ void advance_chunk(GenCtx& g) {
  if (++g.st == static_cast<uint32_t>(g.n_stages)) {
    g.st = 0;
    g.ph ^= 1u;
  }
}
__device__ __forceinline__ void release_chunk(GenCtx& g, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(&g.empty[g.st]);
  advance_chunk(g);
}

// A "mix"-type block: rows x rowf weights against vec[NREP][rowf].  Two ring stages are taken per pass and their rows
// dealt to the 8 warps in one round (up to 8 consecutive rows per warp, all partial sums in flight together, ONE
// multi-value reduction); lanes split the columns.  epi(row, rep, value) runs on exactly one lane per (row, rep).
template <int NREP, typename Epi>
__device__ __forceinline__ void matvec_block(GenCtx& g, const aewn_gen_block& blk, const float* vec, int cap, int warp,
                                             int lane, Epi epi, long long* dbg = nullptr) {
  constexpr int RB = 8;
  constexpr int V = RB * NREP;
  constexpr int sh = multi_shift<V>();
  const int rowq = blk.rowf >> 2;
  const uint32_t rowb = static_cast<uint32_t>(blk.rowf) * 4u;
  const uint32_t vec_a = smem_u32(vec);
  for (int r0 = 0; r0 < blk.rows; r0 += 2 * cap) {
    const int m = min(2 * cap, blk.rows - r0);  // rows of this pass: [0, cap) in stage A, [cap, m) in stage B
    if (dbg) *dbg++ = clock64();
    uint32_t spA, spB;
    bool ok = acquire_chunk(g, spA);
    const uint32_t stA = g.st;
    advance_chunk(g);
    uint32_t stB = stA;
    spB = spA;
    if (m > cap) {
      ok = acquire_chunk(g, spB) && ok;
      stB = g.st;
      advance_chunk(g);
    }
    if (dbg) *dbg++ = clock64();
    const int rbw = (m + kGenComputeWarps - 1) / kGenComputeWarps;  // rows per warp, <= RB by construction of cap
    const int rb = warp * rbw;
    const int nr = min(rbw, m - rb);  // rows of this warp (may be <= 0)
    if (ok && nr > 0) {
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
      uint32_t wrow[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = min(rb + r, m - 1);  // clamp: surplus slots re-read the last row, their sums are discarded
        wrow[r] = (row < cap ? spA + row * rowb : spB + (row - cap) * rowb) + lane * 16u;
      }
      for (int c = lane; c < rowq; c += 32) {
        float4 w[RB], zv[NREP];
#pragma unroll
        for (int q = 0; q < NREP; ++q) zv[q] = lds128(vec_a + q * rowb + c * 16u);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          w[r] = lds128(wrow[r]);
          wrow[r] += 512u;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
          for (int q = 0; q < NREP; ++q) acc[r * NREP + q] = dot4(w[r], zv[q], acc[r * NREP + q]);
      }
      MultiReduce<V, 16>::run(acc, lane);
      const int vi = lane >> sh;  // value index r * NREP + q held by this lane
      const int r = vi / NREP;
      if ((lane & ((1 << sh) - 1)) == 0 && r < nr) epi(r0 + rb + r, vi % NREP, acc[0]);
    }
    if (dbg) *dbg++ = clock64();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&g.empty[stA]);
      if (m > cap) mbar_arrive(&g.empty[stB]);
    }
    if (dbg) *dbg++ = clock64();
  }
}

template <int NREP, int NC>
__global__ void __launch_bounds__(kGenThreads, 1) gen_kernel(const __grid_constant__ aewn_gen_desc p) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~uintptr_t(127));
  const GenLayout L = gen_layout(p);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CL = p.cluster;
  const uint32_t rank = cluster_ctarank();
  const int group = blockIdx.x / CL;

  float* stages = reinterpret_cast<float*>(smem + L.stages);
  float* xbuf = reinterpret_cast<float*>(smem + L.xbuf);
  float* zbuf = reinterpret_cast<float*>(smem + L.zbuf);
  float* h0 = reinterpret_cast<float*>(smem + L.h0);
  float* h1 = reinterpret_cast<float*>(smem + L.h1);
  float* lg = reinterpret_cast<float*>(smem + L.lg);
  float* sacc = reinterpret_cast<float*>(smem + L.sacc);
  float* part = reinterpret_cast<float*>(smem + L.part);
  int* codes = reinterpret_cast<int*>(smem + L.codes);
  volatile int* abort_flag = codes + 8;
  volatile int* stop_flag = codes + 9;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + p.n_stages;
  uint64_t* xbar = empty + p.n_stages;  // [2]
  uint64_t* stepbar = xbar + 2;

  const int Rp = L.Rp, RQ = Rp >> 2;
  const int KAQ = L.KA >> 2;
  const int pairs = L.pairs, nres = L.nres, nskp = L.nskp, np1 = L.np1, np2 = L.np2;

  if (tid == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kGenComputeWarps);
    }
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    mbar_init(stepbar, CL);
    fence_barrier_init();
    *abort_flag = 0;
    *stop_flag = 0;
  }
  // constant parts of the exchanged vectors: bias slot = 1, slice padding = 0 (finite, its weight columns are zero)
  for (int i = tid; i < NREP * L.Dz; i += kGenThreads) zbuf[i] = (i % L.Dz == L.Dp) ? 1.f : 0.f;
  for (int i = tid; i < NREP * L.Sz; i += kGenThreads) h0[i] = (i % L.Sz == L.Sp) ? 1.f : 0.f;
  for (int i = tid; i < NREP * L.Pz; i += kGenThreads) h1[i] = (i % L.Pz == L.Pp) ? 1.f : 0.f;
  for (int i = tid; i < NREP * L.Qp; i += kGenThreads) lg[i] = 0.f;
  for (int i = tid; i < 2 * NREP * Rp; i += kGenThreads) xbuf[i] = 0.f;
  __syncthreads();
  cluster_sync_all();  // barriers initialised cluster-wide before any remote copy / arrive

  const float* ws = p.wstream + static_cast<size_t>(rank) * p.stream_stride;

  if (warp == kGenComputeWarps) {
    // ================================================================ producer: replay the static stream every step
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      bool ok = true;
      for (int t = p.t_begin; t < p.t_end && ok; ++t) {
        const bool samp = (t + 1 >= p.t_prime);
        for (int b = 0; b < p.n_blocks && ok; ++b) {
          const aewn_gen_block blk = p.blocks[b];
          if (blk.kind >= 2 && !samp) continue;
          const int cap = chunk_cap(blk.kind, blk.rowf, p.stage_bytes);
          for (int r0 = 0; r0 < blk.rows; r0 += cap) {
            const int m = min(cap, blk.rows - r0);
            const uint32_t bytes = static_cast<uint32_t>(m) * blk.rowf * 4u;
            if (!mbar_wait(&empty[st], ph ^ 1u, abort_flag)) {
              ok = false;
              break;
            }
            mbar_expect_tx(&full[st], bytes);
            bulk_load_1d(reinterpret_cast<uint8_t*>(stages) + static_cast<size_t>(st) * p.stage_bytes,
                         ws + blk.off + static_cast<size_t>(r0) * blk.rowf, bytes, &full[st]);
            if (++st == static_cast<uint32_t>(p.n_stages)) {
              st = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else {
    // ================================================================ compute warps
    GenCtx g;
    g.stages_a = smem_u32(stages);
    g.full = full;
    g.empty = empty;
    g.xbar = xbar;
    g.stepbar = stepbar;
    g.abort = abort_flag;
    g.st = 0;
    g.ph = 0;
    g.xcnt = 0;
    g.stepph = 0;
    g.stage_floats = p.stage_bytes >> 2;
    g.n_stages = p.n_stages;
    g.cl = CL;
    g.cl_log2 = 31 - __clz(CL);
    g.nrep = NREP;
    g.rank = rank;

    const size_t hist_slots = static_cast<size_t>(p.hist_off[p.n_layers]);
    float* hist_g = p.hist + static_cast<size_t>(group) * NREP * hist_slots * Rp;
    int* wav_g = p.wav + static_cast<size_t>(group) * NREP * p.wav_pitch;
    const float* uni_g = p.uniforms + static_cast<size_t>(group) * NREP * p.wav_pitch;

    float4 xv[NC][NREP];  // this thread's float4 column(s) of v = [x[t-d] | x[t] | cond[t], 1], per replica
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int q = 0; q < NREP; ++q) xv[c][q] = make_float4(0.f, 0.f, 0.f, 0.f);

    // rows per ring stage of the mat-vec blocks; 2 stages are dealt to 8 warps x <= 8 rows in one round
    const int cap_mix = chunk_cap(1, L.Dz, p.stage_bytes);
    const int cap_p1 = chunk_cap(2, L.Sz, p.stage_bytes);
    const int cap_p2 = chunk_cap(3, L.Pz, p.stage_bytes);

    // history-ring slot of step t per layer (t mod (d+1)), kept incrementally: no divisions inside the step loop
    int* hslot = codes + 16;  // [n_layers] ints in the codes/flags area
    for (int l = tid; l < p.n_layers; l += kCT) hslot[l] = p.t_begin % (p.dil[l] + 1);
    compute_bar();

    int cur = 0;
    constexpr int kGroup = 32 / (kGateChunkRows * NREP);  // gate chunks whose partial sums are reduced together

    // Abort protocol: a failed (bounded) wait raises *abort_flag and the math of the affected chunk is skipped, but the
    // control flow of the step -- every CTA barrier, every ring release -- still runs, so no thread is left behind at a
    // bar.sync; all compute threads leave together at the next step boundary.
    for (int t = p.t_begin; t < p.t_end; ++t) {
      const bool samp = (t + 1 >= p.t_prime);
      const bool stamp = p.dbg_clock && blockIdx.x == 0 && tid == 0 && t == p.t_begin + 8;
      int n_stamp = 0;
      auto mark = [&]() {
        if (stamp) p.dbg_clock[n_stamp++] = clock64();
      };
      mark();
      // ---------------------------------------------------------------- step prologue
      if (tid == 0) *stop_flag = *abort_flag;
      if (tid < NREP && (t < p.t_prime || t == p.t_begin)) {
        int c = __ldcg(wav_g + static_cast<size_t>(tid) * p.wav_pitch + t);
        if (c < 0 || c >= p.Q) {
          atomicCAS(p.err, 0, AEWN_ERR_INVALID);
          c = 0;
        }
        codes[tid] = c;
      }
      step_barrier(g, tid);  // also publishes codes / stop_flag CTA-wide
      if (*stop_flag) break;
      for (int i = tid; i < NREP * RQ; i += kCT) {
        const int q = i / RQ, c = i - q * RQ;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.base_t + static_cast<size_t>(codes[q]) * p.base_pitch) + c);
        reinterpret_cast<float4*>(xbuf + (cur * NREP + q) * Rp)[c] = v;
        if (rank == 0) {
          float* dst = hist_g + (static_cast<size_t>(q) * hist_slots + p.hist_off[0] + hslot[0]) * Rp;
          __stcg(reinterpret_cast<float4*>(dst) + c, v);
        }
      }
      for (int i = tid; i < NREP * nskp; i += kCT) sacc[i] = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = tid + c * kCT;
        if (col >= 2 * RQ && col < KAQ) {
          const float4 cv =
              __ldg(reinterpret_cast<const float4*>(p.cond + static_cast<size_t>(t) * p.cond_pitch) + (col - 2 * RQ));
#pragma unroll
          for (int q = 0; q < NREP; ++q) xv[c][q] = cv;
        }
      }
      auto prefetch_hist = [&](int l) {
        const int d = p.dil[l];
        const int slot = (hslot[l] == d) ? 0 : hslot[l] + 1;  // (t + 1) mod (d + 1) == (t - d) mod (d + 1)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int col = tid + c * kCT;
          if (col < RQ) {
#pragma unroll
            for (int q = 0; q < NREP; ++q) {
              const float* src = hist_g + (static_cast<size_t>(q) * hist_slots + p.hist_off[l] + slot) * Rp;
              xv[c][q] = __ldcg(reinterpret_cast<const float4*>(src) + col);
            }
          }
        }
      };
      prefetch_hist(0);
      compute_bar();
      mark();

      for (int l = 0; l < p.n_layers; ++l) {
        const aewn_gen_block blkA = p.blocks[2 * l];
        const aewn_gen_block blkB = p.blocks[2 * l + 1];
        const bool final_layer = (blkB.rows == nskp);
        // ------------------------------------------------------------ gate block: z = tanh(A_f v) * sigmoid(A_g v)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int col = tid + c * kCT;
          if (col >= RQ && col < 2 * RQ) {
#pragma unroll
            for (int q = 0; q < NREP; ++q)
              xv[c][q] = reinterpret_cast<const float4*>(xbuf + (cur * NREP + q) * Rp)[col - RQ];
          }
        }
        for (int g0 = 0, grp = 0; g0 < L.n_gate_chunks; g0 += kGroup, ++grp) {
          float acc[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
          for (int ci = 0; ci < kGroup; ++ci) {
            const int chunk = g0 + ci;
            if (chunk < L.n_gate_chunks) {
              const int m = min(kGateChunkRows, blkA.rows - chunk * kGateChunkRows);
              if (stamp && l == 3) p.dbg_clock[256 + 3 * chunk] = clock64();
              uint32_t sp;
              const bool ok = acquire_chunk(g, sp);
              if (stamp && l == 3) p.dbg_clock[256 + 3 * chunk + 1] = clock64();
              if (ok) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                  const int col = tid + c * kCT;
                  if (col < KAQ) {
                    float4 w[kGateChunkRows];
#pragma unroll
                    for (int r = 0; r < kGateChunkRows; ++r)   // rows >= m of a ragged last chunk re-read row m-1
                      w[r] = lds128(sp + (static_cast<uint32_t>(min(r, m - 1)) * KAQ + col) * 16u);
#pragma unroll
                    for (int r = 0; r < kGateChunkRows; ++r) {
                      if (r < m) {
#pragma unroll
                        for (int q = 0; q < NREP; ++q) {
                          const int ai = (ci * kGateChunkRows + r) * NREP + q;
                          acc[ai] = dot4(w[r], xv[c][q], acc[ai]);
                        }
                      }
                    }
                  }
                }
              }
              release_chunk(g, lane);
              if (stamp && l == 3) p.dbg_clock[256 + 3 * chunk + 2] = clock64();
            }
          }
          MultiReduce<32, 16>::run(acc, lane);
          part[(warp * L.n_gate_groups + grp) * 32 + lane] = acc[0];
        }
        mark();
        compute_bar();
        mark();
        for (int v = tid; v < pairs * NREP; v += kCT) {
          const int j = v / NREP, q = v - j * NREP;
          const int ch = (2 * j) / kGateChunkRows, r = (2 * j) % kGateChunkRows;
          const int grp = ch / kGroup, ci = ch - grp * kGroup;
          const int vi = (ci * kGateChunkRows + r) * NREP + q;
          float f = 0.f, gt = 0.f;
#pragma unroll
          for (int w = 0; w < kGenComputeWarps; ++w) {
            const float* pp = part + (w * L.n_gate_groups + grp) * 32;
            f += pp[vi];
            gt += pp[vi + NREP];
          }
          zbuf[q * L.Dz + rank * L.pairs_p + j] = tanhf(f) * (1.0f / (1.0f + expf(-gt)));
        }
        mark();
        exchange_begin(g, tid, zbuf + rank * L.pairs_p, L.Dz, L.pairs_p, (stamp && l == 3) ? p.dbg_clock + 320 : nullptr);
        if (l + 1 < p.n_layers) prefetch_hist(l + 1);
        if (stamp && l == 3) p.dbg_clock[323] = clock64();
        exchange_wait(g);
        if (stamp && l == 3) p.dbg_clock[324] = clock64();
        mark();
        // ------------------------------------------------------------ mix block: residual rows, then skip rows
        const int nres_l = final_layer ? 0 : nres;
        matvec_block<NREP>(g, blkB, zbuf, cap_mix, warp, lane, [&](int row, int q, float val) {
          if (row < nres_l) {
            const int grow = rank * L.nres_p + row;
            xbuf[((cur ^ 1) * NREP + q) * Rp + grow] = val + xbuf[(cur * NREP + q) * Rp + grow];
          } else {
            const int srow = row - nres_l;
            const float sv = sacc[q * nskp + srow] + val;
            sacc[q * nskp + srow] = sv;
            h0[q * L.Sz + rank * L.nskp_p + srow] = fmaxf(sv, 0.f);  // only the last layer's value is exchanged
          }
        }, (stamp && l == 3) ? p.dbg_clock + 300 : nullptr);
        mark();
        if (!final_layer) {
          exchange_begin(g, tid, xbuf + (cur ^ 1) * NREP * Rp + rank * L.nres_p, Rp, L.nres_p);
          // this CTA's slice of x_{l+1}[t] -> history ring (after the bar.sync inside exchange_begin)
          const size_t slot = static_cast<size_t>(p.hist_off[l + 1] + hslot[l + 1]);
          const int nq = L.nres_p >> 2;
          for (int i = tid; i < NREP * nq; i += kCT) {
            const int q = i / nq, c = i - q * nq;
            const float4 v = reinterpret_cast<const float4*>(xbuf + ((cur ^ 1) * NREP + q) * Rp + rank * L.nres_p)[c];
            __stcg(reinterpret_cast<float4*>(hist_g + (static_cast<size_t>(q) * hist_slots + slot) * Rp +
                                             rank * L.nres_p) + c, v);
          }
          exchange_wait(g);
          cur ^= 1;
        } else {
          exchange_begin(g, tid, h0 + rank * L.nskp_p, L.Sz, L.nskp_p);
          exchange_wait(g);
        }
        mark();
      }

      if (tid < p.n_layers) hslot[tid] = (hslot[tid] == p.dil[tid]) ? 0 : hslot[tid] + 1;  // read again after a bar.sync

      if (samp) {
        // ---------------------------------------------------------------- post-net, softmax, inverse-CDF draw
        const aewn_gen_block blk1 = p.blocks[2 * p.n_layers];
        const aewn_gen_block blk2 = p.blocks[2 * p.n_layers + 1];
        matvec_block<NREP>(g, blk1, h0, cap_p1, warp, lane, [&](int row, int q, float val) {
          h1[q * L.Pz + rank * L.np1_p + row] = fmaxf(val, 0.f);
        });
        exchange_begin(g, tid, h1 + rank * L.np1_p, L.Pz, L.np1_p);
        exchange_wait(g);
        matvec_block<NREP>(g, blk2, h1, cap_p2, warp, lane,
                           [&](int row, int q, float val) { lg[q * L.Qp + rank * L.np2_p + row] = val; });
        exchange_begin(g, tid, lg + rank * L.np2_p, L.Qp, L.np2_p);
        exchange_wait(g);
        if (warp < NREP && !*abort_flag) {
          // every CTA draws redundantly from identical logits, so the new code needs no further exchange
          const int q = warp;
          const float* lq = lg + q * L.Qp;
          const int per = (p.Q + 31) / 32;
          const int k0 = lane * per, k1 = min(p.Q, k0 + per);
          const int pad = L.np2_p - np2;
          const int base = (k0 / np2) * L.np2_p + (k0 % np2), left0 = np2 - (k0 % np2);
          // logit(k) for k = k0, k0+1, ...: walk the slice-padded layout without a division per element
          auto walk = [&](auto&& body) {
            int pos = base, left = left0;
            for (int k = k0; k < k1; ++k) {
              body(lq[pos]);
              ++pos;
              if (--left == 0) {
                pos += pad;
                left = np2;
              }
            }
          };
          float m = -INFINITY;
          walk([&](float v) { m = fmaxf(m, v); });
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
          float mine = 0.f;
          walk([&](float v) { mine += expf(v - m); });
          float incl = mine;
#pragma unroll
          for (int s = 1; s < 32; s <<= 1) {
            const float o = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += o;
          }
          const float total = __shfl_sync(0xffffffffu, incl, 31);
          const float u = __ldg(uni_g + static_cast<size_t>(q) * p.wav_pitch + t + 1);
          const float target = u * total;
          float run = incl - mine;
          int cnt = 0;
          walk([&](float v) {
            run += expf(v - m);
            cnt += (run <= target) ? 1 : 0;
          });
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
          const int idx = min(cnt, p.Q - 1);
          if (lane == 0) {
            codes[q] = idx;
            if (rank == 0) __stcg(wav_g + static_cast<size_t>(q) * p.wav_pitch + t + 1, idx);
          }
          if (rank == 0 && p.logits_out) {
            float* dst = p.logits_out + ((static_cast<size_t>(group) * NREP + q) * p.wav_pitch + t + 1) * p.Q;
            for (int k = lane; k < p.Q; k += 32) dst[k] = lq[(k / np2) * L.np2_p + (k % np2)];
          }
        }
      }
    }
    if (tid == 0 && *abort_flag) atomicCAS(p.err, 0, AEWN_ERR_TIMEOUT);
  }
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still copy into or arrive on its shared memory
}

typedef void (*GenKernel)(const aewn_gen_desc);

GenKernel pick_kernel(int nrep, int nc) {
  if (nc == 1) {
    if (nrep == 1) return gen_kernel<1, 1>;
    if (nrep == 2) return gen_kernel<2, 1>;
    if (nrep == 4) return gen_kernel<4, 1>;
  } else if (nc == 2) {
    if (nrep == 1) return gen_kernel<1, 2>;
    if (nrep == 2) return gen_kernel<2, 2>;
    if (nrep == 4) return gen_kernel<4, 2>;
  }
  return nullptr;
}

int validate(const aewn_gen_desc* d, int* nc_out) {
  if (!d) return set_err(AEWN_ERR_INVALID, "gen: null descriptor");
  const int cl = d->cluster;
  if (!(cl == 1 || cl == 2 || cl == 4 || cl == 8 || cl == 16))
    return set_err(AEWN_ERR_INVALID, "gen: cluster must be 1, 2, 4, 8 or 16 (got %d)", cl);
  if (!(d->n_rep == 1 || d->n_rep == 2 || d->n_rep == 4))
    return set_err(AEWN_ERR_INVALID, "gen: n_rep must be 1, 2 or 4 (got %d)", d->n_rep);
  if (d->n_layers < 1 || d->n_layers > AEWN_GEN_MAX_LAYERS || d->n_blocks != 2 * d->n_layers + 2)
    return set_err(AEWN_ERR_INVALID, "gen: n_layers %d / n_blocks %d out of range", d->n_layers, d->n_blocks);
  if (d->R <= 0 || d->D <= 0 || d->S <= 0 || d->P <= 0 || d->Q <= 0 || d->R % cl || d->D % cl || d->S % cl ||
      d->P % cl || d->Q % cl)
    return set_err(AEWN_ERR_INVALID, "gen: cluster %d must divide R=%d D=%d S=%d P=%d Q=%d", cl, d->R, d->D, d->S,
                   d->P, d->Q);
  if (d->Q > 1024) return set_err(AEWN_ERR_INVALID, "gen: Q=%d > 1024", d->Q);
  if (d->n_groups < 1) return set_err(AEWN_ERR_INVALID, "gen: n_groups must be positive");
  aewn_gen_desc probe = *d;
  probe.n_stages = probe.n_stages > 0 ? probe.n_stages : 2;
  const GenLayout L = gen_layout(probe);
  if ((d->cond_pitch & 3) || d->base_pitch != L.Rp)
    return set_err(AEWN_ERR_INVALID, "gen: cond_pitch %% 4 and base_pitch == cluster*round4(R/cluster) = %d required",
                   L.Rp);
  if (d->t_begin < 0 || d->t_end < d->t_begin || d->t_end >= d->wav_pitch || d->t_end > d->cond_len)
    return set_err(AEWN_ERR_INVALID, "gen: step range [%d, %d) outside wav_pitch %d / cond_len %d", d->t_begin,
                   d->t_end, d->wav_pitch, d->cond_len);
  const int ka = L.KA;
  const int nc = (ka / 4 + kCT - 1) / kCT;
  if (nc > 2) return set_err(AEWN_ERR_INVALID, "gen: gate row of %d floats exceeds %d", ka, 8 * kCT);
  if ((d->stage_bytes & 15) || d->stage_bytes < kGateChunkRows * ka * 4 || d->n_stages < 2 || d->n_stages > 32)
    return set_err(AEWN_ERR_INVALID, "gen: stage_bytes %d (need >= %d, %%16) / n_stages %d invalid", d->stage_bytes,
                   kGateChunkRows * ka * 4, d->n_stages);
  for (int l = 0; l < d->n_layers; ++l) {
    const aewn_gen_block& a = d->blocks[2 * l];
    const aewn_gen_block& b = d->blocks[2 * l + 1];
    if (d->dil[l] < 1 || d->hist_off[l + 1] - d->hist_off[l] != d->dil[l] + 1)
      return set_err(AEWN_ERR_INVALID, "gen: layer %d ring must have dil+1 slots", l);
    if (a.kind != 0 || a.rows != 2 * (d->D / cl) || a.rowf != ka)
      return set_err(AEWN_ERR_INVALID, "gen: layer %d gate block malformed", l);
    const bool fin = (b.rows == d->S / cl);
    if (b.kind != 1 || b.rowf != L.Dz || !(fin || b.rows == d->R / cl + d->S / cl))
      return set_err(AEWN_ERR_INVALID, "gen: layer %d mix block malformed", l);
    if (b.rowf * 4 > d->stage_bytes) return set_err(AEWN_ERR_INVALID, "gen: mix row larger than a stage");
  }
  const aewn_gen_block& p1 = d->blocks[2 * d->n_layers];
  const aewn_gen_block& p2 = d->blocks[2 * d->n_layers + 1];
  if (p1.kind != 2 || p1.rows != d->P / cl || p1.rowf != L.Sz || p2.kind != 3 || p2.rows != d->Q / cl ||
      p2.rowf != L.Pz || p1.rowf * 4 > d->stage_bytes || p2.rowf * 4 > d->stage_bytes)
    return set_err(AEWN_ERR_INVALID, "gen: post-net blocks malformed");
  if (!d->wstream || !d->cond || !d->base_t || !d->hist || !d->wav || !d->uniforms || !d->err)
    return set_err(AEWN_ERR_INVALID, "gen: null device pointer");
  if ((reinterpret_cast<uintptr_t>(d->wstream) & 15u) || (d->stream_stride & 3))
    return set_err(AEWN_ERR_INVALID, "gen: weight stream must be 16-byte aligned with stride %% 4 == 0");
  *nc_out = nc;
  return AEWN_OK;
}

int prepare(const aewn_gen_desc* d, GenKernel* k_out, int* smem_out) {
  int nc = 0;
  int rc = validate(d, &nc);
  if (rc) return rc;
  GenKernel k = pick_kernel(d->n_rep, nc);
  if (!k) return set_err(AEWN_ERR_INVALID, "gen: no kernel for n_rep=%d nc=%d", d->n_rep, nc);
  const GenLayout L = gen_layout(*d);
  if (L.total > 227 * 1024) return set_err(AEWN_ERR_INVALID, "gen: needs %d bytes of shared memory", L.total);
  rc = cuda_err(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total), "gen smem attr");
  if (rc) return rc;
  if (d->cluster > 8) {
    rc = cuda_err(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "gen cluster attr");
    if (rc) return rc;
  }
  *k_out = k;
  *smem_out = L.total;
  return AEWN_OK;
}

void fill_config(const aewn_gen_desc* d, int smem, cudaStream_t stream, cudaLaunchConfig_t* cfg,
                 cudaLaunchAttribute* attr) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(static_cast<unsigned>(d->n_groups * d->cluster), 1, 1);
  cfg->blockDim = dim3(kGenThreads, 1, 1);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = d->cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

}  // namespace
}  // namespace aewn

extern "C" {

int aewn_gen_smem_bytes(const aewn_gen_desc* d) {
  if (!d || d->n_stages < 1 || d->cluster < 1) return aewn::set_err(AEWN_ERR_INVALID, "gen: bad descriptor");
  return aewn::gen_layout(*d).total;
}

int aewn_gen_max_clusters(const aewn_gen_desc* d, int* n_out) {
  using namespace aewn;
  if (!n_out) return set_err(AEWN_ERR_INVALID, "gen: null output");
  GenKernel k;
  int smem;
  int rc = prepare(d, &k, &smem);
  if (rc) return rc;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(d, smem, nullptr, &cfg, attr);
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *n_out = n;
  return AEWN_OK;
}

int aewn_gen_run(const aewn_gen_desc* d, aewn_stream_t stream) {
  using namespace aewn;
  GenKernel k;
  int smem;
  int rc = prepare(d, &k, &smem);
  if (rc) return rc;
  if (d->t_end == d->t_begin) return AEWN_OK;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(d, smem, static_cast<cudaStream_t>(stream), &cfg, attr);
  count_launch();
  return cuda_err(cudaLaunchKernelEx(&cfg, k, *d), "gen_kernel launch");
}
}
