// Persistent incremental sampler (wavenet.py:367-531 WaveNet.forward_test) -- see include/aewn.h for the contract.
//
// One thread-block cluster generates for n_rep replicas.  Every CTA owns 1/cluster of the rows of every matrix of the
// decoder and streams them (fp32, consumption order, prepared once by the host) from L2 into a shared-memory ring with
// 1-D TMA bulk copies issued by a dedicated producer warp, so the ~54 MB of weights an arch.basic step touches flow at
// the aggregate L2 bandwidth of the cluster while the eight compute warps do the mat-vec work out of shared memory.
// The vectors that cross CTAs (z: D floats, x: R floats, h: S/P floats, logits: Q floats per replica) are pushed with
// st.shared::cluster into every CTA and published by a cluster-scope mbarrier (one remote arrive per CTA pair), which
// costs a few hundred cycles where a grid-wide barrier through L2 would cost microseconds.  Per generated sample the
// kernel executes 2*L + 3 such barriers and no launch; the reference issues several hundred launches per sample.
#include "host_util.h"
#include "ptx.cuh"

namespace aewn {
namespace {

constexpr int kGenComputeWarps = 8;
constexpr int kCT = kGenComputeWarps * 32;  // compute threads
constexpr int kGenThreads = kCT + 32;       // + producer warp
constexpr int kGateChunkRows = 4;           // rows per ring stage in a gate block (2 channel pairs)

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

// Shared-memory carve-up, identical on host (size query) and device.
struct GenLayout {
  int stages, xbuf, zbuf, h0, h1, lg, sacc, part, codes, bars, total;
  int Rp, Dz, Sz, Pz, Qp, n_gate_chunks;
};

__host__ __device__ inline GenLayout gen_layout(const aewn_gen_desc& p) {
  GenLayout L;
  const int nrep = p.n_rep, cl = p.cluster;
  L.Rp = round4(p.R);
  L.Dz = round4(p.D) + 4;
  L.Sz = round4(p.S) + 4;
  L.Pz = round4(p.P) + 4;
  L.Qp = round4(p.Q);
  L.n_gate_chunks = (2 * (p.D / cl) + kGateChunkRows - 1) / kGateChunkRows;
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
  L.sacc = take(nrep * (p.S / cl) * 4);
  L.part = take(kGenComputeWarps * L.n_gate_chunks * kGateChunkRows * nrep * 4);
  L.codes = take(64);
  L.bars = take((2 * p.n_stages + 2) * 8);
  L.total = o + 128;  // slack for aligning the dynamic base to 128 B
  return L;
}

// ------------------------------------------------------------------ small PTX helpers local to this kernel
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
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
  return c < 64 ? c : 64;
}

struct GenCtx {
  float* stages;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* clbar;
  volatile int* abort;
  uint32_t st, ph;  // ring stage / phase of the next chunk to consume
  uint32_t clph;    // cluster barrier phase
  int stage_floats, n_stages;
  int cl;
};

// All compute threads: publish this CTA's remote stores and wait for every CTA of the cluster to do the same.
__device__ __forceinline__ void cluster_exchange(GenCtx& g, int tid) {
  compute_bar();
  if (tid < g.cl) {
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    mbar_arrive_remote(mapa_u32(smem_u32(g.clbar), static_cast<uint32_t>(tid)));
  }
  bool ok = false;
  for (uint32_t i = 0; i < kSpinLimit; ++i) {
    if (mbar_try_wait_cluster(g.clbar, g.clph)) {
      ok = true;
      break;
    }
    if ((i & 255u) == 255u && *g.abort) break;
  }
  if (!ok) *g.abort = 1;
  g.clph ^= 1u;
}

__device__ __forceinline__ const float* acquire_chunk(GenCtx& g) {
  if (!mbar_wait(&g.full[g.st], g.ph, g.abort)) return nullptr;
  return g.stages + static_cast<size_t>(g.st) * g.stage_floats;
}
__device__ __forceinline__ void release_chunk(GenCtx& g, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(&g.empty[g.st]);
  if (++g.st == static_cast<uint32_t>(g.n_stages)) {
    g.st = 0;
    g.ph ^= 1u;
  }
}

// A "mix"-type block: rows x rowf weights against vec[NREP][rowf]; warp w takes rows w, w+8, ...; lanes split the
// columns.  epi(row, rep, value) runs on exactly one lane per (row, rep).
template <int NREP, typename Epi>
__device__ __forceinline__ void matvec_block(GenCtx& g, const aewn_gen_block& blk, const float* vec, int stage_bytes,
                                             int warp, int lane, Epi epi) {
  const int rowq = blk.rowf >> 2;
  const int cap = chunk_cap(blk.kind, blk.rowf, stage_bytes);
  constexpr int sh = multi_shift<NREP>();
  for (int r0 = 0; r0 < blk.rows; r0 += cap) {
    const int m = min(cap, blk.rows - r0);
    const float* sp = acquire_chunk(g);  // nullptr once the CTA is aborting: skip the math, keep the control flow
    const float4* w4 = reinterpret_cast<const float4*>(sp);
    for (int r = warp; sp && r < m; r += kGenComputeWarps) {
      float acc[NREP];
#pragma unroll
      for (int q = 0; q < NREP; ++q) acc[q] = 0.f;
      for (int c = lane; c < rowq; c += 32) {
        const float4 w = w4[r * rowq + c];
#pragma unroll
        for (int q = 0; q < NREP; ++q) acc[q] = dot4(w, reinterpret_cast<const float4*>(vec + q * blk.rowf)[c], acc[q]);
      }
      MultiReduce<NREP, 16>::run(acc, lane);
      if ((lane & ((1 << sh) - 1)) == 0) epi(r0 + r, lane >> sh, acc[0]);
    }
    release_chunk(g, lane);
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
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + p.n_stages;
  uint64_t* clbar = empty + p.n_stages;

  const int Rp = L.Rp, RQ = Rp >> 2;
  const int KA = 2 * Rp + p.cond_pitch, KAQ = KA >> 2;
  const int pairs = p.D / CL, nres = p.R / CL, nskp = p.S / CL, np1 = p.P / CL, np2 = p.Q / CL;

  if (tid == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kGenComputeWarps);
    }
    mbar_init(clbar, CL);
    fence_barrier_init();
    *abort_flag = 0;
  }
  // constant parts of the exchanged vectors: bias column = 1, padding = 0
  for (int i = tid; i < NREP * L.Dz; i += kGenThreads) zbuf[i] = (i % L.Dz == round4(p.D)) ? 1.f : 0.f;
  for (int i = tid; i < NREP * L.Sz; i += kGenThreads) h0[i] = (i % L.Sz == round4(p.S)) ? 1.f : 0.f;
  for (int i = tid; i < NREP * L.Pz; i += kGenThreads) h1[i] = (i % L.Pz == round4(p.P)) ? 1.f : 0.f;
  for (int i = tid; i < 2 * NREP * Rp; i += kGenThreads) xbuf[i] = 0.f;
  __syncthreads();
  cluster_sync_all();  // barriers initialised cluster-wide before any remote arrive / store

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
    g.stages = stages;
    g.full = full;
    g.empty = empty;
    g.clbar = clbar;
    g.abort = abort_flag;
    g.st = 0;
    g.ph = 0;
    g.clph = 0;
    g.stage_floats = p.stage_bytes >> 2;
    g.n_stages = p.n_stages;
    g.cl = CL;

    const size_t hist_slots = static_cast<size_t>(p.hist_off[p.n_layers]);
    float* hist_g = p.hist + static_cast<size_t>(group) * NREP * hist_slots * Rp;
    int* wav_g = p.wav + static_cast<size_t>(group) * NREP * p.wav_pitch;
    const float* uni_g = p.uniforms + static_cast<size_t>(group) * NREP * p.wav_pitch;

    float4 xv[NC][NREP];  // this thread's float4 column(s) of v = [x[t-d] | x[t] | cond[t], 1], per replica
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int q = 0; q < NREP; ++q) xv[c][q] = make_float4(0.f, 0.f, 0.f, 0.f);

    int cur = 0;
    constexpr int shA = multi_shift<kGateChunkRows * NREP>();
    volatile int* stop_flag = codes + 9;

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
      compute_bar();
      if (*stop_flag) break;
      for (int i = tid; i < NREP * RQ; i += kCT) {
        const int q = i / RQ, c = i - q * RQ;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.base_t + static_cast<size_t>(codes[q]) * p.base_pitch) + c);
        reinterpret_cast<float4*>(xbuf + (cur * NREP + q) * Rp)[c] = v;
        if (rank == 0) {
          const int d0 = p.dil[0];
          float* dst = hist_g + (static_cast<size_t>(q) * hist_slots + p.hist_off[0] + (t % (d0 + 1))) * Rp;
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
        const int slot = (t + 1) % (d + 1);  // == (t - d) mod (d + 1)
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
        int chunk = 0;
        for (int r0 = 0; r0 < blkA.rows; r0 += kGateChunkRows, ++chunk) {
          const int m = min(kGateChunkRows, blkA.rows - r0);
          const float* sp = acquire_chunk(g);
          const float4* w4 = reinterpret_cast<const float4*>(sp);
          float acc[kGateChunkRows * NREP];
#pragma unroll
          for (int i = 0; i < kGateChunkRows * NREP; ++i) acc[i] = 0.f;
          if (sp) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const int col = tid + c * kCT;
              if (col < KAQ) {
#pragma unroll
                for (int r = 0; r < kGateChunkRows; ++r) {
                  if (r < m) {
                    const float4 w = w4[r * KAQ + col];
#pragma unroll
                    for (int q = 0; q < NREP; ++q) acc[r * NREP + q] = dot4(w, xv[c][q], acc[r * NREP + q]);
                  }
                }
              }
            }
          }
          release_chunk(g, lane);
          MultiReduce<kGateChunkRows * NREP, 16>::run(acc, lane);
          if ((lane & ((1 << shA) - 1)) == 0)
            part[(warp * L.n_gate_chunks + chunk) * (kGateChunkRows * NREP) + (lane >> shA)] = acc[0];
        }
        mark();
        compute_bar();
        mark();
        for (int v = tid; v < pairs * NREP; v += kCT) {
          const int j = v / NREP, q = v - j * NREP;
          const int ch = (2 * j) / kGateChunkRows, r = (2 * j) % kGateChunkRows;
          float f = 0.f, gt = 0.f;
#pragma unroll
          for (int w = 0; w < kGenComputeWarps; ++w) {
            const float* pp = part + (w * L.n_gate_chunks + ch) * (kGateChunkRows * NREP);
            f += pp[r * NREP + q];
            gt += pp[(r + 1) * NREP + q];
          }
          const float z = tanhf(f) * (1.0f / (1.0f + expf(-gt)));
          const uint32_t la = smem_u32(zbuf + q * L.Dz + rank * pairs + j);
          for (int c = 0; c < CL; ++c) st_cluster_f32(mapa_u32(la, c), z);
        }
        mark();
        cluster_exchange(g, tid);
        mark();
        // ------------------------------------------------------------ mix block: residual rows, then skip rows
        if (l + 1 < p.n_layers) prefetch_hist(l + 1);
        const int nres_l = final_layer ? 0 : nres;
        const int dn = (l + 1 < p.n_layers) ? p.dil[l + 1] : 0;
        const size_t ring_next = (l + 1 < p.n_layers) ? static_cast<size_t>(p.hist_off[l + 1] + (t % (dn + 1))) : 0;
        const bool push_h0 = samp && (l + 1 == p.n_layers);
        matvec_block<NREP>(g, blkB, zbuf, p.stage_bytes, warp, lane, [&](int row, int q, float val) {
          if (row < nres_l) {
            const int grow = rank * nres + row;
            const float xn = val + xbuf[(cur * NREP + q) * Rp + grow];
            const uint32_t la = smem_u32(xbuf + ((cur ^ 1) * NREP + q) * Rp + grow);
            for (int c = 0; c < CL; ++c) st_cluster_f32(mapa_u32(la, c), xn);
            __stcg(hist_g + (static_cast<size_t>(q) * hist_slots + ring_next) * Rp + grow, xn);
          } else {
            const int srow = row - nres_l;
            const float sv = sacc[q * nskp + srow] + val;
            sacc[q * nskp + srow] = sv;
            if (push_h0) {
              const uint32_t la = smem_u32(h0 + q * L.Sz + rank * nskp + srow);
              const float h = fmaxf(sv, 0.f);
              for (int c = 0; c < CL; ++c) st_cluster_f32(mapa_u32(la, c), h);
            }
          }
        });
        mark();
        cluster_exchange(g, tid);
        mark();
        if (!final_layer) cur ^= 1;
      }

      if (samp) {
        // ---------------------------------------------------------------- post-net, softmax, inverse-CDF draw
        const aewn_gen_block blk1 = p.blocks[2 * p.n_layers];
        const aewn_gen_block blk2 = p.blocks[2 * p.n_layers + 1];
        matvec_block<NREP>(g, blk1, h0, p.stage_bytes, warp, lane, [&](int row, int q, float val) {
          const uint32_t la = smem_u32(h1 + q * L.Pz + rank * np1 + row);
          const float h = fmaxf(val, 0.f);
          for (int c = 0; c < CL; ++c) st_cluster_f32(mapa_u32(la, c), h);
        });
        cluster_exchange(g, tid);
        matvec_block<NREP>(g, blk2, h1, p.stage_bytes, warp, lane, [&](int row, int q, float val) {
          const uint32_t la = smem_u32(lg + q * L.Qp + rank * np2 + row);
          for (int c = 0; c < CL; ++c) st_cluster_f32(mapa_u32(la, c), val);
        });
        cluster_exchange(g, tid);
        if (warp < NREP && !*abort_flag) {
          // every CTA draws redundantly from identical logits, so the new code needs no further exchange
          const int q = warp;
          const float* lq = lg + q * L.Qp;
          const int per = (p.Q + 31) / 32;
          const int k0 = lane * per, k1 = min(p.Q, k0 + per);
          float m = -INFINITY;
          for (int k = k0; k < k1; ++k) m = fmaxf(m, lq[k]);
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
          float mine = 0.f;
          for (int k = k0; k < k1; ++k) mine += expf(lq[k] - m);
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
          for (int k = k0; k < k1; ++k) {
            run += expf(lq[k] - m);
            cnt += (run <= target) ? 1 : 0;
          }
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
          const int idx = min(cnt, p.Q - 1);
          if (lane == 0) {
            codes[q] = idx;
            if (rank == 0) __stcg(wav_g + static_cast<size_t>(q) * p.wav_pitch + t + 1, idx);
          }
          if (rank == 0 && p.logits_out) {
            float* dst = p.logits_out + ((static_cast<size_t>(group) * NREP + q) * p.wav_pitch + t + 1) * p.Q;
            for (int k = lane; k < p.Q; k += 32) dst[k] = lq[k];
          }
        }
      }
    }
    if (tid == 0 && *abort_flag) atomicCAS(p.err, 0, AEWN_ERR_TIMEOUT);
  }
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still store into or arrive on its shared memory
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
  if ((d->cond_pitch & 3) || d->base_pitch != round4(d->R))
    return set_err(AEWN_ERR_INVALID, "gen: cond_pitch %% 4 and base_pitch == round4(R) required");
  if (d->t_begin < 0 || d->t_end < d->t_begin || d->t_end >= d->wav_pitch || d->t_end > d->cond_len)
    return set_err(AEWN_ERR_INVALID, "gen: step range [%d, %d) outside wav_pitch %d / cond_len %d", d->t_begin,
                   d->t_end, d->wav_pitch, d->cond_len);
  const int ka = 2 * round4(d->R) + d->cond_pitch;
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
    if (b.kind != 1 || b.rowf != round4(d->D) + 4 || !(fin || b.rows == d->R / cl + d->S / cl))
      return set_err(AEWN_ERR_INVALID, "gen: layer %d mix block malformed", l);
    if (b.rowf * 4 > d->stage_bytes) return set_err(AEWN_ERR_INVALID, "gen: mix row larger than a stage");
  }
  const aewn_gen_block& p1 = d->blocks[2 * d->n_layers];
  const aewn_gen_block& p2 = d->blocks[2 * d->n_layers + 1];
  if (p1.kind != 2 || p1.rows != d->P / cl || p1.rowf != round4(d->S) + 4 || p2.kind != 3 || p2.rows != d->Q / cl ||
      p2.rowf != round4(d->P) + 4 || p1.rowf * 4 > d->stage_bytes || p2.rowf * 4 > d->stage_bytes)
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
