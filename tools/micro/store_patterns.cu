// Micro-benchmark: what one SM can push into L2/HBM with the fused layer kernel's epilogue access patterns.
// One CTA per SM (forced by a 200 KB dynamic shared-memory request, which also shrinks L1 like the real kernel),
// 16 warps per CTA, every warp walks "chunks" of [16 channels][32 time steps]:
//   F  fp32 rows      : 16 x st.global (lane = time step: one 128-byte line per instruction, channel stride = Tp floats)
//   H  fp16 ch-last   : 1 x st.global.v8 (lane = time row: 32 bytes per lane, row stride 768 bytes)
//   h  fp16 ch-last as 2 x 16-byte stores
//   L  fp32 row loads : 16 x ld.global (same shape as F, other buffer), consumed one chunk later
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/store_patterns tools/micro/store_patterns.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TP = 18464;        // padded time steps per batch (cfg2)
constexpr int CH = 384;

template <int MODE_F, int MODE_H, int MODE_L, int CS, int LDK = 0>
__global__ void __launch_bounds__(512, 1) k(float* __restrict__ xo, const float* __restrict__ xi, uint32_t* __restrict__ x16,
                                           int tiles_per_cta, int n_cta_tiles, long long* cyc) {
  extern __shared__ uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, h = warp >> 2;
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < tiles_per_cta; ++it) {
    const int tile = (blockIdx.x + it * gridDim.x) % n_cta_tiles;      // 128 time steps
    const int b = tile / (TP / 128), tt = tile % (TP / 128);
    const int tau = tt * 128 + q * 32 + lane;
    // the warp's 96 channels in 6 chunks of 16 (like RES1 + RES0 of one tile)
    float buf[3][16];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int j = 0; j < 16; ++j) buf[d][j] = 0.f;
    auto issue = [&](int c, float (&bb)[16]) {
      const int bl = LDK == 3 ? (b + 3) % 8 : b;
      const int taul = LDK == 3 ? (tau + 4736) % TP : tau;
      const float* sp = xi + ((long long)bl * CH + h * 96 + c * 16) * TP + taul;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (LDK == 1) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(bb[j]) : "l"(sp));
        else if (LDK == 2) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(bb[j]) : "l"(sp));
        else asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(bb[j]) : "l"(sp));
        sp += TP;
      }
    };
    // prologue: the first MODE_L chunks are in flight before the loop (distance = MODE_L chunks)
    if (MODE_L >= 1) issue(0, buf[0]);
    if (MODE_L >= 2) issue(1, buf[1]);
    if (MODE_L >= 3) issue(2, buf[2]);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const int c0 = h * 96 + c * 16;
      float r[16];
      float (&cur)[16] = buf[MODE_L ? c % MODE_L : 0];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = cur[j] + (float)(c0 + j);
      if (MODE_L && c + MODE_L < 6) issue(c + MODE_L, cur);
      if (MODE_F) {
        float* xp = xo + ((long long)b * CH + c0) * TP + tau;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (CS) __stcs(xp, r[j]); else *xp = r[j];
          xp += TP;
        }
      }
      if (MODE_H) {
        uint32_t* hp = x16 + ((long long)b * TP + tau) * (CH / 2) + c0 / 2;
        uint32_t w[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) w[kk] = __float_as_uint(r[2 * kk]) ^ __float_as_uint(r[2 * kk + 1]);
        if (MODE_H == 1)
          asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(hp), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
        else {
          __stcs(reinterpret_cast<uint4*>(hp), make_uint4(w[0], w[1], w[2], w[3]));
          __stcs(reinterpret_cast<uint4*>(hp) + 1, make_uint4(w[4], w[5], w[6], w[7]));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += buf[0][j] + buf[1][j] + buf[2][j];
  }
  if (acc == 12345.678f) sm[0] = 1;
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <int F, int H, int L, int CS, int LDK = 0>
void run(const char* name, int grid, float* xo, float* xi, uint32_t* x16, long long* cyc, int n_cta_tiles) {
  auto kern = k<F, H, L, CS, LDK>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int tiles = 8 * n_cta_tiles / 148 / 8;   // ~ one pass over the tensors with 148 CTAs
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    kern<<<grid, 512, 200 * 1024>>>(xo, xi, x16, tiles, n_cta_tiles, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc[148]; cudaMemcpy(hc, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
  const double st_bytes = (double)tiles * 128 * CH * (F * 4.0 + (H ? 2.0 : 0.0));
  const double ld_bytes = (double)tiles * 128 * CH * (L * 4.0);
  printf("%-34s grid %3d  %8.1f us  %9lld cyc  store %6.2f B/clk/SM  load %6.2f B/clk/SM  chip %6.2f TB/s  (%.0f cycles per tile)\n",
         name, grid, ms * 1e3, mx, st_bytes / mx, ld_bytes / mx, (st_bytes + ld_bytes) * grid / (ms * 1e-3) / 1e12, (double)mx / tiles);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
}

int main() {
  const int B = 8;
  const size_t n32 = (size_t)B * CH * TP;
  float *xo, *xi; uint32_t* x16; long long* cyc;
  cudaMalloc(&xo, n32 * 4); cudaMalloc(&xi, n32 * 4); cudaMalloc(&x16, n32 * 2); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(xi, 0, n32 * 4);
  const int n_cta_tiles = B * (TP / 128);
  for (int grid : {148, 37}) {
    run<1, 0, 0, 1>("F   fp32 rows .cs", grid, xo, xi, x16, cyc, n_cta_tiles);
    run<1, 0, 1, 1, 0>("F+L no_allocate", grid, xo, xi, x16, cyc, n_cta_tiles);
    run<1, 0, 1, 1, 1>("F+L ld.nc", grid, xo, xi, x16, cyc, n_cta_tiles);
    run<1, 0, 1, 1, 2>("F+L ld.cg", grid, xo, xi, x16, cyc, n_cta_tiles);
    run<1, 0, 1, 1, 3>("F+L loads from another tile", grid, xo, xi, x16, cyc, n_cta_tiles);
    run<1, 0, 1, 0, 0>("F+L default-policy stores", grid, xo, xi, x16, cyc, n_cta_tiles);
    run<1, 0, 1, 1, 0>("F+L in place (xi == xo)", grid, xo, xo, x16, cyc, n_cta_tiles);
  }
  return 0;
}
