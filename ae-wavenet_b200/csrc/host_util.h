// Host-side helpers shared by the C-ABI entry points: error reporting, launch counting, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/aewn.h"

namespace aewn {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int set_err(int code, const char* fmt, ...);
int cuda_err(cudaError_t e, const char* what);
int sm_count();

// 3-D fp32 activation map over (time, channel, batch); box = {32 time, box_rows channels, 1}.
// swizzle: CU_TENSOR_MAP_SWIZZLE_128B (K-major operand) or CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (MN-major tf32).
int encode_act_map(CUtensorMap* map, const aewn_act& a, int box_rows, CUtensorMapSwizzle swz);
// 3-D fp32 OUTPUT map over (time, channel, batch) for TMA stores / reduce-adds (and L2 prefetches); box = {box_t time,
// box_rows channels, 1}, no swizzle.
int encode_out_map(CUtensorMap* map, float* ptr, int t_extent, int channels, int batch, long long row_pitch,
                   long long batch_stride, int box_rows = 32, int box_t = 32);
// 16-bit (fp16 / bf16) map of rank 2 or 3, SWIZZLE_128B, inner box extent 64 elements = 128 bytes (csrc/grcc_fwd.cu)
int encode_f16_map(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                   const cuuint32_t* box, CUtensorMapL2promotion promo, const char* what, bool bf16 = false);
// 2-D K-major weight map [rows][kpad]; box = {32 k, box_rows}.
int encode_w_map(CUtensorMap* map, const float* w, int rows, int kpad, int box_rows);

inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace aewn
