#include "host_util.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace aewn {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_err(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return AEWN_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -static_cast<int>(e);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Resolved through the runtime so the library has no link-time dependency on libcuda.so.
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// L2 promotion of activation boxes.  Every box row is one 128-byte piece of a channel row (rows are ~74 KB apart), and
// the next K block / time tile reads the adjacent 128 bytes.  Promoting misses to 256 bytes was measured equal within
// noise (gpurun ab1: 31.6 vs 31.7 ms/step), so 128 stays; AEWN_ACT_L2_PROMOTION=64|256 overrides for A/B runs.
static CUtensorMapL2promotion act_l2_promotion() {
  static const CUtensorMapL2promotion v = []() {
    const char* e = getenv("AEWN_ACT_L2_PROMOTION");
    if (e && atoi(e) == 256) return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (e && atoi(e) == 64) return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;   // 256 measured equal within noise (wgrad2 -10 %, wgrad1 +4 %)
  }();
  return v;
}

int encode_act_map(CUtensorMap* map, const aewn_act& a, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if (!a.ptr || (reinterpret_cast<uintptr_t>(a.ptr) & 15u))
    return set_err(AEWN_ERR_INVALID, "activation pointer null or not 16-byte aligned");
  if (a.t_extent <= 0 || a.channels <= 0 || a.batch <= 0)
    return set_err(AEWN_ERR_INVALID, "activation extents must be positive (t=%d c=%d b=%d)", a.t_extent,
                   a.channels, a.batch);
  if ((a.row_pitch & 3) || (a.batch_stride & 3) || a.row_pitch < a.t_extent)
    return set_err(AEWN_ERR_INVALID, "activation pitch/stride must be multiples of 4 elements (pitch=%lld bs=%lld)",
                   a.row_pitch, a.batch_stride);
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(a.t_extent), static_cast<cuuint64_t>(a.channels),
                        static_cast<cuuint64_t>(a.batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(a.row_pitch) * 4u, static_cast<cuuint64_t>(a.batch_stride) * 4u};
  if (a.batch == 1 && strides[1] < strides[0] * dims[1]) strides[1] = strides[0] * dims[1];
  cuuint32_t box[3] = {32u, static_cast<cuuint32_t>(box_rows), 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<float*>(a.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, act_l2_promotion(),
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled(act) failed: CUresult %d", (int)r);
  return AEWN_OK;
}

int encode_out_map(CUtensorMap* map, float* ptr, int t_extent, int channels, int batch, long long row_pitch,
                   long long batch_stride, int box_rows, int box_t) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(t_extent), static_cast<cuuint64_t>(channels),
                        static_cast<cuuint64_t>(batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_pitch) * 4u, static_cast<cuuint64_t>(batch_stride) * 4u};
  if (strides[1] < strides[0] * dims[1]) strides[1] = strides[0] * dims[1];
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_t), static_cast<cuuint32_t>(box_rows), 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled(out) failed: CUresult %d", (int)r);
  return AEWN_OK;
}

int encode_w_map(CUtensorMap* map, const float* w, int rows, int kpad, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if (!w || (reinterpret_cast<uintptr_t>(w) & 15u))
    return set_err(AEWN_ERR_INVALID, "weight pointer null or not 16-byte aligned");
  if (rows <= 0 || kpad <= 0 || (kpad & 31))
    return set_err(AEWN_ERR_INVALID, "weight matrix needs rows>0 and kpad a positive multiple of 32 (rows=%d kpad=%d)",
                   rows, kpad);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(kpad), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(kpad) * 4u};
  cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(AEWN_ERR_DRIVER, "cuTensorMapEncodeTiled(w) failed: CUresult %d", (int)r);
  return AEWN_OK;
}

}  // namespace aewn

extern "C" {
int aewn_version(void) { return 100; }
const char* aewn_last_error_string(void) { return aewn::g_err; }
long long aewn_launch_count(void) { return aewn::g_launches.load(); }
}
