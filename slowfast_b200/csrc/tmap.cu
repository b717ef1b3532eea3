#include "tmap.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <mutex>

namespace sfb {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_tiled = nullptr;
static EncodeIm2colFn g_im2col = nullptr;
static int g_driver_version = 0;
static std::once_flag g_once;

static void resolve() {
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_tiled = reinterpret_cast<EncodeTiledFn>(f);
  f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_im2col = reinterpret_cast<EncodeIm2colFn>(f);
  cudaDriverGetVersion(&g_driver_version);
}

static CUtensorMapSwizzle to_cu(SwizzleBytes s) {
  switch (s) {
    case SWZ_32: return CU_TENSOR_MAP_SWIZZLE_32B;
    case SWZ_64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case SWZ_128: return CU_TENSOR_MAP_SWIZZLE_128B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows, uint32_t box_cols, SwizzleBytes swz) {
  std::call_once(g_once, resolve);
  if (!g_tiled) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -1;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu(swz), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu pitch=%llu box=[%u,%u] swz=%d",
              (int)r, base, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_elems,
              box_rows, box_cols, (int)swz);
    return -2;
  }
  return 0;
}

int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, int n, int d, int h, int w, int c, int64_t c_pitch,
                          const int lower_whd[3], const int upper_whd[3], const int stride_whd[3],
                          uint32_t channels_per_pixel, uint32_t pixels_per_column, SwizzleBytes swz) {
  std::call_once(g_once, resolve);
  if (!g_im2col) {
    set_error("cuTensorMapEncodeIm2col entry point unavailable (no CUDA driver?)");
    return -1;
  }
  for (int i = 0; i < 3; ++i) {
    if (lower_whd[i] < -16 || lower_whd[i] > 15 || upper_whd[i] < -16 || upper_whd[i] > 15) {
      set_error("im2col corner out of the 5-D TMA range [-16,15]: lower=(%d,%d,%d) upper=(%d,%d,%d)", lower_whd[0],
                lower_whd[1], lower_whd[2], upper_whd[0], upper_whd[1], upper_whd[2]);
      return -3;
    }
    if (stride_whd[i] < 1 || stride_whd[i] > 8) {
      set_error("im2col traversal stride %d outside [1,8]", stride_whd[i]);
      return -3;
    }
  }
  cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
  cuuint64_t strides[4] = {(cuuint64_t)c_pitch * 2, (cuuint64_t)c_pitch * 2 * w, (cuuint64_t)c_pitch * 2 * w * h,
                           (cuuint64_t)c_pitch * 2 * w * h * d};
  cuuint32_t estr[5] = {1, (cuuint32_t)stride_whd[0], (cuuint32_t)stride_whd[1], (cuuint32_t)stride_whd[2], 1};
  CUresult r = g_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, lower_whd,
                        upper_whd, channels_per_pixel, pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        to_cu(swz), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error(
        "cuTensorMapEncodeIm2col failed (%d): base=%p ndhwc=[%d,%d,%d,%d,%d] pitch=%lld lower=(%d,%d,%d) "
        "upper=(%d,%d,%d) stride=(%d,%d,%d) cpp=%u ppc=%u swz=%d",
        (int)r, base, n, d, h, w, c, (long long)c_pitch, lower_whd[0], lower_whd[1], lower_whd[2], upper_whd[0],
        upper_whd[1], upper_whd[2], stride_whd[0], stride_whd[1], stride_whd[2], channels_per_pixel,
        pixels_per_column, (int)swz);
    return -2;
  }
  // Known driver defect (drivers reporting <= 13.1): for tensors smaller than 128 KiB the encoder sets a bit in
  // the second descriptor word that makes im2col loads fault; clearing it is the documented remedy used by the
  // vendor's own template library.
  if (g_driver_version <= 13010) {
    uint64_t bytes = (uint64_t)c_pitch * 2ull * w * h * d * n;
    if (bytes < 131072ull) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  }
  return 0;
}

}  // namespace sfb
