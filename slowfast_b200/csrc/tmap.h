// Host-side construction of TMA tensor maps (tiled and im2col) through the driver entry points,
// resolved at run time so the library links without libcuda.
#pragma once
#include <cstdint>
#include <cuda.h>

namespace sfb {

// Error plumbing shared by the C-ABI: every entry point returns 0 on success or a negative code and leaves a
// message retrievable with sfb_last_error().
void set_error(const char* fmt, ...);
const char* last_error();

enum SwizzleBytes { SWZ_NONE = 0, SWZ_32 = 32, SWZ_64 = 64, SWZ_128 = 128 };

// 2-D row-major bf16 matrix [rows, cols] with row pitch `pitch_elems`; box = [box_rows, box_cols].
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows, uint32_t box_cols, SwizzleBytes swz);

// 5-D im2col map over a channels-last bf16 activation [N, D, H, W, C] (C contiguous, channel pitch
// `c_pitch` >= C elements).  lower/upper corners and traversal strides are given in (W, H, D) order, exactly as
// the driver consumes them.
int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, int n, int d, int h, int w, int c, int64_t c_pitch,
                          const int lower_whd[3], const int upper_whd[3], const int stride_whd[3],
                          uint32_t channels_per_pixel, uint32_t pixels_per_column, SwizzleBytes swz);

}  // namespace sfb
