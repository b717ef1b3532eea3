/*
 * slowfast_b200 — C ABI of the B200-native video-backbone engine.
 *
 * Drop-in boundary: the reference (facebookresearch/SlowFast) has no native layer; every FLOP of
 * `model(inputs)` / `loss.backward()` behind `slowfast.models.build_model` (slowfast/models/build.py:22)
 * is an ATen operator call.  This header declares what a reference-side binding would call instead of
 * those operator call sites.  All entry points:
 *   - take raw device pointers + plain-old-data descriptors (no torch types),
 *   - enqueue work on the CUDA stream passed as `void* stream` (a cudaStream_t) and never synchronise,
 *   - borrow every pointer for the duration of the call only (owner = the caller's allocator),
 *   - return 0 on success, <0 on error; sfb_last_error() returns the message (thread-local).
 *
 * Activation layout: channels-last NDHWC ("N, T, H, W, C").  Tensor-core operands are carried as TWO bf16
 * planes (hi = bf16(x), lo = bf16(x - hi)) so that the 3-term split product hi*hi + hi*lo + lo*hi reproduces
 * fp32 operand precision (parity mode, nsplit = 3); the fast mode uses the hi plane only (nsplit = 1).
 */
#ifndef SLOWFAST_B200_H_
#define SLOWFAST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* sfb_last_error(void);
/* Library/ABI version and build arch string, e.g. "sm_100a". */
int sfb_abi_version(void);
const char* sfb_build_arch(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM 3-D convolution on tcgen05 tensor cores, TMA im2col operand staging.
 * Replaces nn.Conv3d fprop (resnet_helper.py:332,346,362,485; stem_helper.py:182;
 * video_model_builder.py:147) and, with a transposed/flipped filter matrix and an output view, the
 * autograd dgrad of the same layers.  nn.Linear (attention.py:193,195; common.py:20,22;
 * head_helper.py) is the kt=kh=kw=1 case.
 *
 *   out[n, z, p, q, co] (+)= sum_{tap, ci} A[n, low_t + z*str_t + it*dil_t, low_h + ..., low_w + ..., ci]
 *                                          * B[co, tap, ci]            (out-of-range A reads as 0)
 * ---------------------------------------------------------------------------------------------- */
typedef struct sfb_conv_desc {
  /* A operand: activation planes, channels-last [n, d, h, w, c], channel pitch c_pitch >= c */
  const void* a_hi;
  const void* a_lo; /* may be NULL when nsplit == 1 */
  int32_t n, d, h, w, c;
  int64_t c_pitch;
  /* B operand: filter matrix planes, bf16 [cout, kt*kh*kw*c], K index = ((it*kh + ih)*kw + iw)*c + ci */
  const void* b_hi;
  const void* b_lo; /* may be NULL when nsplit == 1 */
  int32_t cout;
  int32_t kt, kh, kw;
  int32_t dil_t, dil_h, dil_w;
  int32_t str_t, str_h, str_w;
  int32_t low_t, low_h, low_w; /* input coordinate of tap 0 for output index 0 (= -padding for fprop) */
  int32_t out_t, out_h, out_w; /* output grid */
  /* output view: fp32, channel stride 1, element strides for (n, t, h, w) */
  float* out;
  int64_t os_n, os_t, os_h, os_w;
  int32_t accumulate; /* 0: out = result, 1: out += result */
  /* optional per-tile BatchNorm partials: [m_tiles][2][cout] = (sum, sum of squares) over the tile's rows */
  float* stats;
  int32_t nsplit; /* 1 (bf16 operands) or 3 (split-bf16, fp32-class operands) */
} sfb_conv_desc;

/* Number of 128-row output tiles (first extent of `stats`). */
int64_t sfb_conv_m_tiles(const sfb_conv_desc* d);
int sfb_conv_igemm(const sfb_conv_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Convolution weight gradient (autograd wgrad of nn.Conv3d / nn.Linear).
 *   dw[co, (tap, ci)] += sum over output positions m of  dY[m, co] * X[input position of (m, tap), ci]
 * dw is fp32 in the SAME [cout, kt*kh*kw*c] matrix layout as the fprop filter matrix and must be zero-filled
 * (or hold a running sum) by the caller; split-K partials are combined with fp32 reductions.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sfb_wgrad_desc {
  const void* x_hi;
  const void* x_lo; /* forward input planes, channels-last [n, d, h, w, c], pitch c_pitch */
  int32_t n, d, h, w, c;
  int64_t c_pitch;
  const void* dy_hi;
  const void* dy_lo; /* output-gradient planes, dense [n*out_t*out_h*out_w, cout] with row pitch dy_pitch */
  int32_t cout;
  int64_t dy_pitch;
  int32_t kt, kh, kw;
  int32_t dil_t, dil_h, dil_w;
  int32_t str_t, str_h, str_w;
  int32_t low_t, low_h, low_w;
  int32_t out_t, out_h, out_w;
  float* dw;
  int32_t nsplit;
} sfb_wgrad_desc;

int sfb_conv_wgrad(const sfb_wgrad_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLOWFAST_B200_H_ */
