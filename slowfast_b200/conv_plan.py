"""Host-side geometry of the convolution problems handed to the implicit-GEMM kernels.

The data-gradient of a strided convolution is decomposed into ``prod(stride)`` stride-1 sub-convolutions over the
output-gradient grid, one per residue class of the input coordinate, each writing an interleaved (strided) view of
dX.  Every sub-problem has exactly the form the fprop kernel executes (TMA im2col with a lower corner, a tap list
and an output view), so dgrad needs no kernel of its own and wastes no MMA work on zero-inserted gradients.

Pure Python, no torch: unit-tested on CPU against autograd (tests/test_conv_plan.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


@dataclass(frozen=True)
class AxisClass:
    """One residue class r of an input axis for dgrad (dilation 1)."""

    r: int          # input coordinates x = stride * x' + r
    count: int      # number of x' (extent of the sub-output along this axis)
    taps: Tuple[int, ...]  # source filter taps, in sub-filter order j' = 0..J-1 (may be empty)
    low: int        # dY coordinate read by sub-tap 0 at x' = 0


def dgrad_axis_classes(x: int, k: int, stride: int, pad: int) -> List[AxisClass]:
    """Residue classes of one axis.  fprop relation: x_in = o*stride - pad + tap."""
    out = []
    for r in range(stride):
        count = (x - r + stride - 1) // stride if x > r else 0
        k0 = (r + pad) % stride
        if k0 >= k:
            out.append(AxisClass(r, count, (), 0))
            continue
        j_n = (k - k0 + stride - 1) // stride
        c = (r + pad - k0) // stride
        # sub-tap j' reads dY[x' + low + j'] and pairs with source tap k0 + stride*(J-1-j')
        taps = tuple(k0 + stride * (j_n - 1 - jp) for jp in range(j_n))
        out.append(AxisClass(r, count, taps, c - (j_n - 1)))
    return out


@dataclass(frozen=True)
class DgradSub:
    """One stride-1 sub-convolution of a dgrad."""

    k: Tuple[int, int, int]            # sub-filter extent (Jt, Jh, Jw)
    tapmap: Tuple[int, ...]            # source tap index (row-major over the ORIGINAL kt,kh,kw) per sub-tap
    low: Tuple[int, int, int]          # lower corner on the dY grid
    out: Tuple[int, int, int]          # sub-output grid (count per axis)
    r: Tuple[int, int, int]            # residue (offset of the view inside dX)


@dataclass(frozen=True)
class DgradPlan:
    subs: Tuple[DgradSub, ...]
    needs_zero_fill: bool              # some input positions receive no contribution at all


def dgrad_plan(in_size: Sequence[int], k: Sequence[int], stride: Sequence[int], pad: Sequence[int]) -> DgradPlan:
    """Decompose dX = conv_transpose(dY, W) for an fprop conv with the given kernel/stride/pad (dilation 1)."""
    axes = [dgrad_axis_classes(x, kk, s, p) for x, kk, s, p in zip(in_size, k, stride, pad)]
    subs = []
    zero = False
    for at in axes[0]:
        for ah in axes[1]:
            for aw in axes[2]:
                if at.count == 0 or ah.count == 0 or aw.count == 0:
                    continue
                if not at.taps or not ah.taps or not aw.taps:
                    zero = True
                    continue
                tapmap = tuple((tt * k[1] + th) * k[2] + tw for tt in at.taps for th in ah.taps for tw in aw.taps)
                subs.append(DgradSub((len(at.taps), len(ah.taps), len(aw.taps)), tapmap, (at.low, ah.low, aw.low),
                                     (at.count, ah.count, aw.count), (at.r, ah.r, aw.r)))
    return DgradPlan(tuple(subs), zero)


def dgrad_out_view(in_size: Sequence[int], stride: Sequence[int], sub: DgradSub, c_pitch: int, c0: int = 0):
    """(offset, (os_n, os_t, os_h, os_w)) of the sub-problem's output inside dX [n, T, H, W, c_pitch]."""
    t, h, w = in_size
    off = ((sub.r[0] * h + sub.r[1]) * w + sub.r[2]) * c_pitch + c0
    return off, (t * h * w * c_pitch, stride[0] * h * w * c_pitch, stride[1] * w * c_pitch, stride[2] * c_pitch)


def im2col_upper_corner(dim: int, low: int, out: int, stride: int) -> int:
    """Upper bounding-box corner the TMA im2col map needs so that exactly ``out`` base positions exist."""
    return low + (out - 1) * stride + 1 - dim


def corners_in_tma_range(dims, low, out, stride) -> bool:
    for d, lo, o, s in zip(dims, low, out, stride):
        up = im2col_upper_corner(d, lo, o, s)
        if not (-16 <= lo <= 15 and -16 <= up <= 15):
            return False
    return True
