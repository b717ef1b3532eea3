"""Host logic: the strided-dgrad decomposition reproduces autograd's data gradient (CPU, fp64)."""
import itertools

import pytest
import torch
import torch.nn.functional as F

from slowfast_b200.conv_plan import corners_in_tma_range, dgrad_out_view, dgrad_plan


def emulate_dgrad(dy, w, in_size, k, stride, pad):
    """dy: [n, ot, oh, ow, co] ; w: [co, ci, kt, kh, kw] -> dx [n, T, H, W, ci] via the sub-problem list,
    each evaluated exactly the way the fprop kernel would (lower corner + taps + strided output view)."""
    n, co = dy.shape[0], dy.shape[-1]
    ci = w.shape[1]
    T, H, W = in_size
    plan = dgrad_plan(in_size, k, stride, pad)
    dx = torch.zeros(n * T * H * W * ci, dtype=dy.dtype)
    P = 8
    dyp = F.pad(dy.permute(0, 4, 1, 2, 3), (P, P, P, P, P, P))  # n, co, ...
    wf = w.reshape(co, ci, -1)
    for sub in plan.subs:
        assert corners_in_tma_range(dy.shape[1:4], sub.low, sub.out, (1, 1, 1))
        jt, jh, jw = sub.k
        wsub = wf[:, :, list(sub.tapmap)].reshape(co, ci, jt, jh, jw).permute(1, 0, 2, 3, 4)  # [ci, co, ...]
        lt, lh, lw = sub.low
        crop = dyp[:, :, P + lt:P + lt + sub.out[0] + jt - 1, P + lh:P + lh + sub.out[1] + jh - 1,
                   P + lw:P + lw + sub.out[2] + jw - 1]
        o = F.conv3d(crop, wsub).permute(0, 2, 3, 4, 1)  # n, x't, x'h, x'w, ci
        off, (sn, st, sh, sw) = dgrad_out_view(in_size, stride, sub, ci)
        idx = (off + torch.arange(n).view(-1, 1, 1, 1, 1) * sn + torch.arange(sub.out[0]).view(1, -1, 1, 1, 1) * st +
               torch.arange(sub.out[1]).view(1, 1, -1, 1, 1) * sh + torch.arange(sub.out[2]).view(1, 1, 1, -1, 1) * sw +
               torch.arange(ci).view(1, 1, 1, 1, -1))
        dx[idx.reshape(-1)] += o.reshape(-1)
    return dx.reshape(n, T, H, W, ci), plan


CASES = [
    ((4, 9, 9), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ((4, 10, 10), (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ((4, 9, 11), (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ((6, 5, 5), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ((32, 3, 3), (7, 1, 1), (4, 1, 1), (3, 0, 0)),
    ((4, 8, 8), (1, 1, 1), (1, 2, 2), (0, 0, 0)),
    ((4, 7, 7), (1, 1, 1), (1, 2, 2), (0, 0, 0)),
    ((4, 12, 12), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    ((8, 6, 6), (5, 3, 3), (2, 2, 1), (2, 1, 1)),
    ((5, 6, 6), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
]


@pytest.mark.parametrize("in_size,k,stride,pad", CASES)
def test_dgrad_plan_matches_autograd(in_size, k, stride, pad):
    torch.manual_seed(0)
    n, ci, co = 2, 3, 4
    x = torch.randn(n, ci, *in_size, dtype=torch.float64, requires_grad=True)
    w = torch.randn(co, ci, *k, dtype=torch.float64)
    y = F.conv3d(x, w, stride=stride, padding=pad)
    dy = torch.randn_like(y)
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    dx, plan = emulate_dgrad(dy.permute(0, 2, 3, 4, 1).contiguous(), w, in_size, k, stride, pad)
    assert torch.allclose(dx.permute(0, 4, 1, 2, 3), dx_ref, atol=1e-10)
    # zero-fill is needed exactly when the kernel is smaller than the stride somewhere
    assert plan.needs_zero_fill == any(kk < s for kk, s in zip(k, stride))
    # total taps across sub-problems == taps of the original filter (no wasted MMA work)
    if not plan.needs_zero_fill:
        assert sum(len(s.tapmap) for s in plan.subs) == k[0] * k[1] * k[2]


def test_stride1_is_single_flipped_conv():
    plan = dgrad_plan((4, 8, 8), (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert len(plan.subs) == 1
    assert plan.subs[0].tapmap == tuple(reversed(range(9)))
    assert plan.subs[0].low == (0, -1, -1)
