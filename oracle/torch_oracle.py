"""TEST INFRASTRUCTURE (parity oracle) — only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference arm may import this file; the product path (slowfast_b200/) never does.

Plain-PyTorch (fp32, CPU by default) restatement of the reference's forward pass for the hot-path models, written
as pure functions over a ``state_dict`` with the reference's key names.  Gradients come from torch autograd over
these functions.  Each function cites the reference lines it follows.

Pinning: ``oracle/make_golden.py`` (run in the build container, where /root/reference exists) checks this file
against the UNMODIFIED reference modules imported through ``oracle/refshim.py`` and stores golden vectors under
``tests/golden``; ``tests/test_oracle.py`` re-checks the restatement against those vectors on any box.
The reference ships no tests or golden vectors of its own, so the reference modules themselves,
run here, are the pin.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

STAGE_DEPTH = {18: (2, 2, 2, 2), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}  # video_model_builder.py:38


def _bn(x, sd: SD, prefix: str, training: bool, momentum=0.1, eps=1e-5):
    """nn.BatchNorm3d forward (batchnorm_helper.py:16 -> torch): batch statistics + running-stat update in train
    mode, running statistics in eval mode."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)


def _conv_geometry(w: torch.Tensor, kind: str, stride: int = 1, alpha: int = 1):
    kt, kh, kw = w.shape[2:]
    if kind == "stem":      # stem_helper.py:182-189: stride [1,2,2], padding [kt//2, 3, 3]
        return (1, 2, 2), (kt // 2, kh // 2, kw // 2)
    if kind == "a":         # resnet_helper.py:332-339 (STRIDE_1X1 False): Tx1x1, stride 1, padding [T//2,0,0]
        return (1, 1, 1), (kt // 2, 0, 0)
    if kind == "b":         # resnet_helper.py:346-355: 1x3x3, stride [1,s,s], padding [0,1,1]
        return (1, stride, stride), (0, 1, 1)
    if kind == "c":         # resnet_helper.py:362-369
        return (1, 1, 1), (0, 0, 0)
    if kind == "branch1":   # resnet_helper.py:485-493
        return (1, stride, stride), (0, 0, 0)
    if kind == "fuse":      # video_model_builder.py:147-154: [k,1,1], stride [alpha,1,1], padding [k//2,0,0]
        return (alpha, 1, 1), (kt // 2, 0, 0)
    raise ValueError(kind)


def _stem(x, sd: SD, prefix: str, training: bool):
    """ResNetBasicStem.forward (stem_helper.py:196-201): conv -> bn -> relu -> maxpool [1,3,3]/[1,2,2]/[0,1,1]."""
    w = sd[prefix + ".conv.weight"]
    s, p = _conv_geometry(w, "stem")
    x = F.conv3d(x, w, None, s, p)
    x = F.relu(_bn(x, sd, prefix + ".bn", training))
    return F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))


def _fuse(xs, xf, sd: SD, prefix: str, alpha: int, training: bool):
    """FuseFastToSlow.forward (video_model_builder.py:162-169)."""
    w = sd[prefix + ".conv_f2s.weight"]
    s, p = _conv_geometry(w, "fuse", alpha=alpha)
    f = F.relu(_bn(F.conv3d(xf, w, None, s, p), sd, prefix + ".bn", training))
    return torch.cat([xs, f], 1), xf


def _res_block(x, sd: SD, prefix: str, stride: int, training: bool):
    """ResBlock.forward (resnet_helper.py:512-521) with BottleneckTransform.forward (:377-392).
    drop_path is a no-op in the reference (called without training=True, §3.3)."""
    b2 = prefix + ".branch2"
    wa, wb, wc = sd[b2 + ".a.weight"], sd[b2 + ".b.weight"], sd[b2 + ".c.weight"]
    s, p = _conv_geometry(wa, "a")
    f = F.relu(_bn(F.conv3d(x, wa, None, s, p), sd, b2 + ".a_bn", training))
    s, p = _conv_geometry(wb, "b", stride=stride)
    f = F.relu(_bn(F.conv3d(f, wb, None, s, p), sd, b2 + ".b_bn", training))
    s, p = _conv_geometry(wc, "c")
    f = _bn(F.conv3d(f, wc, None, s, p), sd, b2 + ".c_bn", training)
    if prefix + ".branch1.weight" in sd:
        w1 = sd[prefix + ".branch1.weight"]
        s, p = _conv_geometry(w1, "branch1", stride=stride)
        x = _bn(F.conv3d(x, w1, None, s, p), sd, prefix + ".branch1_bn", training) + f
    else:
        x = x + f
    return F.relu(x)


def _stage(xs: List[torch.Tensor], sd: SD, prefix: str, depth: int, strides: List[int], training: bool):
    """ResStage.forward (resnet_helper.py:697-726), no Nonlocal."""
    out = []
    for p, x in enumerate(xs):
        for i in range(depth):
            x = _res_block(x, sd, f"{prefix}.pathway{p}_res{i}", strides[p] if i == 0 else 1, training)
        out.append(x)
    return out


def _basic_head(feats: List[torch.Tensor], sd: SD, training: bool, dropout_rate: float, act: str = "softmax"):
    """ResNetBasicHead.forward (head_helper.py:305-350) for pool size == feature size (train crop)."""
    pooled = [f.mean(dim=(2, 3, 4), keepdim=True) for f in feats]  # AvgPool3d over the whole extent
    x = torch.cat(pooled, 1).permute(0, 2, 3, 4, 1)
    if dropout_rate > 0.0:
        x = F.dropout(x, dropout_rate, training)
    x = F.linear(x, sd["head.projection.weight"], sd["head.projection.bias"])
    if not training:
        if act == "softmax":
            x = torch.softmax(x, dim=4)
        x = x.mean([1, 2, 3])
    return x.reshape(x.shape[0], -1)


def slowfast_forward(cfg, sd: SD, inputs: List[torch.Tensor], training: bool = True, record=None) -> torch.Tensor:
    """SlowFast.forward (video_model_builder.py:423-441).  ``sd`` must hold parameters AND BN buffers; the BN
    running statistics are updated in place in training mode, as in the reference."""
    depth = STAGE_DEPTH[cfg.RESNET.DEPTH]
    alpha = cfg.SLOWFAST.ALPHA
    xs, xf = inputs
    xs = _stem(xs, sd, "s1.pathway0_stem", training)
    xf = _stem(xf, sd, "s1.pathway1_stem", training)
    xs, xf = _fuse(xs, xf, sd, "s1_fuse", alpha, training)
    if record is not None:
        record["s1"] = (xs.detach(), xf.detach())
    for i in range(4):
        xs, xf = _stage([xs, xf], sd, f"s{i + 2}", depth[i], cfg.RESNET.SPATIAL_STRIDES[i], training)
        if i < 3:
            xs, xf = _fuse(xs, xf, sd, f"s{i + 2}_fuse", alpha, training)
        if record is not None:
            record[f"s{i + 2}"] = (xs.detach(), xf.detach())
        # pathway{0,1}_pool are MaxPool3d with kernel = stride = [1,1,1] for ARCH slowfast (identity, _POOL1 :107)
    return _basic_head([xs, xf], sd, training, cfg.MODEL.DROPOUT_RATE, cfg.MODEL.HEAD_ACT)


def resnet_forward(cfg, sd: SD, inputs: List[torch.Tensor], training: bool = True) -> torch.Tensor:
    """ResNet.forward (video_model_builder.py:645-661) for the single-pathway archs (c2d / i3d / slow)."""
    depth = STAGE_DEPTH[cfg.RESNET.DEPTH]
    pool1 = {"2d": 1, "c2d": 2, "slow_c2d": 1, "i3d": 2, "slow_i3d": 1, "slow": 1}[cfg.MODEL.ARCH]
    (x,) = inputs
    x = _stem(x, sd, "s1.pathway0_stem", training)
    for i in range(4):
        (x,) = _stage([x], sd, f"s{i + 2}", depth[i], cfg.RESNET.SPATIAL_STRIDES[i], training)
        if i == 0 and pool1 > 1:  # pathway0_pool after res2 (:543-549, :651-653)
            x = F.max_pool3d(x, (pool1, 1, 1), (pool1, 1, 1), 0)
    return _basic_head([x], sd, training, cfg.MODEL.DROPOUT_RATE, cfg.MODEL.HEAD_ACT)


FORWARD = {"SlowFast": slowfast_forward, "ResNet": resnet_forward}


def forward(cfg, sd: SD, inputs, training: bool = True) -> torch.Tensor:
    return FORWARD[cfg.MODEL.MODEL_NAME](cfg, sd, inputs, training)


def forward_backward(cfg, sd: SD, inputs, dlogits: torch.Tensor):
    """logits and d(sum(logits*dlogits))/d(param) for every floating-point parameter in ``sd`` (train mode)."""
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    work = dict(sd)
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    for k in sd:
        if "running_" in k:
            work[k] = sd[k].clone()
    logits = forward(cfg, work, inputs, True)
    grads = torch.autograd.grad(logits, [leaves[k] for k in names], dlogits)
    return logits.detach(), dict(zip(names, grads))


def synthetic_inputs(cfg, batch: int, seed: int, crop: int | None = None, frames: int | None = None):
    """Seeded Kinetics-shaped clips: randn(B,3,T,crop,crop) packed per pathway
    (datasets/utils.py:78-112 pack_pathway_output: slow = frames at linspace(0, T-1, T//ALPHA))."""
    crop = crop or cfg.DATA.TRAIN_CROP_SIZE
    frames = frames or cfg.DATA.NUM_FRAMES
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, frames, crop, crop, generator=g)
    if cfg.MODEL.ARCH == "slowfast":
        idx = torch.linspace(0, frames - 1, frames // cfg.SLOWFAST.ALPHA).long()
        return [x.index_select(2, idx).contiguous(), x]
    return [x]


def fixture_state(template: SD, seed: int) -> SD:
    """Deterministic, non-degenerate values for every entry of a state_dict (shape/dtype from ``template``).
    Fresh reference init has c_bn.weight == 0 (ZERO_INIT_FINAL_BN), which would leave every residual branch
    unexercised; fixtures therefore draw BN scale/shift and running statistics at random."""
    out = {}
    for i, (k, v) in enumerate(template.items()):
        g = torch.Generator().manual_seed(seed * 7919 + i)
        if not v.is_floating_point():
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif "bn" in k.split(".")[-2] and k.endswith(".weight"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif "bn" in k.split(".")[-2] and k.endswith(".bias"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        else:
            out[k] = torch.randn(v.shape, generator=g) * 0.01
    return out
