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

import math
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


def _basic_head(feats: List[torch.Tensor], sd: SD, training: bool, dropout_rate: float, act: str = "softmax",
                pool_sizes=None):
    """ResNetBasicHead.forward (head_helper.py:305-350): AvgPool3d(pool_size, stride=1) per pathway (:250-255; the
    pool size is the TRAIN-crop feature extent, video_model_builder.py:398-416 / :627-637), so a larger test crop yields
    several windows that are projected, soft-maxed per location and averaged (:338-345)."""
    if pool_sizes is None:
        pooled = [f.mean(dim=(2, 3, 4), keepdim=True) for f in feats]  # AdaptiveAvgPool3d((1,1,1))
    else:
        pooled = [F.avg_pool3d(f, tuple(ps), stride=1) for f, ps in zip(feats, pool_sizes)]
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
    c32 = cfg.DATA.TRAIN_CROP_SIZE // 32
    pools = None if cfg.MULTIGRID.SHORT_CYCLE else [[cfg.DATA.NUM_FRAMES // alpha, c32, c32], [cfg.DATA.NUM_FRAMES, c32, c32]]
    return _basic_head([xs, xf], sd, training, cfg.MODEL.DROPOUT_RATE, cfg.MODEL.HEAD_ACT, pools)


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
    c32 = cfg.DATA.TRAIN_CROP_SIZE // 32
    pools = None if cfg.MULTIGRID.SHORT_CYCLE else [[cfg.DATA.NUM_FRAMES // pool1, c32, c32]]
    return _basic_head([x], sd, training, cfg.MODEL.DROPOUT_RATE, cfg.MODEL.HEAD_ACT, pools)


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
    grads = torch.autograd.grad(logits, [leaves[k] for k in names], dlogits, allow_unused=True)
    # (entries the forward never reads - e.g. HOGLayerC's constant buffers - have no gradient)
    return logits.detach(), {k: g for k, g in zip(names, grads) if g is not None}


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
        parent = k.split(".")[-2] if "." in k else ""
        if not v.is_floating_point():
            out[k] = torch.zeros_like(v)
        elif k.endswith("weight_x") or k.endswith("weight_y"):  # HOGLayerC's constant Sobel buffers (operators.py:85-89)
            sob = torch.tensor([[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]]).view(1, 1, 3, 3).repeat(3, 1, 1, 1)
            out[k] = sob if k.endswith("weight_x") else sob.transpose(2, 3).contiguous()
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif "rel_pos" in k or k == "cls_token":
            out[k] = torch.randn(v.shape, generator=g) * 0.2
        elif ("bn" in parent or "norm" in parent) and k.endswith(".weight"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif ("bn" in parent or "norm" in parent) and k.endswith(".bias"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        else:
            out[k] = torch.randn(v.shape, generator=g) * 0.01
    return out


# ================================================================================================ MViT (v2)
def _round_width(width, multiplier, min_width=1, divisor=1):
    """models/utils.py:10-24."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def mvit_block_specs(cfg):
    """Per-block (dim, dim_out, heads, q/kv pooling kernel & stride, input thw) exactly as MViT.__init__ derives them
    (video_model_builder.py:914-1030), including the adaptive KV stride rule (:936-945)."""
    mv = cfg.MVIT
    depth = mv.DEPTH
    dim_mul, head_mul = [1.0] * (depth + 1), [1.0] * (depth + 1)
    for i, m in mv.DIM_MUL:
        dim_mul[int(i)] = m
    for i, m in mv.HEAD_MUL:
        head_mul[int(i)] = m
    pool_q, pool_kv = [[] for _ in range(depth)], [[] for _ in range(depth)]
    stride_q, stride_kv = [[] for _ in range(depth)], [[] for _ in range(depth)]
    for e in mv.POOL_Q_STRIDE:
        stride_q[e[0]] = list(e[1:])
        pool_q[e[0]] = list(mv.POOL_KVQ_KERNEL) if mv.POOL_KVQ_KERNEL is not None else \
            [s + 1 if s > 1 else s for s in e[1:]]
    kv_list = mv.POOL_KV_STRIDE
    if mv.POOL_KV_STRIDE_ADAPTIVE is not None:
        _s = list(mv.POOL_KV_STRIDE_ADAPTIVE)
        kv_list = []
        for i in range(depth):
            if len(stride_q[i]) > 0:
                _s = [max(_s[d] // stride_q[i][d], 1) for d in range(len(_s))]
            kv_list.append([i] + _s)
    for e in kv_list:
        stride_kv[e[0]] = list(e[1:])
        pool_kv[e[0]] = list(mv.POOL_KVQ_KERNEL) if mv.POOL_KVQ_KERNEL is not None else \
            [s + 1 if s > 1 else s for s in e[1:]]
    patch_stride = list(mv.PATCH_STRIDE)
    size = [cfg.DATA.NUM_FRAMES // patch_stride[0], cfg.DATA.TRAIN_CROP_SIZE // patch_stride[1],
            cfg.DATA.TRAIN_CROP_SIZE // patch_stride[2]]
    embed, heads = mv.EMBED_DIM, mv.NUM_HEADS
    specs = []
    for i in range(depth):
        heads = _round_width(heads, head_mul[i])
        if mv.DIM_MUL_IN_ATT:
            dim_out = _round_width(embed, dim_mul[i], divisor=_round_width(heads, head_mul[i]))
        else:
            dim_out = _round_width(embed, dim_mul[i + 1], divisor=_round_width(heads, head_mul[i + 1]))
        specs.append(dict(dim=embed, dim_out=dim_out, heads=heads, kq=pool_q[i], kkv=pool_kv[i], sq=stride_q[i],
                          skv=stride_kv[i], size=list(size)))
        if len(stride_q[i]) > 0:
            size = [s // st for s, st in zip(size, stride_q[i])]
        embed = dim_out
    return specs


def _attention_pool(x, w, stride, thw, norm_w, norm_b, heads_folded=True):
    """attention_pool (attention.py:13-45) with a depthwise Conv3d pool + LayerNorm; x: (B, H, 1+L, D), cls first."""
    if w is None:
        return x, thw
    cls, x = x[:, :, :1], x[:, :, 1:]
    B, H, L, D = x.shape
    T, Hh, W = thw
    x = x.reshape(B * H, T, Hh, W, D).permute(0, 4, 1, 2, 3)
    k = w.shape[2:]
    x = F.conv3d(x, w, None, stride, [kk // 2 for kk in k], 1, D)
    thw = list(x.shape[2:])
    x = x.reshape(B, H, D, -1).transpose(2, 3)
    x = torch.cat((cls, x), 2)
    if norm_w is not None:
        x = F.layer_norm(x, (D,), norm_w, norm_b, 1e-6)
    return x, thw


def _rel_pos_index(nq, nk):
    """dist index of cal_rel_pos_* (attention.py:76-86, 122-127)."""
    qr, kr = max(nk / nq, 1.0), max(nq / nk, 1.0)
    d = torch.arange(nq)[:, None] * qr - torch.arange(nk)[None, :] * kr + (nk - 1) * kr
    return d.long()


def _mvit_attention(x, sd, pre, spec, thw):
    """MultiScaleAttention.forward (attention.py:293-392), pool_first False, separate_qkv False, mode conv,
    cls token on, decomposed relative positions, residual pooling."""
    B, N, _ = x.shape
    H = spec["heads"]
    att_dim = sd[pre + ".qkv.weight"].shape[0] // 3
    hd = att_dim // H
    qkv = F.linear(x, sd[pre + ".qkv.weight"], sd.get(pre + ".qkv.bias")).reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]

    def pool(t, name, stride):
        w = sd.get(f"{pre}.pool_{name}.weight")
        if w is None:
            return t, thw
        return _attention_pool(t, w, stride, thw, sd[f"{pre}.norm_{name}.weight"], sd[f"{pre}.norm_{name}.bias"])

    q, q_thw = pool(q, "q", spec["sq"])
    k, k_thw = pool(k, "k", spec["skv"])
    v, _ = pool(v, "v", spec["skv"])
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    qt, qh, qw = q_thw
    kt, kh, kw = k_thw
    rq = q[:, :, 1:].reshape(B, H, qt, qh, qw, hd)
    if pre + ".rel_pos_h" in sd:  # cal_rel_pos_spatial (:64-108): bias on the non-cls block only, un-scaled q
        Rh = sd[pre + ".rel_pos_h"][_rel_pos_index(qh, kh)]
        Rw = sd[pre + ".rel_pos_w"][_rel_pos_index(qw, kw)]
        assert sd[pre + ".rel_pos_h"].shape[0] == 2 * max(qh, kh) - 1, "rel-pos interpolation is not on this path"
        rel_h = torch.einsum("bythwc,hkc->bythwk", rq, Rh)
        rel_w = torch.einsum("bythwc,wkc->bythwk", rq, Rw)
        a = attn[:, :, 1:, 1:].reshape(B, H, qt, qh, qw, kt, kh, kw)
        a = a + rel_h[:, :, :, :, :, None, :, None] + rel_w[:, :, :, :, :, None, None, :]
        attn = torch.cat([attn[:, :, :1], torch.cat([attn[:, :, 1:, :1], a.reshape(B, H, qt * qh * qw, kt * kh * kw)], 3)], 2)
    if pre + ".rel_pos_t" in sd:  # cal_rel_pos_temporal (:111-147)
        assert sd[pre + ".rel_pos_t"].shape[0] == 2 * max(qt, kt) - 1
        Rt = sd[pre + ".rel_pos_t"][_rel_pos_index(qt, kt)]
        rel_t = torch.einsum("bythwc,tkc->bythwk", rq, Rt)
        a = attn[:, :, 1:, 1:].reshape(B, H, qt, qh, qw, kt, kh, kw) + rel_t[:, :, :, :, :, :, None, None]
        attn = torch.cat([attn[:, :, :1], torch.cat([attn[:, :, 1:, :1], a.reshape(B, H, qt * qh * qw, kt * kh * kw)], 3)], 2)
    attn = attn.softmax(-1)
    o = attn @ v
    o = torch.cat([o[:, :, :1], o[:, :, 1:] + q[:, :, 1:]], 2)  # residual pooling (:381-385)
    o = o.transpose(1, 2).reshape(B, -1, att_dim)
    return F.linear(o, sd[pre + ".proj.weight"], sd[pre + ".proj.bias"]), q_thw


def _mvit_block(x, sd, pre, spec, thw):
    """MultiScaleBlock.forward (attention.py:491-514), no layer scale, drop-path off.  The channel expansion
    dim -> dim_out happens in the attention (DIM_MUL_IN_ATT, MViTv2: att_dim == dim_out, residual = proj(norm1(x)))
    or in the MLP (MViTv1 default: att_dim == dim, residual = proj(norm2(x)))."""
    dim, dim_out = spec["dim"], spec["dim_out"]
    att_dim = sd[pre + ".attn.qkv.weight"].shape[0] // 3
    in_att = att_dim == dim_out
    xn = F.layer_norm(x, (dim,), sd[pre + ".norm1.weight"], sd[pre + ".norm1.bias"], 1e-6)
    xb, thw_new = _mvit_attention(xn, sd, pre + ".attn", spec, thw)
    if in_att and dim != dim_out:
        x = F.linear(xn, sd[pre + ".proj.weight"], sd[pre + ".proj.bias"])
    sq = spec["sq"]
    if len(sq) > 0 and sq[0] * sq[1] * sq[2] > 1:  # pool_skip = MaxPool3d(kernel s+1, stride s, pad k//2) (:485-489)
        ks = [s + 1 if s > 1 else s for s in sq]
        cls, xs = x[:, :1], x[:, 1:]
        B, L, C = xs.shape
        xs = xs.reshape(B, *thw, C).permute(0, 4, 1, 2, 3)
        xs = F.max_pool3d(xs, ks, sq, [kk // 2 for kk in ks])
        x = torch.cat((cls, xs.reshape(B, C, -1).transpose(1, 2)), 1)
    x = x + xb
    xn = F.layer_norm(x, (att_dim,), sd[pre + ".norm2.weight"], sd[pre + ".norm2.bias"], 1e-6)
    h = F.gelu(F.linear(xn, sd[pre + ".mlp.fc1.weight"], sd[pre + ".mlp.fc1.bias"]))
    if not in_att and dim != dim_out:
        x = F.linear(xn, sd[pre + ".proj.weight"], sd[pre + ".proj.bias"])
    x = x + F.linear(h, sd[pre + ".mlp.fc2.weight"], sd[pre + ".mlp.fc2.bias"])
    return x, thw_new


def mvit_forward(cfg, sd: SD, inputs: List[torch.Tensor], training: bool = True, record=None) -> torch.Tensor:
    """MViT.forward (video_model_builder.py:1166-1244) for the MViTv2 configs: cls token, no absolute position
    embedding, relative positions, cls-token readout, TransformerBasicHead (head_helper.py:547-563).
    Stochastic depth and dropout must be configured off (parity runs)."""
    mv = cfg.MVIT
    assert mv.CLS_EMBED_ON and not mv.USE_ABS_POS and not mv.USE_MEAN_POOLING and mv.MODE == "conv"
    assert not mv.POOL_FIRST and not mv.SEPARATE_QKV and not mv.NORM_STEM
    assert not training or (float(mv.DROPPATH_RATE) == 0.0 and float(cfg.MODEL.DROPOUT_RATE) == 0.0 and
                            float(mv.DROPOUT_RATE) == 0.0), "oracle runs without stochastic regularisers"
    (x,) = inputs
    x = F.conv3d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], tuple(mv.PATCH_STRIDE),
                 tuple(mv.PATCH_PADDING))
    B, C, T, H, W = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd["cls_token"].expand(B, -1, -1), x), 1)
    thw = [T, H, W]
    for i, spec in enumerate(mvit_block_specs(cfg)):
        x, thw = _mvit_block(x, sd, f"blocks.{i}", spec, thw)
        if record is not None:
            record[f"b{i}"] = x.detach()
    x = F.layer_norm(x, (x.shape[-1],), sd["norm.weight"], sd["norm.bias"], 1e-6)[:, 0]
    x = F.linear(x, sd["head.projection.weight"], sd["head.projection.bias"])
    if not training and cfg.MODEL.HEAD_ACT == "softmax":
        x = torch.softmax(x, 1)
    return x.reshape(B, -1)


FORWARD["MViT"] = mvit_forward


# ================================================================================================ X3D
def _x3d_block(x, sd: SD, prefix: str, stride: int, training: bool):
    """ResBlock.forward (resnet_helper.py:512-521) with X3DTransform.forward (:253-256, modules in construction
    order :199-251): a 1x1x1 -> BN -> ReLU -> channelwise Tx3x3 (stride on the 3x3, STRIDE_1X1 False) -> BN ->
    [SE on even block indices: operators.py:55-59] -> Swish (pytorchvideo: x*sigmoid(x)) -> c 1x1x1 -> BN."""
    b2 = prefix + ".branch2"
    wa, wb, wc = sd[b2 + ".a.weight"], sd[b2 + ".b.weight"], sd[b2 + ".c.weight"]
    f = F.relu(_bn(F.conv3d(x, wa), sd, b2 + ".a_bn", training))
    kt = wb.shape[2]
    f = _bn(F.conv3d(f, wb, None, (1, stride, stride), (kt // 2, 1, 1), 1, wb.shape[0]), sd, b2 + ".b_bn", training)
    if b2 + ".se.fc1.weight" in sd:
        s = f.mean(dim=(2, 3, 4), keepdim=True)  # AdaptiveAvgPool3d((1,1,1))
        s = F.relu(F.conv3d(s, sd[b2 + ".se.fc1.weight"], sd[b2 + ".se.fc1.bias"]))
        s = torch.sigmoid(F.conv3d(s, sd[b2 + ".se.fc2.weight"], sd[b2 + ".se.fc2.bias"]))
        f = f * s
    f = f * torch.sigmoid(f)
    f = _bn(F.conv3d(f, wc), sd, b2 + ".c_bn", training)
    if prefix + ".branch1.weight" in sd:
        x = _bn(F.conv3d(x, sd[prefix + ".branch1.weight"], None, (1, stride, stride)), sd, prefix + ".branch1_bn",
                training) + f
    else:
        x = x + f
    return F.relu(x)


def x3d_forward(cfg, sd: SD, inputs: List[torch.Tensor], training: bool = True, record=None) -> torch.Tensor:
    """X3D.forward (video_model_builder.py:799-802): X3DStem (stem_helper.py:280-285), four ResStages of
    X3DTransform blocks (every stage's first block has stride 2, :706-711), X3DHead (head_helper.py:461-488) for
    pool size == feature size.  DROPCONNECT_RATE is 0 in the X3D yamls (drop_path inactive)."""
    (x,) = inputs
    p = "s1.pathway0_stem"
    wxy, wt = sd[p + ".conv_xy.weight"], sd[p + ".conv.weight"]
    x = F.conv3d(x, wxy, None, (1, 2, 2), (0, wxy.shape[3] // 2, wxy.shape[4] // 2))
    x = F.conv3d(x, wt, None, (1, 1, 1), (wt.shape[2] // 2, 0, 0), 1, wt.shape[0])
    x = F.relu(_bn(x, sd, p + ".bn", training))
    if record is not None:
        record["s1"] = x
    for stage in range(2, 6):
        i = 0
        while f"s{stage}.pathway0_res{i}.branch2.a.weight" in sd:
            x = _x3d_block(x, sd, f"s{stage}.pathway0_res{i}", 2 if i == 0 else 1, training)
            if record is not None:
                record[f"s{stage}.{i}"] = x
            i += 1
    x = F.relu(_bn(F.conv3d(x, sd["head.conv_5.weight"]), sd, "head.conv_5_bn", training))
    x = x.mean(dim=(2, 3, 4), keepdim=True)  # AvgPool3d(pool_size) over the whole extent
    x = F.relu(F.conv3d(x, sd["head.lin_5.weight"]))
    x = x.permute(0, 2, 3, 4, 1)
    if cfg.MODEL.DROPOUT_RATE > 0.0:
        x = F.dropout(x, cfg.MODEL.DROPOUT_RATE, training)
    x = F.linear(x, sd["head.projection.weight"], sd["head.projection.bias"])
    if not training:
        if cfg.MODEL.HEAD_ACT == "softmax":
            x = torch.softmax(x, dim=4)
        x = x.mean([1, 2, 3])
    return x.reshape(x.shape[0], -1)


FORWARD["X3D"] = x3d_forward


# ================================================================================================ MaskFeat
def _mvit_feature_geometry(cfg):
    """slowfast/models/utils.py:185-214 calc_mvit_feature_geometry."""
    depth = cfg.MVIT.DEPTH
    ps = list(cfg.MVIT.PATCH_STRIDE)
    size = [[cfg.DATA.NUM_FRAMES // ps[0], cfg.DATA.TRAIN_CROP_SIZE // ps[1], cfg.DATA.TRAIN_CROP_SIZE // ps[2]]
            for _ in range(depth)]
    stride = [list(ps) for _ in range(depth)]
    for x in cfg.MVIT.POOL_Q_STRIDE:
        for i in range(depth):
            if i >= x[0]:
                for j in range(3):
                    size[i][j] //= x[j + 1]
                    stride[i][j] *= x[j + 1]
    return size, stride


def hog_layer(x: torch.Tensor, nbins: int = 9, pool: int = 8) -> torch.Tensor:
    """HOGLayerC.forward (operators.py:79-122) without the optional gaussian window: [B,3,H,W] -> [B,3,nbins,H/pool,W/pool]."""
    wx = torch.tensor([[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]]).view(1, 1, 3, 3).repeat(3, 1, 1, 1).to(x)
    wy = wx.transpose(2, 3)
    x = F.pad(x, pad=(1, 1, 1, 1), mode="reflect")
    gx = F.conv2d(x, wx, None, 1, 0, 1, 3)
    gy = F.conv2d(x, wy, None, 1, 0, 1, 3)
    norm = torch.stack([gx, gy], dim=-1).norm(dim=-1)
    phase = torch.atan2(gx, gy) / math.pi * nbins
    b, c, h, w = norm.shape
    out = torch.zeros((b, c, nbins, h, w), dtype=torch.float, device=x.device)
    out.scatter_add_(2, phase.view(b, c, 1, h, w).floor().long() % nbins, norm.view(b, c, 1, h, w))
    out = out.unfold(3, pool, pool).unfold(4, pool, pool).sum(dim=[-1, -2])
    return F.normalize(out, p=2, dim=2)


def maskfeat_masks(cfg, mask: torch.Tensor):
    """_get_multiscale_mask (masked.py:165-176) for the (single) pretrain depth, and the token-grid float mask."""
    size, _ = _mvit_feature_geometry(cfg)
    fm = mask.float()
    out = F.interpolate(fm, size=size[cfg.MASK.PRETRAIN_DEPTH[-1]][-1]).flatten(1).to(torch.bool)
    return out, fm


def maskfeat_labels(cfg, frames: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """_get_hog_label_3d (masked.py:254-281): HOG of every PATCH_STRIDE[0]-th frame regrouped per output token,
    rows selected by the multiscale mask -> [n_mask, 3*nbins*(stride/8)^2]."""
    size, _ = _mvit_feature_geometry(cfg)
    out_mask, _ = maskfeat_masks(cfg, mask)
    fs = size[cfg.MASK.PRETRAIN_DEPTH[-1]][-1]
    x = frames[:, :, ::cfg.MVIT.PATCH_STRIDE[0]].transpose(1, 2)
    B, T = x.shape[:2]
    hog = hog_layer(x.flatten(0, 1)).flatten(1, 2)
    u = hog.shape[-1] // fs
    hog = hog.permute(0, 2, 3, 1).unfold(1, u, u).unfold(2, u, u).flatten(3).view(B, T, fs, fs, -1).flatten(1, 3)
    return hog[out_mask]


def maskfeat_forward(cfg, sd: SD, inputs: List[torch.Tensor], training: bool = True, record=None) -> torch.Tensor:
    """MaskMViT._maskfeat_forward (masked.py:519-612) with HOG targets, HEAD_TYPE "separate", one pretrain depth:
    patch embedding, mask-token substitution, encoder blocks, MSSeparateHead (LayerNorm, drop cls, select masked
    tokens, Linear; head_helper.py:656-672).  inputs = [frames, mask]; returns the predictions [n_mask, classes]."""
    mv = cfg.MVIT
    frames, mask = inputs
    out_mask, fm = maskfeat_masks(cfg, mask)
    x = F.conv3d(frames, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], tuple(mv.PATCH_STRIDE),
                 tuple(mv.PATCH_PADDING))
    B, C, T, H, W = x.shape
    x = x.flatten(2).transpose(1, 2)
    tm = F.interpolate(fm, size=(H, W)).flatten(1).unsqueeze(-1)
    x = x * (1 - tm) + sd["mask_token"].expand(B, x.shape[1], -1) * tm
    x = torch.cat((sd["cls_token"].expand(B, -1, -1), x), 1)
    thw = [T, H, W]
    depth = cfg.MASK.PRETRAIN_DEPTH[-1]
    for i, spec in enumerate(mvit_block_specs(cfg)[:depth + 1]):
        x, thw = _mvit_block(x, sd, f"blocks.{i}", spec, thw)
    x = F.layer_norm(x, (x.shape[-1],), sd["pred_head.transforms.0.0.weight"], sd["pred_head.transforms.0.0.bias"], 1e-6)
    x = x[:, 1:][out_mask]
    return F.linear(x, sd["pred_head.projections.0.weight"], sd["pred_head.projections.0.bias"])


FORWARD["MaskMViT"] = maskfeat_forward


def maskfeat_inputs(cfg, batch: int, seed: int):
    """Seeded clip + cube mask (B, T/2, 7, 7) with ~40 % of the cells masked (AUG.MASK_RATIO 0.4; the reference's
    MaskingGenerator3D draws cuboids - any 0/1 mask exercises the same arithmetic)."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randn(batch, 3, cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE, cfg.DATA.TRAIN_CROP_SIZE, generator=g)
    size, _ = _mvit_feature_geometry(cfg)
    t = cfg.DATA.NUM_FRAMES // cfg.MVIT.PATCH_STRIDE[0]
    ms = max(1, size[cfg.MASK.PRETRAIN_DEPTH[-1]][-1] // 2)
    mask = (torch.rand(batch, t, ms, ms, generator=g) < 0.4).float()
    return [frames, mask]
