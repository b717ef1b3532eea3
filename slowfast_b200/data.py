"""Device-side head of the input pipeline.

``pack_pathways_u8(frames_u8, cfg)`` takes the decoder's uint8 clips already on the GPU - ``[B, T, H, W, 3]`` - and returns
what the reference's loader hands to the model after its host-side work: ``tensor_normalize`` (slowfast/datasets/utils.py
:278-297), the ``(T,H,W,C) -> (C,T,H,W)`` permute (datasets/kinetics.py:375-405) and ``pack_pathway_output``
(datasets/utils.py:78-112: the slow pathway keeps the frames at ``linspace(0, T-1, T // ALPHA).long()``).  One kernel launch
per pathway (csrc/input_pipeline.cu); the host -> device copy carries uint8 instead of fp32.  Spatial augmentation (random
crop / flip / jitter) is data-dependent host logic of the loader and stays there.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import lib as L


def _normalize_pack(frames_u8: torch.Tensor, idx, mean, std, reverse: bool) -> torch.Tensor:
    if frames_u8.device.type != "cuda":
        raise L.NativeLibraryError("slowfast_b200 runs on CUDA devices only (no CPU fallback)")
    assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 5 and frames_u8.shape[-1] == 3 and frames_u8.is_contiguous()
    b, t, h, w, _ = frames_u8.shape
    t_out = t if idx is None else idx.numel()
    out = torch.empty((b, 3, t_out, h, w), dtype=torch.float32, device=frames_u8.device)
    m = (C.c_float * 3)(*[float(x) for x in mean])
    s = (C.c_float * 3)(*[float(x) for x in std])
    L.check(L.load().sfb_clip_normalize_pack(frames_u8.data_ptr(), b, t, h, w, None if idx is None else idx.data_ptr(), t_out,
                                             m, s, 1 if reverse else 0, out.data_ptr(),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sfb_clip_normalize_pack")
    from . import ops
    ops.add_launches(1)
    return out


def pack_pathways_u8(frames_u8: torch.Tensor, cfg) -> List[torch.Tensor]:
    """uint8 ``[B, T, H, W, 3]`` on the GPU -> the model's input list (``[slow, fast]`` or ``[x]``), fp32 NCTHW."""
    mean, std = list(cfg.DATA.MEAN), list(cfg.DATA.STD)
    rev = bool(getattr(cfg.DATA, "REVERSE_INPUT_CHANNEL", False))
    t = frames_u8.shape[1]
    if cfg.MODEL.ARCH == "slowfast":
        idx = torch.linspace(0, t - 1, t // cfg.SLOWFAST.ALPHA).long().to(torch.int32).to(frames_u8.device)
        return [_normalize_pack(frames_u8, idx, mean, std, rev), _normalize_pack(frames_u8, None, mean, std, rev)]
    return [_normalize_pack(frames_u8, None, mean, std, rev)]
