"""A/B of the channelwise 3x3x3 stride-1 kernels on X3D-M's layer shapes (B=16): v2 (register tiles, L1 gathers) vs v3
(shared-memory ring): forward (with producer BN+ReLU and BN partials), stride-1 data gradient, weight gradient.
Writes gpurun_out/dw_probe.json.  Effective GB/s = (read x once + write y once) / time."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slowfast_b200 import lib as L, ops

lib = L.load()
dev = torch.device("cuda")
LAYERS = [("res2 54ch 16x56x56", 16, 16, 56, 56, 54), ("res3 108ch 16x28x28", 16, 16, 28, 28, 108),
          ("res4 216ch 16x14x14", 16, 16, 14, 14, 216), ("res5 432ch 16x7x7", 16, 16, 7, 7, 432)]


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


recs = []
for name, n, t, h, w, c in LAYERS:
    cp = ops.pad8(c)
    torch.manual_seed(0)
    x = torch.randn(n, t, h, w, cp, device=dev)
    x[..., c:] = 0
    wt = torch.randn(c, 1, 3, 3, 3, device=dev) / 5
    sc = torch.zeros(cp, device=dev); sh = torch.zeros(cp, device=dev)
    sc[:c] = torch.rand(c, device=dev) + 0.5
    sh[:c] = torch.randn(c, device=dev) * 0.5
    geom = ops.DwGeom(n, t, h, w, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    dy = torch.randn(n, t, h, w, cp, device=dev)
    dy[..., c:] = 0
    rec = dict(layer=name, mbytes=round(2 * x.numel() * 4 / 1e6, 1))
    outs = {}
    for mode in (0, 1):
        lib.sfb_set_dw3(mode)
        m_tiles, tps = ops.dwconv_tiles(geom, cp, True)
        y = torch.empty(n, t, h, w, cp, device=dev)
        stats = torch.zeros(2, c, m_tiles, device=dev)
        dx = torch.empty(n, t, h, w, cp, device=dev)
        dw = torch.empty_like(wt)
        wp = torch.empty(ops.dwconv_wgrad_blocks(geom) * cp * 27, device=dev)
        xin = dict(x_f32=ops.f32view(x), in_affine=(sc, sh, True))
        f = timeit(lambda: ops.dwconv_fwd(geom, cp, c, wt, ops.f32view(y), stats, **xin))
        d = timeit(lambda: ops.dwconv_bwd(geom, cp, c, wt, ops.f32view(dy), None, None, dx=ops.f32view(dx), **xin))
        g = timeit(lambda: ops.dwconv_bwd(geom, cp, c, wt, ops.f32view(dy), dw, wp, **xin))
        tag = "v3" if mode else "v2"
        rec[tag + "_fwd_us"], rec[tag + "_dgrad_us"], rec[tag + "_wgrad_us"] = round(f, 1), round(d, 1), round(g, 1)
        rec[tag + "_fwd_gbs"] = round(2 * x.numel() * 4 / f / 1e3, 1)
        outs[mode] = (y.clone(), dx.clone(), dw.clone(), stats.sum(2).clone())
    rec["diff"] = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(outs[1], outs[0])]
    print(json.dumps(rec), flush=True)
    recs.append(rec)
lib.sfb_set_dw3(1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(recs, open(os.path.join(ROOT, "gpurun_out", "dw_probe.json"), "w"), indent=1)
