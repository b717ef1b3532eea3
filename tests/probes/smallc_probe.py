"""A/B of the two bodies of sfb_conv_igemm (tensor-core implicit GEMM vs fp32 SIMT, csrc/conv_direct.cu) on the narrow layers
of a SlowFast-8x8-R50 step (fast pathway, B=8) and X3D-M's 1x1x1 layers: per-layer time, achieved GB/s on the algorithmic
bytes, max |difference| between the two outputs.  Writes gpurun_out/smallc_probe.json."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slowfast_b200 import lib as L, ops

lib = L.load()
dev = torch.device("cuda")
# (name, n, t, h, w, cin, cout, k, stride, pad)
LAYERS = [
    ("fast res2 c  8->32 1x1x1", 8, 32, 56, 56, 8, 32, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("fast res2 a 32->8 3x1x1", 8, 32, 56, 56, 32, 8, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("fast res2 b  8->8 1x3x3", 8, 32, 56, 56, 8, 8, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("fast res2 a dgrad 8->32 3x1x1", 8, 32, 56, 56, 8, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("fast res2 s  8->32 1x1x1 (shortcut)", 8, 32, 56, 56, 8, 32, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("fast res3 a 32->16 3x1x1", 8, 32, 56, 56, 32, 16, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("fast res3 b 16->16 1x3x3 s2", 8, 32, 56, 56, 16, 16, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("fast res3 c 16->64 1x1x1", 8, 32, 28, 28, 16, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("fast res3 a 64->16 3x1x1", 8, 32, 28, 28, 64, 16, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("fast res4 b 32->32 1x3x3", 8, 32, 14, 14, 32, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("fuse1 8->16 7x1x1 s4", 8, 32, 56, 56, 8, 16, (7, 1, 1), (4, 1, 1), (3, 0, 0)),
    ("x3d res2 a 24->56(54) 1x1x1", 16, 16, 56, 56, 24, 54, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("x3d res2 c 56->24 1x1x1", 16, 16, 56, 56, 56, 24, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
]


def run(name, n, t, h, w, cin, cout, k, stride, pad, nsplit=3):
    torch.manual_seed(0)
    x = ops.alloc_planes(n, t, h, w, cin, nsplit, dev)
    xf = torch.randn(n, t, h, w, cin, device=dev)
    ops.split_planes(xf, x)
    wt = torch.randn(cout, cin, *k, device=dev) * 0.1
    taps = k[0] * k[1] * k[2]
    fm = ops.alloc_filter(cout, taps, cin, nsplit, dev)
    ops.filter_pack(wt, fm)
    geom = ops.fprop_geom(x, k, stride, pad)
    ot, oh, ow = geom.out
    cp = ops.pad8(cout)
    m_tiles = ops.conv_m_tiles(n, geom)
    res = {}
    outs = {}
    for mode in (0, 1):
        lib.sfb_set_simt_smallc(mode, 1 << 20)
        y = torch.zeros(n, ot, oh, ow, cp, device=dev)
        stats = torch.zeros(2, cout, m_tiles, device=dev)
        strides = (ot * oh * ow * cp, oh * ow * cp, ow * cp, cp)
        for _ in range(3):
            ops.conv_igemm(x, fm, geom, y, strides, stats=stats, nsplit=nsplit)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv_igemm(x, fm, geom, y, strides, stats=stats, nsplit=nsplit)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 20 * 1e3
        outs[mode] = (y.clone(), stats.clone())
    lib.sfb_set_simt_smallc(0, 0)
    M = n * ot * oh * ow
    by = x.rows * cin * 4 + M * cp * 4
    dy = (outs[0][0] - outs[1][0]).abs().max().item() / outs[0][0].abs().max().item()
    ds = (outs[0][1] - outs[1][1]).abs().max().item() / outs[0][1].abs().max().item()
    rec = dict(layer=name, M=M, cin=cin, cout=cout, taps=taps, macs=taps * cin * cout, tc_us=round(res[0], 1),
               simt_us=round(res[1], 1), tc_gbs=round(by / res[0] / 1e3, 1), simt_gbs=round(by / res[1] / 1e3, 1),
               out_rel_diff=dy, stats_rel_diff=ds)
    print(json.dumps(rec), flush=True)
    return rec


recs = [run(*l) for l in LAYERS]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(recs, open(os.path.join(ROOT, "gpurun_out", "smallc_probe.json"), "w"), indent=1)
