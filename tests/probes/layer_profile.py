"""Per-launch timing (CUDA events, eager mode) of every implicit-GEMM launch of one SlowFast train step, with shapes.
Writes gpurun_out/layer_profile.json.  Usage: python tests/probes/layer_profile.py [batch] [nsplit]"""
import json, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slowfast_b200 import ops
from slowfast_b200.config import get_cfg
from slowfast_b200.nets.resnet import B200SlowFast

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nsplit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = get_cfg("SLOWFAST_8x8_R50", B200={"NSPLIT": nsplit, "CUDA_GRAPH": False})
torch.manual_seed(0)
model = B200SlowFast(cfg).cuda().train()
T, A = cfg.DATA.NUM_FRAMES, cfg.SLOWFAST.ALPHA
clip = torch.randn(B, 3, T, 224, 224, device="cuda")
idx = torch.linspace(0, T - 1, T // A).long().cuda()
x = [clip.index_select(2, idx).contiguous(), clip]
y = torch.randint(0, 400, (B,), device="cuda")

def step():
    model.zero_grad(set_to_none=True)
    F.cross_entropy(model(x), y).backward()

step(); step()
recs = []
oc, ow = ops.conv_igemm, ops.conv_wgrad
def conv(xp, f, geom, out, strides, **k):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = oc(xp, f, geom, out, strides, **k); e.record()
    m = xp.n * geom.out[0] * geom.out[1] * geom.out[2]
    recs.append(dict(kind="conv", M=m, N=f.rows, K=f.ntaps * f.cols_pad, C=xp.c, taps=f.ntaps, k=list(geom.k),
                     stride=list(geom.stride), acc=bool(k.get("accumulate", False)), ev=(s, e)))
    return r
def wgrad(xp, dy, geom, dwm, **k):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = ow(xp, dy, geom, dwm, **k); e.record()
    taps = geom.k[0] * geom.k[1] * geom.k[2]
    recs.append(dict(kind="wgrad", M=dy.rows, N=dy.c, K=taps * xp.c, C=xp.c, taps=taps, k=list(geom.k),
                     stride=list(geom.stride), ev=(s, e)))
    return r
ops.conv_igemm, ops.conv_wgrad = conv, wgrad
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record(); step(); s1.record()
torch.cuda.synchronize()
ops.conv_igemm, ops.conv_wgrad = oc, ow
out = []
for r in recs:
    s, e = r.pop("ev")
    r["us"] = s.elapsed_time(e) * 1e3
    r["gflop"] = 2.0 * r["M"] * r["N"] * r["K"] / 1e9
    r["tflops"] = r["gflop"] / r["us"] * 1e-3 * 1e3 / 1e3 * 1e3 if r["us"] > 0 else 0
    r["tflops"] = r["gflop"] * 1e9 / (r["us"] * 1e-6) / 1e12
    out.append(r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(step_ms=s0.elapsed_time(s1), batch=B, nsplit=nsplit, launches=out),
          open(os.path.join(ROOT, "gpurun_out", f"layer_profile_b{B}_n{nsplit}.json"), "w"))
print("step ms", s0.elapsed_time(s1), "conv launches", len(out))
