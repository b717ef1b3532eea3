"""One eager train step of a model family under ncu's profiler range (cudaProfilerStart/Stop), after warm-up steps.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/x.csv python tests/probes/ncu_step.py slowfast [batch] [nsplit]
"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slowfast_b200.config import get_cfg

which = sys.argv[1] if len(sys.argv) > 1 else "slowfast"
B = int(sys.argv[2]) if len(sys.argv) > 2 else {"slowfast": 8, "mvit": 4, "x3d": 16, "mvit_b": 1}[which]
nsplit = int(sys.argv[3]) if len(sys.argv) > 3 else 3
over = {"NSPLIT": nsplit, "CUDA_GRAPH": False}
if which == "slowfast":
    from slowfast_b200.nets.resnet import B200SlowFast as M
    cfg = get_cfg("SLOWFAST_8x8_R50", B200=over)
elif which == "mvit":
    from slowfast_b200.nets.mvit import B200MViT as M
    cfg = get_cfg("MVITv2_S_16x4", B200=over)
elif which == "x3d":
    from slowfast_b200.nets.x3d import B200X3D as M
    cfg = get_cfg("X3D_M", B200=over)
else:
    raise SystemExit(which)
torch.manual_seed(0)
model = M(cfg).cuda().train()
model.cuda_graphs = False
T = cfg.DATA.NUM_FRAMES
clip = torch.randn(B, 3, T, 224, 224, device="cuda")
if which == "slowfast":
    idx = torch.linspace(0, T - 1, T // cfg.SLOWFAST.ALPHA).long().cuda()
    x = [clip.index_select(2, idx).contiguous(), clip]
else:
    x = [clip]
y = torch.randint(0, 400, (B,), device="cuda")


def step():
    model.zero_grad(set_to_none=True)
    F.cross_entropy(model(x), y).backward()


step(); step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", which, B, nsplit)
