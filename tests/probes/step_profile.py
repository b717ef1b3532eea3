"""Run a few EAGER (no CUDA graph) train steps of SlowFast-8x8-R50; the step between cudaProfilerStart/Stop is the one
ncu records (`ncu --profile-from-start off ...`).  Usage: python tests/probes/step_profile.py [batch] [nsplit]"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
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
    loss = F.cross_entropy(model(x), y)
    loss.backward()

step(); step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, batch", B, "nsplit", nsplit)
