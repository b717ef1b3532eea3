"""Quick X3D-M throughput probe (fwd+bwd+SGD).  Usage: python tests/probes/x3d_bench.py [batch] [nsplit] [graphs 0|1]"""
import os, sys, json
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slowfast_b200.config import get_cfg
from slowfast_b200.nets.x3d import B200X3D
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 3
graphs = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
cfg = get_cfg("X3D_M", B200={"NSPLIT": ns, "CUDA_GRAPH": graphs})
torch.manual_seed(0)
model = B200X3D(cfg).cuda().train()
x = [torch.randn(B, 3, 16, 224, 224, device="cuda")]
y = torch.randint(0, 400, (B,), device="cuda")
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=5e-5)
def step():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(x), y); loss.backward(); opt.step(); return loss
for _ in range(4): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps(dict(model="X3D-M", batch=B, nsplit=ns, graphs=graphs, ms_per_step=ms, clips_per_s=B / ms * 1e3, loss=l.item(),
                      mem_gb=torch.cuda.max_memory_allocated() / 2**30)))
