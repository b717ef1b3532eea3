"""GPU probe: MViTv2-S engine vs oracle - per-block activation error and per-parameter gradient error."""
import json, os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as TO
from slowfast_b200.config import get_cfg
from slowfast_b200.nets.mvit import B200MViT

def run(crop, frames, batch, nsplit, out, gentle=False):
    cfg = get_cfg("MVITv2_S_16x4", DATA={"NUM_FRAMES": frames, "TRAIN_CROP_SIZE": crop, "TEST_CROP_SIZE": crop},
                  MODEL={"DROPOUT_RATE": 0.0}, MVIT={"DROPPATH_RATE": 0.0}, B200={"NSPLIT": nsplit, "CUDA_GRAPH": False})
    torch.manual_seed(0)
    model = B200MViT(cfg)
    state = TO.fixture_state(model.state_dict(), 123)
    model.load_state_dict(state)
    inputs = TO.synthetic_inputs(cfg, batch, 321)
    dlogits = torch.randn(batch, 400, generator=torch.Generator().manual_seed(9))
    rec = {}
    TO.mvit_forward(cfg, {k: v.clone() for k, v in state.items()}, inputs, True, record=rec)
    o_logits, o_grads = TO.forward_backward(cfg, state, inputs, dlogits)
    model = model.cuda().train()
    res = dict(crop=crop, frames=frames, batch=batch, nsplit=nsplit)
    try:
        logits = model([t.cuda() for t in inputs])
        torch.cuda.synchronize()
        res["logits_rel_l2"] = ((logits.detach().cpu() - o_logits).norm() / o_logits.norm()).item()
        for i in range(len(model.blocks)):
            x = model.ctx._bufs[("x", i + 1)].cpu()
            r = rec[f"b{i}"]
            res[f"b{i}"] = ((x - r).norm() / r.norm()).item()
        logits.backward(dlogits.cuda())
        torch.cuda.synchronize()
        per = {k: ((p.grad.cpu() - o_grads[k]).norm() / o_grads[k].norm().clamp_min(1e-6 * 1)).item()
               for k, p in model.named_parameters()}
        norms = sorted(v.norm().item() for v in o_grads.values())
        floor = 1e-2 * norms[len(norms) // 2]
        per = {k: ((p.grad.cpu() - o_grads[k]).norm() / o_grads[k].norm().clamp_min(floor)).item()
               for k, p in model.named_parameters()}
        res["grad_median"] = sorted(per.values())[len(per) // 2]
        res["grad_worst"] = sorted(per.items(), key=lambda kv: -kv[1])[:25]
        res["grad_best_examples"] = sorted(per.items(), key=lambda kv: kv[1])[:5]
    except Exception as e:  # noqa: BLE001
        res["error"] = repr(e)[:600] + traceback.format_exc()[-1500:]
    print(json.dumps(res)[:6000], flush=True)
    out.write(json.dumps(res) + "\n"); out.flush()

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mvit_probe.jsonl"), "w") as f:
        run(64, 8, 2, 3, f)
        if len(sys.argv) > 1:
            run(224, 16, 1, 3, f)
