"""GPU probe: per-stage activation error and per-parameter gradient error of the engine vs the oracle."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as TO
from slowfast_b200.config import get_cfg
from slowfast_b200.nets.resnet import B200SlowFast

def run(crop, frames, batch, nsplit, out):
    cfg = get_cfg("SLOWFAST_8x8_R50", DATA={"NUM_FRAMES": frames, "TRAIN_CROP_SIZE": crop}, MODEL={"DROPOUT_RATE": 0.0},
                  B200={"NSPLIT": nsplit})
    torch.manual_seed(0)
    model = B200SlowFast(cfg)
    state = TO.fixture_state(model.state_dict(), 123)
    model.load_state_dict(state)
    inputs = TO.synthetic_inputs(cfg, batch, 321)
    dlogits = torch.randn(batch, 400, generator=torch.Generator().manual_seed(9))
    rec = {}
    work = {k: v.clone() for k, v in state.items()}
    TO.slowfast_forward(cfg, work, inputs, True, record=rec)
    o_logits, o_grads = TO.forward_backward(cfg, state, inputs, dlogits)
    model = model.cuda().train()
    logits = model([t.cuda() for t in inputs])
    logits.backward(dlogits.cuda())
    torch.cuda.synchronize()
    res = dict(crop=crop, frames=frames, batch=batch, nsplit=nsplit)
    res["logits_rel_l2"] = ((logits.detach().cpu() - o_logits).norm() / o_logits.norm()).item()
    res["logits_rel_max"] = ((logits.detach().cpu() - o_logits).abs().max() / o_logits.abs().max()).item()
    for i, (slow, fast, cs) in enumerate(model._trace):
        key = f"s{i + 1}"
        rs, rf = rec[key]
        es = slow.planes.to_float().cpu().permute(0, 4, 1, 2, 3)
        ef = fast.planes.to_float().cpu().permute(0, 4, 1, 2, 3)
        res[key + "_slow"] = ((es - rs).norm() / rs.norm()).item()
        res[key + "_fast"] = ((ef - rf).norm() / rf.norm()).item()
    per = {k: ((p.grad.cpu() - o_grads[k]).norm() / o_grads[k].norm().clamp_min(1e-20)).item()
           for k, p in model.named_parameters()}
    res["grad_median"] = sorted(per.values())[len(per) // 2]
    res["grad_worst"] = sorted(per.items(), key=lambda kv: -kv[1])[:12]
    # error by depth: first conv of each stage
    for k in ["head.projection.weight", "s5.pathway0_res2.branch2.c.weight", "s5.pathway0_res0.branch2.a.weight",
              "s4.pathway0_res0.branch2.a.weight", "s3.pathway0_res0.branch2.a.weight", "s2.pathway0_res0.branch2.a.weight",
              "s2.pathway1_res0.branch2.a.weight", "s1_fuse.conv_f2s.weight", "s1.pathway0_stem.conv.weight",
              "s1.pathway1_stem.conv.weight", "s5.pathway0_res2.branch2.c_bn.weight", "s2.pathway0_res0.branch2.a_bn.bias"]:
        res["g:" + k] = per[k]
    print(json.dumps(res), flush=True)
    out.write(json.dumps(res) + "\n"); out.flush()

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "model_probe.jsonl"), "w") as f:
        run(64, 16, 3, 3, f)
        run(64, 16, 3, 1, f)
        run(224, 32, 2, 3, f)
        run(224, 32, 2, 1, f)
