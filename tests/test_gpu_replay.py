"""GPU parity of the path bench.py actually times: CUDA-graph REPLAY of the forward / backward programs.

``ModelFunction`` switches to graph replay on the third call with the same input signature (engine.py); every other
test file runs one eager step.  Here each model family runs several optimizer steps with graphs on and is compared
  (a) with the same steps executed eagerly (graphs off): logits, every parameter gradient, BN running statistics,
  (b) with the oracle (CPU fp32 restatement of the reference) stepping the same SGD,
and the arena logic (one buffer set per input signature) is exercised by alternating batch sizes and train / eval.
Split-K weight gradients are reduced with floating-point atomics, so eager and replay agree to rounding, not bit-wise.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

FAMILIES = {
    # name: (preset, overrides, batch)
    "slowfast": ("SLOWFAST_8x8_R50", dict(DATA={"NUM_FRAMES": 16, "TRAIN_CROP_SIZE": 64, "TEST_CROP_SIZE": 64},
                                          MODEL={"DROPOUT_RATE": 0.0}), 2),
    "c2d": ("C2D_8x8_R50", dict(DATA={"NUM_FRAMES": 8, "TRAIN_CROP_SIZE": 64, "TEST_CROP_SIZE": 64},
                                MODEL={"DROPOUT_RATE": 0.0}), 2),
    "x3d": ("X3D_M", dict(DATA={"NUM_FRAMES": 4, "TRAIN_CROP_SIZE": 64, "TEST_CROP_SIZE": 64},
                          MODEL={"DROPOUT_RATE": 0.0}), 2),
    "mvit": ("MVITv2_S_16x4", dict(DATA={"NUM_FRAMES": 8, "TRAIN_CROP_SIZE": 64, "TEST_CROP_SIZE": 64},
                                   MODEL={"DROPOUT_RATE": 0.0}, MVIT={"DROPPATH_RATE": 0.0}), 2),
}


def _cls(cfg):
    name = cfg.MODEL.MODEL_NAME
    if name == "SlowFast":
        from slowfast_b200.nets.resnet import B200SlowFast as M
    elif name == "MViT":
        from slowfast_b200.nets.mvit import B200MViT as M
    elif name == "X3D":
        from slowfast_b200.nets.x3d import B200X3D as M
    else:
        from slowfast_b200.nets.resnet_single import B200ResNet as M
    return M


def _build(family, graphs, dev, seed=7, gentle=True):
    from oracle import torch_oracle as TO
    from slowfast_b200.config import get_cfg
    preset, over, batch = FAMILIES[family]
    over = dict(over)
    over["B200"] = {"NSPLIT": 3, "CUDA_GRAPH": graphs}
    cfg = get_cfg(preset, **over)
    torch.manual_seed(0)
    model = _cls(cfg)(cfg)
    state = TO.fixture_state(model.state_dict(), seed)
    if gentle:  # weak residual branches: no chaotic amplification, (almost) no ReLU mask flips between implementations
        for k in state:
            if k.endswith("c_bn.weight"):
                state[k] = state[k] * 0.1
    model.load_state_dict(state)
    return cfg, model.to(dev).train(), state, batch


def _inputs(cfg, batch, seed):
    from oracle import torch_oracle as TO
    x = TO.synthetic_inputs(cfg, batch, seed)
    y = torch.randint(0, cfg.MODEL.NUM_CLASSES, (batch,), generator=torch.Generator().manual_seed(seed + 5))
    return x, y


def _train_steps(model, cfg, batch, dev, n_steps, lr=0.002):
    """n_steps of SGD on per-step DIFFERENT inputs; returns per-step logits and the last step's gradients."""
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9)
    outs = []
    for s in range(n_steps):
        x, y = _inputs(cfg, batch, 100 + s)
        opt.zero_grad(set_to_none=True)
        logits = model([t.to(dev) for t in x])
        loss = torch.nn.functional.cross_entropy(logits, y.to(dev))
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        opt.step()
        outs.append(logits.detach().cpu())
    torch.cuda.synchronize()
    return outs, {k: v.cpu() for k, v in grads.items()}


@pytest.mark.parametrize("family", list(FAMILIES))
def test_replay_matches_eager_over_steps(family, cuda_device):
    steps = 5  # calls 1-2 eager warm-up, call 3 captures + replays, calls 4-5 pure replay
    cfg, mg, _, batch = _build(family, True, cuda_device)
    _, me, _, _ = _build(family, False, cuda_device)
    og, gg = _train_steps(mg, cfg, batch, cuda_device, steps)
    oe, ge = _train_steps(me, cfg, batch, cuda_device, steps)
    # noise floor: a SECOND eager model run the same way (split-K weight gradients use floating-point atomics, and SGD
    # steps amplify the rounding-level differences: measured 4-6e-4 on step 4 at lr 0.02 between any two runs)
    _, me2, _, _ = _build(family, False, cuda_device)
    oe2, ge2 = _train_steps(me2, cfg, batch, cuda_device, steps)
    key = [k for k in mg._graphs]
    assert len(key) == 1 and mg._graphs[key[0]].bwd_graph is not None, "the graphed model never switched to replay"
    assert not me._graphs
    for s in range(steps):
        rel = ((og[s] - oe[s]).abs().max() / oe[s].abs().max()).item()
        noise = ((oe2[s] - oe[s]).abs().max() / oe[s].abs().max()).item()
        print(f"{family}: step {s} logits replay vs eager {rel:.2e} (eager vs eager {noise:.2e})")
        assert rel < max(2e-4, 5 * noise), f"{family}: step {s} logits replay vs eager {rel} (eager vs eager {noise})"
    norms = sorted(v.norm().item() for v in ge.values())
    floor = 1e-2 * norms[len(norms) // 2]
    per = {k: ((gg[k] - ge[k]).norm() / ge[k].norm().clamp_min(floor)).item() for k in ge}
    per2 = {k: ((ge2[k] - ge[k]).norm() / ge[k].norm().clamp_min(floor)).item() for k in ge}
    worst = max(per.items(), key=lambda kv: kv[1])
    med, med2 = sorted(per.values())[len(per) // 2], sorted(per2.values())[len(per2) // 2]
    print(f"{family}: last-step gradients replay vs eager: median {med:.2e} worst {worst} (eager vs eager: median {med2:.2e} "
          f"worst {max(per2.values()):.2e}) - ReLU nets amplify the atomics' rounding noise over the SGD steps")
    # (two eager runs differ by 1e-3 ... 2e-2 median here depending on where the atomics' rounding flipped a ReLU: the
    # gradient comparison is held to the mask-flip plateau the other model tests use, the logits above to the noise itself)
    assert med < max(5e-2, 3 * med2) and worst[1] < max(0.3, 3 * max(per2.values()))
    sg, se, se2 = mg.state_dict(), me.state_dict(), me2.state_dict()
    for k in sg:
        if "running_" in k:
            tol = max(1e-4 * se[k].abs().max().item(), 5 * (se2[k] - se[k]).abs().max().item())
            assert (sg[k] - se[k]).abs().max().item() <= tol, (k, (sg[k] - se[k]).abs().max().item(), tol)
        if k.endswith("num_batches_tracked"):
            assert int(sg[k]) == steps == int(se[k]), k


@pytest.mark.parametrize("family", ["slowfast", "mvit"])
def test_replay_matches_oracle_over_steps(family, cuda_device):
    """4 SGD steps under graph replay vs 4 SGD steps of the oracle (autograd on the CPU) from the same state: the
    step-4 logits see every forward, backward, BN-statistics and parameter-update of the first three steps."""
    from oracle import torch_oracle as TO
    steps, lr = 4, 0.002
    cfg, mg, state, batch = _build(family, True, cuda_device)
    og, _ = _train_steps(mg, cfg, batch, cuda_device, steps, lr)
    sd = {k: v.clone() for k, v in state.items()}
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    mom = {k: torch.zeros_like(sd[k]) for k in names}
    for s in range(steps):
        x, y = _inputs(cfg, batch, 100 + s)
        leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
        work = dict(sd)
        work.update(leaves)
        logits = TO.forward(cfg, work, x, True)     # updates the running statistics in ``work`` in place
        loss = torch.nn.functional.cross_entropy(logits, y)
        grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
        for k, g in zip(names, grads):
            if g is None:
                continue
            mom[k] = 0.9 * mom[k] + g
            sd[k] = sd[k] - lr * mom[k]
        for k in sd:
            if "running_" in k:
                sd[k] = work[k]
        rel = ((og[s] - logits.detach()).abs().max() / logits.detach().abs().max()).item()
        print(f"{family}: step {s} logits engine(replay from step 2) vs oracle {rel:.2e}")
        assert rel < 1e-3, f"step {s}: {rel}"
        assert torch.equal(og[s].argmax(1), logits.detach().argmax(1))
    new = mg.state_dict()
    for k in new:
        if "running_" in k:
            assert (new[k].cpu() - sd[k]).abs().max().item() < 2e-3 * sd[k].abs().max().item(), k


def test_alternating_signatures_do_not_corrupt_captured_programs(cuda_device):
    """buffers used to be keyed without the input shape, so a forward at another batch size freed the
    memory a captured graph replays into.  Alternate train B=2 / eval B=3 / train B=1 / eval B=2 with graphs on and
    compare every output with an eager model fed the same sequence."""
    cfg, mg, _, _ = _build("slowfast", True, cuda_device)
    _, me, _, _ = _build("slowfast", False, cuda_device)
    seq = [("train", 2), ("train", 2), ("train", 2), ("eval", 3), ("eval", 3), ("train", 2), ("eval", 3), ("train", 1),
           ("train", 2), ("eval", 3), ("eval", 2), ("train", 1), ("train", 1), ("train", 2), ("eval", 3)]

    def run(model):
        outs = []
        opt = torch.optim.SGD(model.parameters(), lr=0.002)
        for i, (mode, b) in enumerate(seq):
            x, y = _inputs(cfg, b, 300 + i)
            xd = [t.to(cuda_device) for t in x]
            if mode == "train":
                model.train()
                opt.zero_grad(set_to_none=True)
                out = model(xd)
                torch.nn.functional.cross_entropy(out, y.to(cuda_device)).backward()
                g = torch.cat([p.grad.flatten() for p in model.parameters()]).norm().item()
                opt.step()
                outs.append((out.detach().cpu(), g))
            else:
                model.eval()
                with torch.no_grad():
                    outs.append((model(xd).cpu(), 0.0))
        torch.cuda.synchronize()
        return outs

    a, b = run(mg), run(me)
    assert len(mg._graphs) >= 2, "expected captured programs for at least two signatures"
    for i, ((oa, ga), (ob, gb)) in enumerate(zip(a, b)):
        rel = ((oa - ob).abs().max() / ob.abs().max()).item()
        assert rel < 1e-3, f"call {i} {seq[i]}: outputs differ {rel}"   # (rounding noise amplified by the SGD steps: ~2e-4)
        # (the two models differ by the order of float atomics in the weight-gradient reductions; the batch-1 train steps of
        # this sequence amplify that: measured up to 1.4e-3 on the gradient norm at call 12)
        assert abs(ga - gb) <= 5e-3 * max(gb, 1e-12), f"call {i} {seq[i]}: gradient norms {ga} vs {gb}"


def test_two_forwards_before_backward_is_loud(cuda_device):
    """saved activations are per model; a second forward before the first one's backward must raise
    instead of silently differentiating the wrong activations."""
    cfg, m, _, batch = _build("c2d", False, cuda_device)
    x1, _ = _inputs(cfg, batch, 1)
    x2, _ = _inputs(cfg, batch, 2)
    o1 = m([t.to(cuda_device) for t in x1])
    o2 = m([t.to(cuda_device) for t in x2])
    with pytest.raises(RuntimeError, match="another forward"):
        o1.sum().backward()
    o2.sum().backward()  # the latest forward is still differentiable


def test_gradient_accumulation_under_replay(cuda_device):
    """zero_grad(set_to_none=False) + replayed static gradient slots must accumulate g1 + g2, not 2*g2."""
    cfg, m, _, batch = _build("c2d", True, cuda_device)
    xs = [_inputs(cfg, batch, 40 + i) for i in range(5)]

    def grads_of(i):
        m.zero_grad(set_to_none=True)
        x, y = xs[i]
        torch.nn.functional.cross_entropy(m([t.to(cuda_device) for t in x]), y.to(cuda_device)).backward()
        return [p.grad.detach().clone() for p in m.parameters()]

    for p in m.parameters():     # freeze the running statistics' influence: BN uses batch stats in train mode anyway
        p.requires_grad_(True)
    for i in range(3):
        grads_of(i)              # warm-up + capture
    g3, g4 = grads_of(3), grads_of(4)
    m.zero_grad(set_to_none=True)
    for i in (3, 4):             # accumulate two replayed backward passes into the same .grad
        x, y = xs[i]
        torch.nn.functional.cross_entropy(m([t.to(cuda_device) for t in x]), y.to(cuda_device)).backward()
    torch.cuda.synchronize()
    for p, a, b in zip(m.parameters(), g3, g4):
        want = a + b
        assert torch.allclose(p.grad, want, rtol=1e-3, atol=1e-6 * want.abs().max().item() + 1e-12)


def test_resnet_head_fully_convolutional_eval(cuda_device):
    """TEST_CROP_SIZE > TRAIN_CROP_SIZE: stride-1 window pooling, per-location projection +
    softmax, mean over locations (head_helper.py:305-350), against the oracle (pinned to the reference on this case
    in tests/test_oracle.py)."""
    from oracle import torch_oracle as TO
    for family in ("slowfast", "c2d"):
        cfg, m, state, _ = _build(family, False, cuda_device, gentle=False)
        x = TO.synthetic_inputs(cfg, 2, 9, crop=96)
        m.eval()
        with torch.no_grad():
            probs = m([t.to(cuda_device) for t in x]).cpu()
        ref = TO.forward(cfg, {k: v.clone() for k, v in state.items()}, x, False)
        assert probs.shape == ref.shape == (2, 400)
        assert ((probs - ref).abs().max() / ref.abs().max()).item() < 1e-3
        assert torch.equal(probs.argmax(1), ref.argmax(1))
        assert torch.allclose(probs.sum(1), torch.ones(2), atol=1e-5)
