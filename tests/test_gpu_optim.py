"""GPU: the fused optimizer step on the flat gradient bucket (csrc/optim.cu, slowfast_b200/optim.py)
against torch.optim.SGD(nesterov) / AdamW and torch's clip_grad_norm_ - the optimizers the reference's
construct_optimizer builds (slowfast/models/optimizer.py:105-136) and the norm / clipping of tools/train_net.py:154-172."""
import types

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv3d(3, 5, (1, 3, 3), bias=False)       # 135 elements (not a multiple of 4)
        self.bn = nn.BatchNorm3d(5)
        self.fc = nn.Linear(70001, 3)                            # spans many 8192-element chunks + a ragged tail
        self.ctx = types.SimpleNamespace(flat_grad=None)


def _fill_bucket(model, grads):
    from slowfast_b200.engine import flat_offsets
    params = list(model.parameters())
    offs, total = flat_offsets(params)
    flat = torch.zeros(total, device=params[0].device)
    for p, o, g in zip(params, offs, grads):
        flat[o:o + p.numel()] = g.flatten()
    model.ctx.flat_grad = flat


@pytest.mark.parametrize("kind", ["sgd", "sgd_plain", "adamw"])
@pytest.mark.parametrize("clip", [0.0, 0.5])
def test_flat_optimizer_matches_torch(kind, clip, cuda_device):
    from slowfast_b200.optim import FlatOptimizer
    torch.manual_seed(0)
    a, b = _Toy().to(cuda_device), _Toy().to(cuda_device)
    b.load_state_dict(a.state_dict())

    def groups(m):
        return [dict(params=[m.bn.weight, m.bn.bias], weight_decay=0.0),
                dict(params=[m.conv.weight, m.fc.weight], weight_decay=1e-2),
                dict(params=[m.fc.bias], weight_decay=0.0)]

    if kind == "adamw":
        ref = torch.optim.AdamW(groups(b), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
        opt = FlatOptimizer(a, "adamw", groups(a), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, clip_grad_l2norm=clip)
    else:
        nest = kind == "sgd"
        ref = torch.optim.SGD(groups(b), lr=0.1, momentum=0.9, dampening=0.0, nesterov=nest, weight_decay=1e-4)
        opt = FlatOptimizer(a, "sgd", groups(a), lr=0.1, momentum=0.9, nesterov=nest, clip_grad_l2norm=clip)
    g = torch.Generator(device="cpu").manual_seed(1)
    for step in range(4):
        grads = [torch.randn(p.shape, generator=g).to(cuda_device) for p in a.parameters()]
        if step == 2:  # LR schedule changes the rate between steps
            opt.set_lr(0.03)
            for grp in ref.param_groups:
                grp["lr"] = 0.03
        _fill_bucket(a, grads)
        for p, gr in zip(b.parameters(), grads):
            p.grad = gr.clone()
        want_norm = torch.norm(torch.stack([torch.norm(p.grad, 2.0) for p in b.parameters()]), 2.0)
        if clip > 0:
            torch.nn.utils.clip_grad_norm_(b.parameters(), clip)
        ref.step()
        opt.step()
        assert torch.allclose(opt.grad_norm, want_norm, rtol=1e-5)
        for (n, p), q in zip(a.named_parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-7), (kind, clip, step, n, (p - q).abs().max().item())


def test_flat_optimizer_on_engine_model_matches_torch_sgd(cuda_device):
    """Whole loop: engine backward -> (flat bucket) -> FlatOptimizer with flat_grad_only (no param.grad copies) vs the same
    model stepping torch.optim.SGD on param.grad, 4 steps under CUDA-graph replay."""
    from oracle import torch_oracle as TO
    from slowfast_b200.config import get_cfg
    from slowfast_b200.nets.resnet_single import B200ResNet
    from slowfast_b200.optim import FlatOptimizer
    cfg = get_cfg("C2D_8x8_R50", DATA={"NUM_FRAMES": 8, "TRAIN_CROP_SIZE": 64}, MODEL={"DROPOUT_RATE": 0.0})
    models = []
    for _ in range(2):
        torch.manual_seed(0)
        m = B200ResNet(cfg)
        st = TO.fixture_state(m.state_dict(), 3)
        for k in st:   # weak residual branches + a small rate: the comparison is about the update rule, not about chaos
            if k.endswith("c_bn.weight"):
                st[k] = st[k] * 0.1
        m.load_state_dict(st)
        models.append(m.to(cuda_device).train())
    a, b = models
    a.flat_grad_only = True
    opt_a = FlatOptimizer(a, "sgd", lr=0.001, momentum=0.9, nesterov=True, weight_decay=1e-4)
    opt_b = torch.optim.SGD(b.parameters(), lr=0.001, momentum=0.9, nesterov=True, weight_decay=1e-4)
    for s in range(5):
        x = [t.to(cuda_device) for t in TO.synthetic_inputs(cfg, 2, 50 + s)]
        y = torch.randint(0, 400, (2,), generator=torch.Generator().manual_seed(s)).to(cuda_device)
        opt_a.zero_grad()
        la = torch.nn.functional.cross_entropy(a(x), y)
        la.backward()
        opt_a.step()
        opt_b.zero_grad(set_to_none=True)
        lb = torch.nn.functional.cross_entropy(b(x), y)
        lb.backward()
        opt_b.step()
        assert all(p.grad is None for p in a.parameters())
        assert abs(la.item() - lb.item()) < 5e-4 * abs(lb.item()), (s, la.item(), lb.item())
    worst = max(((p - q).abs().max() / q.abs().max().clamp_min(1e-12)).item() for p, q in zip(a.parameters(), b.parameters()))
    assert worst < 1e-3, worst


def test_allreduce_flat_c_abi(cuda_device):
    """sfb_allreduce_flat: the single in-place ncclAllReduce of the flat gradient bucket through the C
    ABI, on communicators created here with ncclCommInitAll (one per visible GPU, at most 2)."""
    import ctypes as C
    import os

    from slowfast_b200 import lib as L
    lib = L.load()
    import torch.distributed  # noqa: F401  (makes sure torch's bundled libnccl is loaded in this process)
    nccl = None
    cands = [os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib", "libnccl.so.2"),
             "libnccl.so.2"]
    for c in cands:
        try:
            nccl = C.CDLL(c, mode=C.RTLD_GLOBAL)
            break
        except OSError:
            continue
    if nccl is None:
        pytest.skip("libnccl.so.2 not found")
    ndev = min(2, torch.cuda.device_count())
    comms = (C.c_void_p * ndev)()
    devs = (C.c_int * ndev)(*range(ndev))
    assert nccl.ncclCommInitAll(comms, ndev, devs) == 0
    try:
        n = 1_000_003
        bufs = []
        for d in range(ndev):
            with torch.cuda.device(d):
                bufs.append(torch.full((n,), float(d + 1), device=f"cuda:{d}") + torch.arange(n, device=f"cuda:{d}") % 7)
        want_sum = sum(b.to("cuda:0") for b in bufs)
        for average in (1, 0):
            work = [b.clone() for b in bufs]
            nccl.ncclGroupStart()
            for d in range(ndev):
                with torch.cuda.device(d):
                    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                    L.check(lib.sfb_allreduce_flat(work[d].data_ptr(), n, comms[d], average, st), "sfb_allreduce_flat")
            nccl.ncclGroupEnd()
            for d in range(ndev):
                torch.cuda.synchronize(d)
            want = want_sum / ndev if average else want_sum
            for d in range(ndev):
                assert torch.allclose(work[d].to("cuda:0"), want, rtol=1e-6)
    finally:
        for d in range(ndev):
            nccl.ncclCommDestroy(C.c_void_p(comms[d]))
