"""GPU probe: run the implicit-GEMM conv kernel over a matrix of shapes and compare with torch (fp64 on the same
split operands).  Each case runs in a child process so that a trapping kernel cannot take the rest of the sweep
down.  Usage (on a GPU box):  python tests/probes/conv_probe.py [--out gpurun_out/conv_probe.jsonl]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

# name, n, t, h, w, cin, cout, kernel(t,h,w), stride, pad, extra
CASES = [
    dict(name="1x1_c64_o64", n=2, t=4, h=14, w=14, cin=64, cout=64, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0)),
    dict(name="1x1_c256_o512", n=2, t=4, h=14, w=14, cin=256, cout=512, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0)),
    dict(name="1x3x3_c64_o64", n=2, t=4, h=14, w=14, cin=64, cout=64, k=(1, 3, 3), s=(1, 1, 1), p=(0, 1, 1)),
    dict(name="1x3x3_s2_c128", n=2, t=4, h=28, w=28, cin=128, cout=128, k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1)),
    dict(name="3x1x1_c64_o64", n=2, t=8, h=14, w=14, cin=64, cout=64, k=(3, 1, 1), s=(1, 1, 1), p=(1, 0, 0)),
    dict(name="1x3x3_c32", n=2, t=4, h=14, w=14, cin=32, cout=32, k=(1, 3, 3), s=(1, 1, 1), p=(0, 1, 1)),
    dict(name="1x3x3_c16", n=2, t=4, h=14, w=14, cin=16, cout=16, k=(1, 3, 3), s=(1, 1, 1), p=(0, 1, 1)),
    dict(name="1x3x3_c8", n=2, t=4, h=14, w=14, cin=8, cout=8, k=(1, 3, 3), s=(1, 1, 1), p=(0, 1, 1)),
    dict(name="3x1x1_c8_o8", n=2, t=8, h=14, w=14, cin=8, cout=8, k=(3, 1, 1), s=(1, 1, 1), p=(1, 0, 0)),
    dict(name="1x1_c80_o64", n=2, t=4, h=14, w=14, cin=80, cout=64, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0)),
    dict(name="1x1_c64_o24", n=2, t=4, h=14, w=14, cin=64, cout=24, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0)),
    dict(name="fuse_7x1x1_s4", n=2, t=32, h=7, w=7, cin=32, cout=64, k=(7, 1, 1), s=(4, 1, 1), p=(3, 0, 0)),
    dict(name="stem_1x7x7_s2_c8", n=1, t=2, h=32, w=32, cin=8, cout=64, k=(1, 7, 7), s=(1, 2, 2), p=(0, 3, 3)),
    dict(name="1x1_s2_c64", n=2, t=4, h=14, w=14, cin=64, cout=128, k=(1, 1, 1), s=(1, 2, 2), p=(0, 0, 0)),
    dict(name="big_1x1_c1024_o256", n=4, t=8, h=14, w=14, cin=1024, cout=256, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0)),
    dict(name="slice_acc_view", n=2, t=4, h=14, w=14, cin=64, cout=32, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0),
         view=True),
]


def run_case(case: dict, nsplit: int) -> dict:
    import torch
    from slowfast_b200 import lib as L
    import ctypes as C

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = L.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)
    n, t, h, w, cin, cout = (case[k] for k in ("n", "t", "h", "w", "cin", "cout"))
    kt, kh, kw = case["k"]
    st, sh, sw = case["s"]
    pt, ph, pw = case["p"]
    x = torch.randn(n, t, h, w, cin, generator=g).to(dev)            # NDHWC
    wt = (torch.randn(cout, kt, kh, kw, cin, generator=g) / (cin * kt * kh * kw) ** 0.5).to(dev)
    x_hi = x.bfloat16()
    x_lo = (x - x_hi.float()).bfloat16()
    w_hi = wt.bfloat16()
    w_lo = (wt - w_hi.float()).bfloat16()
    ot = (t + 2 * pt - (kt - 1) - 1) // st + 1
    oh = (h + 2 * ph - (kh - 1) - 1) // sh + 1
    ow = (w + 2 * pw - (kw - 1) - 1) // sw + 1
    view = case.get("view", False)
    ctot = cout + 16 if view else cout
    out_full = torch.full((n, ot, oh, ow, ctot), 1.0, device=dev, dtype=torch.float32) if view else \
        torch.empty((n, ot, oh, ow, ctot), device=dev, dtype=torch.float32)
    out_view_ptr = out_full.data_ptr() + (16 * 4 if view else 0)

    d = L.ConvDesc()
    d.a_hi, d.a_lo = x_hi.data_ptr(), x_lo.data_ptr()
    d.n, d.d, d.h, d.w, d.c, d.c_pitch = n, t, h, w, cin, cin
    d.b_hi, d.b_lo = w_hi.data_ptr(), w_lo.data_ptr()
    d.cout = cout
    d.kt, d.kh, d.kw = kt, kh, kw
    d.dil_t = d.dil_h = d.dil_w = 1
    d.str_t, d.str_h, d.str_w = st, sh, sw
    d.low_t, d.low_h, d.low_w = -pt, -ph, -pw
    d.out_t, d.out_h, d.out_w = ot, oh, ow
    d.out = out_view_ptr
    d.os_w = ctot
    d.os_h = ctot * ow
    d.os_t = ctot * ow * oh
    d.os_n = ctot * ow * oh * ot
    d.accumulate = 1 if view else 0
    d.nsplit = nsplit
    m_tiles = lib.sfb_conv_m_tiles(C.byref(d))
    stats = torch.zeros(2, cout, m_tiles, device=dev, dtype=torch.float32)
    d.stats = stats.data_ptr()
    rc = lib.sfb_conv_igemm(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        return dict(ok=False, err=lib.sfb_last_error().decode())
    torch.cuda.synchronize()

    # reference in fp64 on exactly the operands the kernel saw
    if nsplit == 3:
        xr = x_hi.double() + x_lo.double()
        wr = w_hi.double() + w_lo.double()
    else:
        xr = x_hi.double()
        wr = w_hi.double()
    ref = torch.nn.functional.conv3d(xr.permute(0, 4, 1, 2, 3), wr.permute(0, 4, 1, 2, 3), stride=(st, sh, sw),
                                     padding=(pt, ph, pw)).permute(0, 2, 3, 4, 1)
    got = out_full[..., 16:] if view else out_full
    if view:
        ref = ref + 1.0
        untouched = bool((out_full[..., :16] == 1.0).all().item())
    else:
        untouched = True
    err = (got.double() - ref).abs()
    scale = ref.abs().max().item()
    res = dict(ok=True, max_abs=err.max().item(), ref_max=scale, rel=err.max().item() / max(scale, 1e-30),
               untouched=untouched, m_tiles=int(m_tiles))
    if not view:
        ssum = stats[0].double().sum(1)
        ssq = stats[1].double().sum(1)
        rs = ref.reshape(-1, cout)
        res["stat_sum_rel"] = ((ssum - rs.sum(0)).abs().max() / rs.sum(0).abs().max()).item()
        res["stat_sq_rel"] = ((ssq - (rs * rs).sum(0)).abs().max() / (rs * rs).sum(0).abs().max()).item()
    if res["rel"] > 1e-3:
        # error pattern: which rows / columns are wrong
        bad = err > 1e-3 * scale
        rows_bad = bad.reshape(-1, cout).any(1)
        cols_bad = bad.reshape(-1, cout).any(0)
        res["bad_rows"] = int(rows_bad.sum().item())
        res["bad_cols"] = int(cols_bad.sum().item())
        res["first_bad_rows"] = rows_bad.nonzero().flatten()[:12].tolist()
        res["first_bad_cols"] = cols_bad.nonzero().flatten()[:12].tolist()
        res["got_sample"] = got.reshape(-1, cout)[:2, :4].tolist()
        res["ref_sample"] = ref.reshape(-1, cout)[:2, :4].tolist()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "conv_probe.jsonl"))
    ap.add_argument("--child-from", type=int, default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    jobs = [(c, ns) for c in CASES if not args.only or args.only in c["name"] for ns in (1, 3)]
    if args.child_from is not None:
        # worker: run jobs sequentially from the given index; a trap kills this process and the parent resumes
        for i in range(args.child_from, len(jobs)):
            case, ns = jobs[i]
            print(f"BEGIN {i}", flush=True)
            try:
                res = run_case(case, ns)
            except Exception as e:  # noqa: BLE001 - report and let the parent restart (context may be dead)
                print(f"RESULT {i} " + json.dumps(dict(ok=False, err=repr(e)[:1500])), flush=True)
                sys.exit(3)
            print(f"RESULT {i} " + json.dumps(res), flush=True)
        return
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    done = {}
    nxt = 0
    with open(args.out, "w") as f:
        while nxt < len(jobs):
            cmd = [sys.executable, __file__, "--child-from", str(nxt)]
            if args.only:
                cmd += ["--only", args.only]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
                out, err = r.stdout, r.stderr
            except subprocess.TimeoutExpired as e:
                out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                err = "timeout"
            last_begin = nxt - 1
            for line in out.splitlines():
                if line.startswith("BEGIN "):
                    last_begin = int(line.split()[1])
                elif line.startswith("RESULT "):
                    _, idx, payload = line.split(" ", 2)
                    done[int(idx)] = json.loads(payload)
            if last_begin >= 0 and last_begin not in done:
                done[last_begin] = dict(ok=False, err="worker died", stderr=err[-1500:])
            for i in range(nxt, max(done) + 1 if done else nxt):
                if i in done:
                    res = dict(done[i])
                    res.update(case=jobs[i][0]["name"], nsplit=jobs[i][1])
                    f.write(json.dumps(res) + "\n")
                    f.flush()
                    print(json.dumps(res), flush=True)
            new_nxt = (max(done) + 1) if done else nxt + 1
            if new_nxt <= nxt:
                new_nxt = nxt + 1
            nxt = new_nxt


if __name__ == "__main__":
    main()
