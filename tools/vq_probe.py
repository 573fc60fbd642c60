#!/usr/bin/env python
"""Probe for the tensor-core VQ path: measured error of the bf16x3 dot products for several data distributions
(relative to |z||e|), and kernel time at the microbench shape (VQ_CLIPS clips)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
D, n = 64, 256
g = torch.Generator(device="cpu").manual_seed(1)
res = {}
if os.environ.get("VQ_ERR", "1") == "1":
    cases = {
        "normal": (torch.randn(4, 4096, D, generator=g), torch.randn(n, D, generator=g)),
        "uniform_pos": (torch.rand(4, 4096, D, generator=g), torch.rand(n, D, generator=g)),
        "logscale": (torch.randn(4, 4096, D, generator=g) * torch.logspace(-3, 3, 4096)[None, :, None],
                     torch.randn(n, D, generator=g) * torch.logspace(-2, 2, n)[:, None]),
        "ref_init": (torch.randn(4, 4096, D, generator=g), (torch.rand(n, D, generator=g) * 2 - 1) / n),
        "same_sign_heavy": (torch.randn(4, 4096, D, generator=g).abs() + 3, torch.randn(n, D, generator=g).abs() + 3),
    }
    for name, (z, cb) in cases.items():
        B, L, _ = z.shape
        acc = torch.empty(B * L, n, device=dev)
        lib.vq_quantize(z.to(dev), cb.to(dev), n, want_bdl=False, algo=lib.VQ_TENSOR, dbg_acc=acc)
        torch.cuda.synchronize()
        zd, ed = z.view(-1, D).double(), cb.double()
        exact = zd @ ed.t() - 0.5 * (cb.float() ** 2).sum(1).double()[None]
        scale = zd.norm(dim=1, keepdim=True) * ed.norm(dim=1)[None]
        absdot = zd.abs() @ ed.abs().t()
        err = (acc.cpu().double() - exact).abs()
        import math
        res[name] = {"log2_max_err_over_norms": math.log2((err / scale).max().item()),
                     "log2_max_err_over_sum_abs": math.log2((err / absdot).max().item())}
clips = int(os.environ.get("VQ_CLIPS", "256"))
L = 498 * 16
z = torch.randn(clips, L, D, device=dev)
cb = torch.randn(n, D, generator=g).to(dev)
for want in (False, True):
    for _ in range(2):
        lib.vq_quantize(z, cb, n, want_bdl=want, algo=lib.VQ_TENSOR)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.vq_quantize(z, cb, n, want_bdl=want, algo=lib.VQ_TENSOR)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rows = clips * L
    res[f"ms_zq{int(want)}"] = ms
    res[f"GBps_zq{int(want)}"] = rows * (4 * D + 8 + (4 * D if want else 0)) / ms / 1e6
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
lib.vq_quantize(z, cb, n, want_bdl=False, algo=lib.VQ_TENSOR, recheck_rows=cnt)
res["recheck_frac"] = cnt.item() / (clips * L)
res["window"] = os.environ.get("FDM_B200_VQ_WINDOW", "1")
print(json.dumps(res))
