#!/usr/bin/env python
"""Which hot-loop kernels are power-capped? Loops each kernel alone for ~2 s while sampling nvidia-smi."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
M, d, H, T, S = 25344, 1024, 8, 198, 128
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(M, d, generator=g).to(dev).bfloat16()
wq = (torch.randn(3 * d, d, generator=g) / 32).to(dev).bfloat16()
w2 = (torch.randn(d, 2 * d, generator=g) / 45).to(dev).bfloat16()
h2 = torch.randn(M, 2 * d, generator=g).to(dev).bfloat16()
bq = torch.zeros(3 * d, device=dev); b1 = torch.zeros(d, device=dev)
qkv = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
out = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
gam = torch.ones(d, device=dev); bet = torch.zeros(d, device=dev)
ttab = torch.randn(1000, d, device=dev); tidx = torch.tensor([5], dtype=torch.int32, device=dev)
slopes = torch.tensor([2.0 ** -(i + 1) for i in range(H)], device=dev)
kern = {
    "gemm_qkv(25344x3072x1024)": (lambda: lib.gemm(x, wq, qkv, bias=bq), 2.0 * M * 3 * d * d),
    "gemm_ffn2(25344x1024x2048,+res)": (lambda: lib.gemm(h2, w2, out, bias=b1, residual=x), 2.0 * M * d * 2 * d),
    "attention(128 seq x 8 heads x 198)": (lambda: lib.self_attention(qkv[:, 0:], qkv[:, d:], qkv[:, 2 * d:], out, S, T, T, H, 128, 0.0884, slopes=slopes, period=30), 4.0 * S * H * T * T * 128),
    "layernorm(25344x1024)": (lambda: lib.layernorm(out, out, g1=gam, b1=bet), 0.0),
    "layernorm_double(25344x1024,+cross,+time)": (lambda: lib.layernorm(x, out, g1=gam, b1=bet, r2=x[:M // 2], vec2=ttab, vec_index_dev=tidx, g2=gam, b2=bet), 0.0),
}
Q = "clocks.sm,power.draw"
res = {}
for name, (fn, flops) in kern.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=f)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 2.5:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    p.terminate(); p.wait()
    f.flush(); f.seek(0)
    rows = [l.split(",") for l in f.read().splitlines() if "," in l]
    rows = rows[len(rows) // 3:]
    clk = sorted(float(r[0]) for r in rows); pw = sorted(float(r[1]) for r in rows)
    us = e0.elapsed_time(e1) * 1e3 / n
    res[name] = {"us": us, "TFLOPs": flops / us / 1e6 if flops else None, "sm_mhz_median": clk[len(clk) // 2], "power_w_median": pw[len(pw) // 2]}
print(json.dumps(res, indent=1))
