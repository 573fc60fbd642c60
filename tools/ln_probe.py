#!/usr/bin/env python
"""Times the denoiser step's two LayerNorm launches alone (plain norm3 / fused pair norm1 -> + cross + time -> norm2) at the
VOCASET (25344 x 1024) and MEAD (12736 x 512) step shapes: CUDA events over 50 calls, a 256 MB fill between repeats so the
inputs come from HBM like inside the step. FDM_B200_LN_HOT=0 selects the round-1 warp-per-two-rows kernel, 1 the new plain
kernel only, 2 (default) the new plain and pair kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rows, d in ((25344, 1024), (12736, 512)):
    x = torch.randn(rows, d, device=dev).bfloat16()
    r2 = torch.randn(rows // 2, d, device=dev).bfloat16()
    g1, b1, g2, b2 = (torch.randn(d, device=dev) for _ in range(4))
    vec = torch.randn(1000, d, device=dev)
    idx = torch.tensor([500], dtype=torch.int32, device=dev)
    out = torch.empty_like(x)
    res = torch.randn(rows, d, device=dev).bfloat16()
    for name, f, nbytes in (("plain", lambda: lib.layernorm(x, out, g1=g1, b1=b1), 2 * rows * d * 2),
                            ("pair", lambda: lib.layernorm(x, out, g1=g1, b1=b1, r2=r2, vec2=vec, vec_index_dev=idx, g2=g2, b2=b2),
                             2 * rows * d * 2 + rows // 2 * d * 2),
                            ("plain+res", lambda: lib.layernorm(x, out, r1=res, g1=g1, b1=b1), 3 * rows * d * 2),
                            ("pair+res", lambda: lib.layernorm(x, out, r1=res, g1=g1, b1=b1, r2=r2, vec2=vec, vec_index_dev=idx, g2=g2, b2=b2),
                             3 * rows * d * 2 + rows // 2 * d * 2)):
        for _ in range(3):
            f()
        ts = []
        for _ in range(20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            f()
        e1.record(); torch.cuda.synchronize()
        hot = e0.elapsed_time(e1) / 50 * 1e3
        print(f"rows={rows} d={d} {name}: cold {us:.1f} us = {nbytes / us / 1e3:.0f} GB/s; back-to-back (L2-warm) {hot:.1f} us = {nbytes / hot / 1e3:.0f} GB/s")
