#!/usr/bin/env python
"""Times fdm_gemm_bf16 variants (bias / activation / residual / row count) alone, CUDA events over 20 calls."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
N, K = int(os.environ.get("N", "1024")), int(os.environ.get("K", "1024"))
for M in (12672, 25344):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for act, use_res in itertools.product((lib.ACT_NONE, lib.ACT_MISH, lib.ACT_RELU), (False, True)):
        f = lambda: lib.gemm(a, w, out, bias=bias, act=act, residual=res if use_res else None)
        for _ in range(3):
            f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"M={M} N={N} K={K} act={act} residual={use_res}: {us:.1f} us, {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
