#!/usr/bin/env python
"""The EVQ-VAE decoder's last GEMM, vertice_map_reverse (1024 -> 15069 fp32, models/vq_vae_vocaset.py:257): N * 4 bytes is
not a multiple of 16, so C cannot go through the TMA store engine. Times it against the same GEMM with N padded to 15072
(TMA path) and with a bf16 output, 64 clips x 10 s (M = 31872)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
M, K = int(os.environ.get("M", "31872")), 1024
a = torch.randn(M, K, device=dev).bfloat16()
for N, dt in ((15069, torch.float32), (15072, torch.float32), (15072, torch.bfloat16), (70110, torch.float32), (70112, torch.float32)):
    if N > 20000 and M > 20000:
        a2 = a[:9536]  # BIWI: 64 clips x 6 s
    else:
        a2 = a
    w = (torch.randn(N, K, device=dev) / 32).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(a2.shape[0], N, device=dev, dtype=dt)
    f = lambda: lib.gemm(a2, w, out, bias=bias)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    Mm = a2.shape[0]
    print(f"M={Mm} N={N} K={K} out={str(dt)[6:]}: {us:.1f} us, {2.0 * Mm * N * K / us / 1e6:.0f} TFLOP/s, C write {Mm * N * out.element_size() / us / 1e3:.0f} GB/s")
