#!/usr/bin/env python
"""Where the time of the public VQAutoEncoder.quant() goes at the configs[4] size (1024 clips x 10 s = 8.16 M rows):
the whole call, the quantiser kernel with its output sets, the fused by-product pass (fdm_vq_stats)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
os.environ["FDM_B200_RANDOM_AUDIO_ENCODER"] = "1"
import torch
import bench
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
fdm, ae, diff = bench.build_models("vocaset", dev, "bf16")
z = torch.randn(int(os.environ.get("VQ_CLIPS", "1024")), 498 * 16, 64, device=dev)
cb = ae.quantize.embedding.weight.detach().float().contiguous()


def t(f, n=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("quant() API:", round(t(lambda: ae.quant(z)), 3), "ms")
idx, zq, zr = lib.vq_quantize(z, cb, 256, want_bdl=True, want_rows=True)
print("kernel idx + (B,D,L) + (B,L,D):", round(t(lambda: lib.vq_quantize(z, cb, 256, want_bdl=True, want_rows=True)), 3), "ms")
print("kernel idx + (B,D,L):", round(t(lambda: lib.vq_quantize(z, cb, 256, want_bdl=True)), 3), "ms")
print("fdm_vq_stats:", round(t(lambda: lib.vq_stats(z, cb, idx.view(-1), 256)), 3), "ms")
