#!/usr/bin/env python
"""One EVQ-VAE quantise + decode of BATCH clips (configs[4] shape: 10 s, T = 498) - run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list, or alone for the CUDA-event time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
import bench
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
os.environ["FDM_B200_RANDOM_AUDIO_ENCODER"] = "1"
BATCH, T = int(os.environ.get("BATCH", "64")), int(os.environ.get("FRAMES", "498"))
preset = os.environ.get("PRESET", "vocaset")
fdm, ae, diff = bench.build_models(preset, dev, "bf16")
P = fdm.preset
z = torch.randn(BATCH, T * P.fq, P.zdim, device=dev)
emo = torch.eye(7, device=dev)[[i % 7 for i in range(BATCH)]] if P.emotion else None
def run():
    zq, _, _ = ae.quant(z, emo) if P.emotion else ae.quant(z)
    return ae.decode(zq)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); v = run(); e1.record(); torch.cuda.synchronize()
print("quant + decode ms", e0.elapsed_time(e1), tuple(v.shape))
