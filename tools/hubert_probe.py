#!/usr/bin/env python
"""One HuBERT-large encode of BATCH 10 s clips (configs[4] shape) — run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel launch list, or alone for the CUDA-event time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
import bench
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
BATCH = int(os.environ.get("BATCH", "64"))
SECONDS = float(os.environ.get("SECONDS", "10"))
fdm, ae, diff = bench.build_models(os.environ.get("PRESET", "vocaset"), dev, "bf16")
a = bench.synthetic_audio(BATCH, int(16000 * SECONDS), 0).to(dev)
REPS = int(os.environ.get("REPS", "1"))
if os.environ.get("AUDIO_PRECISION"):
    fdm.audio_precision = os.environ["AUDIO_PRECISION"]
for _ in range(2 if REPS > 1 else 1):
    fdm.encode_audio(a.clone())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(REPS):
    h = fdm.encode_audio(a.clone())
e1.record()
torch.cuda.synchronize()
print("encode ms", e0.elapsed_time(e1) / REPS, tuple(h.shape))
