#!/usr/bin/env python
"""Quick timing of the tensor-core VQ kernel at the configs[4] size (1024 clips x 10 s = 8.16 M rows): indices only,
+ z_q (B, D, L), + row layout; env knobs (FDM_B200_VQ_L2_AHEAD, FDM_B200_VQ_WINDOW) are read by the library."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
from fdm_b200 import lib
lib.require_device()
dev = torch.device("cuda:0")
clips, L, D, codes = int(os.environ.get("VQ_CLIPS", "1024")), 498 * 16, 64, 256
cb = torch.randn(codes, D, device=dev)
z = torch.randn(clips, L, D, device=dev)
rows = clips * L
out = {k: os.environ.get(k) for k in ("FDM_B200_VQ_L2_AHEAD", "FDM_B200_VQ_WINDOW")}
for name, kw, per_row in (("idx", dict(want_bdl=False), 4 * D + 8), ("idx+zq", dict(want_bdl=True), 8 * D + 8)):
    for _ in range(3):
        lib.vq_quantize(z, cb, codes, algo=lib.VQ_TENSOR, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.vq_quantize(z, cb, codes, algo=lib.VQ_TENSOR, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[name] = {"ms": round(ms, 4), "GBps": round(rows * per_row / ms / 1e6, 1)}
print(json.dumps(out))
