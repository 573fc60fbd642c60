#!/usr/bin/env python
"""BASELINE.json configs[4]: EVQ-VAE quantise microbench — 1024 clips x 10 s (VOCASET preset: 8 159 232 latent rows
x 64 dims against 256 codes). Reports kernel time, algorithmic HBM GB/s (4*D read + 8 B index + 4*D z_q per row)
and checks a slice of rows bit-exactly against the CPU oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch  # noqa: E402
from fdm_b200 import lib  # noqa: E402

lib.require_device()
dev = torch.device("cuda:0")
clips, T, fq, D, codes = 1024, 498, 16, 64, 256
L = T * fq
g = torch.Generator(device="cpu").manual_seed(0)
cb = torch.randn(codes, D, generator=g).to(dev)
z = torch.randn(clips, L, D, device=dev)
for _ in range(3):
    idx, zq, _ = lib.vq_quantize(z, cb, codes, want_bdl=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    idx, zq, _ = lib.vq_quantize(z, cb, codes, want_bdl=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
rows = clips * L
alg = rows * (4 * D + 8 + 4 * D)
from oracle import reference_ops as R  # checker
n_chk = 20000
oi, ozq, _ = R.vq_quantize(z[0, :n_chk].cpu(), cb.cpu())
ok = bool(torch.equal(oi, idx.view(clips, L)[0, :n_chk].cpu()))
peak = 6556.8
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
print(json.dumps({"workload": "VQ quantise microbench, 1024 clips x 10 s (8159232 rows x 64, 256 codes)", "ms": ms,
                  "rows_per_s": rows / (ms / 1e3), "algorithmic_GBps": alg / (ms / 1e3) / 1e9, "hbm_peak_GBps": peak,
                  "frac": alg / (ms / 1e3) / 1e9 / peak, "fp32_ffma_TFLOPs": 2.0 * rows * D * codes / (ms / 1e3) / 1e12,
                  "bit_exact_vs_oracle_first_rows": ok, "rows_checked": n_chk}))
