#!/usr/bin/env python
"""BASELINE.json configs[4]: "EVQ-VAE quantize+decode and HuBERT encode microbench: 1024 clips x 10 s".

Runs, on one B200, the three once-per-clip stages of the VOCASET preset at the 10 s shape (N = T = 498 frames,
7 968 latent rows per clip) for CLIPS clips in batches of BATCH:
  * HuBERT-large audio encoder (fdm_b200.audio, tcgen05 GEMMs incl. the implicit-GEMM conv stack),
  * EVQ-VAE quantise (fdm_vq_quantize, tensor-core filter + exact recheck) on random latents,
  * EVQ-VAE decode to 5023 x 3 vertices (fdm_b200.vqvae).
Prints one JSON line: CUDA-event time per stage, clips/s, and the executed TFLOP/s (HuBERT, decode) or algorithmic
GB/s (quantise) against the measured peaks. Random-init weights, synthetic audio (see bench.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from fdm_b200 import lib  # noqa: E402

lib.require_device()
dev = torch.device("cuda:0")
CLIPS = int(os.environ.get("CLIPS", "1024"))
BATCH = int(os.environ.get("BATCH", "64"))
SECONDS = 10.0
fdm, ae, diff = bench.build_models("vocaset", dev, "bf16")
P = fdm.preset
n_samples = int(16000 * SECONDS)
from fdm_b200.presets import conv_out_len  # noqa: E402
N = conv_out_len(n_samples)
N -= N % 2
T = N
pk = bench.peaks()


def timed(fn, n_batches):
    fn(0)  # warm-up (packs weights, builds tensor maps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_batches):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


nb = CLIPS // BATCH
audios = [bench.synthetic_audio(BATCH, n_samples, 0).to(dev) for _ in range(2)]
hidden_shape = []


def run_hubert(i):
    a = audios[i % 2].clone()  # fresh tensor identity: no cache hit
    h = fdm.encode_audio(a)
    hidden_shape[:] = list(h.shape)


ms_hubert = timed(run_hubert, nb)
# HuBERT-large FLOPs per clip at N frames (383.1 GFLOP at 10 s)
flops_hubert = 383.1e9 * CLIPS
g = torch.Generator(device="cpu").manual_seed(0)
lat = torch.randn(BATCH, T * P.fq, P.zdim, device=dev)
zq_keep = []


def run_quant(i):
    zq, _, _ = ae.quant(lat)
    zq_keep[:] = [zq]


ms_quant = timed(run_quant, nb)
rows = CLIPS * T * P.fq
bytes_quant = rows * (4 * P.zdim + 8 + 2 * 4 * P.zdim)  # z read, int64 index, z_q in both layouts (quant() API + decoder input)


def run_decode(i):
    ae.decode(zq_keep[0])


ms_decode = timed(run_decode, nb)
flops_decode = 71.6e9 * CLIPS  # SURVEY section 8(a) D1 at T = 498
print(json.dumps({
    "workload": f"configs[4]: {CLIPS} clips x 10 s, VOCASET preset, batches of {BATCH}", "frames_per_clip": T,
    "hubert_encode": {"ms": ms_hubert, "clips_per_s": CLIPS / ms_hubert * 1e3, "TFLOPs": flops_hubert / ms_hubert / 1e9,
                      "frac_of_sustained_bf16_peak": flops_hubert / ms_hubert / 1e9 / pk["tf_sustained"], "out_shape": hidden_shape},
    "vq_quantize": {"ms": ms_quant, "rows": rows, "algorithmic_GBps": bytes_quant / ms_quant / 1e6,
                    "frac_of_hbm_peak": bytes_quant / ms_quant / 1e6 / pk["hbm"],
                    "note": "quant() API: indices + z_q (B,D,L) + z_q rows + loss/perplexity by-products (torch reductions included)"},
    "vq_decode": {"ms": ms_decode, "clips_per_s": CLIPS / ms_decode * 1e3, "TFLOPs": flops_decode / ms_decode / 1e9,
                  "frac_of_sustained_bf16_peak": flops_decode / ms_decode / 1e9 / pk["tf_sustained"]},
    "peaks": pk}))
