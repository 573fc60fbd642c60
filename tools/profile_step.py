#!/usr/bin/env python
"""Run a few eager denoising steps of the bench workload (VOCASET, 64 clips x 4 s, guidance) with synthetic
audio features, for ncu captures of the hot-loop kernels:

  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --steps 2
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 20 -c 3 -o gpurun_out/prof_gemm \
      python tools/profile_step.py --steps 1
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--frames", type=int, default=198)
ap.add_argument("--preset", default="vocaset")
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()

import bench  # noqa: E402
from fdm_b200 import lib  # noqa: E402
from utiles.classifierfree import ClassifierFreeSampleModel  # noqa: E402

dev = torch.device("cuda:0")
lib.require_device()
import fdm_b200.modules as M
import models.hubert as H
from oracle import reference_ops as R  # tiny audio encoder config only (no compute): keeps start-up short
H.HubertModel.from_pretrained = classmethod(lambda cls, *a, **k: cls(R.audio_encoder_config("hubert", True)))
fdm, ae, diff = bench.build_models(args.preset, dev, args.precision)
P = fdm.preset
B, T = args.clips, args.frames
audio = torch.zeros(B, 16, device=dev)
N = T * (2 if P.pair_audio else 1)
fdm.set_audio_features(audio, torch.randn(B, N, P.audio_dim, device=dev))
ids = torch.eye(P.n_id, device=dev)[[i % P.n_id for i in range(B)]]
emo = torch.eye(7, device=dev)[[i % 7 for i in range(B)]] if P.emotion else None
eng = fdm.prepare(audio, T, ids, emo, guidance="emotion" if P.emotion else "id")
x = torch.randn(B, T * P.d, device=dev)
xin = x.view(B * T, P.d).to(eng.dtype)
xbf = torch.empty(B * T, P.d, device=dev, dtype=torch.bfloat16)
t_dev = torch.tensor([500], dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for _ in range(args.steps):
    x0 = eng.denoise(xin, t_dev)
    lib.ddpm_step(x0[0], x, x, diff.posterior_mean_coef1, diff.posterior_mean_coef2, diff._sigma_table(), x0_uncond=x0[1],
                  guidance=2.5, noise=None, out_bf16=xbf, t_dev=t_dev, seed=1, clip_index0=0)
torch.cuda.synchronize()
print("done", lib.launch_count)
