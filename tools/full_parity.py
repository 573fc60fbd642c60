#!/usr/bin/env python
"""BASELINE.json configs[0] as a parity run: VOCASET shapes, full-size FDM + EVQ-VAE (tiny audio encoder so the CPU
oracle stays fast), ONE clip, `--steps` DDPM steps (default the full 1000: t = 999..0) with host-generated noise
shared by both implementations. fp32 mode on the GPU vs the CPU oracle: latent max-abs, VQ index agreement, vertex
max-abs, lip-vertex error; then the bf16 mode against the same oracle run."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--preset", default="vocaset")
ap.add_argument("--samples", type=int, default=32000)
args = ap.parse_args()

import helpers  # noqa: E402
from helpers import build_product, hf_audio_model, oracle_inputs  # noqa: E402
from oracle import reference_ops as R  # noqa: E402
from oracle.metrics import lip_vertex_error  # noqa: E402
from oracle.weights import host_noise  # noqa: E402

helpers.N_SAMPLES = args.samples
dev = torch.device("cuda:0")
preset = args.preset
P = R.PRESETS[preset]
steps = list(range(999, 999 - args.steps, -1))
out = {"preset": preset, "ddpm_steps": args.steps, "audio_samples": args.samples}
ref = None
for precision in ("fp32", "bf16"):
    fdm, ae, diff = build_product(preset, device=dev, codebook="normal")
    fdm.set_precision(precision)
    ae.set_precision(precision)
    sd, audio, idh, emo = oracle_inputs(preset, fdm)
    a = audio[None].to(dev)
    hidden_gpu = fdm.encode_audio(a)
    T = hidden_gpu.shape[1] // (2 if P["pair"] else 1)
    shape = (1, T * P["fq"], P["zdim"])
    xT = host_noise(99, 0, 1000, shape).to(dev)
    diff.noise_source = lambda t: host_noise(99, 0, t, shape)
    conds = (emo.to(dev), idh.to(dev)) if P["emotion"] else (idh.to(dev),)
    t0 = time.time()
    lat = diff.p_sample_loop(shape, a, *conds, x_T=xT, steps=steps)
    zq, _, (_, _, idx) = ae.quant(lat, emo.to(dev)) if P["emotion"] else ae.quant(lat)
    verts = ae.decode(zq)
    torch.cuda.synchronize()
    gpu_s = time.time() - t0
    if ref is None:
        torch.set_num_threads(os.cpu_count())
        hidden = R.audio_encode(hf_audio_model(preset, sd), audio)
        tabs = R.diffusion_tables(1000)
        t0 = time.time()
        with torch.no_grad():
            rl = R.p_sample_loop(tabs, lambda z, t: R.fdm_forward(sd, preset, hidden, t, z, idh, emo), xT[0].cpu(),
                                 lambda t: host_noise(99, 0, t, shape)[0], steps=steps)
            emo_pos = int(emo.argmax()) if P["emotion"] else None
            ri, rzq, margin = R.vq_quantize(rl, ae.quantize.embedding.weight.detach().cpu(), emo_pos)
            rv = R.vq_decode({k: v.detach().cpu() for k, v in ae.state_dict().items()}, preset, rzq)
        ref = (rl, ri, rv, margin, time.time() - t0)
    rl, ri, rv, margin, cpu_s = ref
    lip = np.load(os.path.join(helpers.GOLDEN, "lip_vertices.npy"))
    same = (idx[:, 0].cpu() == ri)
    r = {"latent_max_abs": float((lat[0].cpu() - rl).abs().max()), "latent_rel": float((lat[0].cpu() - rl).norm() / rl.norm()),
         "vq_index_agreement": float(same.float().mean()), "vq_mismatch_max_margin": float(margin[~same].max()) if (~same).any() else 0.0,
         "vertex_max_abs": float((verts[0].cpu() - rv).abs().max()), "vertex_ref_max": float(rv.abs().max()),
         "gpu_seconds": gpu_s, "frames": T}
    if rv.shape[-1] == 15069:
        z = np.zeros_like(rv.numpy())
        lg, lr = lip_vertex_error(z, verts[0].cpu().numpy(), lip), lip_vertex_error(z, rv.numpy(), lip)
        r["lve_rel_diff"] = abs(lg - lr) / lr
    out[precision] = r
out["oracle_cpu_seconds"] = ref[4]
print(json.dumps(out))
