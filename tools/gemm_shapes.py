#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMMs of one eager denoiser step of the bench workload (CUDA events around every
lib.gemm call, three repeats): which of QKV / out-proj / FFN1 / FFN2 is how far from the peak."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch
import bench
from fdm_b200 import lib
from utiles.classifierfree import ClassifierFreeSampleModel
lib.require_device()
dev = torch.device("cuda:0")
preset = os.environ.get("PRESET", "vocaset")
B = int(os.environ.get("CLIPS", "64")); seconds = float(os.environ.get("SECONDS", "4"))
fdm, ae, diff = bench.build_models(preset, dev, "bf16")
diff.denoise_fn = ClassifierFreeSampleModel(fdm, level=2.5)
P = fdm.preset
from fdm_b200.presets import conv_out_len
N = conv_out_len(int(16000 * seconds)); N -= N % 2
T = N // 2 if P.pair_audio else N
audio = bench.synthetic_audio(B, int(16000 * seconds), 0).to(dev)
ids = torch.eye(P.n_id)[[i % P.n_id for i in range(B)]].to(dev)
emo = torch.eye(7)[[i % 7 for i in range(B)]].to(dev) if P.emotion else None
eng = fdm.engine()
fdm.prepare(audio, T, ids, emo, guidance=diff.denoise_fn.guidance_cond)
xin = torch.randn(B * T, P.d, device=dev).to(eng.dtype)
t_dev = torch.tensor([500], dtype=torch.int32, device=dev)
rec = []
orig = lib.gemm
def wrapped(a, w, out, *args, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(a, w, out, *args, **kw); e1.record()
    M = kw.get("M") or a.shape[0]
    rec.append(((M, w.shape[0], w.shape[1], kw.get("residual") is not None, str(out.dtype)[6:]), e0, e1))
    return r
lib.gemm = wrapped
for _ in range(2):
    eng.denoise(xin, t_dev)
rec.clear()
for _ in range(5):
    eng.denoise(xin, t_dev)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for k, e0, e1 in rec:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
for (M, Nn, K, res, od), (n, ms) in agg.items():
    us = ms / n * 1e3
    print(f"M={M} N={Nn} K={K} residual={res} out={od}: {n // 5} per step, {us:.1f} us, {2.0 * M * Nn * K / us / 1e6:.0f} TFLOP/s")
