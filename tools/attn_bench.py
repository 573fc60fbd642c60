#!/usr/bin/env python
"""Self-attention kernel timing at the benchmark shapes (CUDA events, 50 launches after warm-up): the FDM step's
(B = 128 sequences = 64 clips x 2 guidance passes, 8 heads x 128, T = 198, causal + periodic ALiBi), MEAD's, and the
EVQ-VAE decoder's unmasked attention, BIWI's 4 x 256 heads, the audio encoders' 16 / 12 x 64 heads and the 10 s shapes.
FDM_B200_ATTN_TC selects the kernel for head dim 128 / T <= 208 (2 = attention_tc2.cu, 1 = attention_tc.cu, 0 = neither),
FDM_B200_ATTN_TC3=0 sends every other shape to the mma.sync kernel instead of attention_tc3.cu."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "face-diffusion-model_b200")]
import torch  # noqa: E402
from fdm_b200 import lib  # noqa: E402

lib.require_device()
dev = torch.device("cuda:0")
out = {"FDM_B200_ATTN_TC": os.environ.get("FDM_B200_ATTN_TC", "2"), "FDM_B200_ATTN_TC3": os.environ.get("FDM_B200_ATTN_TC3", "1")}
for name, B, H, dh, T, causal in (("vocaset_step", 128, 8, 128, 198, True), ("mead_step_32x8s", 64, 4, 128, 199, True),
                                  ("mead_256", 512, 4, 128, 199, True), ("vq_decode_4s", 64, 8, 128, 198, False),
                                  ("biwi_step_64clips", 128, 4, 256, 149, True), ("biwi_step_128clips", 256, 4, 256, 149, True),
                                  ("hubert_4s", 64, 16, 64, 198, False), ("hubert_10s", 64, 16, 64, 498, False),
                                  ("wav2vec2_6s", 128, 12, 64, 299, False), ("vq_decode_10s", 64, 8, 128, 498, False),
                                  ("vocaset_step_10s", 32, 8, 128, 498, True)):
    d = H * dh
    qkv = torch.randn(B * T, 3 * d, device=dev).bfloat16()
    o = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
    slopes = torch.tensor([2.0 ** (-(2.0 ** -(math.log2(H) - 3)) * (i + 1)) for i in range(H)], device=dev) if causal else None
    run = lambda: lib.self_attention(qkv[:, 0:], qkv[:, d:], qkv[:, 2 * d:], o, B, T, T, H, dh, 1.0 / math.sqrt(dh), slopes=slopes, period=30)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    fl = 4.0 * B * H * T * T * dh
    out[name] = {"us": round(us, 2), "dense_TFLOPs": round(fl / us / 1e6, 1)}
print(json.dumps(out))
