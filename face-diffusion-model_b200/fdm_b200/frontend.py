"""Device-side pieces either side of the sampling path.

* `prepare_audio`: what the demos do on the host before `diffusion.sample` (demo/demo_3d_mead.py:85-97): the
  Wav2Vec2Processor zero-mean / unit-variance normalisation and the one second of trailing zeros, on the GPU.
* `vertex_metrics`: the vertex-error formulas of metric/metric.py:115-138 (LVE, FVE, EME, all-vertex error) on the GPU,
  so that evaluation does not round-trip (frames, V, 3) tensors through numpy.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import lib


def prepare_audio(speech: torch.Tensor, pad_seconds: float = 1.0, sample_rate: int = 16000) -> torch.Tensor:
    """speech (B, L) or (L,) raw 16 kHz samples on the GPU -> (B, L + pad) normalised per clip and zero-padded."""
    x = speech.detach().float()
    if x.dim() == 1:
        x = x[None]
    return lib.audio_normalize_pad(x.contiguous(), int(round(pad_seconds * sample_rate)))


def vertex_metrics(pred: torch.Tensor, gt: torch.Tensor, lip_idx: Optional[torch.Tensor] = None,
                   face_idx: Optional[torch.Tensor] = None, emotion_idx: Optional[torch.Tensor] = None) -> Dict[str, float]:
    """pred, gt (frames, V*3) fp32 on the GPU. Returns the metrics metric/metric.py prints, for the index sets given."""
    p, g = pred.detach().float().contiguous(), gt.detach().float().contiguous()
    out = {"all": float(lib.vertex_error(p, g, None, "max").mean())}
    for name, idx, mode in (("lve", lip_idx, "max"), ("fve", face_idx, "max"), ("eme", emotion_idx, "mean")):
        if idx is not None:
            out[name] = float(lib.vertex_error(p, g, idx.to(p.device, torch.int64).contiguous(), mode).mean())
    return out
