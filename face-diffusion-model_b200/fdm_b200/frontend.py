"""Device-side pieces either side of the sampling path.

* `prepare_audio` / `resample`: what the demos do on the host before `diffusion.sample` (demo/demo_3d_mead.py:83-97):
  resampling to 16 kHz, the Wav2Vec2Processor zero-mean / unit-variance normalisation and the one second of trailing
  zeros, on the GPU.
* `vertex_metrics`: the vertex-error formulas of metric/metric.py:115-138 (LVE, FVE, EME, all-vertex error) on the GPU,
  so that evaluation does not round-trip (frames, V, 3) tensors through numpy.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import lib


def resample_filter(orig_sr: int, target_sr: int, beta: float = 5.0):
    """Low-pass design of scipy.signal.resample_poly (Kaiser-windowed sinc, half length 10 * max(up, down), cutoff at the
    lower Nyquist rate), evaluated in float64 on the host. Returns (taps incl. the front padding, up, down, pre) with
    `pre` the number of leading output samples upfirdn produces before the first wanted one."""
    import math
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    max_rate = max(up, down)
    half_len = 10 * max_rate
    n = 2 * half_len + 1
    fc = 1.0 / max_rate
    m = torch.arange(n, dtype=torch.float64) - half_len
    h = fc * torch.sinc(fc * m) * torch.kaiser_window(n, periodic=False, beta=beta, dtype=torch.float64)
    h = h / h.sum() * up  # firwin(scale=True): unit gain at DC, then the zero-stuffing gain
    n_pre_pad = down - half_len % down
    taps = torch.cat([torch.zeros(n_pre_pad, dtype=torch.float64), h])
    pre = (half_len + n_pre_pad) // down
    return taps, up, down, pre


def resample(speech: torch.Tensor, orig_sr: int, target_sr: int = 16000) -> torch.Tensor:
    """speech (B, L) or (L,) on the GPU at `orig_sr` -> (B, ceil(L * target / orig)) at `target_sr`: the demos'
    librosa.load(path, sr=16000) step (demo/demo_3d_mead.py:83) as a polyphase FIR on the device. The arithmetic is
    scipy.signal.resample_poly's (pinned in tests against scipy); librosa's default soxr_hq filter is a different low-pass,
    so the samples agree with the reference's only to the resamplers' stop-band level."""
    x = speech.detach().float()
    if x.dim() == 1:
        x = x[None]
    if int(orig_sr) == int(target_sr):
        return x.contiguous()
    taps, up, down, pre = resample_filter(orig_sr, target_sr)
    n_out = -(-x.shape[1] * up // down)
    return lib.resample_poly(x.contiguous(), taps.to(x.device, torch.float32).contiguous(), up, down, pre, n_out)


def prepare_audio(speech: torch.Tensor, pad_seconds: float = 1.0, sample_rate: int = 16000, target_rate: int = 16000) -> torch.Tensor:
    """speech (B, L) or (L,) raw samples on the GPU at `sample_rate` -> resampled to 16 kHz if needed, normalised per clip
    (Wav2Vec2Processor) and zero-padded by `pad_seconds`: everything demo/demo_3d_mead.py:83-97 does on the host."""
    x = speech.detach().float()
    if x.dim() == 1:
        x = x[None]
    if int(sample_rate) != int(target_rate):
        x = resample(x, sample_rate, target_rate)
    return lib.audio_normalize_pad(x.contiguous(), int(round(pad_seconds * target_rate)))


def vertex_metrics(pred: torch.Tensor, gt: torch.Tensor, lip_idx: Optional[torch.Tensor] = None,
                   face_idx: Optional[torch.Tensor] = None, emotion_idx: Optional[torch.Tensor] = None) -> Dict[str, float]:
    """pred, gt (frames, V*3) fp32 on the GPU. Returns the metrics metric/metric.py prints, for the index sets given."""
    p, g = pred.detach().float().contiguous(), gt.detach().float().contiguous()
    out = {"all": float(lib.vertex_error(p, g, None, "max").mean())}
    for name, idx, mode in (("lve", lip_idx, "max"), ("fve", face_idx, "max"), ("eme", emotion_idx, "mean")):
        if idx is not None:
            out[name] = float(lib.vertex_error(p, g, idx.to(p.device, torch.int64).contiguous(), mode).mean())
    return out


class VertexWriter:
    """`.npy` writer of the sample scripts (np.save of a (1, T, V*3) fp32 array per clip,
    samples/sample_diffusion_mead.py:86), off the critical path: the vertices are copied device -> pinned host on a
    side stream (overlapping the all-gather / the next job) and written by a worker thread."""

    def __init__(self):
        import queue
        import threading
        self._q = queue.Queue()
        self._stream = None
        self._err = None
        self._t = threading.Thread(target=self._work, daemon=True)
        self._t.start()

    def _work(self):
        import numpy as np
        while True:
            item = self._q.get()
            if item is None:
                return
            host, ev, paths = item
            try:
                ev.synchronize()
                arr = host.numpy()
                for i, path in enumerate(paths):
                    np.save(path, arr[i:i + 1])
            except Exception as e:  # surfaced by close()
                self._err = e

    def submit(self, verts: torch.Tensor, paths) -> None:
        """verts (B, T, V*3) fp32 on the GPU; paths: B file names. Returns immediately."""
        assert verts.is_cuda and verts.dim() == 3 and len(paths) == verts.shape[0]
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=verts.device)
        host = torch.empty(verts.shape, dtype=torch.float32, pin_memory=True)
        self._stream.wait_stream(torch.cuda.current_stream(verts.device))
        with torch.cuda.stream(self._stream):
            host.copy_(verts.detach().float(), non_blocking=True)
            verts.record_stream(self._stream)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._q.put((host, ev, list(paths)))

    def close(self) -> None:
        self._q.put(None)
        self._t.join()
        if self._err is not None:
            raise self._err
