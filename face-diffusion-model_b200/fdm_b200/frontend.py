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
