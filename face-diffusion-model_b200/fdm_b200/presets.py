"""Architecture presets of the three dataset variants of the reference (SURVEY.md §8 preset table).

The reference has no config object for these: the numbers are constructor literals and argparse defaults
(models/fdm_vocaset.py:9-51, models/fdm_vqvae_mead.py:9-52, models/fdm.py:10-52, models/utils/config.py:4-80,
utiles/args.py:4-20)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List


@dataclass(frozen=True)
class Preset:
    name: str
    d: int               # FDM feature_dim
    heads: int
    period: int          # periodic ALiBi / periodic PE period
    fq: int              # face_quan_num: latent tokens per frame
    zdim: int            # zquant_dim
    pair_audio: bool     # concatenate pairs of audio frames (50 fps -> 25 fps)
    audio_dim: int       # audio-encoder hidden size
    periodic_pe: bool
    latent_mish: bool    # latent_encoder = Linear+Mish (else Linear only)
    style_mish: bool     # style_embedd = Linear+Mish (else Linear only)
    emotion: bool
    n_id: int
    audio_encoder: str   # "hubert" | "wav2vec2"
    ddpm_range: tuple    # (hi, lo): p_sample_loop runs t = hi-1 .. lo in the reference file
    layers: int = 8

    @property
    def dh(self) -> int:
        return self.d // self.heads

    @property
    def audio_in(self) -> int:
        return self.audio_dim * (2 if self.pair_audio else 1)


PRESETS = {
    "vocaset": Preset("vocaset", 1024, 8, 30, 16, 64, False, 1024, True, True, False, False, 8, "hubert", (1000, 500)),
    "mead": Preset("mead", 512, 4, 30, 8, 64, True, 1024, False, True, False, True, 25, "hubert", (1000, 0)),
    "biwi": Preset("biwi", 1024, 4, 25, 8, 128, True, 768, False, False, True, False, 6, "wav2vec2", (1000, 500)),
}


def alibi_slopes(n_head: int) -> List[float]:
    """Head slopes of the FaceFormer-style temporal bias (closed form of get_slopes in
    models/fdm_vocaset.py:95-105): 2^(-8 i / n) for power-of-two n."""
    def pow2(n):
        start = 2.0 ** (-(2.0 ** -(math.log2(n) - 3)))
        return [start * start ** i for i in range(n)]
    if math.log2(n_head).is_integer():
        return pow2(n_head)
    c = 2 ** math.floor(math.log2(n_head))
    return pow2(c) + alibi_slopes(2 * c)[0::2][: n_head - c]


def conv_out_len(n: int, kernels=(10, 3, 3, 3, 3, 2, 2), strides=(5, 2, 2, 2, 2, 2, 2)) -> int:
    for k, s in zip(kernels, strides):
        n = (n - k) // s + 1
    return n
