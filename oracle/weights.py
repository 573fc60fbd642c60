"""Deterministic, module-order-independent random weights for parity runs (TEST INFRASTRUCTURE).

Each tensor of a state_dict is drawn from a torch CPU generator seeded by a hash of (seed, key name), so the
imported reference modules, the oracle restatement and the CUDA product can be given identical weights without
shipping them: the golden fixtures only hold outputs. Scales keep activations O(1) through the stacks; the
FDM's zero-initialised latent_decoder (models/fdm_vocaset.py:49-51) is given N(0, 0.02) weights because a
zero output layer would make every parity check trivially true."""
import hashlib
import math
from typing import Dict

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return torch.Generator(device="cpu").manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)


def fill_state_dict(template: Dict[str, torch.Tensor], seed: int = 0, codebook: str = "reference") -> Dict[str, torch.Tensor]:
    """template: name -> tensor (only shapes/dtypes are used). Returns name -> new tensor.
    Buffers that are functions of the architecture (positional tables) are passed through unchanged."""
    out = {}
    for k, v in template.items():
        if not v.is_floating_point() or k.endswith(".pe") or "num_batches_tracked" in k or "." not in k:
            # integer buffers, positional tables and the top-level diffusion schedule buffers are not weights
            out[k] = v.clone()
            continue
        g = _gen(seed, k)
        shape = tuple(v.shape)
        if k.endswith("quantize.embedding.weight"):
            if codebook == "reference":  # uniform(+-1/n_e): models/lib/quantizer.py:33
                t = (torch.rand(shape, generator=g) * 2 - 1) / shape[0]
            else:
                t = torch.randn(shape, generator=g)
        elif k.endswith("latent_decoder.weight"):
            t = torch.randn(shape, generator=g) * 0.02
        elif k.endswith("parametrizations.weight.original0") or k.endswith("weight_g"):
            t = 1.0 + 0.1 * torch.rand(shape, generator=g)
        elif v.dim() >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) / math.sqrt(max(fan_in, 1))
        elif k.endswith("norm.weight") or "layer_norm.weight" in k or "norm1.weight" in k or "norm2.weight" in k or "norm3.weight" in k:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.1 * torch.randn(shape, generator=g)
        out[k] = t.to(v.dtype)
    return out


def synthetic_audio(clip: int, n_samples: int) -> torch.Tensor:
    """Zero-mean / unit-variance Gaussian audio, what Wav2Vec2Processor hands the model."""
    g = torch.Generator(device="cpu").manual_seed(1234 + clip)
    a = torch.randn(n_samples, generator=g)
    return (a - a.mean()) / a.std()


def host_noise(seed: int, clip: int, t: int, shape) -> torch.Tensor:
    """Host-generated sampler noise shared by both implementations in parity runs (t = 1000 draws x_T)."""
    g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + clip) * 1009 + t)
    return torch.randn(shape, generator=g)
