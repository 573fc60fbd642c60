"""Shared test plumbing: build product modules / oracle inputs with the deterministic parity weights."""
import os
import warnings

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 7
N_SAMPLES = 8000


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def build_product(preset: str, tiny_audio: bool = True, device="cpu", codebook="reference"):
    """Product FDM + EVQ-VAE + diffusion (same constructors a reference user calls), parity weights loaded."""
    from oracle import reference_ops as R
    from oracle.weights import fill_state_dict
    import fdm_b200.modules as M
    warnings.simplefilter("ignore")
    kind = R.PRESETS[preset]["audio"]
    cfg_fn = (lambda: R.audio_encoder_config(kind, tiny_audio))
    # the audio checkpoints of the reference live at hard-coded absolute paths that do not exist here:
    # from_pretrained falls back to "config + random init", with the (tiny) config chosen by the test
    old_h, old_w = M.hubert_large_config, M.wav2vec2_base_config
    import models.hubert as H
    import models.wav2vec as W
    H.HubertModel.from_pretrained = classmethod(lambda cls, *a, **k: cls(cfg_fn()))
    W.Wav2Vec2Model.from_pretrained = classmethod(lambda cls, *a, **k: cls(cfg_fn()))
    if preset == "vocaset":
        from models.fdm_vocaset import FDM
        from models.vq_vae_vocaset import VQAutoEncoder
        from models.utils.config import vocaset_vq_vae_args as vargs
        from video_diffusion_pytorch.diffusion_BIWI_encoder_decoder import GaussianDiffusion
        fdm = FDM(feature_dim=1024)
    elif preset == "mead":
        from models.fdm_vqvae_mead import FDM
        from models.vq_vae_emotion import VQAutoEncoder
        from utiles.args import vq_vae_args as vargs
        from video_diffusion_pytorch.diffusion_mead_encoder_decoder import GaussianDiffusion
        fdm = FDM(feature_dim=512, vertice_dim=5023 * 3, struct="Dec")
    else:
        from models.fdm import FDM
        from models.vq_vae import VQAutoEncoder
        from models.utils.config import biwi_vq_vae_args as vargs
        from video_diffusion_pytorch.diffusion_BIWI_encoder_decoder import GaussianDiffusion
        fdm = FDM(feature_dim=1024, struct="Dec")
    ae = VQAutoEncoder(vargs())
    diff = GaussianDiffusion(fdm, timesteps=1000, loss_type="l2")
    diff.load_state_dict(fill_state_dict(diff.state_dict(), SEED))
    ae.load_state_dict(fill_state_dict(ae.state_dict(), SEED, codebook=codebook))
    return fdm.eval().to(device), ae.eval().to(device), diff.eval().to(device)


def oracle_inputs(preset: str, fdm, clip: int = 0):
    """(state_dict of the FDM on CPU, audio (L,), id one-hot, emotion one-hot | None)."""
    from oracle import reference_ops as R
    from oracle.weights import synthetic_audio
    P = R.PRESETS[preset]
    sd = {k: v.detach().cpu() for k, v in fdm.state_dict().items()}
    audio = synthetic_audio(clip, N_SAMPLES)
    idh = torch.eye(P["n_id"])[(1 + clip) % P["n_id"]][None]
    emo = torch.eye(7)[(4 + clip) % 7][None] if P["emotion"] else None
    return sd, audio, idh, emo


def hf_audio_model(preset: str, sd, tiny=True):
    """HF audio encoder (third-party arithmetic) holding the FDM's audio_encoder.* weights."""
    from oracle import reference_ops as R
    from transformers import HubertModel, Wav2Vec2Model
    kind = R.PRESETS[preset]["audio"]
    cls = HubertModel if kind == "hubert" else Wav2Vec2Model
    m = cls(R.audio_encoder_config(kind, tiny)).eval()
    m.load_state_dict({k[len("audio_encoder."):]: v for k, v in sd.items() if k.startswith("audio_encoder.")})
    return m


def build_vqvae(preset: str, device="cpu", codebook="normal"):
    """Product EVQ-VAE alone (same constructor a reference user calls), parity weights loaded."""
    from oracle.weights import fill_state_dict
    warnings.simplefilter("ignore")
    if preset == "vocaset":
        from models.vq_vae_vocaset import VQAutoEncoder
        from models.utils.config import vocaset_vq_vae_args as vargs
    elif preset == "mead":
        from models.vq_vae_emotion import VQAutoEncoder
        from utiles.args import vq_vae_args as vargs
    else:
        from models.vq_vae import VQAutoEncoder
        from models.utils.config import biwi_vq_vae_args as vargs
    ae = VQAutoEncoder(vargs())
    ae.load_state_dict(fill_state_dict(ae.state_dict(), SEED, codebook=codebook))
    return ae.eval().to(device)


def encoder_case(preset: str, ae=None):
    """(EVQ-VAE state_dict on CPU, motion (T, in_dim), emotion one-hot | None, reference encode() output (fq*T, D)) —
    the seeded input of oracle/gen_golden_encode.py and the output the real reference produced for it."""
    from oracle import reference_ops as R
    g = golden("encode")
    T = int(g["frames"])
    gen = torch.Generator(device="cpu").manual_seed(4321 + len(preset))
    x = torch.randn(T, R.PRESETS[preset]["in_dim"], generator=gen)
    emo = torch.eye(7)[4][None] if R.PRESETS[preset]["emotion"] else None
    ae = ae if ae is not None else build_vqvae(preset)
    sd = {k: v.detach().cpu() for k, v in ae.state_dict().items()}
    return sd, x, emo, torch.from_numpy(g[f"{preset}_h"])
