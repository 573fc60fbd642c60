"""CPU: the drop-in boundary — state_dict layout, schedule bits, C-ABI symbols, loud failure without a GPU."""
import json
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_meta(preset):
    with torch.device("meta"):
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
        return GaussianDiffusion(fdm, timesteps=1000, loss_type="l2"), VQAutoEncoder(vargs())


@pytest.mark.parametrize("preset", ["vocaset", "mead", "biwi"])
def test_state_dict_layout_matches_reference(preset):
    """Reference checkpoints must load unchanged: identical keys and shapes (golden: the real reference modules)."""
    import warnings
    warnings.simplefilter("ignore")
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        layout = json.load(f)[preset]
    diff, ae = _build_meta(preset)
    for mod, want in ((diff, layout["diffusion"]), (ae, layout["vqvae"])):
        got = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert set(got) == set(want), (sorted(set(want) - set(got))[:5], sorted(set(got) - set(want))[:5])
        for k in want:
            assert got[k] == want[k], k


def test_product_schedule_bit_exact():
    from fdm_b200.modules import cosine_tables
    g = golden("schedule")
    for name, val in cosine_tables(1000):
        assert np.array_equal(val.to(torch.float32).numpy(), g[name]), name


def test_c_abi_exports_every_declared_symbol():
    from fdm_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "fdm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fdm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/fdm_b200.h but not exported"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    assert handle.fdm_abi_version() == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from fdm_b200 import lib
    with pytest.raises(lib.FdmError):
        lib.require_device()
    with pytest.raises((lib.FdmError, AssertionError)):
        lib.gemm(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, 4))
