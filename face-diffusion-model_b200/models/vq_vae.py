"""models.vq_vae — EVQ-VAE, BIWI / generic variant (reference models/vq_vae.py:8-347): 256 codes x 128,
8 latent tokens per frame, pre-embedding Linear, output Linear without bias."""
from fdm_b200.modules import VQAutoEncoderBase


class VQAutoEncoder(VQAutoEncoderBase):
    def __init__(self, args):
        super().__init__()
        self._build(args)
