"""models.vq_vae_vocaset — EVQ-VAE, VOCASET variant (reference models/vq_vae_vocaset.py:9-258,
models/lib/quantizer.py:13-64): 256 codes x 64, 16 latent tokens per frame, no pre-embedding Linear,
output Linear with bias."""
from fdm_b200.modules import VQAutoEncoderBase


class VQAutoEncoder(VQAutoEncoderBase):
    pre_linear = False
    out_bias = True

    def __init__(self, args):
        super().__init__()
        self._build(args)
