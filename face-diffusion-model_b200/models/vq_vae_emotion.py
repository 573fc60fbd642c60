"""models.vq_vae_emotion — emotion-sliced EVQ-VAE, 3D-MEAD variant (reference models/vq_vae_emotion.py:8-352):
7 x 256 codes x 64; quant() takes the emotion one-hot and searches that emotion's 256-code slice."""
from fdm_b200.modules import VQAutoEncoderBase


class VQAutoEncoder(VQAutoEncoderBase):
    emotion_sliced = True

    def __init__(self, args):
        super().__init__()
        self._build(args)
