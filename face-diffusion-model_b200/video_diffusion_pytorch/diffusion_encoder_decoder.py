"""video_diffusion_pytorch.diffusion_encoder_decoder — generic sampler (reference diffusion_encoder_decoder.py:
550-675): requires image_size / num_frames like the reference, p_sample_loop runs t = 499 .. 0."""
from fdm_b200.modules import GaussianDiffusionBase


class Unet3D:
    def __init__(self, *a, **k):
        raise NotImplementedError("Unet3D is dead code in the reference's sampling path")


class GaussianDiffusion(GaussianDiffusionBase):
    n_cond = 1
    default_range = (500, 0)

    def __init__(self, denoise_fn, *, image_size, num_frames, text_use_bert_cls=False, channels=3, timesteps=1000,
                 loss_type='l1', use_dynamic_thres=False, dynamic_thres_percentile=0.9):
        super().__init__()
        self.image_size, self.num_frames = image_size, num_frames
        self._build(denoise_fn, timesteps, loss_type, channels, text_use_bert_cls, use_dynamic_thres,
                    dynamic_thres_percentile)
