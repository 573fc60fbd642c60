"""video_diffusion_pytorch.diffusion_BIWI_encoder_decoder — sampler used by the VOCASET / BIWI scripts
(reference diffusion_BIWI_encoder_decoder.py:532-710): conditioning = (id_one_hot,), p_sample_loop runs
t = 999 .. 500 as in the reference file (pass step_range=(1000, 0) for the full chain), plus ddim_sample."""
from fdm_b200.modules import GaussianDiffusionBase


class Unet3D:
    def __init__(self, *a, **k):
        raise NotImplementedError("Unet3D is dead code in the reference's sampling path")


class GaussianDiffusion(GaussianDiffusionBase):
    n_cond = 1
    default_range = (1000, 500)

    def __init__(self, denoise_fn, *, text_use_bert_cls=False, channels=3, timesteps=1000, loss_type='l1',
                 use_dynamic_thres=False, dynamic_thres_percentile=0.9):
        super().__init__()
        self._build(denoise_fn, timesteps, loss_type, channels, text_use_bert_cls, use_dynamic_thres,
                    dynamic_thres_percentile)
