"""models.utils.config — EVQ-VAE hyper-parameter namespaces (reference models/utils/config.py:4-80).

The reference builds these with argparse and calls parse_args() on sys.argv (which breaks its own demo
scripts); here they are plain namespaces with the same attribute names and defaults, and
unknown command-line flags are left alone."""
from types import SimpleNamespace

_COMMON = dict(vqvae_pretrained_path='/data/WX/video-diffusion-pytorch/checkpoints/vqvae/vqvae_100.pt',
               hidden_size=1024, neg=0.2, quant_factor=0, INaffine=False, num_hidden_layers=6,
               num_attention_heads=8, intermediate_size=1536)


def _ns(**kw):
    return SimpleNamespace(**{**_COMMON, **kw})


def vq_vae_args():
    """emotion-conditioned EVQ-VAE (3D MEAD): 7 x 256 codes."""
    return _ns(n_embed=256 * 7, zquant_dim=64, in_dim=5023 * 3, face_quan_num=8)


def origin_vq_vae_args():
    return _ns(n_embed=256, zquant_dim=64, in_dim=5023 * 3, face_quan_num=8)


def biwi_vq_vae_args():
    return _ns(n_embed=256, zquant_dim=128, in_dim=70110, face_quan_num=8)


def vocaset_vq_vae_args():
    return _ns(n_embed=256, zquant_dim=64, in_dim=15069, face_quan_num=16)
