"""utiles.args — emotion EVQ-VAE hyper-parameters (reference utiles/args.py:4-20)."""
from models.utils.config import vq_vae_args  # noqa: F401  (same defaults in the reference)
