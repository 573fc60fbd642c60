"""models.fdm_vqvae_mead — drop-in FDM denoiser, 3D-MEAD emotional variant (reference models/fdm_vqvae_mead.py:8-104).

d = 512, 4 heads, HuBERT-large audio with frame pairing (2048-d), latent regroup (B,8T,64) <-> (B,T,512),
sinusoidal positional encoding, 25 identities + 7 emotions."""
from fdm_b200.modules import FDMBase
from models.hubert import HubertModel

HUBERT_PATH = '/data/WX/hubert-large-ls960-ft'


class FDM(FDMBase):
    preset_name = "mead"

    def __init__(self, feature_dim=512, vertice_dim=70110, n_head=4, num_layers=8, struct='Enc',
                 audio_encoder_path=HUBERT_PATH):
        super().__init__()
        self.struct = struct
        self.vertice_dim = vertice_dim
        self._build(feature_dim, n_head, num_layers, HubertModel.from_pretrained(audio_encoder_path))

    def forward(self, audio, t, vertice, emotion_one_hot, id_one_hot, mask_cond=False, train=True):
        if mask_cond:  # force_mask semantics of mask_cond (reference :54-62): the null (all-zero) emotion condition
            emotion_one_hot = self.mask_cond(emotion_one_hot, force_mask=True)
        return self._forward(audio, t, vertice, id_one_hot, emotion_one_hot, None)[0].clone()  # (not a view of the engine's buffer)
