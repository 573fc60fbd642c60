"""models.fdm_vocaset — drop-in FDM denoiser, VOCASET variant (reference models/fdm_vocaset.py:8-91).

d = 1024, 8 heads, HuBERT-large audio (unpaired 50 fps frames), latent regroup (B,16T,64) <-> (B,T,1024),
periodic positional encoding (period 30), 8 identities. Same constructor, forward signature and state_dict
keys as the reference; the arithmetic runs on libfdm_b200 (see fdm_b200/denoiser.py)."""
from fdm_b200.modules import FDMBase
from models.hubert import HubertModel

HUBERT_PATH = '/data/WX/hubert-large-ls960-ft'  # the reference's hard-coded checkpoint location


class FDM(FDMBase):
    preset_name = "vocaset"

    def __init__(self, feature_dim=512, n_head=8, num_layers=8, struct='Enc', audio_encoder_path=HUBERT_PATH):
        super().__init__()
        self.struct = struct
        self._build(feature_dim, n_head, num_layers, HubertModel.from_pretrained(audio_encoder_path))

    def forward(self, audio, t, vertice, id_one_hot):
        return self._forward(audio, t, vertice, id_one_hot)[0].clone()  # (not a view of the engine's reused output buffer)
