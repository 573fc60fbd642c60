"""models.fdm — drop-in FDM denoiser, BIWI variant (reference models/fdm.py:9-98, struct='Dec').

d = 1024, 4 heads (head dim 256), wav2vec2-base audio with frame pairing (1536-d), period 25, 6 identities.
The reference file omits the (B,8T,128) <-> (B,T,1024) latent regroup its own sample script relies on
(SURVEY.md §8(c) item 6); the regroup is applied here like in the other two variants. Only the working
struct='Dec' decoder variant is provided."""
from fdm_b200.modules import FDMBase
from models.wav2vec import Wav2Vec2Model

WAV2VEC_PATH = '/data/WX/wav2vec2-base-960h'


class FDM(FDMBase):
    preset_name = "biwi"

    def __init__(self, feature_dim=1024, vertice_dim=70110, n_head=4, num_layers=8, struct='Dec',
                 audio_encoder_path=WAV2VEC_PATH):
        super().__init__()
        if struct != 'Dec':
            raise NotImplementedError("struct='Enc' is a stale single-token variant in the reference; use struct='Dec'")
        self.struct = struct
        self.vertice_dim = vertice_dim
        self._build(feature_dim, n_head, num_layers, Wav2Vec2Model.from_pretrained(audio_encoder_path))

    def forward(self, audio, t, vertice, one_hot):
        return self._forward(audio, t, vertice, one_hot)[0].clone()  # (not a view of the engine's reused output buffer)
