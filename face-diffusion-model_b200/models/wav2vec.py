"""models.wav2vec — drop-in for the reference's wav2vec 2.0 wrapper (reference models/wav2vec.py:69-143)."""
from transformers import Wav2Vec2Model as _HFWav2Vec2Model

from fdm_b200.modules import make_audio_encoder_class, wav2vec2_base_config

Wav2Vec2Model = make_audio_encoder_class(_HFWav2Vec2Model, wav2vec2_base_config)
Wav2Vec2Model.__name__ = Wav2Vec2Model.__qualname__ = "Wav2Vec2Model"
