"""models.hubert — drop-in for the reference's HuBERT wrapper (reference models/hubert.py:72-146).

`HubertModel` keeps the HF parameter layout (checkpoints load unchanged) but its forward runs on the
libfdm_b200 audio-encoder engine (fdm_b200/audio.py), once per clip batch."""
from transformers import HubertModel as _HFHubertModel

from fdm_b200.modules import hubert_large_config, make_audio_encoder_class

HubertModel = make_audio_encoder_class(_HFHubertModel, hubert_large_config)
HubertModel.__name__ = HubertModel.__qualname__ = "HubertModel"
