"""l3embedding/audio_model.py of the reference -> l3embedding_b200.model (tower builders :8,118,225,335,490 and
convert_audio_model_to_embedding :445)."""
from l3embedding_b200.model import (construct_cnn_L3_kapredbinputbn_audio_model, construct_cnn_L3_melspec1_audio_model,  # noqa: F401
                                    construct_cnn_L3_melspec2_audio_model, construct_cnn_L3_orig_audio_model,
                                    construct_tiny_L3_audio_model, convert_audio_model_to_embedding)

__all__ = ["construct_cnn_L3_orig_audio_model", "construct_cnn_L3_kapredbinputbn_audio_model",
           "construct_cnn_L3_melspec1_audio_model", "construct_cnn_L3_melspec2_audio_model",
           "convert_audio_model_to_embedding", "construct_tiny_L3_audio_model"]
