"""l3embedding/vision_model.py of the reference -> l3embedding_b200.model (tower builders :7,102,221 and
construct_cnn_l3_orig_vision_embedding_model :198)."""
from l3embedding_b200.model import (construct_cnn_L3_orig_inputbn_vision_model, construct_cnn_L3_orig_vision_model,  # noqa: F401
                                    construct_cnn_l3_orig_vision_embedding_model, construct_tiny_L3_vision_model)

__all__ = ["construct_cnn_L3_orig_vision_model", "construct_cnn_L3_orig_inputbn_vision_model",
           "construct_cnn_l3_orig_vision_embedding_model", "construct_tiny_L3_vision_model"]
