"""l3embedding/model.py of the reference (MODELS, construct_cnn_L3_*, load_model, load_embedding, convert_num_gpus,
L3_merge_audio_vision_models, gpu_wrapper) -> l3embedding_b200.model.  Like the reference module (model.py:2-4) it also
re-exports the tower builders and multi_gpu_model."""
from l3embedding_b200.model import *          # noqa: F401,F403
from l3embedding_b200.model import (MODELS, Adam, L3_merge_audio_vision_models, convert_num_gpus, gpu_wrapper,  # noqa: F401
                                    load_embedding, load_model, multi_gpu_model)
from .audio_model import *                    # noqa: F401,F403
from .vision_model import *                   # noqa: F401,F403
