"""l3embedding/training_utils.py of the reference (multi_gpu_model :21-170) -> l3embedding_b200.model.multi_gpu_model:
the replicas are one process per GPU, the gradient exchange lives in libl3b200 (l3_dp_*)."""
from l3embedding_b200.model import multi_gpu_model  # noqa: F401
