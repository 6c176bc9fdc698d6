"""l3embedding/train.py of the reference (train, data_generator, single_epoch_data_generator, get_restart_info,
LossHistory, TimeHistory ...) -> l3embedding_b200.train; `from l3embedding.train import *` (03_train_embedding.py:4)
and `from l3embedding.train import LossHistory` (classifier/train.py:28) keep working."""
from l3embedding_b200.train import *          # noqa: F401,F403
from l3embedding_b200.train import (CSVLogger, LossHistory, ModelCheckpoint, TimeHistory, data_generator,  # noqa: F401
                                    get_restart_info, keras_tuples, single_epoch_data_generator, train)
