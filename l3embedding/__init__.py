"""Import shim: the reference's package name over the B200-native implementation.

The reference's callers import `l3embedding.*` (05_generate_embedding_samples.py:5 `from l3embedding.model import
load_embedding`; 03_train_embedding.py:4 `from l3embedding.train import *`; classifier/train.py:28 `from
l3embedding.train import LossHistory`; l3embedding/train.py:17-18).  With this directory on the path in place of the
reference's, those imports resolve to l3embedding_b200 unchanged.  No code lives here.
"""
