"""Data-parallel bootstrap replacing l3embedding/training_utils.py:21-170 (multi_gpu_model).

The reference builds one graph spanning N GPUs inside one process; here every GPU has its own process (torchrun)
holding a full replica.  Per step each rank takes the contiguous slice `get_slice` (training_utils.py:121-133) would
give its replica and runs forward/backward with its own BN batch statistics (reference semantics: no sync-BN).  The
gradient exchange itself lives in the CUDA library (csrc/dp.cu, l3_dp_*): NCCL all-reduces on a communication stream,
in buckets that leave while the backward pass is still running.  What is left here is the bootstrap -- rank, world
size and handing rank 0's NCCL unique id to the other ranks through the torch.distributed process group torchrun users
already have (gloo or nccl; any other 128-byte broadcast would do) -- and the sum of the validation scalars.
"""
from __future__ import annotations


def replica_slice(batch: int, rank: int, parts: int) -> slice:
    """training_utils.py:121-133: step = batch // parts; the last replica also takes the remainder."""
    step = batch // parts
    if rank == parts - 1:
        return slice(rank * step, batch)
    return slice(rank * step, (rank + 1) * step)


class SingleReplica:
    world_size, rank = 1, 0

    def slice(self, batch):
        return slice(0, batch)

    def attach(self, engine):
        return None

    def average_bn_state(self, engine):
        return None

    def sum_scalars(self, *xs):
        return xs


class LibraryReplicas:
    """One replica per torch.distributed rank; gradients are exchanged inside libl3b200 (attach -> l3_dp_init)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world_size = dist.get_world_size()
        self.rank = dist.get_rank()

    def slice(self, batch):
        return replica_slice(batch, self.rank, self.world_size)

    def attach(self, engine):
        """Joins `engine` to the job once (collective: every rank calls it for its engine at the same point)."""
        if engine.dp_world is None:
            box = [engine.dp_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(box, src=0)
            engine.dp_init(box[0], self.rank, self.world_size)

    def average_bn_state(self, engine):
        engine.dp_average_bn_state()

    def sum_scalars(self, *xs):
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(xs, dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return tuple(t.tolist())


def current(num_gpus: int):
    """Replica set for a model built with `num_gpus` (model.py:184-195)."""
    if num_gpus <= 1:
        return SingleReplica()
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("num_gpus=%d needs one process per GPU: launch with torchrun and call "
                           "torch.distributed.init_process_group first" % num_gpus)
    if dist.get_world_size() != num_gpus:
        raise RuntimeError("model built for %d GPUs but the process group has %d ranks" % (num_gpus, dist.get_world_size()))
    return LibraryReplicas()
