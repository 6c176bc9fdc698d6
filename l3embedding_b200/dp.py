"""Data-parallel glue replacing l3embedding/training_utils.py:21-170 (multi_gpu_model).

The reference builds one graph spanning N GPUs inside one process; here every GPU has its own process (torchrun)
holding a full replica.  Per step each rank takes the contiguous slice `get_slice` (training_utils.py:121-133)
would give its replica, runs forward/backward with its own BN batch statistics (reference semantics: no sync-BN),
and the flat gradient arena is summed over ranks with ONE all-reduce (NCCL over NVLink on GPUs; gloo in the CPU
tests).  Gradients are already scaled by 1/global_batch on the device, so the sum is the global-batch gradient.
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

    def allreduce_grads(self, engine):
        return None

    def average_bn_state(self, engine):
        return None

    def sum_scalars(self, *xs):
        return xs


class TorchDistReplicas:
    """One replica per torch.distributed rank."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world_size = dist.get_world_size()
        self.rank = dist.get_rank()

    def slice(self, batch):
        return replica_slice(batch, self.rank, self.world_size)

    def allreduce_tensor(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def attach(self, engine):
        return None

    def allreduce_grads(self, engine):
        # one flat fp32 arena (38 MB for cnn_L3_melspec2): a single NCCL ring/NVLS all-reduce
        return self.allreduce_tensor(engine.grads)

    def average_bn_state(self, engine):
        self.allreduce_tensor(engine.bn_state)
        engine.bn_state.div_(self.world_size)

    def sum_scalars(self, *xs):
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(xs, dtype=torch.float64, device=dev)
        self.allreduce_tensor(t)
        return tuple(t.tolist())


def current(num_gpus: int):
    """Replica set for a model built with `num_gpus` (model.py:184-195)."""
    if num_gpus <= 1:
        return SingleReplica()
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("num_gpus=%d needs one process per GPU: launch with torchrun and call "
                           "torch.distributed.init_process_group('nccl') first" % num_gpus)
    if dist.get_world_size() != num_gpus:
        raise RuntimeError("model built for %d GPUs but the process group has %d ranks" % (num_gpus, dist.get_world_size()))
    return TorchDistReplicas()
