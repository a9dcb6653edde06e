"""Mirror of /root/reference/utils/distribution_utils.py for the one strategy the reference uses:
single-host synchronous data parallelism. One process per GPU (torchrun), NCCL all-reduce of the
flat gradient arena; the 'mirrored' graph replication of TF becomes identical engines per rank."""
from __future__ import annotations

import os


def per_device_batch_size(batch_size, num_gpus):
    """distribution_utils.py:107-134 -- same arithmetic, same error text."""
    if num_gpus <= 1:
        return batch_size
    remainder = batch_size % num_gpus
    if remainder:
        err = ('When running with multiple GPUs, batch size '
               'must be a multiple of the number of available GPUs. Found {} '
               'GPUs with a batch size of {}; try --batch_size={} instead.'
               ).format(num_gpus, batch_size, batch_size - remainder)
        raise ValueError(err)
    return int(batch_size / num_gpus)


class DataParallel:
    """What `get_distribution_strategy` returns here: rank / world and the side channel for the NCCL id."""

    def __init__(self, rank: int, world: int, local_rank: int):
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dist = None

    def init_control_plane(self, backend: str = "gloo"):
        if self.world > 1 and self.dist is None:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if not dist.is_initialized():
                dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world)
            self.dist = dist
        return self

    def broadcast_bytes(self, payload: bytes | None) -> bytes:
        if self.world == 1:
            return payload
        box = [payload]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def mean_scalar(self, v: float) -> float:
        """strategy.reduce(MEAN, per-replica loss) -- /root/reference/core/estimator.py:576-577."""
        if self.world == 1:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        self.dist.all_reduce(t)
        return float(t[0]) / self.world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()


def get_distribution_strategy(distribution_strategy="default", num_gpus=0, num_workers=1, all_reduce_alg=None):
    """distribution_utils.py:27-104. 'off'/'one_device' and num_gpus <= 1 -> None; 'mirrored'/'default' ->
    DataParallel over the torchrun environment; multi-worker and parameter-server are not offered
    (the reference raises / leaves them untested, :68-69, :100-101)."""
    if num_gpus < 0:
        raise ValueError("`num_gpus` can not be negative.")
    distribution_strategy = distribution_strategy.lower()
    if distribution_strategy == "off":
        if num_gpus > 1 or num_workers > 1:
            raise ValueError("When {} GPUs and  {} workers are specified, distribution_strategy flag cannot be set to "
                             "'off'.".format(num_gpus, num_workers))
        return None
    if distribution_strategy == "multi_worker_mirrored" or num_workers > 1:
        raise NotImplementedError
    if distribution_strategy == "one_device" or num_gpus <= 1:
        if distribution_strategy == "one_device" and num_gpus > 1:
            raise ValueError("When num_gpus is larger than 1, distribution_strategy flag cannot be set to 'one_device'.")
        return None
    if distribution_strategy in ("mirrored", "default"):
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world != num_gpus:
            raise ValueError(f"--num_gpus {num_gpus} needs one process per GPU: launch with "
                             f"`python -m torch.distributed.run --nproc-per-node {num_gpus} ...` (WORLD_SIZE={world})")
        return DataParallel(int(os.environ.get("RANK", "0")), world, int(os.environ.get("LOCAL_RANK", "0")))
    if distribution_strategy == "parameter_server":
        raise NotImplementedError("parameter_server is untested in the reference and not offered here")
    raise ValueError("Unrecognized Distribution Strategy: %r" % distribution_strategy)
