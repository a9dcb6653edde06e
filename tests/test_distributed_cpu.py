"""The N > 1 path on CPU: two real processes over gloo drive the repo's data-parallel host logic
(boxsegliver_b200/distribution_utils.py, the mirror of /root/reference/utils/distribution_utils.py) with the
oracle standing in for the device, and must land on exactly the weights of the single-process emulation of
MirroredStrategy (oracle.unet_ref.mirrored_train_step: loss / R, SUM of gradients in rank order, identical Adam
update on every mirror, batch-norm moving statistics MEAN -- /root/reference/core/estimator.py:570-613)."""
import os
import socket

import numpy as np
import pytest

from boxsegliver_b200 import distribution_utils as DU
from boxsegliver_b200 import synthetic
from oracle import unet_ref as R

CFG = dict(height=16, width=16, channel=3, init_channels=8, num_down_samples=2, normalizer="batch_norm",
           weight_decay_rate=1e-5, loss_type="xentropy", loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4))
GLOBAL_BATCH, WORLD, LR = 4, 2, 1e-3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shards():
    per = DU.per_device_batch_size(GLOBAL_BATCH, WORLD)
    return [synthetic.make_batch(per, CFG["height"], CFG["width"], 3, seed=1357 + r) for r in range(WORLD)]


def _worker(rank, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dp = DU.get_distribution_strategy("mirrored", num_gpus=WORLD)
    assert isinstance(dp, DU.DataParallel) and dp.rank == rank and dp.world == WORLD
    dp.init_control_plane("gloo")
    # the 128-byte communicator id travels through this side channel on the GPU path
    token = dp.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    assert token == bytes(range(128))
    cfg = R.UNetCfg(**CFG)
    params = R.init_params(cfg, seed=5)   # mirrored variables: same seed on every rank
    slots = {}
    images, labels = _shards()[rank]
    losses = []
    for step in (1, 2):
        grads, moving, loss = R.replica_grads(params, images, labels, cfg, WORLD)
        names = sorted(grads)
        flat = torch.from_numpy(np.concatenate([grads[k].astype(np.float64).ravel() for k in names]))
        dist.all_reduce(flat)                                     # SUM, what bsl_allreduce_sum_f32 does on the device
        mnames = sorted(moving)
        mflat = torch.from_numpy(np.concatenate([np.asarray(moving[k], np.float64).ravel() for k in mnames]))
        dist.all_reduce(mflat)
        mflat /= WORLD                                            # bsl_scale_f32(1 / R): cross-replica MEAN
        gsum, off = {}, 0
        for k in names:
            n = grads[k].size
            gsum[k] = flat[off:off + n].numpy().reshape(grads[k].shape)
            off += n
        mmean, off = {}, 0
        for k in mnames:
            n = np.asarray(moving[k]).size
            mmean[k] = mflat[off:off + n].numpy().reshape(np.asarray(moving[k]).shape)
            off += n
        R.mirrored_apply(params, slots, step, gsum, mmean, cfg, LR)
        losses.append(dp.mean_scalar(loss) + R.regularization_loss(params, cfg) * 0)
    dp.barrier()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array(losses), **{k.replace("/", "|"): v for k, v in params.items()})
    dist.destroy_process_group()


def test_two_rank_gloo_matches_virtual_replicas(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    cfg = R.UNetCfg(**CFG)
    ref = R.init_params(cfg, seed=5)
    slots = {}
    ref_losses = []
    for step in (1, 2):
        reg = R.regularization_loss(ref, cfg)
        ref_losses.append(R.mirrored_train_step(ref, slots, step, _shards(), cfg, LR) - reg)
    r0 = dict(np.load(tmp_path / "rank0.npz"))
    r1 = dict(np.load(tmp_path / "rank1.npz"))
    assert np.allclose(r0["losses"], ref_losses, rtol=1e-12)
    for k, v in ref.items():
        a, b = r0[k.replace("/", "|")], r1[k.replace("/", "|")]
        assert np.array_equal(a, b), f"mirrors diverged on {k}"
        assert np.allclose(a, v, rtol=1e-6, atol=1e-9), k


def test_per_device_batch_size_and_strategy_errors(monkeypatch):
    assert DU.per_device_batch_size(8, 1) == 8 and DU.per_device_batch_size(256, 8) == 32
    with pytest.raises(ValueError, match="must be a multiple"):
        DU.per_device_batch_size(10, 4)
    assert DU.get_distribution_strategy("off", num_gpus=1) is None
    assert DU.get_distribution_strategy("mirrored", num_gpus=1) is None
    with pytest.raises(ValueError):
        DU.get_distribution_strategy("off", num_gpus=2)
    with pytest.raises(NotImplementedError):
        DU.get_distribution_strategy("multi_worker_mirrored", num_gpus=2)
    monkeypatch.setenv("WORLD_SIZE", "1")
    with pytest.raises(ValueError, match="one process per GPU"):
        DU.get_distribution_strategy("mirrored", num_gpus=2)
