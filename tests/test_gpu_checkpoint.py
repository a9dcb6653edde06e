"""Engine state through the TF Saver V2 wire format (boxsegliver_b200/checkpoint.py): save -> restore reproduces the
weights, moving statistics and optimizer slots bit for bit, a restored engine continues training exactly like the one
that was saved, and --load_weights / --weights_scope renaming follows /root/reference/core/models.py:151-185."""
import argparse

import numpy as np
import pytest

from boxsegliver_b200 import checkpoint as K
from boxsegliver_b200 import models, synthetic
from boxsegliver_b200.engine import EngineConfig, UNetEngine

pytestmark = pytest.mark.gpu


def _engine(ctx):
    return UNetEngine(ctx, EngineConfig(batch=2, height=64, width=64, weight_decay_rate=1e-5,
                                        loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4)))


def test_save_restore_continue_training(ctx, tmp_path):
    images, labels = synthetic.make_batch(2, 64, 64, 3, seed=1402)
    a = _engine(ctx)
    a.init_weights(seed=2)
    a.set_inputs(images, labels)
    for _ in range(2):
        a.train_step(1e-3)
    path = K.save_engine(a, tmp_path / "model.ckpt")
    assert path.endswith("model.ckpt-2") and K.latest_checkpoint(tmp_path) == path
    r = K.load_checkpoint(path)
    names = r.get_variable_to_shape_map()
    assert names["UNet/Encode1/Repeat/convolution2d_1/weights"] == [3, 3, 3, 64]
    assert "UNet/Encode1/Repeat/convolution2d_1/BatchNorm/moving_mean" in names
    assert "Optimizer/UNet/AdjustChannels/weights/Adam_1" in names and "Optimizer/beta2_power" in names
    assert int(r.get_tensor("global_step")) == 2 and K.find_root_scope(r) == "UNet"
    a.train_step(1e-3)
    loss_a, w_a = a.read_loss(), a.get_weights()
    a.close()

    b = _engine(ctx)
    b.init_weights(seed=77)                                   # different weights, overwritten by the restore
    assert K.restore_engine(b, tmp_path, with_slots=True) == 2   # directory -> CheckpointState -> latest prefix
    assert b.step_count == 2
    b.set_inputs(images, labels)
    b.train_step(1e-3)
    loss_b, w_b = b.read_loss(), b.get_weights()
    b.close()
    assert loss_a == loss_b
    for k in w_a:
        assert np.array_equal(w_a[k], w_b[k]), k


def test_load_weights_with_scope_renaming(ctx, tmp_path):
    a = _engine(ctx)
    w = a.init_weights(seed=4)
    K.save_checkpoint(tmp_path / "other" / "pretrained", {k.replace("UNet", "Backbone", 1): v for k, v in a.get_weights().items()})
    a.close()
    b = _engine(ctx)
    b.init_weights(seed=5)
    with pytest.raises(KeyError):                             # no Optimizer/* entry to infer the scope from
        K.restore_engine(b, tmp_path / "other" / "pretrained")
    K.restore_engine(b, tmp_path / "other" / "pretrained", weights_scope="Backbone")
    for k, v in b.get_weights().items():
        if k in w:
            assert np.array_equal(v, w[k]), k
    # through the host mirror of core/models.py:init_model
    class M:  # noqa: E301
        engine = b
    args = argparse.Namespace(load_weights=str(tmp_path / "other" / "pretrained"), weights_scope="Backbone",
                              model_dir=str(tmp_path / "run1"), load_weights_version="checkpoint")
    assert models.init_model(M, args) is None                 # file has no global_step
    assert models.init_model(M, argparse.Namespace(load_weights=None)) is None
    with pytest.raises(FileNotFoundError, match="doesn't exist"):
        K.restore_engine(b, tmp_path / "missing.ckpt")
    b.close()


def test_unet3d_save_restore_with_padded_slots(ctx, tmp_path):
    """UNet3D stores channels zero-padded (30 -> 64) and permuted ([encoder | up] concat halves): the optimizer slots
    go through the engine's own pack / unpack, are saved in TF variable shapes, and a restored engine continues
    training bit-identically."""
    from boxsegliver_b200.unet3d_engine import UNet3DConfig, UNet3DEngine
    cfg = UNet3DConfig(batch=1, depth=4, height=32, width=32, weight_decay_rate=3e-5)
    images, labels = synthetic.make_volume_batch(1, 4, 32, 32, seed=5)
    a = UNet3DEngine(ctx, cfg)
    a.init_weights(seed=1)
    a.set_inputs(images, labels)
    for _ in range(2):
        a.train_step(1e-3)
    path = K.save_engine(a, tmp_path / "model.ckpt")
    r = K.load_checkpoint(path)
    shapes = r.get_variable_to_shape_map()
    assert shapes["UNet3D/conv_e0/conv2/weights"] == [1, 3, 3, 30, 30]
    assert shapes["Optimizer/UNet3D/conv_e0/conv2/weights/Adam_1"] == [1, 3, 3, 30, 30]
    assert shapes["Optimizer/UNet3D/conv_d0/conv1/weights/Adam"] == [1, 3, 3, 60, 30]
    assert np.abs(r.get_tensor("Optimizer/UNet3D/conv_d0/conv1/weights/Adam")[:, :, :, 30:, :]).max() > 0  # up half kept
    a.train_step(1e-3)
    loss_a, w_a = a.read_loss(), a.get_weights()
    a.close()
    b = UNet3DEngine(ctx, cfg)
    b.init_weights(seed=9)
    assert K.restore_engine(b, tmp_path, with_slots=True) == 2 and b.step_count == 2
    b.set_inputs(images, labels)
    b.train_step(1e-3)
    loss_b, w_b = b.read_loss(), b.get_weights()
    b.close()
    assert loss_a == loss_b
    for k in w_a:
        assert np.array_equal(w_a[k], w_b[k]), k


def test_relative_model_dir_and_adamw_slot_names(ctx, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    eng = UNetEngine(ctx, EngineConfig(batch=2, height=32, width=32, optimizer="adamw", weight_decay_rate=1e-4))
    eng.init_weights(seed=2)
    images, labels = synthetic.make_batch(2, 32, 32, 3, seed=1)
    eng.set_inputs(images, labels)
    eng.train_step(1e-3)
    K.save_engine(eng, "out/model.ckpt")
    K.save_engine(eng, "out/model.ckpt", global_step=7)
    assert K.latest_checkpoint("out") is not None
    names = K.load_checkpoint(K.latest_checkpoint("out")).get_variable_to_shape_map()
    assert "Optimizer/UNet/ED-Bridge/ED-Bridge_1/weights/AdamW_1" in names
    assert K.restore_engine(eng, "out", with_slots=True) == 7
    eng.close()
