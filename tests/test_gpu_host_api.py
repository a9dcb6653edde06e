"""The drop-in call sequence of the reference (core/models.py:224-281 model_fn -> core/solver.py:204-243 Solver ->
core/estimator.py:756-757 `sess.run([train_op, loss])`) driven through the host mirror, for every model class of
MODEL_ZOO that a shipped script trains. The step it runs must be the step the engine tests gate: same loss as the
engine driven directly with the same seed, loss decreasing over a few steps, predictions / metrics populated."""
import argparse

import numpy as np
import pytest

from boxsegliver_b200 import models, solver, synthetic
from boxsegliver_b200.networks.base import ModeKeys

pytestmark = pytest.mark.gpu


def _args(model, **kw):
    p = argparse.ArgumentParser()
    models.add_arguments(p)
    solver.add_arguments(p)
    a = p.parse_args(["--model", model, "--classes", "Liver", "Tumor", "--batch_size", "2",
                      "--normalizer", kw.pop("normalizer", "batch_norm"), "--learning_rate", "1e-3"])
    base = dict(im_height=64, im_width=64, im_channel=3, num_gpus=1, seed=3, weight_decay_rate=1e-5, bias_decay=False,
                loss_type="xentropy", loss_weight_type="numerical", loss_numeric_w=[0.2, 0.4, 4.4],
                loss_proportion_decay=1000.0, metrics_train=["Dice"], use_spatial=False, use_context=False)
    base.update(kw)
    for k, v in base.items():
        setattr(a, k, v)
    return a


def _drive(ctx, model_name, feed_kw, **kw):
    args = _args(model_name, **kw)
    params = models.get_model_params(args, build_metrics=True)
    params.update(args=args, solver=solver.Solver, ctx=ctx, world=1)
    images, labels = synthetic.make_batch(2, 64, 64, 3, seed=1400)
    feats = dict(images=images, **feed_kw(images, labels))
    spec = models.model_fn(feats, labels, ModeKeys.TRAIN, params)
    model = spec.model
    assert spec.train_op is not None and spec.loss is not None
    model.feed(images, labels, **feed_kw(images, labels))
    losses = []
    for _ in range(4):
        spec.train_op.run(with_metrics=True)
        losses.append(spec.train_op.fetch_loss())
    ctx.check_device()
    assert params["solver_instance"].global_step == 4
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    md = model.collect_metrics()
    assert set(md) == {"Liver/Dice", "Tumor/Dice"} and all(0.0 <= v <= 1.0 for v in md.values())
    # forward-only in EVAL mode through the same object (core/models.py:262-281)
    model.mode = ModeKeys.EVAL
    prob = model.run_forward()
    assert prob.shape == (2, 64, 64, 3) and np.allclose(prob.sum(-1), 1.0, atol=1e-5)
    for cls in ("Liver", "Tumor"):
        m = model.predictions[cls + "Pred"]
        assert m.dtype == np.uint8 and m.shape == (2, 64, 64, 1)
        assert np.array_equal(m[..., 0], (prob[..., ("Liver", "Tumor").index(cls) + 1] > 0.5).astype(np.uint8))
    eng = model.engine
    first = losses[0]
    model.engine = None
    eng.close()
    return first, args


def test_unet_through_model_fn_matches_engine(ctx):
    from boxsegliver_b200.engine import EngineConfig, UNetEngine
    first, args = _drive(ctx, "UNet", lambda im, lb: {})
    cfg = EngineConfig(batch=2, height=64, width=64, channel=3, classes=("Background", "Liver", "Tumor"),
                       weight_decay_rate=1e-5, loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4))
    eng = UNetEngine(ctx, cfg)
    eng.init_weights(seed=3)
    images, labels = synthetic.make_batch(2, 64, 64, 3, seed=1400)
    eng.set_inputs(images, labels)
    eng.train_step(1e-3)
    direct = sum(eng.read_loss())
    eng.close()
    assert first == pytest.approx(direct, rel=1e-6)


def test_unetinter_through_model_fn(ctx):
    def feed(images, labels):
        return dict(sp_guide=synthetic.make_guides(images, labels, 200, 2, seed=1)[1])
    _drive(ctx, "UNetInter", feed, normalizer="instance_norm", use_spatial=True, guide_channel=2,
           loss_type="xentropy+dice", mid_cat=False)


def test_unetinter_mid_cat_through_model_fn(ctx):
    """--mid_cat as in run_scripts/template: the guide joins the first block's output in front of the first max-pool."""
    def feed(images, labels):
        return dict(sp_guide=synthetic.make_guides(images, labels, 200, 2, seed=1)[1])
    _drive(ctx, "UNetInter", feed, normalizer="instance_norm", use_spatial=True, guide_channel=2,
           loss_type="xentropy+dice", mid_cat=True, img_grad=True)    # --img_grad: a no-op in UNetInter.py:82-86


def test_gunet_through_model_fn(ctx):
    def feed(images, labels):
        c, g = synthetic.make_guides(images, labels, 200, 1, seed=1)
        return dict(context=c, sp_guide=g)
    _drive(ctx, "GUNet", feed, normalizer="instance_norm", use_spatial=True, use_context=True, guide_channel=1,
           side_dropout=0.5)


def test_gunet_backbone_dropout_through_model_fn(ctx):
    """--dropout (core/models.py:87, GUNet.py:189-190) reaches the engine through the reference's flag."""
    def feed(images, labels):
        c, g = synthetic.make_guides(images, labels, 200, 1, seed=1)
        return dict(context=c, sp_guide=g)
    _drive(ctx, "GUNet", feed, normalizer="instance_norm", use_spatial=True, use_context=True, guide_channel=1,
           side_dropout=0.5, dropout=0.2)


def test_unsupported_flags_raise(ctx):
    args = _args("UNetInter", normalizer="instance_norm", without_norm=True)
    params = models.get_model_params(args)
    params.update(args=args, solver=solver.Solver, ctx=ctx, world=1)
    with pytest.raises(NotImplementedError, match="without_norm"):
        models.model_fn(dict(images=None), None, ModeKeys.TRAIN, params)


def test_engine_variable_names_follow_slim_scoping(ctx):
    from boxsegliver_b200.engine import EngineConfig, UNetEngine
    from tests.slim_names import unet_variable_names
    for norm in ("batch_norm", "instance_norm"):
        eng = UNetEngine(ctx, EngineConfig(batch=1, height=32, width=32, normalizer=norm, training=False))
        assert set(eng.params) == set(unet_variable_names(4, norm)), norm
        eng.close()


def _one_step(ctx, **kw):
    from boxsegliver_b200.engine import EngineConfig, UNetEngine
    cfg = EngineConfig(batch=2, height=32, width=32, weight_decay_rate=1e-4, loss_weight_type="numerical",
                       loss_numeric_w=(0.2, 0.4, 4.4), **kw)
    eng = UNetEngine(ctx, cfg)
    w0 = eng.init_weights(seed=5)
    images, labels = synthetic.make_batch(2, 32, 32, 3, seed=1400)
    eng.set_inputs(images, labels)
    eng.forward(True)
    eng.loss_backward()
    g = eng.get_grads()
    eng.optimizer_step(1e-3)
    w1 = eng.get_weights()
    eng.close()
    return cfg, w0, g, w1


@pytest.mark.parametrize("kw", [
    dict(optimizer="adam", adam_beta1=0.5, adam_beta2=0.999, adam_eps=1e-4),
    dict(optimizer="adamw", adamw_weight_decay=0.05),
    dict(optimizer="momentum", momentum=0.7, use_nesterov=True),
])
def test_engine_honours_optimizer_hyperparameters(ctx, kw):
    """--adam_beta1/2/eps, --mm_mm, --mm_nesterov and AdamW are not silently ignored: the engine's post-step weights
    equal oracle.tf_ops applied to the engine's own gradients with the SAME hyper-parameters."""
    from oracle import tf_ops as O
    cfg, w0, g, w1 = _one_step(ctx, **kw)
    for name in ("UNet/ED-Bridge/ED-Bridge_1/weights", "UNet/Decode1/Conv2d_transpose/biases",
                 "UNet/Encode2/Repeat/convolution2d_1/BatchNorm/gamma"):
        w = w0[name].astype(np.float64)
        reg = not name.endswith(("gamma", "beta"))
        ge = g[name].astype(np.float64) + (cfg.weight_decay_rate * w if reg else 0.0)
        if cfg.optimizer == "momentum":
            want, _ = O.momentum_step(w, ge, 0.0 * w, 1e-3, cfg.momentum, cfg.use_nesterov)
        else:
            decay = cfg.adamw_weight_decay if cfg.optimizer == "adamw" else 0.0
            want, _, _ = O.adam_step(w, ge, 0.0 * w, 0.0 * w, 1, 1e-3, cfg.adam_beta1, cfg.adam_beta2, cfg.adam_eps,
                                     decoupled_decay=decay)
        upd, ref = w1[name].astype(np.float64) - w, want - w
        # the update is rounded onto the fp32 grid of w (half an ulp per element on top of the 1e-4 gate)
        assert np.linalg.norm(upd - ref) <= 1e-4 * np.linalg.norm(ref) + 2.0 ** -24 * np.linalg.norm(w) + 1e-9, (name, kw)


def test_trunc_norm_initializer_and_flag_plumbing(ctx):
    """--weight_init trunc_norm (NetworksV2/base.py:137-139): N(0, 0.01) truncated at 2 sigma, through the model class."""
    args = _args("UNet", weight_init="trunc_norm", adam_beta1=0.8)
    params = models.get_model_params(args)
    params.update(args=args, solver=solver.Solver, ctx=ctx, world=1)
    images, labels = synthetic.make_batch(2, 64, 64, 3, seed=1400)
    spec = models.model_fn(dict(images=images), labels, ModeKeys.TRAIN, params)
    eng = spec.model.engine
    assert (eng.cfg.weight_init, eng.cfg.adam_beta1, eng.cfg.adam_beta2) == ("trunc_norm", 0.8, 0.999)
    w = eng.get_weights()["UNet/Decode2/Repeat/convolution2d_1/weights"]
    assert np.abs(w).max() <= 0.02 + 1e-7 and 0.0085 < w.std() < 0.0090        # sd of N(0,1) cut at 2 sigma = 0.8796
    assert abs(w.mean()) < 1e-4
    spec.model.engine = None
    eng.close()
    with pytest.raises(TypeError, match="momentum"):
        bad = _args("UNet", optimizer="Momentum", mm_nesterov=True)
        p2 = models.get_model_params(bad)
        p2.update(args=bad, solver=solver.Solver, ctx=ctx, world=1)
        models.model_fn(dict(images=images), labels, ModeKeys.TRAIN, p2)
