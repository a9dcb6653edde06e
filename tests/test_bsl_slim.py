"""bsl_slim: the reference's slim layer-function seam. The body of UNet._build_network
(/root/reference/NetworksV2/UNet.py:41-100) is restated here call for call -- same functions, same arguments, same
scopes -- against `boxsegliver_b200.bsl_slim`, and must (CPU) produce the variable names slim's scoping rules give and
the engine's own layer table, and (GPU) lower to exactly the C-ABI call sequence of engine.UNetEngine with bit-identical
results."""
import types

import numpy as np
import pytest

from boxsegliver_b200 import bsl_slim as slim
from tests.slim_names import unet_variable_names

tf = slim.tf


def build_unet(images, num_classes, normalizer="batch_norm", weight_decay_rate=1e-5, bias_decay=False, is_training=True,
               init_channels=64, num_down_samples=4, name="UNet"):
    """UNet._net_arg_scope + UNet._build_network (UNet.py:41-100) with `slim` / `tf` bound to bsl_slim."""
    w_reg = slim.l2_regularizer(weight_decay_rate) if weight_decay_rate > 0 else None     # base.py:128-135
    b_reg = None if bias_decay else w_reg
    if normalizer == "batch_norm":                                                        # base.py:153-165
        norm_fn, norm_params = slim.batch_norm, {"scale": True, "is_training": is_training}
    else:
        norm_fn, norm_params = slim.instance_norm, {}
    layers = {}
    with slim.arg_scope([slim.conv2d, slim.conv2d_transpose], weights_regularizer=w_reg,
                        weights_initializer=slim.xavier_initializer(), biases_regularizer=b_reg,
                        outputs_collections=["EPts"]):
        with slim.arg_scope([slim.conv2d], normalizer_fn=norm_fn, normalizer_params=norm_params):
            out_channels = init_channels
            tensor_out = images
            with tf.variable_scope(name, "UNet"):
                for i in range(num_down_samples):
                    with tf.variable_scope("Encode{:d}".format(i + 1)):
                        tensor_out = slim.repeat(tensor_out, 2, slim.conv2d, out_channels, 3)
                        layers["Encode{:d}".format(i + 1)] = tensor_out
                        tensor_out = slim.max_pool2d(tensor_out, [2, 2])
                    out_channels *= 2
                tensor_out = slim.repeat(tensor_out, 2, slim.conv2d, out_channels, 3, scope="ED-Bridge")
                for i in reversed(range(num_down_samples)):
                    out_channels /= 2
                    with tf.variable_scope("Decode{:d}".format(i + 1)):
                        tensor_out = slim.conv2d_transpose(tensor_out, tensor_out.get_shape()[-1] // 2, 2, 2)
                        tensor_out = tf.concat((layers["Encode{:d}".format(i + 1)], tensor_out), axis=-1)
                        tensor_out = slim.repeat(tensor_out, 2, slim.conv2d, out_channels, 3)
                with slim.arg_scope([slim.conv2d], activation_fn=None, normalizer_fn=None, normalizer_params=None):
                    logits = slim.conv2d(tensor_out, num_classes, 1, scope="AdjustChannels")
    return logits


def _engine_table(cfg):
    from boxsegliver_b200.engine import UNetEngine
    return [(L.kind, L.scope, L.cin, L.cout, L.h, L.w, L.level, L.role)
            for L in UNetEngine._layer_specs(types.SimpleNamespace(cfg=cfg))]


@pytest.mark.parametrize("normalizer,depth,hw", [("batch_norm", 4, 64), ("instance_norm", 3, 96)])
def test_captured_graph_gives_slim_names_and_the_engine_layer_table(normalizer, depth, hw):
    from boxsegliver_b200.engine import EngineConfig
    with slim.graph_as_default() as g:
        images = slim.placeholder((2, hw, hw, 3))
        logits = build_unet(images, 3, normalizer=normalizer, num_down_samples=depth)
        assert logits.shape == (2, hw, hw, 3)
        table, info = slim.lower_to_layer_table(g, logits)
    names = set()
    ns = "BatchNorm" if normalizer == "batch_norm" else "InstanceNorm"
    for L in table:
        names.add(L["scope"] + "/weights")
        if L["kind"] in ("stem", "conv"):
            names |= {f"{L['scope']}/{ns}/gamma", f"{L['scope']}/{ns}/beta"}
        else:
            names.add(L["scope"] + "/biases")
    assert names == set(unet_variable_names(depth, normalizer, with_moving=False))
    cfg = EngineConfig(batch=2, height=hw, width=hw, num_down_samples=depth, normalizer=normalizer)
    assert [(L["kind"], L["scope"], L["cin"], L["cout"], L["h"], L["w"], L["level"], L["role"]) for L in table] == \
        _engine_table(cfg)
    assert info["normalizer"] == normalizer and info["num_down_samples"] == depth and info["init_channels"] == 64
    assert info["regularizers"] == {(("l2", 1e-5), ("l2", 1e-5))}


def test_scope_uniquification_and_arg_scope_rules():
    with slim.graph_as_default() as g:
        x = slim.placeholder((1, 16, 16, 8))
        with slim.arg_scope([slim.conv2d], normalizer_fn=slim.batch_norm, normalizer_params={"scale": True}) as sc:
            a = slim.conv2d(x, 8, 3)                       # default scope "Conv"
            b = slim.conv2d(a, 8, 3)                       # uniquified: "Conv_1"
            with tf.variable_scope("blk"):
                c = slim.repeat(b, 2, slim.conv2d, 8, 3)   # blk/Repeat/convolution2d_{1,2}
                d = slim.repeat(c, 1, slim.conv2d, 8, 3)   # second Repeat in the same scope: blk/Repeat_1/...
        with slim.arg_scope(sc):                           # re-entering a captured scope
            e = slim.conv2d(d, 8, 3, scope="named")
        f = slim.conv2d(e, 8, 3, scope="plain")            # outside: no normaliser -> bias
        assert [n.name for n in g.nodes if n.kind == "conv"] == [
            "Conv", "Conv_1", "blk/Repeat/convolution2d_1", "blk/Repeat/convolution2d_2", "blk/Repeat_1/convolution2d_1",
            "named", "plain"]
        assert e.node.attrs["normalizer"] == "batch_norm" and not e.node.attrs["has_bias"]
        assert f.node.attrs["normalizer"] is None and f.node.attrs["has_bias"]
        with pytest.raises(ValueError, match="kwargs must be empty"):
            with slim.arg_scope(sc, activation_fn=None):
                pass


def test_graphs_outside_the_family_are_rejected():
    with slim.graph_as_default() as g:
        x = slim.placeholder((1, 32, 32, 3))
        with slim.arg_scope([slim.conv2d], normalizer_fn=slim.batch_norm, normalizer_params={"scale": True}):
            y = slim.conv2d(x, 64, 3, stride=2)
            with pytest.raises(NotImplementedError, match="stride"):
                slim.lower_to_layer_table(g, y)
    with slim.graph_as_default() as g:
        x = slim.placeholder((1, 32, 32, 3))
        with slim.arg_scope([slim.conv2d], normalizer_fn=slim.batch_norm, normalizer_params={"scale": True}):
            y = slim.repeat(x, 3, slim.conv2d, 64, 3)
            with pytest.raises(NotImplementedError, match="more than two"):
                slim.lower_to_layer_table(g, y)
    with slim.graph_as_default() as g:
        x = slim.placeholder((1, 32, 32, 3))
        y = slim.conv2d(x, 64, 5)
        with pytest.raises(NotImplementedError, match="kernel"):
            slim.lower_to_layer_table(g, y)


@pytest.mark.gpu
def test_lowered_graph_issues_the_engine_launch_sequence(ctx):
    """`same launches as UNetEngine`: the C-ABI call sequence of a training step, logits, loss and every gradient."""
    from boxsegliver_b200 import synthetic
    from boxsegliver_b200.engine import EngineConfig, UNetEngine
    n, hw = 2, 64
    kw = dict(loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4))
    with slim.graph_as_default() as g:
        logits = build_unet(slim.placeholder((n, hw, hw, 3)), 3, weight_decay_rate=1e-5)
        geng = slim.lower(g, logits, ctx, ("Background", "Liver", "Tumor"), **kw)
    ref = UNetEngine(ctx, EngineConfig(batch=n, height=hw, width=hw, weight_decay_rate=1e-5, **kw))
    assert geng.cfg == ref.cfg and set(geng.params) == set(ref.params)
    images, labels = synthetic.make_batch(n, hw, hw, 3, seed=1403)
    out = []
    for eng in (ref, geng):
        eng.init_weights(seed=9)
        eng.set_inputs(images, labels)
        ctx._trace = []
        eng.train_step(1e-3, with_metrics=True)
        trace, ctx._trace = ctx._trace, None
        ctx.check_device()
        out.append((trace, eng.logits.download(np.float32, (n, hw, hw, 3)), eng.read_loss(),
                    eng.G.download(np.float32, (eng.n_train,)), eng.W.download(np.float32, (eng.n_train,))))
        eng.close()
    assert out[0][0] == out[1][0] and len(out[0][0]) > 100
    assert any(name == "bsl_conv2d_fprop_stats" for name in out[1][0])
    for a, b in zip(out[0][1:], out[1][1:]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
