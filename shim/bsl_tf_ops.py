"""Python half of the TensorFlow 1.13 custom-op binding (shim/tf_custom_op.cc): loads libbsl_tf_ops.so, attaches the
gradients `optimizer.minimize` (/root/reference/core/solver.py:239) looks up, and offers the slim-signature layer
functions so that `NetworksV2/UNet.py:45-117` runs with `import bsl_tf_ops as slim`-style indirection.

NOT importable in this repo's image (TensorFlow 1.13 needs Python <= 3.7; requirements.txt:2): it is the file a
maintainer of the reference adds next to `NetworksV2/`. tests/test_shim_syntax.py checks that it parses, that every op it
uses is registered in tf_custom_op.cc and that every differentiable op has a gradient here. The same C entry points are
exercised for real through ctypes (boxsegliver_b200/_lib.py) by the -m gpu tests, and the same slim signatures through
boxsegliver_b200/bsl_slim.py.
"""
import os

import tensorflow as tf
from tensorflow.python.framework import ops

_lib = tf.load_op_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbsl_tf_ops.so"))


def _bf16(x):
    return tf.cast(x, tf.bfloat16)


# ------------------------------------------------------------------------------------------------ gradients
@ops.RegisterGradient("BslConv2D")
def _conv2d_grad(op, dy):
    x, w = op.inputs
    kh, kw = w.shape[0].value, w.shape[1].value
    return (_lib.bsl_conv2d_backprop_input(dy, w),
            _bf16(_lib.bsl_conv2d_backprop_filter(x, dy, kh=kh, kw=kw)))   # cast back: the variable's bf16 shadow


@ops.RegisterGradient("BslConv2DStats")
def _conv2d_stats_grad(op, dy, dsums):
    # the statistics output feeds only BslNormRelu, whose gradient op returns the complete dy (FusedBatchNormGrad):
    # no gradient flows through `sums` itself
    return _conv2d_grad(op, dy)


@ops.RegisterGradient("BslHeadConv")
def _head_grad(op, dlogits):
    x, w, _ = op.inputs
    dx, dw, db = _lib.bsl_head_conv_grad(x, w, dlogits)
    return dx, dw, db


@ops.RegisterGradient("BslConv2DTranspose")
def _conv2d_transpose_grad(op, dy):
    x, w, _ = op.inputs
    dx, dw, db = _lib.bsl_conv2d_transpose_grad(x, w, op.outputs[0], dy)
    return dx, _bf16(dw), db


@ops.RegisterGradient("BslConv3D")
def _conv3d_grad(op, dy):
    x, w = op.inputs
    dx, dw = _lib.bsl_conv3d_grad(x, w, dy, ksize=op.get_attr("ksize"), strides=op.get_attr("strides"))
    return dx, _bf16(dw)


@ops.RegisterGradient("BslConv3DTranspose")
def _conv3d_transpose_grad(op, dy):
    x, w = op.inputs
    dx, dw = _lib.bsl_conv3d_transpose_grad(x, w, op.outputs[0], dy, sd=op.get_attr("sd"))
    return dx, _bf16(dw)


@ops.RegisterGradient("BslNormRelu")
def _norm_relu_grad(op, da, dpooled, *unused):
    y, _, gamma, beta = op.inputs[:4]
    a = op.outputs[0]
    if op.get_attr("pool"):
        # MaxPoolGrad + AddN of the skip gradient in one pass (UNet.py:81,91-93)
        da = _lib.bsl_max_pool_grad_add(a, dpooled, da)
    dy, dgamma, dbeta = _lib.bsl_norm_relu_grad(
        y, da, op.outputs[2], op.outputs[3], op.outputs[4], op.outputs[5], mode=op.get_attr("mode"),
        epsilon=op.get_attr("epsilon"), decay=op.get_attr("decay"), center=op.get_attr("center"),
        scale=op.get_attr("scale"))
    return dy, None, dgamma, dbeta, None, None, None


@ops.RegisterGradient("BslWeightedXent")
def _wxent_grad(op, dloss, _):
    return op.outputs[1] * dloss, None


@ops.RegisterGradient("BslDiceLoss")
def _dice_grad(op, dloss, _):
    return op.outputs[1] * dloss, None


ops.NotDifferentiable("BslSoftmaxThreshold")
ops.NotDifferentiable("BslStemIm2col")
ops.NotDifferentiable("BslNormStats")
ops.NotDifferentiable("BslAvgPool2x2")


# ------------------------------------------------------------------------------------------------ slim signatures
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", rate=1, activation_fn=tf.nn.relu,
           normalizer_fn=None, normalizer_params=None, weights_initializer=None, weights_regularizer=None,
           biases_initializer=tf.zeros_initializer(), biases_regularizer=None, outputs_collections=None, scope=None,
           pool=False):
    """slim.conv2d on the sm_100a kernels: BslConv2DStats -> BslNormRelu (norm + ReLU, optionally + 2x2 max-pool), or
    BslHeadConv for the 1x1 logits layer. Variables keep slim's names ("weights", "BatchNorm/gamma", ...)."""
    k = kernel_size if isinstance(kernel_size, (list, tuple)) else (kernel_size, kernel_size)
    if stride != 1 or rate != 1 or padding != "SAME":
        raise NotImplementedError("accelerated path: stride 1, rate 1, SAME")
    with tf.variable_scope(scope, "Conv", [inputs]):
        cin = inputs.shape[-1].value
        w = tf.get_variable("weights", [k[0], k[1], cin, num_outputs], initializer=weights_initializer,
                            regularizer=weights_regularizer)
        if normalizer_fn is None:
            b = tf.get_variable("biases", [num_outputs], initializer=biases_initializer, regularizer=biases_regularizer)
            return _lib.bsl_head_conv(inputs, tf.reshape(w, [cin, num_outputs]), b)
        params = dict(normalizer_params or {})
        batch = normalizer_fn.__name__ == "batch_norm"
        y, sums = _lib.bsl_conv2d_stats(inputs, _bf16(w), imgs_per_group=0 if batch else 1)
        with tf.variable_scope("BatchNorm" if batch else "InstanceNorm"):
            beta = tf.get_variable("beta", [num_outputs], initializer=tf.zeros_initializer())
            gamma = tf.get_variable("gamma", [num_outputs], initializer=tf.ones_initializer())
            mm = tf.get_variable("moving_mean", [num_outputs], initializer=tf.zeros_initializer(), trainable=False)
            mv = tf.get_variable("moving_variance", [num_outputs], initializer=tf.ones_initializer(), trainable=False)
        out = _lib.bsl_norm_relu(y, sums, gamma, beta, mm, mv, params.get("is_training", True), mode=0 if batch else 1,
                                 epsilon=params.get("epsilon", 1e-3 if batch else 1e-6), decay=params.get("decay", 0.999),
                                 center=int(params.get("center", True)), scale=int(params.get("scale", not batch)),
                                 pool=int(pool))
        return (out[0], out[1]) if pool else out[0]


def conv2d_transpose(inputs, num_outputs, kernel_size, stride=1, activation_fn=tf.nn.relu, weights_initializer=None,
                     weights_regularizer=None, biases_initializer=tf.zeros_initializer(), biases_regularizer=None,
                     outputs_collections=None, scope=None, **unused):
    if kernel_size != 2 or stride != 2:
        raise NotImplementedError("accelerated path: conv2d_transpose 2x2 stride 2")
    with tf.variable_scope(scope, "Conv2d_transpose", [inputs]):
        cin = inputs.shape[-1].value
        w = tf.get_variable("weights", [2, 2, num_outputs, cin], initializer=weights_initializer,
                            regularizer=weights_regularizer)
        b = tf.get_variable("biases", [num_outputs], initializer=biases_initializer, regularizer=biases_regularizer)
        return _lib.bsl_conv2d_transpose(inputs, _bf16(w), b)


def weighted_sparse_softmax_cross_entropy(logits, labels, w_type="none", numeric_w=(), proportion_decay=1000.0,
                                          loss_scale=1.0):
    """loss_metrics.weighted_sparse_softmax_cross_entropy (loss_metrics.py:172-177) as one op."""
    wt = {"none": 0, "numerical": 1, "proportion": 2}[w_type]
    loss, _ = _lib.bsl_weighted_xent(logits, labels, weight_type=wt, numeric_w=list(numeric_w),
                                     proportion_decay=proportion_decay, loss_scale=loss_scale)
    tf.losses.add_loss(loss)
    return loss
