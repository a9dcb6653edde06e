"""Pixel-pair packing of UNet3D's full-resolution level (boxsegliver_b200/unet3d_engine.py): the 30-channel (1,3,3)
convolutions of /root/reference/NetworksV2/UNet3D.py:31-91 run as 64-channel convolutions over pairs of horizontally
adjacent voxels with a re-arranged "super" filter. CPU check of the index tables against the oracle's conv3d: forward,
data gradient and the folded filter gradient are EXACTLY those of the real convolution (fp64), for the stride-1 layers,
the strided layer that leaves the level (TF SAME padding: extra pad on the far side) and the stem's im2col GEMM."""
from types import SimpleNamespace as NS

import numpy as np

from boxsegliver_b200.unet3d_engine import UNet3DEngine as E
from oracle import tf_ops as O


def _super(idx, w):
    return np.where(idx >= 0, w.ravel()[np.maximum(idx, 0)], 0.0)


def _fold(idx, g_super, msize):
    fold = E._fold_table(idx, msize)
    gs = np.append(g_super.ravel(), 0.0)       # index -1 -> the appended zero
    return gs[fold[:, 0]] + gs[fold[:, 1]]


def test_stride1_super_conv_equals_real_conv():
    rng = np.random.default_rng(0)
    n, d, h, w, cinp, coutp = 2, 2, 6, 8, 4, 3
    x, wt = rng.standard_normal((n, d, h, w, cinp)), rng.standard_normal((1, 3, 3, cinp, coutp))
    dy = rng.standard_normal((n, d, h, w, coutp))
    idx = E._super_index(NS(pair="conv", cinp=cinp, coutp=coutp))
    ws = _super(idx, wt)                                              # (3, 3, 2 cinp, 2 coutp)
    xs, dys = x.reshape(n * d, h, w // 2, 2 * cinp), dy.reshape(n * d, h, w // 2, 2 * coutp)
    assert np.allclose(O.conv2d(xs, ws).reshape(n, d, h, w, coutp), O.conv3d(x, wt), atol=1e-12)
    assert np.allclose(O.conv2d_backprop_input(xs.shape, ws, dys).reshape(x.shape),
                       O.conv3d_backprop_input(x.shape, wt, dy), atol=1e-12)
    g = _fold(idx, O.conv2d_backprop_filter(xs, ws.shape, dys), wt.size)
    assert np.allclose(g, O.conv3d_backprop_filter(x, wt.shape, dy).ravel(), atol=1e-12)
    # every stored tap is copied exactly twice, and a quarter of the outer super taps is populated
    assert (np.bincount(idx[idx >= 0], minlength=wt.size) == 2).all()
    assert (idx.reshape(3, 3, 2, cinp, 2, coutp)[:, 0] >= 0).mean() == 0.25


def test_strided_layer_leaving_the_packed_level():
    rng = np.random.default_rng(1)
    n, d, h, w, cinp, coutp = 2, 2, 8, 8, 4, 5
    x, wt = rng.standard_normal((n, d, h, w, cinp)), rng.standard_normal((1, 3, 3, cinp, coutp))
    idx = E._super_index(NS(pair="strided", cinp=cinp, coutp=coutp))
    ws = _super(idx, wt).reshape(1, 3, 2, 2 * cinp, coutp)            # 2-tap rows: super voxels X and X + 1
    xs = x.reshape(n, d, h, w // 2, 2 * cinp)
    y = O.conv3d(x, wt, (1, 2, 2))
    assert np.allclose(O.conv3d(xs, ws, (1, 2, 1)), y, atol=1e-12)    # stride 1 over super voxels = stride 2 over voxels
    dy = rng.standard_normal(y.shape)
    assert np.allclose(O.conv3d_backprop_input(xs.shape, ws, dy, (1, 2, 1)).reshape(x.shape),
                       O.conv3d_backprop_input(x.shape, wt, dy, (1, 2, 2)), atol=1e-12)
    g = _fold(idx, O.conv3d_backprop_filter(xs, ws.shape, dy, (1, 2, 1)), wt.size)
    assert np.allclose(g, O.conv3d_backprop_filter(x, wt.shape, dy, (1, 2, 2)).ravel(), atol=1e-12)
    assert not (idx.reshape(3, 2, 2, cinp, coutp)[:, 1, 1] >= 0).any()   # voxel 2X + 3 is outside the 3-tap row


def test_super_conv_over_the_concat_buffer():
    """conv_d0/conv1 reads the level's concat buffer, 128 lanes per voxel pair: [enc even | enc odd | up even | up odd]."""
    rng = np.random.default_rng(3)
    n, d, h, w, half, coutp = 1, 2, 4, 8, 3, 2
    enc, up = rng.standard_normal((n, d, h, w, half)), rng.standard_normal((n, d, h, w, half))
    wt = rng.standard_normal((1, 3, 3, 2 * half, coutp))
    idx = E._super_index(NS(pair="conv", catpair=True, cinp=2 * half, coutp=coutp))
    ws = _super(idx, wt)
    cat = np.concatenate((enc.reshape(n * d, h, w // 2, 2 * half), up.reshape(n * d, h, w // 2, 2 * half)), axis=-1)
    ref = O.conv3d(np.concatenate((enc, up), axis=-1), wt)
    assert np.allclose(O.conv2d(cat, ws).reshape(ref.shape), ref, atol=1e-12)


def test_stem_block_diagonal():
    rng = np.random.default_rng(2)
    m = rng.standard_normal((4, 3))
    idx = E._super_index(NS(pair="stem", cinp=4, coutp=3))
    ws = _super(idx, m)
    cols = rng.standard_normal((10, 2, 4))                            # im2col rows of 10 voxel pairs
    assert np.allclose((cols.reshape(10, 8) @ ws).reshape(10, 2, 3), cols @ m, atol=1e-12)
    g = _fold(idx, rng.standard_normal(ws.shape) * (idx >= 0), m.size)
    assert g.shape == (12,)
