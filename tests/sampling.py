"""Sampled-output oracle checks for tensors too large to run the full numpy oracle on (BASELINE cfg2 / cfg3 / cfg4).

A convolution output element depends on one k x k x Cin input patch, so the oracle's arithmetic can be evaluated exactly
(fp64) at a few hundred random output positions from the device's own stored input tensor, without ever forming the
full-size reference. TF SAME padding as in oracle/tf_ops.same_pads: out = ceil(in / s),
pad_total = max((out - 1) * s + k - in, 0), pad_before = pad_total // 2.
TEST INFRASTRUCTURE ONLY (like oracle/).
"""
import numpy as np


def bf16_bits_to_f64(u16):
    return (np.ascontiguousarray(u16).astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def same_pad_before(size, k, s):
    out = -(-size // s)
    return max((out - 1) * s + k - size, 0) // 2, out


def sample_positions(rng, shape, count):
    """`count` distinct-ish random positions of an index space `shape`, always including the corners / borders."""
    idx = [rng.integers(0, s, count) for s in shape]
    corners = min(count, 2 ** len(shape))
    for c in range(corners):                      # border handling (padding) is where conv kernels go wrong
        for d, s in enumerate(shape):
            idx[d][c] = 0 if (c >> d) & 1 == 0 else s - 1
    return [np.asarray(i, np.int64) for i in idx]


def conv_at(x_bits, x_c0, cin, w, pos, stride=(1, 1, 1), cin_map=None):
    """fp64 convolution outputs at `pos` = (n, od, oh, ow) index arrays.
    x_bits: uint16 bf16 bit patterns [N, D, H, W, ld]; channels [x_c0, x_c0 + cin) are the conv input (or the
    columns listed in cin_map, for zero-padded / permuted storage); w: [kd, kh, kw, Cin, Cout] fp64."""
    n, od, oh, ow = pos
    kd, kh, kw, wcin, cout = w.shape
    dims = x_bits.shape[1:4]
    pads = [same_pad_before(dims[i], (kd, kh, kw)[i], stride[i])[0] for i in range(3)]
    cols = (np.arange(cin) + x_c0) if cin_map is None else (np.asarray(cin_map) + x_c0)
    assert len(cols) == wcin, (len(cols), wcin)
    out = np.zeros((len(n), cout), np.float64)
    for a in range(kd):
        idz = od * stride[0] + a - pads[0]
        for b in range(kh):
            idy = oh * stride[1] + b - pads[1]
            for c in range(kw):
                idx = ow * stride[2] + c - pads[2]
                ok = (idz >= 0) & (idz < dims[0]) & (idy >= 0) & (idy < dims[1]) & (idx >= 0) & (idx < dims[2])
                if not ok.any():
                    continue
                patch = np.zeros((len(n), wcin), np.float64)
                sel = np.nonzero(ok)[0]
                rows = x_bits[n[sel], idz[sel], idy[sel], idx[sel]]            # [m, ld]
                patch[sel] = bf16_bits_to_f64(rows[:, cols])
                out += patch @ w[a, b, c]
    return out


def gather(bits, c0, c, pos):
    """fp64 values of channels [c0, c0 + c) at pos (n, d, h, w)."""
    n, d, h, w = pos
    return bf16_bits_to_f64(bits[n, d, h, w][:, c0:c0 + c])


def conv_transpose_at(x_bits, x_c0, cin, w, bias, pos, stride):
    """relu(conv_transpose(x) + bias) with kernel == stride at output positions `pos`; w: [kd, kh, kw, Cout, Cin]."""
    n, od, oh, ow = pos
    sd, sh, sw = stride
    i_pos = (n, od // sd, oh // sh, ow // sw)
    taps = (od % sd, oh % sh, ow % sw)
    xin = gather(x_bits, x_c0, cin, i_pos)                                     # [m, cin]
    wt = w[taps[0], taps[1], taps[2]]                                          # [m, Cout, Cin]
    y = np.einsum("mi,moi->mo", xin, wt)
    if bias is not None:
        y = y + bias[None, :]
    return np.maximum(y, 0.0)


def channel_moments(bits, c0, c, per_sample: bool, chunk=4):
    """fp64 mean and biased variance per channel over all positions (per sample when per_sample) of a bf16 tensor."""
    nimg = bits.shape[0]
    s1 = np.zeros((nimg, c), np.float64)
    s2 = np.zeros((nimg, c), np.float64)
    for i in range(0, nimg, chunk):
        v = bf16_bits_to_f64(bits[i:i + chunk, ..., c0:c0 + c]).reshape(min(chunk, nimg - i), -1, c)
        s1[i:i + chunk] = v.sum(axis=1)
        s2[i:i + chunk] = (v * v).sum(axis=1)
    m = int(np.prod(bits.shape[1:4]))
    if not per_sample:
        s1, s2, m = s1.sum(axis=0, keepdims=True), s2.sum(axis=0, keepdims=True), m * nimg
    mean = s1 / m
    var = np.maximum(s2 / m - mean * mean, 0.0)
    return mean, var


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
