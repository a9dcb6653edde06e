"""TEST INFRASTRUCTURE (oracle): numpy float32 restatement of the reference's per-sample input map function. Only
tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Follows /root/reference/DataLoader/Liver/input_pipeline.py:243-284 (data_processing_train),
DataLoader/Liver/input_pipeline_g.py:357-412 (sp_guide, flips of the guide) and utils/image_ops.py:209-238,245-320,
396-434 (random_noise, random_flip, create_spatial_guide_2d). The TF-1.13 kernels it calls are third-party
(tensorflow-gpu==1.13, requirements.txt:2, not vendored): their published CPU algorithms are restated here --
resize_bilinear_op.cc (compute_interpolation_weights / compute_lerp) and resize_nearest_neighbor_op.cc with
align_corners=True. PARITY UNPINNED by the reference (it ships no fixtures); pinned instead against
torch.nn.functional.interpolate(align_corners=True) and hand-computed cases in tests/test_oracle_input.py.
Every operation is a single float32 operation in the order the kernels use, so csrc/augment.cu matches bit for bit.
"""
from __future__ import annotations

import numpy as np

from . import tf_ops as O

F = np.float32


def _interp_axis(out_size: int, in_size: int):
    """compute_interpolation_weights with align_corners=True: (lower, upper, lerp) per output index."""
    scale = F(in_size - 1) / F(out_size - 1) if out_size > 1 else F(0)
    pos = np.arange(out_size, dtype=F) * scale
    lo = pos.astype(np.int64)
    hi = np.minimum(lo + 1, in_size - 1)
    return lo, hi, (pos - lo.astype(F)).astype(F)


def resize_bilinear_align(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """tf.image.resize_bilinear(img, (out_h, out_w), align_corners=True) for [h, w] or [h, w, c]; returns float32."""
    x = img.astype(F)
    ylo, yhi, yl = _interp_axis(out_h, x.shape[0])
    xlo, xhi, xl = _interp_axis(out_w, x.shape[1])
    ex = (slice(None), slice(None)) + (None,) * (x.ndim - 2)
    xl_, yl_ = xl[None, :][ex], yl[:, None][ex]
    tl, tr = x[ylo][:, xlo], x[ylo][:, xhi]
    bl, br = x[yhi][:, xlo], x[yhi][:, xhi]
    top = tl + (tr - tl) * xl_
    bot = bl + (br - bl) * xl_
    return (top + (bot - top) * yl_).astype(F)


def _roundf(v: np.ndarray) -> np.ndarray:
    """C roundf for non-negative float32 (half away from zero)."""
    t = np.trunc(v)
    return t + ((v - t) >= F(0.5))


def resize_nearest_align(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """tf.image.resize_nearest_neighbor(img, (out_h, out_w), align_corners=True) for [h, w]."""
    def idx(out_size, in_size):
        scale = F(in_size - 1) / F(out_size - 1) if out_size > 1 else F(0)
        return np.minimum(_roundf(np.arange(out_size, dtype=F) * scale).astype(np.int64), in_size - 1)
    return img[idx(out_h, img.shape[0])][:, idx(out_w, img.shape[1])]


def uniform_noise(shape, scale: float, seed: int, offset: int, start: int = 0) -> np.ndarray:
    """tf.random_uniform(shape, -|s|, |s|) on this repo's Philox stream: element e uses counter e // 4, lane e % 4;
    value = u * (2|s|) + (-|s|) with u = Uint32ToFloat(bits)."""
    n = int(np.prod(shape))
    idx = np.arange(start, start + n, dtype=np.uint64)
    blk = idx >> np.uint64(2)
    ctr = np.stack([blk & np.uint64(0xFFFFFFFF), blk >> np.uint64(32),
                    np.full(n, offset & 0xFFFFFFFF, np.uint64), np.full(n, (offset >> 32) & 0xFFFFFFFF, np.uint64)],
                   axis=-1).astype(np.uint32)
    words = O.philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32))
    bits = words[np.arange(n), (idx & np.uint64(3)).astype(np.int64)]
    u = ((bits & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)).view(F) - F(1.0)
    s = F(abs(scale))
    return (u * (F(2.0) * s) + (-s)).astype(F).reshape(shape)


def create_spatial_guide_2d(shape, centers: np.ndarray, stddevs: np.ndarray) -> np.ndarray:
    """image_ops.create_spatial_guide_2d (Gaussian branch): max over centres of exp(-sum((coords - c)^2 / (2 s^2)))."""
    yy, xx = np.meshgrid(np.arange(shape[0], dtype=F), np.arange(shape[1], dtype=F), indexing="ij")
    out = None
    for (cy, cx), (sy, sx) in zip(centers.astype(F), stddevs.astype(F)):
        dy, dx = yy - cy, xx - cx
        t = (dy * dy) / (F(2.0) * sy * sy) + (dx * dx) / (F(2.0) * sx * sx)
        g = np.exp(-t).astype(F)
        out = g if out is None else np.maximum(out, g)
    return out


def data_processing_train(slices: np.ndarray, seg: np.ndarray | None, bbox, clip, lab_scale: int, out_hw,
                          present=None, noise_scale: float = 0.0, seed: int = 0, offset: int = 0, sample: int = 0,
                          flip: int = 0, centers=None, stddevs=None, min_std: float = 1.0, with_guide: bool = False):
    """One sample. slices uint16 [C, src_h, src_w]; seg uint8 [src_h, src_w]; bbox = (off_row, off_col, h, w).
    Returns (images fp32 [H, W, C], labels int32 [H, W] or None, sp_guide fp32 [H, W, 1] or None)."""
    r0, c0, h, w = (int(v) for v in bbox)
    H, W = out_hw
    C = slices.shape[0]
    img = np.stack([resize_bilinear_align(slices[c, r0:r0 + h, c0:c0 + w], H, W) for c in range(C)], axis=-1)
    cmin, cmax = F(clip[0]), F(clip[1])
    img = ((np.clip(img, cmin, cmax) - cmin) / (cmax - cmin)).astype(F)
    if noise_scale:
        img = img + uniform_noise((H, W, C), noise_scale, seed, offset, start=sample * H * W * C)
        pres = np.ones(C, F) if present is None else np.asarray(present, F)
        img = (img * pres).astype(F)
    labels = None
    if seg is not None:
        lab = resize_nearest_align(seg[r0:r0 + h, c0:c0 + w], H, W)
        labels = (lab.astype(F) / F(lab_scale)).astype(np.int32)
    guide = None
    if with_guide:
        if centers is not None and len(centers) > 0:
            sd = np.maximum(np.asarray(stddevs, F), F(min_std))
            gd = create_spatial_guide_2d((h, w), np.asarray(centers, F), sd)
            guide = (resize_bilinear_align(gd, H, W) / F(2.0) + F(0.5)).astype(F)[..., None]
        else:
            guide = np.full((H, W, 1), 0.5, F)
    if flip & 1:
        img, labels, guide = (None if a is None else a[:, ::-1] for a in (img, labels, guide))
    if flip & 2:
        img, labels, guide = (None if a is None else a[::-1] for a in (img, labels, guide))
    return (np.ascontiguousarray(img), None if labels is None else np.ascontiguousarray(labels),
            None if guide is None else np.ascontiguousarray(guide))
