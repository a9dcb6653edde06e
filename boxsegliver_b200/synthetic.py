"""Seeded synthetic CT slices with the reference's input contract (no dataset ships with the reference).

Contract restated from /root/reference/DataLoader/Liver/input_pipeline.py:243-284: images are fp32
[N,H,W,3] -- three adjacent slices of a windowed volume mapped to [0,1] (:256-258) plus
U(-noise, +noise) (--noise_scale 0.05, run_scripts/template/001_unet.sh:18) -- and labels are int32
[N,H,W] in {0 background, 1 liver, 2 tumor}. The phantom is a body ellipse, a liver ellipse and up
to two tumour blobs per slice; slice 0 of a batch (when N >= 3) has no tumour and slice 1 is all
background, so the zero-count branches of _compute_weights and the Dice sums are exercised.
GUNet-style extras (`context` histogram, `sp_guide`) follow input_pipeline_g.py:374-394.
"""
from __future__ import annotations

import numpy as np

FOLD_SEED = 1357  # /root/reference/DataLoader/Liver/input_pipeline.py:139


def _ellipse(yy, xx, cy, cx, ry, rx, rot=0.0):
    c, s = np.cos(rot), np.sin(rot)
    y, x = yy - cy, xx - cx
    u, v = c * x + s * y, -s * x + c * y
    return (u / rx) ** 2 + (v / ry) ** 2 <= 1.0


def make_batch(n: int, h: int, w: int, channels: int = 3, seed: int = FOLD_SEED, noise: float = 0.05,
               num_classes: int = 3):
    """Returns (images fp32 [n,h,w,channels] in ~[0,1], labels int32 [n,h,w])."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    images = np.empty((n, h, w, channels), np.float32)
    labels = np.zeros((n, h, w), np.int32)
    for i in range(n):
        body = _ellipse(yy, xx, h * 0.5, w * 0.5, h * rng.uniform(0.36, 0.45), w * rng.uniform(0.40, 0.48))
        vol = np.where(body, 0.45, 0.05).astype(np.float32)
        lab = np.zeros((h, w), np.int32)
        empty = (n >= 3 and i == 1)
        if not empty:
            lcy, lcx = h * rng.uniform(0.42, 0.55), w * rng.uniform(0.30, 0.42)
            lry, lrx = h * rng.uniform(0.14, 0.22), w * rng.uniform(0.16, 0.24)
            liver = _ellipse(yy, xx, lcy, lcx, lry, lrx, rng.uniform(-0.5, 0.5)) & body
            vol[liver] = 0.62
            lab[liver] = 1
            if num_classes > 2 and not (n >= 3 and i == 0):
                for _ in range(rng.integers(1, 3)):
                    ty, tx = lcy + lry * rng.uniform(-0.5, 0.5), lcx + lrx * rng.uniform(-0.5, 0.5)
                    tr = max(2.0, min(h, w) * rng.uniform(0.02, 0.05))
                    tumor = _ellipse(yy, xx, ty, tx, tr, tr * rng.uniform(0.7, 1.3)) & liver
                    vol[tumor] = 0.50
                    lab[tumor] = 2
        labels[i] = lab
        for c in range(channels):  # neighbouring slices: the same phantom, slightly shifted
            sh = c - channels // 2
            images[i, :, :, c] = np.roll(vol, sh, axis=0)
    images += rng.uniform(-noise, noise, size=images.shape).astype(np.float32)
    return images, labels


def make_context(labels: np.ndarray, images: np.ndarray, hist_scale: float = 20.0, bins: int = 100):
    """GUNet `context` [n, 2*bins]: density histograms of liver and tumour intensities x hist_scale."""
    n = labels.shape[0]
    ctx = np.zeros((n, 2 * bins), np.float32)
    mid = images[..., images.shape[-1] // 2]
    for i in range(n):
        for k, cls in enumerate((1, 2)):
            v = mid[i][labels[i] == cls]
            if v.size:
                hist, _ = np.histogram(v, bins=bins, range=(0.0, 1.0), density=True)
                ctx[i, k * bins:(k + 1) * bins] = np.nan_to_num(hist) * hist_scale / bins
    return ctx


def make_sp_guide(labels: np.ndarray, sigma: float = 6.0):
    """GUNet `sp_guide` [n,h,w,1]: 0.5 + 0.5 * max of Gaussians at tumour centres (0.5 when no tumour)."""
    n, h, w = labels.shape
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.full((n, h, w, 1), 0.5, np.float32)
    for i in range(n):
        ys, xs = np.nonzero(labels[i] == 2)
        if ys.size:
            g = np.exp(-((yy - ys.mean()) ** 2 + (xx - xs.mean()) ** 2) / (2 * sigma * sigma))
            out[i, :, :, 0] = 0.5 + 0.5 * g
    return out


def make_guides(images: np.ndarray, labels: np.ndarray, context_dim: int = 200, guide_channel: int = 1, seed: int = 0):
    """(context [n, context_dim], sp_guide [n,h,w,guide_channel]) as the reference's pipeline feeds GUNet
    (/root/reference/DataLoader/Liver/input_pipeline_g.py:374-394,549-567): histograms + N(0, 0.002) noise, the second
    guide channel (when present) is a Gaussian at the liver centre."""
    rng = np.random.default_rng(seed)
    ctx = make_context(labels, images, bins=context_dim // 2)
    ctx = (ctx + rng.normal(0, 0.002, ctx.shape) * (ctx.sum(axis=1, keepdims=True) > 0)).astype(np.float32)
    g = make_sp_guide(labels)
    if guide_channel == 2:
        n, h, w = labels.shape
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        g2 = np.full((n, h, w, 1), 0.5, np.float32)
        for i in range(n):
            ys, xs = np.nonzero(labels[i] == 1)
            if ys.size:
                g2[i, :, :, 0] = 0.5 + 0.5 * np.exp(-((yy - ys.mean()) ** 2 + (xx - xs.mean()) ** 2) / (2 * 12.0 ** 2))
        g = np.concatenate((g, g2), axis=-1)
    return ctx, np.ascontiguousarray(g, np.float32)


def make_volume_batch(n: int, d: int, h: int, w: int, seed: int = FOLD_SEED, guide_channel: int = 0):
    """UNet3D inputs as /root/reference/DataLoader/NF/input_pipeline_3d.py:352-408 feeds them: z-scored volume
    [n,d,h,w,1] (statistics over non-zero voxels), labels {0,1} [n,d,h,w], optional Gaussian click guides."""
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.mgrid[0:d, 0:h, 0:w].astype(np.float32)
    images = np.zeros((n, d, h, w, 1), np.float32)
    labels = np.zeros((n, d, h, w), np.int32)
    guide = np.zeros((n, d, h, w, max(guide_channel, 1)), np.float32)
    for i in range(n):
        body = ((yy - h / 2) / (0.45 * h)) ** 2 + ((xx - w / 2) / (0.42 * w)) ** 2 < 1
        vol = np.where(body, 0.35, 0.0).astype(np.float32)
        lab = np.zeros((d, h, w), np.int32)
        for _ in range(rng.integers(0, 4) if i else 0, 4):          # sample 0 has no lesion (all-background edge case)
            cz, cy, cx = rng.uniform(0.2, 0.8) * d, rng.uniform(0.25, 0.75) * h, rng.uniform(0.25, 0.75) * w
            rz, ry, rx = rng.uniform(0.1, 0.3) * d + 1, rng.uniform(0.05, 0.15) * h, rng.uniform(0.05, 0.15) * w
            blob = ((zz - cz) / rz) ** 2 + ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1
            vol[blob & body] = 0.7
            lab[blob & body] = 1
            if guide_channel:
                g = np.exp(-(((zz - cz) / 1.0) ** 2 + ((yy - cy) / 5.0) ** 2 + ((xx - cx) / 5.0) ** 2) / 2)
                guide[i, ..., 0] = np.maximum(guide[i, ..., 0], g)
        vol += rng.normal(0, 0.03, vol.shape).astype(np.float32) * body
        nz = vol[vol != 0]
        vol = np.where(vol != 0, (vol - nz.mean()) / (nz.std() + 1e-8), 0.0)
        images[i, ..., 0] = vol
        labels[i] = lab
    return (images, labels, guide) if guide_channel else (images, labels)
