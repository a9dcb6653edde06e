"""Pins oracle/input_ref.py (CPU only): bilinear resize against torch's align_corners=True interpolation, nearest
resize and the label scaling against hand-computed cases, the Gaussian guide against a direct formula, flips and the
noise stream's range / determinism."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import input_ref as R


def test_bilinear_align_corners_matches_torch():
    rng = np.random.default_rng(0)
    for (h, w), (H, W) in (((37, 53), (64, 48)), ((300, 280), (256, 256)), ((16, 16), (16, 16)), ((5, 7), (1, 9))):
        x = rng.integers(0, 4000, (h, w)).astype(np.uint16)
        got = R.resize_bilinear_align(x, H, W)
        ref = F.interpolate(torch.tensor(x.astype(np.float64))[None, None], size=(H, W), mode="bilinear",
                            align_corners=True)[0, 0].numpy()
        if H == 1:      # TF: scale 0 for a single output row (first row); torch does the same with align_corners
            ref = F.interpolate(torch.tensor(x[:1].astype(np.float64))[None, None], size=(1, W), mode="bilinear",
                                align_corners=True)[0, 0].numpy()
        assert got.dtype == np.float32 and np.allclose(got, ref, rtol=2e-6, atol=0.4)   # fp32 sample positions (err ~3e-5 px) x neighbour differences up to 4000
    # corners are reproduced exactly and identity resize is the identity
    assert got.shape == (1, 9)
    x = rng.integers(0, 65535, (9, 11)).astype(np.uint16)
    y = R.resize_bilinear_align(x, 20, 30)
    assert y[0, 0] == x[0, 0] and y[-1, -1] == x[-1, -1] and y[0, -1] == x[0, -1]
    assert np.array_equal(R.resize_bilinear_align(x, 9, 11), x.astype(np.float32))


def test_nearest_align_corners_hand_cases():
    x = np.arange(5, dtype=np.uint8)[None, :].repeat(2, 0)          # width 5 -> 9: scale 0.5, roundf half away
    assert R.resize_nearest_align(x, 2, 9)[0].tolist() == [0, 1, 1, 2, 2, 3, 3, 4, 4]
    x = np.arange(4, dtype=np.uint8)[None, :].repeat(2, 0)          # width 4 -> 2: scale 3 -> {0, 3}
    assert R.resize_nearest_align(x, 2, 2)[0].tolist() == [0, 3]
    x = np.arange(10, dtype=np.uint8)[:, None].repeat(3, 1)         # height 10 -> 4: scale 3 -> {0, 3, 6, 9}
    assert R.resize_nearest_align(x, 4, 3)[:, 0].tolist() == [0, 3, 6, 9]


def test_sample_pipeline_properties():
    rng = np.random.default_rng(1)
    slices = rng.integers(0, 3000, (3, 64, 64)).astype(np.uint16)
    seg = (rng.integers(0, 3, (64, 64)) * 64).astype(np.uint8)       # LiTS labels stored as k * 64 (lab_scale 64)
    kw = dict(bbox=(5, 9, 40, 48), clip=(900.0, 1300.0), lab_scale=64, out_hw=(32, 32))
    img, lab, _ = R.data_processing_train(slices, seg, **kw)
    assert img.shape == (32, 32, 3) and img.dtype == np.float32 and 0.0 <= img.min() and img.max() <= 1.0
    assert lab.dtype == np.int32 and set(np.unique(lab)) <= {0, 1, 2}
    assert lab[0, 0] == seg[5, 9] // 64 and lab[-1, -1] == seg[5 + 39, 9 + 47] // 64
    # flips act on image and label together
    img_lr, lab_lr, _ = R.data_processing_train(slices, seg, flip=1, **kw)
    img_ud, lab_ud, _ = R.data_processing_train(slices, seg, flip=2, **kw)
    assert np.array_equal(img_lr, img[:, ::-1]) and np.array_equal(lab_lr, lab[:, ::-1])
    assert np.array_equal(img_ud, img[::-1]) and np.array_equal(lab_ud, lab[::-1])
    # noise: bounded by the scale, deterministic in (seed, offset), absent in empty slices
    a, _, _ = R.data_processing_train(slices, seg, noise_scale=0.05, seed=7, offset=3, present=(1, 1, 0), **kw)
    b, _, _ = R.data_processing_train(slices, seg, noise_scale=0.05, seed=7, offset=3, present=(1, 1, 0), **kw)
    assert np.array_equal(a, b) and np.all(a[..., 2] == 0)
    d = a[..., :2] - img[..., :2]
    assert np.abs(d).max() <= 0.05 + 1e-6 and np.abs(d).max() > 0.04 and abs(d.mean()) < 5e-3
    c, _, _ = R.data_processing_train(slices, seg, noise_scale=0.05, seed=7, offset=4, **kw)
    assert not np.array_equal(a[..., 0], c[..., 0])


def test_spatial_guide_matches_direct_formula():
    centers = np.array([[10.0, 12.0], [30.0, 5.0]], np.float32)
    stddevs = np.array([[3.0, 0.5], [4.0, 6.0]], np.float32)        # 0.5 is raised to min_std = 1
    slices = np.zeros((1, 64, 64), np.uint16)
    _, _, g = R.data_processing_train(slices, None, (0, 0, 40, 40), (0.0, 1.0), 1, (40, 40), centers=centers,
                                      stddevs=stddevs, with_guide=True)
    yy, xx = np.mgrid[0:40, 0:40].astype(np.float64)
    sd = np.maximum(stddevs.astype(np.float64), 1.0)
    ref = np.max([np.exp(-((yy - c[0]) ** 2 / (2 * s[0] ** 2) + (xx - c[1]) ** 2 / (2 * s[1] ** 2)))
                  for c, s in zip(centers.astype(np.float64), sd)], axis=0) / 2 + 0.5
    assert g.shape == (40, 40, 1) and np.allclose(g[..., 0], ref, atol=1e-6)
    assert g[10, 12, 0] == 1.0 and g.min() >= 0.5
    _, _, g0 = R.data_processing_train(slices, None, (0, 0, 40, 40), (0.0, 1.0), 1, (40, 40), centers=np.zeros((0, 2)),
                                       stddevs=np.zeros((0, 2)), with_guide=True)
    assert np.all(g0 == 0.5)
