"""Device-side input stage (csrc/augment.cu) against oracle/input_ref.py through the C ABI: images and labels
bit-exact (every fp32 operation is a single rounded operation on both sides, the noise bits are Philox4x32-10),
the Gaussian guide to 1e-6 (expf). Covers crops of different sizes per sample, up- and down-scaling, empty
neighbour slices, all four flip combinations, samples without tumour centres, and a full 512 -> 256 batch."""
import numpy as np
import pytest

from boxsegliver_b200.input_pipeline import DeviceInputStage
from oracle import input_ref as R

pytestmark = pytest.mark.gpu


def _batch(rng, n, c, src, k):
    slices = rng.integers(0, 2500, (n, c, src, src)).astype(np.uint16)
    yy, xx = np.mgrid[0:src, 0:src]
    slices += (300 * np.sin(yy / 17.0) * np.cos(xx / 23.0) + 400).astype(np.uint16)[None, None]
    seg = (rng.integers(0, 3, (n, src, src)) * 64).astype(np.uint8)
    bbox = np.stack([rng.integers(0, src // 4, n), rng.integers(0, src // 4, n),
                     rng.integers(src // 3, src - src // 4, n), rng.integers(src // 3, src - src // 4, n)], axis=1)
    clip = np.stack([rng.uniform(300, 900, n), rng.uniform(1500, 2600, n)], axis=1).astype(np.float32)
    present = np.ones((n, c), np.uint8)
    present[0, 0] = 0
    slices[0, 0] = 0
    flips = np.arange(n) % 4
    centers = [rng.uniform(0, src // 3, (i % (k + 1), 2)).astype(np.float32) for i in range(n)]
    stddevs = [rng.uniform(0.3, 9.0, (len(ci), 2)).astype(np.float32) for ci in centers]
    return slices, seg, bbox, clip, present, flips, centers, stddevs


@pytest.mark.parametrize("n,c,src,out,noise", [(6, 3, 96, (64, 80), 0.05), (5, 1, 64, (96, 96), 0.0),
                                               (8, 3, 512, (256, 256), 0.05)])
def test_input_stage_matches_oracle(ctx, n, c, src, out, noise):
    rng = np.random.default_rng(100 + n)
    slices, seg, bbox, clip, present, flips, centers, stddevs = _batch(rng, n, c, src, 3)
    st = DeviceInputStage(ctx, n, c, (src, src), out, noise_scale=noise, seed=0x1234ABCD5, max_centers=4, with_guide=True)
    st.step = 3
    st.stage(slices, seg, bbox, clip, 64, present, flips, centers, stddevs)
    H, W = out
    bi, bl, bg = ctx.alloc(n * H * W * c * 4), ctx.alloc(n * H * W * 4), ctx.alloc(n * H * W * 4)
    st.run(bi, bl, bg)
    ctx.check_device()
    img = bi.download(np.float32, (n, H, W, c))
    lab = bl.download(np.int32, (n, H, W))
    gd = bg.download(np.float32, (n, H, W, 1))
    for i in range(n):
        ri, rl, rg = R.data_processing_train(slices[i], seg[i], bbox[i], clip[i], 64, out, present=present[i],
                                             noise_scale=noise, seed=0x1234ABCD5, offset=3, sample=i, flip=int(flips[i]),
                                             centers=centers[i], stddevs=stddevs[i], with_guide=True)
        assert np.array_equal(img[i], ri), (i, np.abs(img[i] - ri).max())
        assert np.array_equal(lab[i], rl), i
        assert np.allclose(gd[i], rg, rtol=0, atol=1e-6), (i, np.abs(gd[i] - rg).max())
    assert np.all(img[0, ..., 0] == 0)            # empty neighbour slice stays empty (no noise)
    assert np.all(gd[0] == 0.5)                   # sample 0 has no centres
    for b in (bi, bl, bg):
        b.free()
    st.close()


def test_input_stage_feeds_engine_and_rejects_bad_boxes(ctx):
    """The stage writes the engine's own input buffers; one training step runs on them."""
    from boxsegliver_b200.engine import EngineConfig, UNetEngine
    rng = np.random.default_rng(5)
    n, src, hw = 2, 128, 64
    slices, seg, bbox, clip, present, flips, _, _ = _batch(rng, n, 3, src, 0)
    eng = UNetEngine(ctx, EngineConfig(batch=n, height=hw, width=hw, loss_weight_type="numerical",
                                       loss_numeric_w=(0.2, 0.4, 4.4)))
    eng.init_weights(0)
    st = DeviceInputStage(ctx, n, 3, (src, src), (hw, hw), noise_scale=0.05, seed=1)
    st.stage(slices, seg, bbox, clip, 64, present, flips)
    st.run(eng.images, eng.labels)
    eng.train_step(1e-3)
    ctx.check_device()
    ref = np.stack([R.data_processing_train(slices[i], seg[i], bbox[i], clip[i], 64, (hw, hw), present=present[i],
                                            noise_scale=0.05, seed=1, offset=0, sample=i, flip=int(flips[i]))[0]
                    for i in range(n)])
    assert np.array_equal(eng.images.download(np.float32, (n, hw, hw, 3)), ref)
    assert np.isfinite(sum(eng.read_loss()))
    bad = bbox.copy()
    bad[0, 2] = src
    bad[0, 0] = 1
    with pytest.raises(ValueError):
        st.stage(slices, seg, bad, clip, 64)
    st.close()
    eng.close()
