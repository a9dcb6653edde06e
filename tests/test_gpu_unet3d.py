"""Parity of the UNet3D engine against oracle/unet3d_ref.py through the C ABI: every layer on the device's stored
input (fp64 oracle, rel <= 1e-2), loss, dlogits, and all gradients over the device's stored tape (every tensor whose
layer normalises over >= 256 voxels <= 1e-2, north_star bf16 tolerance), with the true channel counts 30/60/120/240/320 (zero-padded storage)."""
import numpy as np
import pytest

from boxsegliver_b200 import synthetic
from boxsegliver_b200.device import round_bf16
from boxsegliver_b200.unet3d_engine import UNet3DConfig, UNet3DEngine
from oracle import unet3d_ref as U
from tests.gpu_util import rel, report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,hw,kw", [
    (2, 4, 32, dict()),                                                     # UNet3D.yml channels, 4 pools
    (1, 2, 64, dict(num_pool_layers=5, use_spatial=True, guide_channel=2, loss_numeric_w=(1.0, 10.0))),
    (2, 6, 32, dict(init_channels=16, max_channels=128, loss_weight_type="proportion", loss_numeric_w=())),
    # the 64-lane zero-padded storage of the full-resolution level (what shapes outside the pixel-pair packing's reach
    # run on: W % 16 != 0, H % 32 != 0, more than 32 channels), forced here through the tuning switch
    (2, 4, 32, dict(_pair="0")),
    (1, 2, 32, dict(init_channels=48, max_channels=96, _pair=None)),         # 48 > 32 channels: never packed
])
def test_unet3d_train_step_parity(ctx, n, d, hw, kw, monkeypatch):
    base = dict(depth=d, height=hw, width=hw, channel=1, weight_decay_rate=3e-5)
    kw = dict(kw)
    pair = kw.pop("_pair", "1")
    if pair is not None:
        monkeypatch.setenv("BSL_UNET3D_PAIR", pair)
    base.update(kw)
    ecfg, rcfg = UNet3DConfig(batch=n, **base), U.UNet3DCfg(**base)
    if rcfg.use_spatial:
        images, labels, guide = synthetic.make_volume_batch(n, d, hw, hw, seed=9, guide_channel=rcfg.guide_channel)
    else:
        images, labels = synthetic.make_volume_batch(n, d, hw, hw, seed=9)
        guide = None
    params = U.init_params(rcfg, seed=6)
    rng = np.random.default_rng(1)
    for k in params:
        if k.endswith(("beta", "biases")):
            params[k] = (0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
        if k.endswith("gamma"):
            params[k] = (1 + 0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
    eng = UNet3DEngine(ctx, ecfg)
    assert eng._pair == (pair == "1")
    assert set(eng.params) == set(params)
    eng.set_weights(params)
    back = eng.get_weights()
    for k in params:
        assert np.array_equal(back[k], params[k]), f"pack/unpack round trip: {k}"
    eng.set_inputs(images, labels, guide)
    eng.forward(True)
    eng.predict_outputs(True)
    eng.loss_backward()
    ctx.check_device()
    kcls = rcfg.num_classes
    logits = eng.logits.download(np.float32, (n, d, hw, hw, kcls))
    dlogits = eng.dlogits.download(np.float32, (n, d, hw, hw, kcls))
    grads = eng.get_grads()
    stored = eng.get_stored_forward()         # also asserts the pad lanes are exactly zero
    stored["logits"] = logits
    counts = eng.read_counts()
    masks = eng.masks.download(np.uint8, (kcls - 1, n, d, hw, hw))
    eng.optimizer_step(1e-3)
    ctx.check_device()
    data_loss, reg_loss = eng.read_loss()
    new_w = eng.get_weights()
    eng.close()

    inputs = dict(images=round_bf16(images).astype(np.float64))
    if guide is not None:
        inputs["sp_guide"] = round_bf16(guide).astype(np.float64)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    tft = U.forward(p64, inputs, rcfg, wrnd=round_bf16, stored=stored)
    assert max(tft.errs.values()) < 1e-2, max(tft.errs.items(), key=lambda t: t[1])
    loss_o, dl = U.loss_and_dlogits(tft, labels, rcfg)
    assert abs(data_loss - loss_o) < 1e-4 * abs(loss_o)
    assert abs(reg_loss - U.regularization_loss(params, rcfg)) < 1e-6
    assert rel(dlogits, dl) < 1e-5
    g_ref = U.backward(tft, dl, rcfg, rnd=round_bf16)
    errs = {name: rel(grads[name], g) for name, g in g_ref.items()}
    assert np.median(list(errs.values())) < 1e-2, errs
    # north_star's 1e-2 is gated on every gradient whose layer normalises over >= 256 voxels. The deepest blocks of these
    # deliberately tiny test volumes normalise over 2 .. 128 voxels: a single bf16 rounding of the incoming gradient moves
    # such statistics by percents in ANY bf16 implementation, so those are reported and bounded at 5e-2 (the full-size
    # volumes of BASELINE cfg4, where every layer has >= 4096 voxels, are covered by tests/test_gpu_baseline_shapes.py).
    vox = {}
    for sp in U.layer_specs(rcfg):
        o = [-(-sp["dhw"][i] // sp["s"][i]) for i in range(3)] if sp["kind"] == "conv" else [sp["dhw"][i] * sp["s"][i] for i in range(3)]
        vox[sp["scope"]] = int(np.prod(o))
    for name, e in errs.items():
        scope = name.rsplit("/InstanceNorm", 1)[0].rsplit("/weights", 1)[0].rsplit("/biases", 1)[0]
        assert e < (1e-2 if vox[scope] >= 256 else 5e-2), (name, e, vox[scope])
    big = [e for nm, e in errs.items() if vox[nm.rsplit("/InstanceNorm", 1)[0].rsplit("/weights", 1)[0].rsplit("/biases", 1)[0]] >= 256]
    report(f"unet3d {n}x{d}x{hw}x{hw}", layer_worst=max(tft.errs.values()), grad_median=float(np.median(list(errs.values()))),
           grad_worst_ge256_voxels=max(big), grad_worst_tiny_statistics=max(errs.values()))
    # masks and integer Dice counts: bit-exact functions of the device's logits
    prob = U.O.softmax(logits)
    decided = np.abs(prob[..., 1] - 0.5) > 1e-6
    assert not ((masks[0] != (prob[..., 1] > 0.5)) & decided).any()
    i_, l_, r_ = U.O.seg_counts(masks[0].reshape(n, -1, 1), labels.reshape(n, -1), 1)
    assert np.array_equal(counts[:, 0, 0], i_) and np.array_equal(counts[:, 0, 1], l_) and np.array_equal(counts[:, 0, 2], r_)
    # Adam on the device gradients reproduces the device update; padded lanes never leak into real weights
    tg = U.total_grads(params, grads, rcfg)
    for name in params:
        w, _, _ = U.O.adam_step(params[name].astype(np.float64), tg[name].astype(np.float64), 0.0, 0.0, 1, 1e-3)
        assert rel(new_w[name].astype(np.float64) - params[name], w - params[name]) < 1e-3, name
