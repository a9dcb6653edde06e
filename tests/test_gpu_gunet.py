"""Parity of the guided U-Net engine (context MLP + dropout, guide pyramid, modulated instance-norm blocks, trunk,
loss, backward) against oracle/gunet_ref.py, through the C ABI.

Gates (relative L2 per tensor; bf16 path, north_star tolerance 1e-2):
  * every layer evaluated by the fp64 oracle on the tensor the device stored as its input ........ <= 1e-2
  * gradients: oracle backward over the device's stored forward tape (same ReLU masks): tests/gpu_util.gate_gradients
    (median and 90th percentile <= 1e-2, worst <= 1.25e-2 at these tiny shapes; every tensor <= 1e-2 at BASELINE shapes)
  * dropout multipliers (Philox4x32-10) ........................................................... bit-exact
"""
import ctypes as C

import numpy as np
import pytest

from boxsegliver_b200 import _lib, synthetic
from boxsegliver_b200.device import round_bf16
from boxsegliver_b200.gunet_engine import GUNetConfig, GUNetEngine
from oracle import gunet_ref as G
from oracle import tf_ops as O
from tests.gpu_util import gate_gradients, rel, report

pytestmark = pytest.mark.gpu


def test_dropout_mask_bit_exact(ctx):
    n = 12345
    for keep, seed, off in ((0.5, 0, 17), (0.8, 0xDEADBEEFCAFE, 1 << 33), (1.0, 3, 3)):
        buf = ctx.alloc(n * 4)
        d = _lib.DropoutDesc(keep, seed, off)
        ctx.call("bsl_dropout_mask", C.byref(d), C.c_size_t(n), buf.p, ctx.stream)
        got = buf.download(np.float32, (n,))
        buf.free()
        assert np.array_equal(got, O.dropout_multipliers(n, keep, seed, off))


def test_dropout_bf16_bit_exact(ctx):
    """bsl_dropout_bf16 (backbone --dropout and its gradient): bf16(x * multiplier(flat index)), strided in / out."""
    rng = np.random.default_rng(3)
    pixels, c, ld = 777, 24, 32
    x = round_bf16(rng.standard_normal((pixels, ld), dtype=np.float32))
    bx, bo = ctx.bf16_from_f32(x), ctx.alloc(pixels * c * 2)
    for keep, seed, off in ((0.5, 1, 24), (0.75, 0xABCDEF0123, (1 << 34) + 5)):
        d = _lib.DropoutDesc(keep, seed, off)
        ctx.call("bsl_dropout_bf16", C.byref(d), C.c_longlong(pixels), C.c_int(c), bx.p, C.c_int(ld), bo.p, C.c_int(c),
                 ctx.stream)
        got = ctx.bf16_to_f32(bo, (pixels, c))
        want = round_bf16(x[:, :c] * O.dropout_multipliers(pixels * c, keep, seed, off).reshape(pixels, c))
        assert np.array_equal(got, want)
    bx.free(), bo.free()


def test_fc_and_avgpool_ops(ctx):
    rng = np.random.default_rng(0)
    n, cin, cout = 5, 37, 50
    x = rng.standard_normal((n, cin), dtype=np.float32)
    w = rng.standard_normal((cin, cout), dtype=np.float32) * 0.2
    b = rng.standard_normal(cout, dtype=np.float32)
    dy = rng.standard_normal((n, cout), dtype=np.float32)
    d = _lib.FcDesc(n, cin, cout, 1, 1, _lib.DropoutDesc(0.5, 9, 4))
    bx, bw, bb, bdy = (ctx.from_numpy(a) for a in (x, w, b, dy))
    by, bdx, bdw, bdb = ctx.alloc(n * cout * 4), ctx.alloc(n * cin * 4), ctx.alloc(cin * cout * 4), ctx.alloc(cout * 4)
    ws_bytes = ctx.lib.bsl_fc_bwd_workspace(ctx.h, C.byref(d))
    ws = ctx.alloc(ws_bytes)
    ctx.call("bsl_fc_fwd", C.byref(d), bx.p, bw.p, bb.p, by.p, ctx.stream)
    ctx.call("bsl_fc_bwd", C.byref(d), bx.p, bw.p, by.p, bdy.p, bdx.p, bdw.p, bdb.p, ws.p, C.c_size_t(ws_bytes), ctx.stream)
    y = by.download(np.float32, (n, cout))
    mult = O.dropout_multipliers(n * cout, 0.5, 9, 4).reshape(n, cout)
    pre = x.astype(np.float64) @ w + b
    assert rel(y, np.maximum(pre, 0) * mult) < 1e-6
    dpre = dy * mult * (pre > 0)
    dx, dw, db = O.fully_connected_grad(x.astype(np.float64), w.astype(np.float64), dpre.astype(np.float64))
    assert rel(bdx.download(np.float32, (n, cin)), dx) < 1e-5
    assert rel(bdw.download(np.float32, (cin, cout)), dw) < 1e-5
    assert rel(bdb.download(np.float32, (cout,)), db) < 1e-5
    g = rng.uniform(0.5, 1, (3, 8, 12, 2)).astype(np.float32)
    bg, bo = ctx.from_numpy(g), ctx.alloc(3 * 4 * 6 * 2 * 4)
    ctx.call("bsl_avgpool2x2_f32", C.c_int(3), C.c_int(8), C.c_int(12), C.c_int(2), bg.p, bo.p, ctx.stream)
    assert rel(bo.download(np.float32, (3, 4, 6, 2)), O.avg_pool_2x2(g)) < 1e-7
    for buf in (bx, bw, bb, bdy, by, bdx, bdw, bdb, ws, bg, bo):
        buf.free()


def _make(n, hw, **kw):
    base = dict(height=hw, width=hw, channel=3, init_channels=64, num_down_samples=4, weight_decay_rate=1e-5,
                loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4), dropout_seed=21)
    base.update(kw)
    ecfg = GUNetConfig(batch=n, **base)
    rcfg = G.GUNetCfg(**base)
    images, labels = synthetic.make_batch(n, hw, hw, 3, seed=1360 + n)
    context, guide = synthetic.make_guides(images, labels, rcfg.context_dim, rcfg.guide_channel, seed=5)
    return ecfg, rcfg, dict(images=images, context=context, sp_guide=guide), labels


@pytest.mark.parametrize("n,hw,kw", [
    (2, 64, dict(use_context=True, use_spatial=True, guide_channel=1, loss_type="xentropy+dice")),      # GUNet.yml
    (3, 96, dict(use_context=True, use_spatial=True, guide_channel=2, norm_with_center=False,
                 context_fc_channels=(200, 200), loss_type="xentropy")),                                # GUNet_BOTH.yml
                 # (96 x 96: ragged 6 x 6 / 12 x 12 deep levels on the halo-tile kernels)
    (2, 64, dict(use_context=False, use_spatial=True, guide_channel=2, norm_with_scale=True, mod_layers=(0, 2, 4),
                 loss_type="dice")),
    (2, 64, dict(use_context=True, use_spatial=False, side_dropout=0.0, loss_type="xentropy")),
    # after_affine (slim_nets.channel_wise_affine before every encoder ReLU): ext_config/GUNet_BOTH_AA.yml,
    # GUNet_DE_AA.yml (context only), and an un-modulated first block that keeps its own centre / scale
    (2, 64, dict(use_context=True, use_spatial=True, guide_channel=1, norm_with_center=True, norm_with_scale=True,
                 after_affine=True, context_fc_channels=(200, 200), loss_type="xentropy+dice")),
    (2, 64, dict(use_context=True, use_spatial=False, after_affine=True, loss_type="xentropy")),
    (2, 64, dict(use_context=False, use_spatial=True, guide_channel=2, after_affine=True, mod_layers=(0, 1),
                 loss_type="xentropy")),
    # --normalizer batch_norm (GUNet.py:301,321-325): batch statistics with decay 0.99 on the modulated blocks,
    # per-sample gamma_mod / guide on top of them
    (3, 64, dict(use_context=True, use_spatial=True, guide_channel=1, norm_with_center=True, norm_with_scale=True,
                 normalizer="batch_norm", loss_type="xentropy+dice")),
    (2, 64, dict(use_context=True, use_spatial=False, normalizer="batch_norm", loss_type="xentropy")),
    # backbone --dropout (GUNet.py:189-190): Philox mask behind the normaliser of the first conv of every encoder block,
    # in front of the modulation; the second case drops un-modulated blocks (0 and the bridge) too, keep_prob 0.75
    (2, 64, dict(use_context=True, use_spatial=True, guide_channel=1, loss_type="xentropy+dice", dropout=0.5)),
    (2, 64, dict(use_context=True, use_spatial=True, guide_channel=2, norm_with_scale=True, mod_layers=(1, 2, 3),
                 dropout=0.25, loss_type="xentropy")),
    # --img_grad (GUNet.py:333-337, scripts/103_grad.sh): 9 input channels = concat(images, dy, dx)
    (2, 64, dict(use_context=True, use_spatial=True, guide_channel=1, loss_type="xentropy+dice", img_grad=True)),
])
def test_gunet_train_step_parity(ctx, n, hw, kw):
    ecfg, rcfg, inputs, labels = _make(n, hw, **kw)
    params = G.init_params(rcfg, seed=4)
    rng = np.random.default_rng(2)
    for k in params:                      # non-trivial affine / bias values so every term is exercised
        if k.endswith(("beta", "biases")):
            params[k] = (0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
        if k.endswith("gamma"):
            params[k] = (1 + 0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
    eng = GUNetEngine(ctx, ecfg)
    assert set(eng.params) == set(params)
    eng.set_weights(params)
    eng.set_inputs(inputs["images"], labels)
    eng.set_guides(inputs.get("context"), inputs.get("sp_guide"))
    eng.forward(True)
    eng.predict_outputs(True)
    eng.loss_backward()
    ctx.check_device()
    k = rcfg.num_classes
    logits = eng.logits.download(np.float32, (n, hw, hw, k))
    dlogits = eng.dlogits.download(np.float32, (n, hw, hw, k))
    grads = eng.get_grads()
    stored = eng.get_stored_forward()
    stored["logits"] = logits
    ctxp = eng.get_context_params() if rcfg.use_context else None
    dctx = eng.get_context_grad() if rcfg.use_context else None
    eng.optimizer_step(1e-3)
    ctx.check_device()
    data_loss, reg_loss = eng.read_loss()
    new_w = eng.get_weights()
    eng.close()

    # (--img_grad: the device differences the fp32 images and rounds the packed input once; the oracle does the same)
    rin = dict(inputs, images=inputs["images"] if rcfg.img_grad else round_bf16(inputs["images"]))
    # layer by layer on the device's stored inputs (fp64 oracle ops), and the tape for the backward check
    tft = G.forward({k_: v.astype(np.float64) for k_, v in params.items()},
                    {k_: v.astype(np.float64) for k_, v in rin.items()}, rcfg, True, wrnd=round_bf16, stored=stored, step=1)
    assert max(tft.errs.values()) < 1e-2, max(tft.errs.items(), key=lambda t: t[1])
    if rcfg.use_context:
        assert rel(ctxp, tft.ctx_params) < 1e-5
        if rcfg.side_dropout:
            assert any((f["mult"] == 0).any() for f in tft.fc if f["mult"] is not None)
    if rcfg.dropout:
        dropped = [L for L in tft.layers if L.get("mult") is not None]
        assert len(dropped) == rcfg.num_down_samples + 1 and all((L["mult"] == 0).any() for L in dropped)
    loss_o, dl = G.loss_and_dlogits(tft, labels, rcfg)
    assert abs(data_loss - loss_o) < 1e-4 * abs(loss_o)
    assert abs(reg_loss - G.regularization_loss(params, rcfg)) < 1e-6
    assert rel(dlogits, dl) < 1e-5
    for name, v in tft.new_moving.items():     # batch norm: moving statistics (decay 0.99 / 0.999) from the stored outputs
        assert rel(new_w[name], v) < 1e-4, name
    g_ref = G.backward(tft, dl, rcfg, rnd=round_bf16)
    assert set(g_ref) == {k_ for k_ in grads}
    # The context MLP's parameter gradients are a chain: per-layer d(gamma_mod) slices (products of the bf16 trunk
    # backward, gated at 1e-2 here like every trunk gradient) -> fp32 FC backward (gated at 1e-4 on the DEVICE's own
    # d(gamma_mod), i.e. op by op on identical inputs). The end-to-end figure of the FC gradients -- sums over all
    # modulated layers with cancellation -- is reported.
    errs = {name: rel(grads[name], g) for name, g in g_ref.items() if "/context/" not in name}
    e_fc_end_to_end = max([rel(grads[name], g) for name, g in g_ref.items() if "/context/" in name] or [0.0])
    if rcfg.use_context:
        for s_ in G.layer_specs(rcfg):
            if s_.get("mod_off") is not None:
                sl = slice(s_["mod_off"], s_["mod_off"] + s_["cout"])
                errs[f"{s_['scope']}/d_gamma_mod"] = rel(dctx[:, sl], tft.dctx[:, sl])
        for name, g in G.fc_backward(tft, dctx.astype(np.float64)).items():
            assert rel(grads[name], g) < 1e-4, name
    worst = max(errs.items(), key=lambda t: t[1])
    report("gunet step", layer_worst=max(tft.errs.values()), grad_median=float(np.median(list(errs.values()))),
           grad_worst=worst[1], grad_worst_name=worst[0], fc_grads_end_to_end=e_fc_end_to_end)
    gate_gradients(errs)
    assert e_fc_end_to_end < 3e-2


@pytest.mark.parametrize("mid_cat", [False, True])
def test_unetinter_train_step_parity(ctx, mid_cat):
    """UNetInter (/root/reference/NetworksV2/UNetInter.py:73-146): image + 2-channel click guide as a 5-channel
    input, variables under "UNetInter/", no modulation; same gates as the GUNet parity test. --mid_cat (:87-92,124-125):
    the guide joins the first block's output in front of the first max-pool (66 -> 128 channel conv, stored 128 -> 128)."""
    from boxsegliver_b200.gunet_engine import UNetInterConfig, UNetInterEngine
    n, hw = 2, 64
    base = dict(height=hw, width=hw, init_channels=64, num_down_samples=4, weight_decay_rate=1e-5,
                loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4), loss_type="xentropy+dice", mid_cat=mid_cat)
    ecfg = UNetInterConfig(batch=n, channel=3, guide_channel=2, **base)
    rcfg = G.unetinter_cfg(channel=3, guide_channel=2, **base)
    images, labels = synthetic.make_batch(n, hw, hw, 3, seed=1377)
    _, guide = synthetic.make_guides(images, labels, 200, 2, seed=6)
    params = G.init_params(rcfg, seed=9)
    rng = np.random.default_rng(4)
    for k in params:
        if k.endswith(("beta", "biases")):
            params[k] = (0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
        if k.endswith("gamma"):
            params[k] = (1 + 0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
    eng = UNetInterEngine(ctx, ecfg)
    assert set(eng.params) == set(params)
    eng.set_weights(params)
    eng.set_inputs(images, labels, guide)
    eng.forward(True)
    eng.predict_outputs(True)
    eng.loss_backward()
    ctx.check_device()
    logits = eng.logits.download(np.float32, (n, hw, hw, 3))
    dlogits = eng.dlogits.download(np.float32, (n, hw, hw, 3))
    grads = eng.get_grads()
    stored = eng.get_stored_forward()
    stored["logits"] = logits
    eng.optimizer_step(1e-3)          # sum(w^2) of the step's weights is accumulated by the optimizer kernel
    ctx.check_device()
    data_loss, reg_loss = eng.read_loss()
    eng.close()
    rin = {k_: round_bf16(v).astype(np.float64) for k_, v in G.unetinter_inputs(images, guide, mid_cat).items()}
    if mid_cat:
        assert params["UNetInter/Encode/down_conv2/mod_conv1/Conv/weights"].shape == (3, 3, 66, 128)
    tft = G.forward({k_: v.astype(np.float64) for k_, v in params.items()}, rin, rcfg, True, wrnd=round_bf16,
                    stored=stored)
    assert max(tft.errs.values()) < 1e-2, max(tft.errs.items(), key=lambda t: t[1])
    loss_o, dl = G.loss_and_dlogits(tft, labels, rcfg)
    assert abs(data_loss - loss_o) < 1e-4 * abs(loss_o)
    assert abs(reg_loss - G.regularization_loss(params, rcfg)) < 1e-6
    assert rel(dlogits, dl) < 1e-5
    g_ref = G.backward(tft, dl, rcfg, rnd=round_bf16)
    errs = {name: rel(grads[name], g) for name, g in g_ref.items()}
    assert set(g_ref) == set(grads)
    worst = max(errs.items(), key=lambda t: t[1])
    report("unetinter step" + (" mid_cat" if mid_cat else ""), layer_worst=max(tft.errs.values()), grad_median=float(np.median(list(errs.values()))),
           grad_worst=worst[1], grad_worst_name=worst[0])
    gate_gradients(errs)
