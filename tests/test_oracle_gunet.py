"""Pins oracle/gunet_ref.py (CPU only): an independent torch-CPU autograd implementation of the guided U-Net
(/root/reference/NetworksV2/GUNet.py:162-392) must give the same logits, loss and gradients in fp64; the Philox
stream behind the dropout masks is checked against the published Random123 known-answer vectors."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gunet_ref as G
from oracle import tf_ops as O


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        r = O.philox4x32_10(np.array([ctr], np.uint32), np.array(key, np.uint32))
        assert tuple(int(v) for v in r[0]) == out


def test_dropout_multipliers_rate_and_determinism():
    m = O.dropout_multipliers(100000, 0.5, seed=11, offset=3)
    assert set(np.unique(m)) == {0.0, 2.0}
    assert abs((m > 0).mean() - 0.5) < 0.01
    assert np.array_equal(m, O.dropout_multipliers(100000, 0.5, seed=11, offset=3))
    assert not np.array_equal(m, O.dropout_multipliers(100000, 0.5, seed=11, offset=4))
    assert np.all(O.dropout_multipliers(1000, 1.0, 1, 1) == 1.0)


def _inputs(cfg, n, rng):
    images = rng.uniform(0, 1, (n, cfg.height, cfg.width, cfg.channel))
    context = np.abs(rng.normal(0, 1, (n, cfg.context_dim)))
    context[0, cfg.context_dim // 2:] = 0.0           # tumour-free slice: second histogram all zero
    guide = 0.5 + 0.5 * rng.uniform(0, 1, (n, cfg.height, cfg.width, cfg.guide_channel))
    labels = rng.integers(0, cfg.num_classes, (n, cfg.height, cfg.width)).astype(np.int32)
    return dict(images=images, context=context, sp_guide=guide), labels


def _torch_gunet(params, inputs, labels, cfg, mults):
    """Independent implementation: NCHW torch functional ops + autograd."""
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items() if "/moving_" not in k}
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    x = t(inputs["images"]).permute(0, 3, 1, 2)
    if getattr(cfg, "img_grad", False) and cfg.prefix == "GUNet":     # tf.image.image_gradients: forward differences
        dy = F.pad(x[:, :, 1:] - x[:, :, :-1], (0, 0, 0, 1))
        dx = F.pad(x[:, :, :, 1:] - x[:, :, :, :-1], (0, 1))
        x = torch.cat((x, dy, dx), dim=1)
    ctxp = None
    if cfg.use_context:
        h = t(inputs["context"])
        fcs = G.fc_specs(cfg)
        for k, (sc, _, _, hidden) in enumerate(fcs):
            h = h @ P[f"{sc}/weights"] + P[f"{sc}/biases"]
            if hidden:
                h = torch.relu(h)
                if mults is not None:
                    h = h * t(mults[k])
        ctxp = h
    guides = []
    if cfg.use_spatial:
        g = t(inputs["sp_guide"]).permute(0, 3, 1, 2)
        for i in range(cfg.num_down_samples + 1):
            guides.append(g)
            g = F.avg_pool2d(g, 2)

    def conv(x, s):
        sc, co = s["scope"], s["cout"]
        y = F.conv2d(x, P[f"{sc}/weights"].permute(3, 2, 0, 1), padding=1)
        if cfg.normalizer == "batch_norm":
            y = F.batch_norm(y, None, None, weight=P[f"{sc}/BatchNorm/gamma"] if s["scale"] else None,
                             bias=P[f"{sc}/BatchNorm/beta"] if s["center"] else None, training=True, eps=cfg.bn_eps)
        else:
            y = F.instance_norm(y, weight=P[f"{sc}/InstanceNorm/gamma"] if s["scale"] else None,
                                bias=P[f"{sc}/InstanceNorm/beta"] if s["center"] else None, eps=cfg.in_eps)
        if getattr(cfg, "_drop_mults", None) and s.get("drop") and s["scope"] in cfg._drop_mults:
            y = y * t(cfg._drop_mults[s["scope"]]).permute(0, 3, 1, 2)
        if s["mod_off"] is not None:
            y = y * ctxp[:, s["mod_off"]:s["mod_off"] + co][:, :, None, None]
        if s["sp_off"] is not None:
            ssc = f"GUNet/spatial/conv{s['level'] + 1}"
            full = F.conv2d(guides[s["level"]], P[f"{ssc}/weights"].permute(3, 2, 0, 1), P[f"{ssc}/biases"])
            y = y + full[:, s["sp_off"]:s["sp_off"] + co]
        if s.get("affine"):
            y = y * P[f"{s['affine']}/gamma"][None, :, None, None] + P[f"{s['affine']}/beta"][None, :, None, None]
        return torch.relu(y)

    it = iter(G.layer_specs(cfg))
    skips = []
    for i in range(cfg.num_down_samples + 1):
        x = conv(conv(x, next(it)), next(it))
        if i < cfg.num_down_samples:
            skips.append(x)
            if getattr(cfg, "mid_cat", False) and i == 0:
                x = torch.cat((x, t(inputs["sp_guide"]).permute(0, 3, 1, 2)), dim=1)
            x = F.max_pool2d(x, 2)
    for i in reversed(range(cfg.num_down_samples)):
        s = next(it)
        up = torch.relu(F.conv_transpose2d(x, P[f"{s['scope']}/weights"].permute(3, 2, 0, 1), P[f"{s['scope']}/biases"],
                                           stride=2))
        x = torch.cat((skips[i], up), dim=1)
        x = conv(conv(x, next(it)), next(it))
    s = next(it)
    logits = F.conv2d(x, P[f"{s['scope']}/weights"].permute(3, 2, 0, 1), P[f"{s['scope']}/biases"])
    lab = torch.tensor(labels, dtype=torch.long)
    loss = 0.0
    if "xentropy" in cfg.loss_type:
        assert cfg.loss_weight_type == "numerical"
        oh = F.one_hot(lab, cfg.num_classes).double()
        w = (oh * t(np.array(cfg.loss_numeric_w))).sum(-1)
        w = w / w.sum(dim=(1, 2), keepdim=True) * (cfg.height * cfg.width)
        ce = F.cross_entropy(logits, lab, reduction="none")
        loss = loss + (w * ce).sum() / (w != 0).sum()
    if "dice" in cfg.loss_type:
        prob = torch.softmax(logits, dim=1)[:, 1:]
        oh = F.one_hot(lab, cfg.num_classes).double().permute(0, 3, 1, 2)[:, 1:]
        inter = (oh * prob).sum(dim=(1, 2, 3))
        union = (oh + prob).sum(dim=(1, 2, 3))
        loss = loss + 1.0 - (2.0 * inter / (union + 1e-8)).mean()
    loss.backward()
    return logits.detach().permute(0, 2, 3, 1).numpy(), float(loss), {k: v.grad.numpy() for k, v in P.items()}


@pytest.mark.parametrize("kw", [
    dict(use_context=True, use_spatial=True, guide_channel=2, norm_with_center=True, loss_type="xentropy+dice"),
    dict(use_context=True, use_spatial=False, norm_with_center=False, side_dropout=0.0, loss_type="xentropy"),
    dict(use_context=False, use_spatial=True, guide_channel=1, norm_with_scale=True, mod_layers=(0, 2), loss_type="dice"),
    # ext_config/GUNet_BOTH_AA.yml: channel-wise affine after the modulation of every encoder block
    dict(use_context=True, use_spatial=True, guide_channel=1, norm_with_center=True, norm_with_scale=True,
         after_affine=True, loss_type="xentropy+dice"),
    dict(use_context=True, use_spatial=False, after_affine=True, side_dropout=0.0, mod_layers=(0, 1), loss_type="xentropy"),
    # --normalizer batch_norm: batch statistics shared by all samples, per-sample modulation on top
    dict(use_context=True, use_spatial=True, guide_channel=1, norm_with_center=True, norm_with_scale=True,
         normalizer="batch_norm", loss_type="xentropy+dice"),
    dict(use_context=True, use_spatial=False, normalizer="batch_norm", side_dropout=0.0, loss_type="xentropy"),
])
def test_oracle_matches_torch_autograd(kw):
    cfg = G.GUNetCfg(height=16, width=16, init_channels=4, num_down_samples=2, mod_layers=kw.pop("mod_layers", (1, 2)),
                     context_fc_channels=(12, 10), context_dim=20, loss_weight_type="numerical",
                     loss_numeric_w=(0.2, 0.4, 4.4), weight_decay_rate=0.0, dropout_seed=5, **kw)
    rng = np.random.default_rng(3)
    n = 3
    inputs, labels = _inputs(cfg, n, rng)
    params = {k: v.astype(np.float64) + (0.1 * rng.standard_normal(v.shape) if k.endswith(("beta", "gamma", "biases")) else 0)
              for k, v in G.init_params(cfg, seed=1, dtype=np.float64).items()}
    tape = G.forward(params, inputs, cfg, True, step=2)
    loss, dl = G.loss_and_dlogits(tape, labels, cfg)
    grads = G.backward(tape, dl, cfg)
    mults = [f["mult"] if f["mult"] is not None else np.ones_like(f["pre"]) for f in tape.fc] if cfg.use_context else None
    if cfg.use_context and cfg.side_dropout:
        assert any((f["mult"] == 0).any() for f in tape.fc if f["mult"] is not None)
    t_logits, t_loss, t_grads = _torch_gunet(params, inputs, labels, cfg, mults)
    assert np.allclose(tape.logits, t_logits, rtol=1e-9, atol=1e-10)
    assert abs(loss - t_loss) < 1e-10
    assert set(grads) == {k for k in params if "/moving_" not in k}
    for k, g in grads.items():
        assert np.allclose(g, t_grads[k], rtol=1e-7, atol=1e-10), k
    if cfg.normalizer == "batch_norm":      # moving statistics: decay 0.99 on the modulated blocks, 0.999 elsewhere
        k0 = "GUNet/Encode/down_conv2/mod_conv1/Conv/BatchNorm/moving_mean"
        k1 = "GUNet/Decode/up_conv1/up_conv1_1/BatchNorm/moving_variance"
        y0 = next(L for L in tape.layers if L.get("spec", {}).get("scope", "").endswith("down_conv2/mod_conv1/Conv"))
        mean_y = O.conv2d(y0["x"], y0["w"]).mean(axis=(0, 1, 2))
        assert np.allclose(tape.new_moving[k0], 0.01 * mean_y, rtol=1e-9, atol=1e-12)       # decay 0.99, moving mean was 0
        assert tape.new_moving[k1].min() > 0.999 - 1e-9      # moving variance starts at 1: 0.999 + 0.001 * var
        assert set(tape.new_moving) == {k for k in params if "/moving_" in k}


def test_backbone_dropout_oracle_matches_torch_autograd():
    """--dropout (/root/reference/NetworksV2/GUNet.py:189-190): slim.dropout behind the normaliser of the first conv of
    every encoder block, in front of gamma_mod and the guide map; the same multipliers in the backward pass."""
    cfg = G.GUNetCfg(height=16, width=16, init_channels=4, num_down_samples=2, mod_layers=(1, 2), context_fc_channels=(8,),
                     context_dim=10, loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4), weight_decay_rate=0.0,
                     side_dropout=0.0, dropout=0.25, dropout_seed=5)
    rng = np.random.default_rng(10)
    n = 2
    inputs = dict(images=rng.uniform(0, 1, (n, 16, 16, 3)), context=rng.uniform(0, 1, (n, 10)),
                  sp_guide=rng.uniform(0.5, 1, (n, 16, 16, 1)))
    labels = rng.integers(0, 3, (n, 16, 16)).astype(np.int32)
    params = {k: v.astype(np.float64) + (0.1 * rng.standard_normal(v.shape) if k.endswith(("beta", "gamma", "biases")) else 0)
              for k, v in G.init_params(cfg, seed=3, dtype=np.float64).items()}
    tape = G.forward(params, inputs, cfg, True, step=2)
    drops = {L["spec"]["scope"]: L["mult"] for L in tape.layers if L.get("mult") is not None}
    assert len(drops) == 3 and all(sc.endswith("mod_conv1/Conv") for sc in drops)      # first conv of blocks 1, 2, 3
    for m in drops.values():
        assert set(np.unique(m)) == {0.0, np.float32(1 / 0.75)} and 0.6 < (m > 0).mean() < 0.9
    loss, dl = G.loss_and_dlogits(tape, labels, cfg)
    grads = G.backward(tape, dl, cfg)
    cfg._drop_mults = drops
    t_logits, t_loss, t_grads = _torch_gunet(params, inputs, labels, cfg, None)
    assert np.allclose(tape.logits, t_logits, rtol=1e-9, atol=1e-10)
    assert abs(loss - t_loss) < 1e-10
    for k, g in grads.items():
        assert np.allclose(g, t_grads[k], rtol=1e-7, atol=1e-10), k
    # inference: no dropout
    assert all(L.get("mult") is None for L in G.forward(params, inputs, cfg, False).layers)


def test_img_grad_oracle_matches_torch_autograd():
    """--img_grad (GUNet.py:333-337, scripts/103_grad.sh): concat(images, dy, dx) of tf.image.image_gradients feeds the
    first conv (9 input channels)."""
    cfg = G.GUNetCfg(height=16, width=16, init_channels=4, num_down_samples=2, mod_layers=(1, 2), context_fc_channels=(8,),
                     context_dim=10, loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4), weight_decay_rate=0.0,
                     side_dropout=0.0, img_grad=True)
    rng = np.random.default_rng(11)
    n = 2
    inputs = dict(images=rng.uniform(0, 1, (n, 16, 16, 3)), context=rng.uniform(0, 1, (n, 10)),
                  sp_guide=rng.uniform(0.5, 1, (n, 16, 16, 1)))
    labels = rng.integers(0, 3, (n, 16, 16)).astype(np.int32)
    dy, dx = O.image_gradients(inputs["images"])
    assert not dy[:, -1].any() and not dx[:, :, -1].any()
    assert np.allclose(dy[:, 3, 5], inputs["images"][:, 4, 5] - inputs["images"][:, 3, 5])
    params = {k: v.astype(np.float64) + (0.1 * rng.standard_normal(v.shape) if k.endswith(("beta", "gamma", "biases")) else 0)
              for k, v in G.init_params(cfg, seed=3, dtype=np.float64).items()}
    assert params["GUNet/Encode/down_conv1/mod_conv1/Conv/weights"].shape == (3, 3, 9, 4)
    tape = G.forward(params, inputs, cfg, True)
    loss, dl = G.loss_and_dlogits(tape, labels, cfg)
    grads = G.backward(tape, dl, cfg)
    t_logits, t_loss, t_grads = _torch_gunet(params, inputs, labels, cfg, None)
    assert np.allclose(tape.logits, t_logits, rtol=1e-9, atol=1e-10)
    assert abs(loss - t_loss) < 1e-10
    for k, g in grads.items():
        assert np.allclose(g, t_grads[k], rtol=1e-7, atol=1e-10), k


def test_unetinter_oracle_matches_torch_autograd():
    """UNetInter (/root/reference/NetworksV2/UNetInter.py:73-146): guide concatenated to the images, every conv
    normalised with centre + scale, variables under "UNetInter/"."""
    cfg = G.unetinter_cfg(channel=3, guide_channel=2, height=16, width=16, init_channels=4, num_down_samples=2,
                          loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4), weight_decay_rate=0.0,
                          loss_type="xentropy+dice")
    rng = np.random.default_rng(8)
    n = 2
    images = rng.uniform(0, 1, (n, 16, 16, 3))
    guide = rng.uniform(0, 1, (n, 16, 16, 2))
    labels = rng.integers(0, 3, (n, 16, 16)).astype(np.int32)
    inputs = G.unetinter_inputs(images, guide)
    assert inputs["images"].shape == (n, 16, 16, 5) and np.array_equal(inputs["images"][..., 3:], guide)
    params = {k: v.astype(np.float64) + (0.1 * rng.standard_normal(v.shape) if k.endswith(("beta", "gamma", "biases")) else 0)
              for k, v in G.init_params(cfg, seed=1, dtype=np.float64).items()}
    assert params["UNetInter/Encode/down_conv1/mod_conv1/Conv/weights"].shape == (3, 3, 5, 4)
    assert "UNetInter/Encode/down_conv3/mod_conv2/Conv/InstanceNorm/gamma" in params
    assert "UNetInter/Decode/up2/biases" in params and "UNetInter/AdjustChannels/biases" in params
    assert not any("context" in k or "spatial" in k for k in params)
    tape = G.forward(params, inputs, cfg, True)
    loss, dl = G.loss_and_dlogits(tape, labels, cfg)
    grads = G.backward(tape, dl, cfg)
    t_logits, t_loss, t_grads = _torch_gunet(params, inputs, labels, cfg, None)
    assert np.allclose(tape.logits, t_logits, rtol=1e-9, atol=1e-10)
    assert abs(loss - t_loss) < 1e-10
    assert set(grads) == set(params)
    for k, g in grads.items():
        assert np.allclose(g, t_grads[k], rtol=1e-7, atol=1e-10), k


def test_unetinter_mid_cat_oracle_matches_torch_autograd():
    """--mid_cat (/root/reference/NetworksV2/UNetInter.py:87-92,124-125): the images alone enter the first block; the guide
    is concatenated to its output in front of the first max-pool only (the skip connection keeps the block's channels),
    so the second block's first conv has init_channels + guide_channel inputs."""
    cfg = G.unetinter_cfg(channel=3, guide_channel=2, mid_cat=True, height=16, width=16, init_channels=4,
                          num_down_samples=2, loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4),
                          weight_decay_rate=0.0, loss_type="xentropy")
    rng = np.random.default_rng(9)
    n = 2
    images, guide = rng.uniform(0, 1, (n, 16, 16, 3)), rng.uniform(0, 1, (n, 16, 16, 2))
    labels = rng.integers(0, 3, (n, 16, 16)).astype(np.int32)
    inputs = G.unetinter_inputs(images, guide, mid_cat=True)
    assert inputs["images"].shape == (n, 16, 16, 3)
    params = {k: v.astype(np.float64) + (0.1 * rng.standard_normal(v.shape) if k.endswith(("beta", "gamma", "biases")) else 0)
              for k, v in G.init_params(cfg, seed=2, dtype=np.float64).items()}
    assert params["UNetInter/Encode/down_conv1/mod_conv1/Conv/weights"].shape == (3, 3, 3, 4)
    assert params["UNetInter/Encode/down_conv2/mod_conv1/Conv/weights"].shape == (3, 3, 4 + 2, 8)
    assert params["UNetInter/Decode/up_conv1/up_conv1_1/weights"].shape == (3, 3, 8, 4)     # the skip has no guide lanes
    tape = G.forward(params, inputs, cfg, True)
    loss, dl = G.loss_and_dlogits(tape, labels, cfg)
    grads = G.backward(tape, dl, cfg)
    t_logits, t_loss, t_grads = _torch_gunet(params, inputs, labels, cfg, None)
    assert np.allclose(tape.logits, t_logits, rtol=1e-9, atol=1e-10)
    assert abs(loss - t_loss) < 1e-10
    assert set(grads) == set(params)
    for k, g in grads.items():
        assert np.allclose(g, t_grads[k], rtol=1e-7, atol=1e-10), k


def test_parameter_inventory_matches_reference_naming():
    cfg = G.GUNetCfg(height=32, width=32)
    p = G.init_params(cfg)
    assert cfg.n_modulator_param == 3840                                  # GUNet.py:44-45 with GUNet.yml
    assert p["GUNet/context/fc3/weights"].shape == (256, 3840)
    assert p["GUNet/spatial/conv5/weights"].shape == (1, 1, 1, 2048)
    assert "GUNet/Encode/down_conv1/mod_conv1/Conv/InstanceNorm/gamma" in p       # block 0 is not modulated
    assert "GUNet/Encode/down_conv2/mod_conv1/Conv/InstanceNorm/gamma" not in p   # norm_with_scale: false
    assert "GUNet/Encode/down_conv2/mod_conv1/Conv/InstanceNorm/beta" in p        # norm_with_center: true
    assert "GUNet/Decode/up_conv1/up_conv1_2/InstanceNorm/beta" in p
    reg = G.regularized_names(cfg, p)
    assert "GUNet/spatial/conv2/weights" in reg and "GUNet/context/fc1/weights" not in reg
    # after_affine (ext_config/GUNet_BOTH_AA.yml): every encoder block gets ChannelWiseAffine/{gamma,beta} next to its
    # Conv scope; the modulated blocks' normaliser loses centre and scale (GUNet.py:318-319), block 0 keeps both
    pa = G.init_params(G.GUNetCfg(height=32, width=32, after_affine=True, norm_with_scale=True))
    assert "GUNet/Encode/down_conv3/mod_conv2/ChannelWiseAffine/gamma" in pa
    assert "GUNet/Encode/down_conv1/mod_conv1/ChannelWiseAffine/beta" in pa
    assert "GUNet/Encode/down_conv3/mod_conv2/Conv/InstanceNorm/beta" not in pa
    assert "GUNet/Encode/down_conv3/mod_conv2/Conv/InstanceNorm/gamma" not in pa
    assert "GUNet/Encode/down_conv1/mod_conv1/Conv/InstanceNorm/gamma" in pa
    assert not any("ChannelWiseAffine" in k for k in G.regularized_names(G.GUNetCfg(after_affine=True), pa))
    assert not any("Decode" in k and "ChannelWiseAffine" in k for k in pa)
