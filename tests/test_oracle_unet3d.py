"""Pins oracle/unet3d_ref.py (CPU only) against an independent torch-CPU autograd implementation of
/root/reference/NetworksV2/UNet3D.py:123-202 in fp64, and checks the layer table against the reference's
_ModelConfig (UNet3D.py:31-91) restated literally."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tf_ops as O
from oracle import unet3d_ref as U


def test_layer_table_matches_reference_config():
    t4 = U.model_config(4)
    assert [(b, l) for b, l, _, _ in t4][:4] == [("conv_e0", "conv1"), ("conv_e0", "conv2"), ("conv_e1", "conv1"),
                                                  ("conv_e1", "conv2")]
    d = {(b, l): (k, s) for b, l, k, s in t4}
    assert d[("conv_e1", "conv1")] == ((1, 3, 3), (1, 2, 2))
    assert d[("conv_e2", "conv1")] == ((3, 3, 3), (1, 2, 2))
    assert d[("bridge", "conv1")] == ((3, 3, 3), (2, 2, 2))
    assert d[("conv_d3", "up")] == ((2, 2, 2), (2, 2, 2)) and d[("conv_d2", "up")] == ((1, 2, 2), (1, 2, 2))
    assert d[("conv_d1", "conv1")][0] == (1, 3, 3) and d[("conv_d2", "conv1")][0] == (3, 3, 3)
    d5 = {(b, l): (k, s) for b, l, k, s in U.model_config(5)}
    assert d5[("conv_d4", "up")][0] == (2, 2, 2) and d5[("conv_d3", "up")][0] == (1, 2, 2)
    assert d5[("conv_e4", "conv2")][0] == (3, 3, 3)
    cfg = U.UNet3DCfg()
    specs = {s["scope"]: s for s in U.layer_specs(cfg)}
    assert [specs[f"UNet3D/conv_e{i}/conv2"]["cout"] for i in range(4)] == [30, 60, 120, 240]
    assert specs["UNet3D/bridge/conv2"]["cout"] == 320 and specs["UNet3D/bridge/conv2"]["dhw"] == (32, 8, 8)
    assert specs["UNet3D/conv_d3/conv1"]["cin"] == 480 and specs["UNet3D/conv_d0/conv1"]["cin"] == 60
    assert specs["UNet3D/logits"]["cin"] == 30 and specs["UNet3D/logits"]["dhw"] == (64, 128, 128)
    # BASELINE.md / SURVEY.md section 8d: 637.387 GFLOP forward per 64x128x128 patch
    fl = 0
    for s in U.layer_specs(cfg):
        o = s["dhw"] if s["kind"] != "conv" else tuple(-(-s["dhw"][i] // s["s"][i]) for i in range(3))
        taps = 1 if s["kind"] == "convT" else int(np.prod(s["k"]))
        vox = np.prod(o) if s["kind"] != "convT" else np.prod([s["dhw"][i] * s["s"][i] for i in range(3)])
        fl += 2.0 * vox * taps * s["cin"] * s["cout"]
    assert abs(fl / 1e9 - 637.387) < 0.01, fl / 1e9


def _torch_unet3d(params, inputs, labels, cfg):
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    x = t(inputs["images"])
    if cfg.use_spatial:
        x = torch.cat((x, t(inputs["sp_guide"])), dim=-1)
    x = x.permute(0, 4, 1, 2, 3)
    skips = {}
    for s in U.layer_specs(cfg):
        sc = s["scope"]
        if s["kind"] == "conv":
            pads = [O.same_pads(x.shape[2 + i], s["k"][i], s["s"][i]) for i in range(3)]
            xp = F.pad(x, (pads[2][1], pads[2][2], pads[1][1], pads[1][2], pads[0][1], pads[0][2]))
            y = F.conv3d(xp, P[f"{sc}/weights"].permute(4, 3, 0, 1, 2), stride=s["s"])
            y = F.instance_norm(y, weight=P[f"{sc}/InstanceNorm/gamma"], bias=P[f"{sc}/InstanceNorm/beta"], eps=cfg.in_eps)
            x = torch.relu(y)
            if s["layer"] == "conv2" and s["block"].startswith("conv_e"):
                skips[s["block"]] = x
        elif s["kind"] == "convT":
            up = torch.relu(F.conv_transpose3d(x, P[f"{sc}/weights"].permute(4, 3, 0, 1, 2), stride=s["s"]))
            x = torch.cat((skips[s["block"].replace("d", "e")], up), dim=1)
        else:
            logits = F.conv3d(x, P[f"{sc}/weights"].permute(4, 3, 0, 1, 2), P[f"{sc}/biases"])
    lab = torch.tensor(labels, dtype=torch.long)
    oh = F.one_hot(lab, cfg.num_classes).double()
    w = (oh * t(np.array(cfg.loss_numeric_w))).sum(-1)
    w = w / w.sum(dim=(1, 2, 3), keepdim=True) * float(np.prod(labels.shape[1:]))
    ce = F.cross_entropy(logits, lab, reduction="none")
    loss = (w * ce).sum() / (w != 0).sum()
    loss.backward()
    return logits.detach().permute(0, 2, 3, 4, 1).numpy(), float(loss.detach()), {k: v.grad.numpy() for k, v in P.items()}


@pytest.mark.parametrize("pools,use_spatial,depth", [(4, False, 4), (4, True, 2), (5, False, 2)])
def test_oracle_matches_torch_autograd(pools, use_spatial, depth):
    hw = 16 * 2 ** (pools - 4) * 2
    cfg = U.UNet3DCfg(depth=depth, height=hw, width=hw, channel=1, init_channels=3, max_channels=20,
                      num_pool_layers=pools, use_spatial=use_spatial, loss_numeric_w=(1.0, 10.0), weight_decay_rate=0.0)
    rng = np.random.default_rng(pools + depth)
    n = 2
    inputs = dict(images=rng.standard_normal((n, depth, hw, hw, 1)))
    if use_spatial:
        inputs["sp_guide"] = rng.uniform(0, 1, (n, depth, hw, hw, 2))
    labels = rng.integers(0, 2, (n, depth, hw, hw)).astype(np.int32)
    params = {k: v.astype(np.float64) + (0.1 * rng.standard_normal(v.shape) if k.endswith(("beta", "gamma", "biases")) else 0)
              for k, v in U.init_params(cfg, seed=2, dtype=np.float64).items()}
    tape = U.forward(params, inputs, cfg)
    loss, dl = U.loss_and_dlogits(tape, labels, cfg)
    grads = U.backward(tape, dl, cfg)
    t_logits, t_loss, t_grads = _torch_unet3d(params, inputs, labels, cfg)
    assert np.allclose(tape.logits, t_logits, rtol=1e-8, atol=1e-9)
    assert abs(loss - t_loss) < 1e-10
    assert set(grads) == set(params)
    for k, g in grads.items():
        assert np.allclose(g, t_grads[k], rtol=1e-6, atol=1e-9), k
