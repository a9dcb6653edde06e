"""fp64 central finite differences against the numpy oracle's hand-derived backward (third pin)."""
import numpy as np
import pytest

from oracle import unet_ref as R


@pytest.mark.parametrize("normalizer,loss_type,wtype", [
    ("batch_norm", "xentropy", "numerical"), ("instance_norm", "xentropy", "proportion"), ("batch_norm", "dice", "none")])
def test_finite_differences(normalizer, loss_type, wtype):
    cfg = R.UNetCfg(height=16, width=16, channel=3, init_channels=4, num_down_samples=2, normalizer=normalizer,
                    loss_type=loss_type, loss_weight_type=wtype,
                    loss_numeric_w=(0.2, 0.4, 4.4) if wtype == "numerical" else (), weight_decay_rate=1e-3)
    p = R.init_params(cfg, 0, np.float64)
    rng = np.random.default_rng(1)
    for k in p:  # zero biases put ReLU exactly on its kink wherever the input patch is all-zero
        if k.endswith(("/biases", "/beta")):
            p[k] = rng.normal(0, 0.1, p[k].shape)
    x = rng.uniform(0, 1, (2, 16, 16, 3))
    lab = rng.integers(0, 3, (2, 16, 16)).astype(np.int32)

    def total(q):
        t = R.forward(q, x, cfg, True)
        l, _ = R.loss_and_dlogits(t, lab, cfg)
        return float(l) + R.regularization_loss(q, cfg)

    t = R.forward(p, x, cfg, True)
    _, dl = R.loss_and_dlogits(t, lab, cfg)
    g = R.total_grads(p, R.backward(t, dl, cfg), cfg)
    for k in R.trainable_names(cfg, p):
        idx = tuple(rng.integers(0, s) for s in p[k].shape)
        e = 1e-6
        q = {a: b.copy() for a, b in p.items()}
        q[k][idx] += e
        lp = total(q)
        q[k][idx] -= 2 * e
        lm = total(q)
        fd, an = (lp - lm) / (2 * e), g[k][idx]
        assert abs(fd - an) <= 1e-4 * (abs(fd) + abs(an)) + 1e-9, (k, idx, fd, an)
