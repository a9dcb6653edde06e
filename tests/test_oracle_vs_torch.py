"""The numpy oracle (oracle/unet_ref.py, hand-derived backward) against the independent torch-CPU
restatement (oracle/unet_torch.py, autograd). The reference ships no tests or golden vectors
(SURVEY.md section 4), so two independent restatements agreeing is what pins the oracle."""
import numpy as np
import pytest
import torch

from boxsegliver_b200 import synthetic
from oracle import tf_ops as O
from oracle import unet_ref as R
from oracle import unet_torch as T


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


CASES = [
    dict(normalizer="batch_norm", loss_type="xentropy", loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4)),
    dict(normalizer="instance_norm", loss_type="xentropy", loss_weight_type="proportion"),
    dict(normalizer="batch_norm", loss_type="dice", loss_weight_type="none"),
    dict(normalizer="batch_norm", loss_type="xentropy", loss_weight_type="none", bias_decay=True),
]


@pytest.mark.parametrize("kw", CASES)
def test_forward_loss_and_gradients_match_torch(kw):
    cfg = R.UNetCfg(height=32, width=32, channel=3, init_channels=8, num_down_samples=3, weight_decay_rate=1e-3, **kw)
    params = R.init_params(cfg, seed=3, dtype=np.float64)
    images, labels = synthetic.make_batch(3, 32, 32, 3, seed=11)
    images = images.astype(np.float64)
    tape = R.forward(params, images, cfg, True)
    loss, dl = R.loss_and_dlogits(tape, labels, cfg)
    grads = R.total_grads(params, R.backward(tape, dl, cfg), cfg)
    total = float(loss) + R.regularization_loss(params, cfg)

    tp = T.to_torch_params(params, torch.float64)
    tl, tlogits, tmov = T.total_loss(tp, torch.tensor(images).permute(0, 3, 1, 2), torch.tensor(labels, dtype=torch.long),
                                     cfg)
    tl.backward()
    assert rel(tape.logits, tlogits.permute(0, 2, 3, 1).detach().numpy()) < 1e-10
    assert abs(total - float(tl)) < 1e-10 * max(1.0, abs(total))
    for k, g in grads.items():
        assert rel(g, tp[k].grad.numpy()) < 1e-8, k
    for k, v in tape.new_moving.items():
        assert rel(v, tmov[k].numpy()) < 1e-12, k


def test_train_step_matches_torch_adam():
    cfg = R.UNetCfg(height=16, width=16, channel=3, init_channels=8, num_down_samples=2, weight_decay_rate=1e-4,
                    loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4))
    params = R.init_params(cfg, seed=5, dtype=np.float64)
    images, labels = synthetic.make_batch(2, 16, 16, 3, seed=2)
    images = images.astype(np.float64)
    tp = T.to_torch_params(params, torch.float64)
    slots, tslots = {}, {}
    for step in (1, 2, 3):
        l0, _, _ = R.train_step(params, slots, step, images, labels, cfg, 1e-3)
        l1, _ = T.train_step(tp, tslots, step, torch.tensor(images).permute(0, 3, 1, 2),
                             torch.tensor(labels, dtype=torch.long), cfg, 1e-3)
        assert abs(l0 - l1) < 1e-9 * max(1.0, abs(l0))
    for k in params:
        assert rel(params[k], tp[k].detach().numpy()) < 1e-8, k


def test_eval_mode_uses_moving_statistics():
    cfg = R.UNetCfg(height=16, width=16, channel=3, init_channels=8, num_down_samples=2)
    params = R.init_params(cfg, seed=1, dtype=np.float64)
    rng = np.random.default_rng(0)
    for k in params:
        if k.endswith("moving_mean"):
            params[k] = rng.normal(0, 0.1, params[k].shape)
        if k.endswith("moving_variance"):
            params[k] = rng.uniform(0.5, 1.5, params[k].shape)
    images, _ = synthetic.make_batch(2, 16, 16, 3, seed=4)
    tape = R.forward(params, images.astype(np.float64), cfg, False)
    tp = T.to_torch_params(params, torch.float64)
    tlogits, _ = T.forward(tp, torch.tensor(images.astype(np.float64)).permute(0, 3, 1, 2), cfg, False)
    assert rel(tape.logits, tlogits.permute(0, 2, 3, 1).detach().numpy()) < 1e-10


def test_op_level_against_torch_functional():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 12, 6))
    w = rng.standard_normal((3, 3, 6, 4))
    # stride-2 SAME: TF pads (0 before, 1 after) for even sizes, symmetric for odd ones
    for (h, wd) in [(9, 11), (8, 12)]:
        xs = x[:, :h, :wd]
        y = O.conv2d(xs, w, stride=2)
        ho, pt, pb = O.same_pads(h, 3, 2)
        wo, pl, pr = O.same_pads(wd, 3, 2)
        tx = torch.nn.functional.pad(torch.tensor(xs).permute(0, 3, 1, 2), (pl, pr, pt, pb))
        ty = torch.nn.functional.conv2d(tx, torch.tensor(w).permute(3, 2, 0, 1), stride=2)
        assert rel(y, ty.permute(0, 2, 3, 1).numpy()) < 1e-12
    # max-pool gradient: first maximum wins ties (all-equal window)
    a = np.zeros((1, 2, 2, 1))
    g = O.max_pool_2x2_grad(a, np.ones((1, 1, 1, 1)))
    assert g[0, 0, 0, 0] == 1 and g.sum() == 1


def test_metrics_and_counts():
    labels = np.array([[[0, 1], [2, 2]], [[0, 0], [0, 0]]], np.int32)
    pred_liver = np.array([[[0, 1], [1, 0]], [[0, 0], [0, 0]]], np.uint8)[..., None]
    i, l, r = O.seg_counts(pred_liver, labels, 1)
    assert i.tolist() == [1, 0] and l.tolist() == [2, 0] and r.tolist() == [1, 0]
    d = O.metric_dice(pred_liver, labels, 1)
    exp = np.mean([(2 * 1 + 1e-5) / (2 + 1 + 1e-5), (0 + 1e-5) / (0 + 0 + 1e-5)])
    assert abs(float(d) - exp) < 1e-6
