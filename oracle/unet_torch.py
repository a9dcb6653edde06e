"""CPU ORACLE (second, independent restatement) -- TEST INFRASTRUCTURE ONLY.

torch-CPU (oneDNN/MKL) functional restatement of the reference's 2-D U-Net training step, written
against the same reference lines as oracle/unet_ref.py (NetworksV2/UNet.py:41-135,
loss_metrics.py:115-177, core/solver.py:204-243) but sharing no code with it: autograd supplies the
backward pass that unet_ref.py derives by hand. Two uses, both outside the product path:
  * tests/test_oracle_vs_torch.py cross-checks the numpy oracle against it (parity is unpinned by
    the reference itself, so two independent restatements + finite differences are the pin);
  * bench.py's `--impl reference` / `cpu_baseline` leg times it on the host cores, as the stand-in
    for the reference's TF-1.13 CPU path, which cannot be installed in this image.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import unet_ref as R


def to_torch_params(params: dict, dtype=torch.float32):
    """numpy {tf name: array} -> leaf tensors (requires_grad for trainables)."""
    out = {}
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if not k.endswith(("moving_mean", "moving_variance")):
            t.requires_grad_(True)
        out[k] = t
    return out


def _conv_w(w):      # HWIO -> OIHW
    return w.permute(3, 2, 0, 1)


def _convT_w(w):     # [kh, kw, Cout, Cin] -> torch conv_transpose2d layout [Cin, Cout, kh, kw]
    return w.permute(3, 2, 0, 1)


def forward(p: dict, images: torch.Tensor, cfg: R.UNetCfg, is_training: bool):
    """images NCHW. Returns (logits NCHW, {moving-stat name: new value})."""
    ns = R.norm_scope(cfg)
    new_moving = {}

    def block(x, scope):
        y = F.conv2d(x, _conv_w(p[f"{scope}/weights"]), padding=1)
        g, b = p[f"{scope}/{ns}/gamma"], p[f"{scope}/{ns}/beta"]
        if cfg.normalizer == "batch_norm":
            mm, mv = p[f"{scope}/{ns}/moving_mean"], p[f"{scope}/{ns}/moving_variance"]
            if is_training:
                mean = y.mean(dim=(0, 2, 3))
                var = y.var(dim=(0, 2, 3), unbiased=False)
                m = y.numel() // y.shape[1]
                new_moving[f"{scope}/{ns}/moving_mean"] = (mm * cfg.bn_decay + mean * (1 - cfg.bn_decay)).detach()
                new_moving[f"{scope}/{ns}/moving_variance"] = (
                    mv * cfg.bn_decay + var * (m / max(m - 1, 1)) * (1 - cfg.bn_decay)).detach()
            else:
                mean, var = mm, mv
            z = (y - mean[None, :, None, None]) * torch.rsqrt(var + cfg.bn_eps)[None, :, None, None]
        else:
            mean = y.mean(dim=(2, 3), keepdim=True)
            var = y.var(dim=(2, 3), unbiased=False, keepdim=True)
            z = (y - mean) * torch.rsqrt(var + cfg.in_eps)
        return F.relu(z * g[None, :, None, None] + b[None, :, None, None])

    x = images
    skips = []
    for i in range(cfg.num_down_samples):
        for j in (1, 2):
            x = block(x, f"UNet/Encode{i + 1}/Repeat/convolution2d_{j}")
        skips.append(x)
        x = F.max_pool2d(x, 2)
    for j in (1, 2):
        x = block(x, f"UNet/ED-Bridge/ED-Bridge_{j}")
    for i in reversed(range(cfg.num_down_samples)):
        scope = f"UNet/Decode{i + 1}/Conv2d_transpose"
        up = F.relu(F.conv_transpose2d(x, _convT_w(p[f"{scope}/weights"]), p[f"{scope}/biases"], stride=2))
        x = torch.cat((skips[i], up), dim=1)
        for j in (1, 2):
            x = block(x, f"UNet/Decode{i + 1}/Repeat/convolution2d_{j}")
    scope = "UNet/AdjustChannels"
    return F.conv2d(x, _conv_w(p[f"{scope}/weights"]), p[f"{scope}/biases"]), new_moving


def data_loss(logits: torch.Tensor, labels: torch.Tensor, cfg: R.UNetCfg):
    """Weighted sparse softmax cross-entropy / dice of loss_metrics.py on NCHW logits, labels [N,H,W] int64."""
    n, c, h, w = logits.shape
    oh = F.one_hot(labels, c).to(logits.dtype)                       # [N,H,W,C]
    if cfg.loss_type == "dice":
        prob = torch.softmax(logits, dim=1).permute(0, 2, 3, 1)[..., 1:]
        ohf = oh[..., 1:]
        inter = (ohf * prob).sum(dim=(1, 2, 3))
        union = (ohf + prob).sum(dim=(1, 2, 3))
        return 1.0 - (2.0 * inter / (union + 1e-8)).mean()
    if cfg.loss_weight_type == "none":
        wmap = torch.ones((n, h, w), dtype=logits.dtype)
    else:
        if cfg.loss_weight_type == "numerical":
            wc = torch.tensor(cfg.loss_numeric_w, dtype=logits.dtype).expand(n, c)
        else:
            num = oh.sum(dim=(1, 2))
            if cfg.loss_proportion_decay > 0:
                num = num + cfg.loss_proportion_decay
            prop = 1.0 / num
            wc = prop / prop.sum(dim=1, keepdim=True)
        wmap = (wc[:, None, None, :] * oh).sum(dim=-1)
        wmap = wmap / wmap.sum(dim=(1, 2), keepdim=True) * (h * w)
    ce = F.cross_entropy(logits, labels, reduction="none")
    nz = (wmap != 0).sum().clamp(min=1)
    return (wmap * ce).sum() / nz


def total_loss(p: dict, images, labels, cfg: R.UNetCfg, is_training: bool = True):
    logits, new_moving = forward(p, images, cfg, is_training)
    loss = data_loss(logits, labels, cfg)
    if cfg.weight_decay_rate > 0:
        names = R.regularized_names(cfg, p)
        loss = loss + cfg.weight_decay_rate * 0.5 * sum((p[k] ** 2).sum() for k in names)
    return loss, logits, new_moving


def train_step(p: dict, slots: dict, step: int, images, labels, cfg: R.UNetCfg, lr: float):
    """One Adam(beta1=.9, beta2=.99, eps=1e-8) training step, TF formulation (solver.py:206). Mutates p / slots."""
    for t in p.values():
        if t.requires_grad and t.grad is not None:
            t.grad = None
    loss, logits, new_moving = total_loss(p, images, labels, cfg, True)
    loss.backward()
    lr_t = lr * np.sqrt(1.0 - 0.99 ** step) / (1.0 - 0.9 ** step)
    with torch.no_grad():
        for k, t in p.items():
            if not t.requires_grad:
                continue
            m, v = slots.setdefault(k, (torch.zeros_like(t), torch.zeros_like(t)))
            m.mul_(0.9).add_(t.grad, alpha=0.1)
            v.mul_(0.99).addcmul_(t.grad, t.grad, value=0.01)
            t.addcdiv_(m, v.sqrt().add_(1e-8), value=-lr_t)
        for k, v in new_moving.items():
            p[k].copy_(v)
    return float(loss), logits
