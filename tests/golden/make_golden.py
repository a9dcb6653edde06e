"""Generates tests/golden/{unet2d,gunet,unet3d}_step.npz -- golden vectors for one training step of each model.

The reference (Jarvis73/BoxSegLiver) cannot be imported here: it is TensorFlow 1.13 code and TF cannot be
installed in this image, and it ships no tests or golden vectors of its own (section 4).
These vectors are therefore minted by this repo's CPU oracle (oracle/unet_ref.py, a restatement of
NetworksV2/UNet.py:58-155 + loss_metrics.py:115-339 + core/solver.py:204-243 with TF-1.13/slim semantics) in
fp64, after that oracle has been cross-checked against an independent torch-CPU implementation
(tests/test_oracle_vs_torch.py) and central finite differences (tests/test_oracle_gradcheck.py).
PARITY UNPINNED by the reference itself; pinned by two independent implementations.

    python tests/golden/make_golden.py          # rewrites the .npz next to this file

The fixture stores the INPUTS (images, labels) verbatim and the weights by seed + per-tensor checksums (the
31 M parameters do not belong in git); outputs are logits, loss, masks, argmax, integer Dice counts and, per
trainable tensor, the gradient's L2 norm, sum and leading entries.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from boxsegliver_b200 import synthetic  # noqa: E402
from oracle import unet_ref as R  # noqa: E402

CFG = dict(height=32, width=32, channel=3, init_channels=64, num_down_samples=4, normalizer="batch_norm",
           weight_decay_rate=1e-5, loss_type="xentropy", loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4))
N, WEIGHT_SEED, DATA_SEED, LR = 3, 11, 1357, 1e-3


def build(dtype=np.float64):
    cfg = R.UNetCfg(**CFG)
    params = R.init_params(cfg, seed=WEIGHT_SEED)
    images, labels = synthetic.make_batch(N, CFG["height"], CFG["width"], 3, seed=DATA_SEED)
    p = {k: v.astype(dtype) for k, v in params.items()}
    tape = R.forward(p, images.astype(dtype), cfg, True)
    loss, dl = R.loss_and_dlogits(tape, labels, cfg)
    grads = R.backward(tape, dl, cfg)
    preds = R.predictions(tape, cfg)
    mets = R.metrics(tape, labels, cfg, ("Dice", "VOE", "VD"))
    out = {
        "images": images, "labels": labels,
        "logits": tape.logits.astype(np.float32), "loss": np.float64(loss),
        "reg_loss": np.float64(R.regularization_loss(params, cfg)),
        "argmax": np.argmax(tape.prob, axis=-1).astype(np.uint8),
    }
    for k, v in preds.items():
        out["pred/" + k] = np.asarray(v)
    for k, v in mets.items():
        out["metric/" + k] = np.float32(v)
    for c in range(1, cfg.num_classes):
        i_, l_, r_ = R.O.seg_counts(np.asarray(preds[cfg.classes[c] + "Pred"]), labels, c)
        out[f"counts/{c}"] = np.stack([i_, l_, r_], axis=1).astype(np.uint32)
    names = sorted(grads)
    out["grad_names"] = np.array(names)
    out["grad_norm"] = np.array([np.linalg.norm(grads[k].astype(np.float64)) for k in names])
    out["grad_sum"] = np.array([grads[k].astype(np.float64).sum() for k in names])
    out["grad_head"] = np.stack([np.resize(grads[k].astype(np.float64).ravel()[:8], 8) for k in names])
    wn = sorted(params)
    out["weight_names"] = np.array(wn)
    out["weight_sumsq"] = np.array([float((params[k].astype(np.float64) ** 2).sum()) for k in wn])
    for k, v in tape.new_moving.items():
        out["moving/" + k] = np.asarray(v, np.float32)
    return out


# ---------------------------------------------------------------------------------------------------- GUNet / UNet3D
GCFG = dict(height=32, width=32, channel=3, init_channels=64, num_down_samples=4, mod_layers=(1, 2, 3, 4),
            context_fc_channels=(256, 256), norm_with_center=True, norm_with_scale=False, use_context=True,
            use_spatial=True, guide_channel=1, context_dim=200, side_dropout=0.5, dropout_seed=3,
            weight_decay_rate=1e-5, loss_type="xentropy+dice", loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4))
GN = 2
VCFG = dict(depth=2, height=32, width=32, channel=1, init_channels=30, max_channels=320, num_pool_layers=4,
            weight_decay_rate=3e-5, loss_weight_type="numerical", loss_numeric_w=(1.0, 10.0))
VN = 2


def _grad_summary(out, grads, params):
    names = sorted(grads)
    out["grad_names"] = np.array(names)
    out["grad_norm"] = np.array([np.linalg.norm(grads[k].astype(np.float64)) for k in names])
    out["grad_sum"] = np.array([grads[k].astype(np.float64).sum() for k in names])
    wn = sorted(params)
    out["weight_names"] = np.array(wn)
    out["weight_sumsq"] = np.array([float((params[k].astype(np.float64) ** 2).sum()) for k in wn])


def gunet_inputs():
    images, labels = synthetic.make_batch(GN, GCFG["height"], GCFG["width"], 3, seed=DATA_SEED + 1)
    context, guide = synthetic.make_guides(images, labels, GCFG["context_dim"], GCFG["guide_channel"], seed=2)
    return dict(images=images, context=context, sp_guide=guide), labels


def build_gunet(dtype=np.float64):
    """One GUNet training step (GUNet.yml, --use_context --use_spatial, instance_norm, dropout 0.5): oracle/gunet_ref.py,
    pinned by the torch-autograd cross-check in tests/test_oracle_gunet.py."""
    from oracle import gunet_ref as GU
    cfg = GU.GUNetCfg(**GCFG)
    params = GU.init_params(cfg, seed=WEIGHT_SEED)
    inputs, labels = gunet_inputs()
    tape = GU.forward({k: v.astype(dtype) for k, v in params.items()}, {k: v.astype(dtype) for k, v in inputs.items()},
                      cfg, True, step=1)
    loss, dl = GU.loss_and_dlogits(tape, labels, cfg)
    grads = GU.backward(tape, dl, cfg)
    out = {"labels": labels, "logits": tape.logits.astype(np.float32), "loss": np.float64(loss),
           "reg_loss": np.float64(GU.regularization_loss(params, cfg)), "ctx_params": tape.ctx_params.astype(np.float32),
           "dropout_kept": np.array([int((f["mult"] > 0).sum()) for f in tape.fc if f["mult"] is not None])}
    out.update({"in/" + k: v for k, v in inputs.items()})
    _grad_summary(out, grads, params)
    return out


def build_unet3d(dtype=np.float64):
    """One UNet3D training step (UNet3D.yml channels 30..320, instance_norm): oracle/unet3d_ref.py, pinned by the
    torch-autograd cross-check in tests/test_oracle_unet3d.py."""
    from oracle import unet3d_ref as U3
    cfg = U3.UNet3DCfg(**VCFG)
    params = U3.init_params(cfg, seed=WEIGHT_SEED)
    images, labels = synthetic.make_volume_batch(VN, VCFG["depth"], VCFG["height"], VCFG["width"], seed=DATA_SEED + 2)
    tape = U3.forward({k: v.astype(dtype) for k, v in params.items()}, dict(images=images.astype(dtype)), cfg)
    loss, dl = U3.loss_and_dlogits(tape, labels, cfg)
    grads = U3.backward(tape, dl, cfg)
    out = {"images": images, "labels": labels, "logits": tape.logits.astype(np.float32), "loss": np.float64(loss),
           "reg_loss": np.float64(U3.regularization_loss(params, cfg)),
           "argmax": np.argmax(tape.prob, axis=-1).astype(np.uint8)}
    _grad_summary(out, grads, params)
    return out


ICFG = dict(height=32, width=32, init_channels=64, num_down_samples=4, weight_decay_rate=1e-5,
            loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4), loss_type="xentropy+dice")
IN_ = 2


def build_unetinter(dtype=np.float64):
    """One UNetInter training step (image + 2-channel click guide as 5 input channels, instance_norm, xentropy + dice):
    oracle/gunet_ref.py with prefix "UNetInter", pinned by tests/test_oracle_gunet.py."""
    from oracle import gunet_ref as GU
    cfg = GU.unetinter_cfg(channel=3, guide_channel=2, **ICFG)
    params = GU.init_params(cfg, seed=WEIGHT_SEED)
    images, labels = synthetic.make_batch(IN_, ICFG["height"], ICFG["width"], 3, seed=DATA_SEED + 3)
    _, guide = synthetic.make_guides(images, labels, 200, 2, seed=4)
    inputs = GU.unetinter_inputs(images, guide)
    tape = GU.forward({k: v.astype(dtype) for k, v in params.items()}, {k: v.astype(dtype) for k, v in inputs.items()},
                      cfg, True)
    loss, dl = GU.loss_and_dlogits(tape, labels, cfg)
    grads = GU.backward(tape, dl, cfg)
    out = {"images": images, "sp_guide": guide, "labels": labels, "logits": tape.logits.astype(np.float32),
           "loss": np.float64(loss), "reg_loss": np.float64(GU.regularization_loss(params, cfg))}
    _grad_summary(out, grads, params)
    return out


def build_input_stage():
    """The per-sample input map function (crop, align-corners resize, window, noise, flips, Gaussian guide) on a
    seeded batch of synthetic 16-bit slices: oracle/input_ref.py, pinned by tests/test_oracle_input.py."""
    from oracle import input_ref as IR
    rng = np.random.default_rng(DATA_SEED + 4)
    n, c, src, out = 3, 3, 96, (48, 64)
    yy, xx = np.mgrid[0:src, 0:src]
    slices = (rng.integers(0, 1500, (n, c, src, src)) + 600 + 300 * np.sin(yy / 11.0) * np.cos(xx / 7.0)).astype(np.uint16)
    seg = (rng.integers(0, 3, (n, src, src)) * 64).astype(np.uint8)
    bbox = np.array([[3, 5, 70, 60], [10, 0, 50, 96], [0, 20, 96, 40]], np.int32)
    clip = np.array([[700, 1900], [650, 2100], [800, 1800]], np.float32)
    present = np.array([[1, 1, 1], [0, 1, 1], [1, 1, 0]], np.uint8)
    slices[1, 0] = 0
    slices[2, 2] = 0
    flips = np.array([0, 1, 3], np.int32)
    centers = [np.array([[20.0, 30.0]], np.float32), np.zeros((0, 2), np.float32),
               np.array([[10.0, 5.0], [60.0, 30.0]], np.float32)]
    stddevs = [np.array([[4.0, 6.0]], np.float32), np.zeros((0, 2), np.float32),
               np.array([[0.5, 3.0], [8.0, 2.0]], np.float32)]
    res = [IR.data_processing_train(slices[i], seg[i], bbox[i], clip[i], 64, out, present=present[i], noise_scale=0.05,
                                    seed=77, offset=5, sample=i, flip=int(flips[i]), centers=centers[i],
                                    stddevs=stddevs[i], with_guide=True) for i in range(n)]
    return {"slices": slices, "seg": seg, "bbox": bbox, "clip": clip, "present": present, "flips": flips,
            "centers": np.stack([np.pad(ci, ((0, 2 - len(ci)), (0, 0)), constant_values=-1) for ci in centers]),
            "stddevs": np.stack([np.pad(si, ((0, 2 - len(si)), (0, 0)), constant_values=1) for si in stddevs]),
            "n_centers": np.array([len(ci) for ci in centers], np.int32),
            "images": np.stack([r[0] for r in res]), "labels": np.stack([r[1] for r in res]),
            "guide": np.stack([r[2] for r in res])}


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, fn in (("unet2d_step.npz", build), ("gunet_step.npz", build_gunet), ("unet3d_step.npz", build_unet3d),
                     ("unetinter_step.npz", build_unetinter), ("input_stage.npz", build_input_stage)):
        path = os.path.join(here, name)
        np.savez_compressed(path, **fn())
        print("wrote", path, os.path.getsize(path), "bytes")
