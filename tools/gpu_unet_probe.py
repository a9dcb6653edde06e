"""GPU probe: one U-Net training step on the engine vs the numpy oracle (bf16-storage emulation and fp64).

`python tools/gpu_unet_probe.py [--time]`. Exit code 0 iff the parity gates hold.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from boxsegliver_b200.device import Context, round_bf16  # noqa: E402
from boxsegliver_b200.engine import EngineConfig, UNetEngine  # noqa: E402
from boxsegliver_b200 import synthetic  # noqa: E402
from oracle import unet_ref as R  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def parity_case(ctx, n, hw, normalizer, loss_type="xentropy", wtype="numerical", verbose=True):
    kw = dict(height=hw, width=hw, channel=3, init_channels=64, num_down_samples=4, normalizer=normalizer,
              weight_decay_rate=1e-5, loss_type=loss_type, loss_weight_type=wtype,
              loss_numeric_w=(0.2, 0.4, 4.4) if wtype == "numerical" else ())
    ecfg = EngineConfig(batch=n, **kw)
    rcfg = R.UNetCfg(**kw)
    images, labels = synthetic.make_batch(n, hw, hw, 3, seed=1357 + n)
    params = R.init_params(rcfg, seed=7)
    eng = UNetEngine(ctx, ecfg)
    eng.set_weights(params)
    eng.set_inputs(images, labels)
    lr = 1e-3
    eng.forward(True)
    eng.predict_outputs(True)
    eng.loss_backward()
    ctx.check_device()
    logits = eng.logits.download(np.float32, (n, hw, hw, rcfg.num_classes))
    grads = eng.get_grads()
    stored = eng.get_stored_forward()
    dlogits_dev = eng.dlogits.download(np.float32, (n, hw, hw, rcfg.num_classes))
    p64s = {k: v.astype(np.float64) for k, v in params.items()}
    masks = eng.masks.download(np.uint8, (rcfg.num_classes - 1, n, hw, hw))
    counts = eng.read_counts()
    eng.optimizer_step(lr)
    ctx.check_device()
    data_loss, reg_loss = eng.read_loss()
    new_w = eng.get_weights()

    out = {"case": f"n={n} hw={hw} {normalizer} {loss_type}/{wtype}"}
    # ---- oracle with bf16 storage emulation (isolates implementation errors)
    p32 = {k: v.copy() for k, v in params.items()}
    tape = R.forward(p32, round_bf16(images), rcfg, True, rnd=round_bf16, stem_fp32=False)
    loss_o, dl = R.loss_and_dlogits(tape, labels, rcfg)
    g_o = R.backward(tape, dl, rcfg, rnd=round_bf16)
    out["logits_rel_emul"] = rel(logits, tape.logits)
    out["loss_dev"], out["loss_emul"] = data_loss, float(loss_o)
    out["reg_dev"], out["reg_oracle"] = reg_loss, R.regularization_loss(params, rcfg)
    worst = ("", 0.0)
    gr = {}
    for k, g in g_o.items():
        r = rel(grads[k], g)
        gr[k] = r
        if r > worst[1]:
            worst = (k, r)
    out["grad_rel_freerun_emul_worst"] = worst
    out["grad_rel_freerun_emul_median"] = float(np.median(list(gr.values())))
    # ---- backward on identical forward tensors: oracle backward over the DEVICE's stored tape
    tf_tape = R.tape_from_stored(p32, round_bf16(images), stored, logits, rcfg, wrnd=round_bf16)
    loss_tf, dl_tf = R.loss_and_dlogits(tf_tape, labels, rcfg)
    out["dlogits_rel"] = rel(dlogits_dev, dl_tf)
    g_tf = R.backward(tf_tape, dl_tf, rcfg, rnd=round_bf16)
    gr = {k: rel(grads[k], g) for k, g in g_tf.items()}
    worst = max(gr.items(), key=lambda t: t[1])
    out["grad_rel_worst"] = worst
    out["grad_rel_median"] = float(np.median(list(gr.values())))
    t64s = R.tape_from_stored(p64s, round_bf16(images).astype(np.float64), stored, logits, rcfg, wrnd=round_bf16)
    _, dl64s = R.loss_and_dlogits(t64s, labels, rcfg)
    g64s = R.backward(t64s, dl64s, rcfg)
    gr64 = {k: rel(grads[k], g) for k, g in g64s.items()}
    out["grad_rel_fp64bwd_worst"] = max(gr64.items(), key=lambda t: t[1])
    out["grad_rel_fp64bwd_median"] = float(np.median(list(gr64.values())))
    # ---- pure fp64 oracle (precision of the bf16 path)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    t64 = R.forward(p64, images.astype(np.float64), rcfg, True)
    l64, dl64 = R.loss_and_dlogits(t64, labels, rcfg)
    g64 = R.backward(t64, dl64, rcfg)
    out["logits_rel_fp64"] = rel(logits, t64.logits)
    out["loss_fp64"] = float(l64)
    w64 = max(((k, rel(grads[k], g)) for k, g in g64.items()), key=lambda t: t[1])
    out["grad_rel_fp64_worst"] = w64
    out["grad_rel_fp64_median"] = float(np.median([rel(grads[k], g) for k, g in g64.items()]))
    # ---- masks / counts: bit-exact given the device's own logits
    prob_dev = R.O.softmax(logits)
    m_ref = np.stack([(prob_dev[..., i] > 0.5).astype(np.uint8) for i in range(1, rcfg.num_classes)])
    margin = np.abs(prob_dev[..., 1:] - 0.5).transpose(3, 0, 1, 2)
    decided = margin > 1e-6
    out["mask_mismatch_decided"] = int(((masks != m_ref) & decided).sum())
    out["mask_near_ties"] = int((~decided).sum())
    ilr_ref = np.zeros_like(counts)
    for c in range(1, rcfg.num_classes):
        i_, l_, r_ = R.O.seg_counts(masks[c - 1][..., None], labels, c)
        ilr_ref[:, c - 1, 0], ilr_ref[:, c - 1, 1], ilr_ref[:, c - 1, 2] = i_, l_, r_
    out["counts_equal"] = bool(np.array_equal(counts, ilr_ref))
    # ---- optimizer: apply the oracle's Adam to the DEVICE gradients, compare updated weights
    worst_w = ("", 0.0)
    tg = R.total_grads(params, grads, rcfg)
    for k in R.trainable_names(rcfg, params):
        w, _, _ = R.O.adam_step(params[k].astype(np.float64), tg[k].astype(np.float64), 0.0, 0.0, 1, lr)
        dw_ref = w - params[k]
        dw_dev = new_w[k].astype(np.float64) - params[k]
        r = rel(dw_dev, dw_ref)
        if r > worst_w[1]:
            worst_w = (k, r)
    out["adam_update_rel_worst"] = worst_w
    if normalizer == "batch_norm":
        mm = max(rel(new_w[k], v) for k, v in tape.new_moving.items())
        out["moving_stats_rel_worst"] = mm
    ok = (out["logits_rel_emul"] < 1e-2 and worst[1] < 1e-2 and out["mask_mismatch_decided"] == 0
          and out["counts_equal"] and worst_w[1] < 1e-3 and abs(data_loss - float(loss_o)) < 1e-2 * abs(float(loss_o)))
    out["ok"] = bool(ok)
    if verbose:
        print(json.dumps(out, indent=1, default=str))
        for k, v in gr.items():
            print(f"   {k.replace('UNet/', ''):60s} bf16-bwd {v:.2e}  fp64-bwd {gr64[k]:.2e}")
    eng.close()
    return out


def time_case(ctx, n, hw, steps=10, warm=3):
    ecfg = EngineConfig(batch=n, height=hw, width=hw, loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4))
    eng = UNetEngine(ctx, ecfg)
    eng.init_weights(0)
    images, labels = synthetic.make_batch(n, hw, hw, 3)
    eng.set_inputs(images, labels)
    for _ in range(warm):
        eng.train_step(1e-3)
    ctx.sync()
    e0, e1 = ctx.new_event(), ctx.new_event()
    t0 = time.time()
    ctx.record(e0)
    for _ in range(steps):
        eng.train_step(1e-3)
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1) / steps
    wall = (time.time() - t0) / steps * 1e3
    ctx.check_device()
    loss = eng.read_loss()
    flop = 288.828e9 * n * (hw / 256.0) ** 2
    res = {"n": n, "hw": hw, "ms_per_step": ms, "wall_ms": wall, "slices_per_s": n / ms * 1e3,
           "tflops": flop / ms / 1e9, "loss": loss}
    print("TIME", json.dumps(res))
    eng.close()
    return res


def main():
    ctx = Context(0)
    rep = []
    fails = 0
    for args in [(2, 128, "batch_norm"), (3, 64, "instance_norm"), (2, 64, "batch_norm", "dice", "none"),
                 (2, 64, "batch_norm", "xentropy", "proportion")]:
        try:
            r = parity_case(ctx, *args)
            fails += 0 if r["ok"] else 1
            rep.append(r)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            fails += 1
            rep.append({"case": str(args), "exc": repr(e)})
    if "--time" in sys.argv:
        for n, hw in [(8, 256), (32, 256), (64, 256)]:
            try:
                rep.append(time_case(ctx, n, hw))
            except Exception as e:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                rep.append({"time": (n, hw), "exc": repr(e)})
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/unet_probe.json", "w") as f:
        json.dump(rep, f, indent=1, default=str)
    print("FAILURES", fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
