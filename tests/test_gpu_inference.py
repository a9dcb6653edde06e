"""Device-side inference consumers (TTA mean of probabilities, argmax, confusion counts) must be BIT-EXACT against the
reference's host arithmetic (run_TTA: entry/infer_2d.py:60-78, entry/main_eval_3d.py:246-287; ConfusionMatrix:
loss_metrics.py:542-556) applied to the device's own per-variant probabilities."""
import numpy as np
import pytest

from boxsegliver_b200 import synthetic
from boxsegliver_b200.engine import EngineConfig, UNetEngine
from boxsegliver_b200.inference import GlobalDice, Predictor, tta_variants
from boxsegliver_b200.unet3d_engine import UNet3DConfig, UNet3DEngine

pytestmark = pytest.mark.gpu

AX2 = {1: (2,), 2: (1,), 3: (1, 2)}                          # mask -> numpy axes of [n,h,w,c]
AX3 = {1: (3,), 2: (2,), 3: (2, 3), 4: (1,), 5: (1, 3), 6: (1, 2), 7: (1, 2, 3)}


def _host_tta(variant_probs, variants, axmap, squeeze_d):
    probs = None
    for p, m in zip(variant_probs, variants):
        if squeeze_d:
            p = p[:, 0]
        p = np.flip(p, axis=axmap[m]) if m else p
        probs = p.copy() if probs is None else probs + p      # probs += np.flip(prob, ...)
    avg = probs / np.float32(len(variants))
    return np.argmax(avg, axis=-1).astype(np.uint8)


def _conf(test, ref):                                          # ConfusionMatrix.compute
    return dict(tp=int(((test != 0) * (ref != 0)).sum()), fp=int(((test != 0) * (ref == 0)).sum()),
                tn=int(((test == 0) * (ref == 0)).sum()), fn=int(((test == 0) * (ref != 0)).sum()))


@pytest.mark.parametrize("random_flip,eval_mirror", [(3, True), (1, True), (3, False)])
def test_tta_2d_and_global_dice(ctx, random_flip, eval_mirror):
    n, hw = 3, 64
    eng = UNetEngine(ctx, EngineConfig(batch=n, height=hw, width=hw, training=False))
    eng.init_weights(seed=3)
    images, labels = synthetic.make_batch(n, hw, hw, 3, seed=77)
    eng.set_inputs(images, labels)
    pr = Predictor(eng, random_flip, eval_mirror)
    assert pr.variants == tta_variants(random_flip, eval_mirror)
    pred = pr.predict(images, keep_variant_probs=True)
    ctx.check_device()
    ref = _host_tta(pr.variant_probs, pr.variants, AX2, True)
    assert np.array_equal(pred, ref)
    gd = GlobalDice(ctx, ["Liver", "Tumor"])
    for _ in range(2):                                          # two "batches": counts accumulate
        gd.update_from_pred(pr.pred, eng.labels, n * hw * hw, eng.stream)
    got = gd.read()
    for i, cls in enumerate(("Liver", "Tumor")):
        want = _conf((pred == i + 1).astype(int), (labels == i + 1).astype(int))
        assert got[cls] == {k: 2 * v for k, v in want.items()}
    # masks variant (evaluator_liver.py:316-318): test = <Cls>Pred mask of the un-mirrored forward
    eng.set_inputs(images, labels)
    eng.forward(False)
    eng.predict_outputs(False)
    masks = eng.masks.download(np.uint8, (2, n, hw, hw))
    gm = GlobalDice(ctx, ["Liver", "Tumor"])
    gm.update_from_masks(eng.masks, eng.labels, n * hw * hw, eng.stream)
    for i, cls in enumerate(("Liver", "Tumor")):
        assert gm.read()[cls] == _conf(masks[i].astype(int), (labels == i + 1).astype(int))
    m = gm.read()["Liver"]
    den = 2 * m["tp"] + m["fn"] + m["fp"]
    assert gm.results()["Liver/Dice"] == (2 * m["tp"] / den if den else gm.results()["Liver/Dice"])
    pr.close()
    eng.close()


def test_tta_3d_all_mirrors(ctx):
    n, d, hw = 1, 4, 32
    eng = UNet3DEngine(ctx, UNet3DConfig(batch=n, depth=d, height=hw, width=hw, training=False))
    eng.init_weights(seed=1)
    images, labels = synthetic.make_volume_batch(n, d, hw, hw, seed=4)
    eng.set_inputs(images, labels)
    pr = Predictor(eng, random_flip=7, eval_mirror=True)
    assert pr.variants == [0, 1, 2, 3, 4, 5, 6, 7]
    pred = pr.predict(images, keep_variant_probs=True)
    ctx.check_device()
    assert np.array_equal(pred, _host_tta(pr.variant_probs, pr.variants, AX3, False))
    assert tta_variants(4, True, three_d=True) == [0, 4, 5, 6, 7]      # the reference's `flip & m > 0` quirk
    pr.close()
    eng.close()
