"""Forward-only consumers of the model output, on the device.

  * `tta_variants` / `Predictor.predict`: run_TTA of the reference -- /root/reference/entry/infer_2d.py:60-78 (2-D) and
    /root/reference/entry/main_eval_3d.py:246-287 (3-D): the prediction is argmax over classes of the mean of the
    probabilities of the original and mirrored inputs (each mirrored back), as uint8.
  * `GlobalDice`: the "global dice" accumulation of /root/reference/evaluators/evaluator_liver.py:304-327 over
    ConfusionMatrix.compute (/root/reference/loss_metrics.py:542-556): integer tp / fp / fn summed over batches,
    Dice = 2 tp / (2 tp + fn + fp).
The host never sees the probabilities: only the uint8 prediction (if asked for) and 4 integers per class come back.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def tta_variants(random_flip: int, eval_mirror: bool, three_d: bool = False) -> list[int]:
    """Axis masks (1 = W, 2 = H, 4 = D) of the forward passes run_TTA makes, in its order. The reference tests
    `random_flip & m > 0` for m = 1, 2, 3 (2-D) or 1..7 (3-D): a combined flip runs when ANY of its bits is set."""
    out = [0]
    if eval_mirror:
        out += [m for m in range(1, 8 if three_d else 4) if (random_flip & m) > 0]
    return out


class Predictor:
    """TTA prediction around a planned engine (UNetEngine / GUNetEngine / UNet3DEngine in eval mode)."""

    def __init__(self, engine, random_flip: int = 0, eval_mirror: bool = False):
        self.eng, self.ctx = engine, engine.ctx
        cfg = engine.cfg
        self.three_d = hasattr(cfg, "depth")
        self.variants = tta_variants(random_flip, eval_mirror, self.three_d)
        self.n = cfg.batch
        self.d = cfg.depth if self.three_d else 1
        self.h, self.w = cfg.height, cfg.width
        self.cin = cfg.in_channels if self.three_d else cfg.channel
        self.k = cfg.num_classes
        self.pixels = self.n * self.d * self.h * self.w
        self.orig = self.ctx.alloc(self.pixels * self.cin * 4)
        self.acc = self.ctx.alloc(self.pixels * self.k * 4)
        self.pred = self.ctx.alloc(self.pixels)
        self.guide0 = None
        if getattr(cfg, "use_spatial", False) and not self.three_d:
            self.guide0 = self.ctx.alloc(self.pixels * cfg.guide_channel * 4)

    def _flip(self, src, dst, c, axes, accumulate=0):
        self.ctx.call("bsl_flip_f32", C.c_longlong(self.n), C.c_int(self.d), C.c_int(self.h), C.c_int(self.w), C.c_int(c),
                      C.c_int(axes), C.c_int(accumulate), src.p, dst.p, self.eng.stream)

    def predict(self, images: np.ndarray, context: np.ndarray | None = None, sp_guide: np.ndarray | None = None,
                download: bool = True, keep_variant_probs: bool = False):
        """images as the engine's set_inputs takes them. Returns the uint8 prediction [n,(d),h,w] (or None)."""
        eng, cfg = self.eng, self.eng.cfg
        if self.three_d:
            if cfg.use_spatial:
                images = np.concatenate((images, sp_guide), axis=-1)
        elif hasattr(eng, "set_guides"):
            eng.set_guides(context, sp_guide)
            if self.guide0 is not None:
                self.ctx.call("bsl_memcpy_d2d", self.guide0.p, eng.guides[0][0].p, C.c_size_t(self.guide0.nbytes), eng.stream)
        self.orig.upload(np.ascontiguousarray(images, np.float32))
        self.variant_probs = []
        for i, axes in enumerate(self.variants):
            self._flip(self.orig, eng.images, self.cin, axes)
            if self.guide0 is not None:
                self._flip(self.guide0, eng.guides[0][0], cfg.guide_channel, axes)
            eng.forward(False)
            eng.predict_outputs(False)
            self._flip(eng.prob, self.acc, self.k, axes, accumulate=int(i > 0))
            if keep_variant_probs:
                self.variant_probs.append(eng.prob.download(np.float32, (self.n, self.d, self.h, self.w, self.k)))
        self.ctx.call("bsl_tta_finalize", C.c_longlong(self.pixels), C.c_int(self.k), C.c_int(len(self.variants)),
                      self.acc.p, None, self.pred.p, eng.stream)
        if not download:
            return None
        shp = (self.n, self.d, self.h, self.w) if self.three_d else (self.n, self.h, self.w)
        return self.pred.download(np.uint8, shp)

    def close(self):
        for b in (self.orig, self.acc, self.pred, self.guide0):
            if b is not None:
                b.free()


class GlobalDice:
    """tp / fp / tn / fn per foreground class accumulated on the device over any number of batches."""

    def __init__(self, ctx, classes):
        self.ctx, self.classes = ctx, list(classes)          # foreground class names, label value = index + 1
        self.counts = ctx.alloc(len(self.classes) * 4 * 8).zero()

    def update_from_pred(self, pred_buf, labels_buf, n: int, stream=None):
        """`pred_buf`: uint8 label map (argmax). test = (pred == cls)."""
        for i in range(len(self.classes)):
            self.ctx.call("bsl_confusion_counts", C.c_longlong(n), pred_buf.p, C.c_int(i + 1), labels_buf.p, C.c_int(i + 1),
                          self.counts.at(i * 32), self.ctx.stream_arg(stream))

    def update_from_masks(self, masks_buf, labels_buf, n: int, stream=None):
        """`masks_buf`: the `<Cls>Pred` uint8 masks [classes-1][n] (evaluator_liver.py:316-318)."""
        for i in range(len(self.classes)):
            self.ctx.call("bsl_confusion_counts", C.c_longlong(n), masks_buf.at(i * n), C.c_int(-1), labels_buf.p,
                          C.c_int(i + 1), self.counts.at(i * 32), self.ctx.stream_arg(stream))

    def read(self) -> dict:
        c = self.counts.download(np.uint64, (len(self.classes), 4))
        return {cls: dict(tp=int(c[i, 0]), fp=int(c[i, 1]), tn=int(c[i, 2]), fn=int(c[i, 3]))
                for i, cls in enumerate(self.classes)}

    def results(self) -> dict:
        out = {}
        for cls, m in self.read().items():
            den = 2 * m["tp"] + m["fn"] + m["fp"]
            out[cls + "/Dice"] = 2 * m["tp"] / den if den else float("nan")   # evaluator_liver.py:324-326
        return out
