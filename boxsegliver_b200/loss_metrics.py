"""Host-side pieces of /root/reference/loss_metrics.py: the loss/metric flags, the float formulas applied to the
integer I/L/R counts the device returns, and ConfusionMatrix. (The per-pixel work is in csrc/loss.cu.)"""
from __future__ import annotations

import numpy as np


def add_arguments(parser):
    """loss_metrics.py:26-67 -- same flags, same defaults."""
    group = parser.add_argument_group(title="Loss Arguments")
    group.add_argument("--weight_decay_rate", type=float, default=1e-5)
    group.add_argument("--bias_decay", action="store_true")
    group.add_argument("--loss_type", type=str, default="xentropy", choices=["xentropy", "dice", "xentropy+dice"])
    group.add_argument("--loss_weight_type", type=str, default="none",
                       choices=["none", "numerical", "proportion", "boundary"])
    group.add_argument("--loss_numeric_w", type=float, nargs="+")
    group.add_argument("--loss_proportion_decay", type=float, default=1000)
    group.add_argument("--metrics_train", type=str, default=["Dice"], choices=["Dice", "VOE", "VD"], nargs="+")
    group.add_argument("--metrics_eval", type=str, default=["Dice"],
                       choices=["Dice", "VOE", "RVD", "ASSD", "RMSD", "MSD"], nargs="+")


def metrics_from_counts(ilr: np.ndarray, classes, names=("Dice",)):
    """{"<Cls>/<Met>": batch mean} from uint32 counts [n, classes-1, (I, L, R)]; fp32 arithmetic in the
    reference's operation order (metric_dice :261-296, metric_voe :299-317, metric_vd :320-339)."""
    out = {}
    f = np.float32
    for ci in range(1, len(classes)):
        i, l, r = (ilr[:, ci - 1, k].astype(np.float32) for k in range(3))
        for m in names:
            key = f"{classes[ci]}/{m}"
            ml = m.lower()
            if ml == "dice":
                v = (f(2) * i + f(1e-5)) / (l + r + f(1e-5))
            elif ml == "voe":
                v = f(100) * (f(1) - i / ((l + r - i) + f(1e-5)))   # sum(clip(p + l, 0, 1)) = L + R - I for 0/1 masks
            elif ml == "vd":
                v = f(100) * (np.abs(l - r) / (r + f(1e-5)))
            else:
                raise ValueError("unknown metric " + m)
            out[key] = f(np.mean(v, dtype=np.float32))
    return out


class ConfusionMatrix(object):
    """Same interface as loss_metrics.ConfusionMatrix (:506-580): integer tp / fp / tn / fn of a binary mask pair.
    `from_counts` builds it from the device's (I, L, R) sums without touching the masks on the host."""

    def __init__(self, test=None, reference=None):
        self.reference = reference
        self.test = test
        self.reset()

    @classmethod
    def from_counts(cls, inter: int, left: int, right: int, size: int):
        cm = cls()
        cm.tp = int(inter)
        cm.fp = int(left) - int(inter)
        cm.fn = int(right) - int(inter)
        cm.size = int(size)
        cm.tn = cm.size - cm.tp - cm.fp - cm.fn
        cm.test_empty, cm.test_full = left == 0, left == size
        cm.reference_empty, cm.reference_full = right == 0, right == size
        return cm

    def set_test(self, test):
        self.test = test
        self.reset()

    def set_reference(self, reference):
        self.reference = reference
        self.reset()

    def reset(self):
        self.tp = self.fp = self.tn = self.fn = None
        self.size = None
        self.test_empty = self.test_full = self.reference_empty = self.reference_full = None

    def compute(self):
        if self.test is None or self.reference is None:
            raise ValueError("'test' and 'reference' must both be set to compute confusion matrix.")
        assert self.test.shape == self.reference.shape, "Shape mismatch: {} and {}".format(
            self.test.shape, self.reference.shape)
        t, r = self.test != 0, self.reference != 0
        self.tp = int((t & r).sum())
        self.fp = int((t & ~r).sum())
        self.tn = int((~t & ~r).sum())
        self.fn = int((~t & r).sum())
        self.size = self.reference.size
        self.test_empty = not np.any(self.test)
        self.test_full = bool(np.all(self.test))
        self.reference_empty = not np.any(self.reference)
        self.reference_full = bool(np.all(self.reference))

    def get_matrix(self):
        if any(e is None for e in (self.tp, self.fp, self.tn, self.fn)):
            self.compute()
        return self.tp, self.fp, self.tn, self.fn

    def get_size(self):
        if self.size is None:
            self.compute()
        return self.size

    def get_existence(self):
        if any(c is None for c in (self.test_empty, self.test_full, self.reference_empty, self.reference_full)):
            self.compute()
        return self.test_empty, self.test_full, self.reference_empty, self.reference_full
