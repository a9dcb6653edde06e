"""Host-side mirror of the reference's model base class (/root/reference/NetworksV2/base.py:33-197).

Same build order (network -> loss -> metrics), same argument handling (`_get_regularizer`,
`_get_normalization`, `_get_weights_params`), same public attributes (`classes`, `num_classes`,
`mode`, `is_training`, `layers`, `probability`, `predictions`, `metrics_dict`). What the reference
expresses as TF graph construction is expressed here as the plan of a UNetEngine; what TF evaluates
in `sess.run` is evaluated by enqueuing the C-ABI kernels.
"""
from __future__ import annotations


class ModeKeys:
    TRAIN = "train"
    EVAL = "eval"
    PREDICT = "infer"


class BaseNet:
    def __init__(self, args, name=None):
        self._name = name
        self._args = args
        self._mode = None
        self._is_training = False
        self._inputs = {}
        self._layers = {}
        self.classes = ["Background"]      # base.py:44
        self.metrics_dict = {}
        self.predictions = {}
        self.probability = None
        self.engine = None

    # -- properties with the reference's names ------------------------------------------------
    @property
    def name(self):
        return self._name

    @property
    def args(self):
        return self._args

    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, new_mode):
        if new_mode in (ModeKeys.TRAIN, ModeKeys.EVAL, ModeKeys.PREDICT):
            self._mode = new_mode
            # base.py:77-78: `is_training` is a runtime value fed per run, not a graph constant
            self._is_training = new_mode == ModeKeys.TRAIN

    @property
    def is_training(self):
        return self._is_training

    @property
    def num_classes(self):
        return len(self.classes)

    @property
    def layers(self):
        return self._layers

    # -- helpers restated from base.py:128-178 ---------------------------------------------------
    def _get_regularizer(self):
        """(weights rate, biases rate). NOTE the inverted flag: biases are regularised UNLESS --bias_decay."""
        rate = getattr(self.args, "weight_decay_rate", 0.0) or 0.0
        if rate > 0:
            return rate, (None if getattr(self.args, "bias_decay", False) else rate)
        return None, None

    def _get_initializer(self):
        wi = getattr(self.args, "weight_init", "xavier")
        if wi not in ("xavier", "trunc_norm"):
            raise ValueError("Not supported weight initializer: " + wi)
        return wi

    def _get_normalization(self):
        n = getattr(self.args, "normalizer", "batch_norm")
        if n not in ("batch_norm", "instance_norm"):
            raise ValueError("Not supported normalization function: " + n)
        return n

    def _get_weights_params(self):
        w = {"tag": getattr(self.args, "tag", "")}
        if self.args.loss_weight_type == "numerical":
            w["numeric_w"] = self.args.loss_numeric_w
        elif self.args.loss_weight_type == "proportion":
            if self.args.loss_proportion_decay > 0:
                w["proportion_decay"] = self.args.loss_proportion_decay
        return w

    def _build_network(self, *args, **kwargs):
        raise NotImplementedError

    def _build_loss(self):
        raise NotImplementedError

    def _build_metrics(self):
        raise NotImplementedError

    def __call__(self, inputs, mode, *args, **kwargs):
        """base.py:180-197: network, then loss (TRAIN only), then metrics. Returns the loss handle or None."""
        self._inputs = inputs
        self.mode = mode
        self._build_network(*args, **kwargs)
        ret = None
        if self.mode == ModeKeys.TRAIN:
            ret = self._build_loss()
        if kwargs.get("build_metrics", False):
            self._build_metrics()
        return ret
