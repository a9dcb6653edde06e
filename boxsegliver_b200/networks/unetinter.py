"""`UNetInter` with the reference's constructor / call contract (/root/reference/NetworksV2/UNetInter.py:30-209) on the
sm_100a engine: `model(inputs, mode, **yaml)` with inputs {images, sp_guide, labels}; the click guide enters as extra
input channels (UNetInter.py:89-90), or -- with --mid_cat -- in front of the first max-pool (UNetInter.py:124-125)."""
from __future__ import annotations

import numpy as np

from ..gunet_engine import UNetInterConfig, UNetInterEngine
from ..solver import engine_optimizer_kwargs
from .base import ModeKeys
from .unet import LossHandle, UNet


class UNetInter(UNet):
    def __init__(self, args, name=None):
        super().__init__(args, name or "UNetInter")
        self.use_spatial_guide = getattr(args, "use_spatial", False)        # UNetInter.py:41

    def _build_network(self, *args, **kwargs):
        # --img_grad (5 shipped 101_unetinter*.sh scripts): UNetInter.py:82-86 computes tf.image.image_gradients into
        # self.dy / self.dx, but the concat that would feed them to the network is commented out there -- the graph that
        # trains is the one without them, so the flag is accepted and changes nothing here either.
        if getattr(self.args, "without_norm", False):
            raise NotImplementedError("--without_norm is outside the accelerated path")
        if self.ctx is None:
            from ..device import Context
            self.ctx = Context(0)
            self.world = 1
        w_rate, b_rate = self._get_regularizer()
        cfg = UNetInterConfig(
            batch=self.bs, height=int(self.height), width=int(self.width), channel=self.channel,
            classes=tuple(self.classes), init_channels=kwargs.get("init_channels", 64),
            num_down_samples=kwargs.get("num_down_samples", 4), normalizer=self._get_normalization(),
            weight_decay_rate=w_rate or 0.0, bias_decay=(w_rate is not None and b_rate is None),
            loss_type=getattr(self.args, "loss_type", "xentropy"),
            loss_weight_type=getattr(self.args, "loss_weight_type", "none"),
            loss_numeric_w=tuple(getattr(self.args, "loss_numeric_w", None) or ()),
            loss_proportion_decay=getattr(self.args, "loss_proportion_decay", 1000.0),
            **engine_optimizer_kwargs(self.args), weight_init=self._get_initializer(),
            training=self.mode == ModeKeys.TRAIN, world=getattr(self, "world", 1),
            guide_channel=getattr(self.args, "guide_channel", 2), dropout_seed=getattr(self.args, "seed", 0),
            mid_cat=bool(getattr(self.args, "mid_cat", False)))
        if self.engine is None or self.engine.user_cfg != cfg:
            if self.engine is not None:
                self.engine.close()
            self.engine = UNetInterEngine(self.ctx, cfg)
            self.engine.init_weights(seed=getattr(self.args, "seed", 0))
        self.ret_prob = kwargs.get("ret_prob", False)
        self.ret_pred = kwargs.get("ret_pred", False)
        self._layers["logits"] = self.engine.logits

    def _build_loss(self):
        lt = self.args.loss_type
        if "xentropy" not in lt and "dice" not in lt:
            raise ValueError("Not supported loss_type: {}".format(lt))   # UNetInter.py:177-178
        self._loss = LossHandle(self)
        return self._loss

    def feed(self, images: np.ndarray, labels: np.ndarray | None = None, sp_guide: np.ndarray | None = None):
        self.engine.set_inputs(images, labels, sp_guide)
