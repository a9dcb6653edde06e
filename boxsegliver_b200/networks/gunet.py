"""`GUNet` with the reference's constructor / call contract (/root/reference/NetworksV2/GUNet.py:220-413) on the
sm_100a engine: `model(inputs, mode, **yaml)` with inputs {images, labels, [context], [sp_guide]}."""
from __future__ import annotations

import numpy as np

from ..gunet_engine import GUNetConfig, GUNetEngine
from ..solver import engine_optimizer_kwargs
from .base import ModeKeys
from .unet import LossHandle, UNet


class GUNet(UNet):
    def __init__(self, args, name=None):
        super().__init__(args, name or "GUNet")
        self.use_context_guide = getattr(args, "use_context", False)
        self.use_spatial_guide = getattr(args, "use_spatial", False)
        self.side_dropout = getattr(args, "side_dropout", 0.5)
        self.dropout = getattr(args, "dropout", None)
        self.use_se = getattr(args, "use_se", False)
        if hasattr(args, "ct_conv"):
            raise NotImplementedError("ct_conv (convolutional context sub-network) is outside the accelerated path")

    def _build_network(self, *args, **kwargs):
        if self.ctx is None:
            from ..device import Context
            self.ctx = Context(0)
            self.world = 1
        w_rate, b_rate = self._get_regularizer()
        cfg = GUNetConfig(
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
            mod_layers=tuple(kwargs.get("mod_layers", [])),
            context_fc_channels=tuple(kwargs.get("context_fc_channels", [256])),
            context_model=kwargs.get("context_model", "fc"),
            norm_with_center=kwargs.get("norm_with_center", False), norm_with_scale=kwargs.get("norm_with_scale", False),
            after_affine=kwargs.get("after_affine", False),
            use_context=self.use_context_guide, use_spatial=self.use_spatial_guide,
            guide_channel=getattr(self.args, "guide_channel", 1), side_dropout=self.side_dropout or 0.0,
            dropout_seed=getattr(self.args, "seed", 0), use_se=self.use_se, fix=getattr(self.args, "fix", False),
            without_norm=getattr(self.args, "without_norm", False), dropout=self.dropout,
            img_grad=bool(getattr(self.args, "img_grad", False)))
        if self.engine is None or self.engine.cfg != cfg:
            if self.engine is not None:
                self.engine.close()
            self.engine = GUNetEngine(self.ctx, cfg)
            self.engine.init_weights(seed=getattr(self.args, "seed", 0))
        self.ret_prob = kwargs.get("ret_prob", False)
        self.ret_pred = kwargs.get("ret_pred", False)
        self._layers["logits"] = self.engine.logits

    def _build_loss(self):
        lt = self.args.loss_type
        if "xentropy" not in lt and "dice" not in lt:
            raise ValueError("Not supported loss_type: {}".format(lt))   # GUNet.py:409-410
        self._loss = LossHandle(self)
        return self._loss

    def feed(self, images: np.ndarray, labels: np.ndarray | None = None, context: np.ndarray | None = None,
             sp_guide: np.ndarray | None = None):
        self.engine.set_inputs(images, labels)
        self.engine.set_guides(context, sp_guide)
