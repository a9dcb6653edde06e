"""`UNet3D` with the reference's constructor / call contract (/root/reference/NetworksV2/UNet3D.py:94-202) on the
sm_100a engine: `model(inputs, mode, **yaml)` with inputs {images [n,d,h,w,c], labels [n,d,h,w], [sp_guide]}."""
from __future__ import annotations

import numpy as np

from .. import distribution_utils
from ..loss_metrics import metrics_from_counts
from ..solver import engine_optimizer_kwargs
from ..unet3d_engine import UNet3DConfig, UNet3DEngine
from .base import BaseNet, ModeKeys
from .unet import LossHandle


class UNet3D(BaseNet):
    def __init__(self, args, name=None):
        super().__init__(args)
        self._name = name or "UNet3D"
        self.classes.extend(self.args.classes)
        self.bs = distribution_utils.per_device_batch_size(args.batch_size, getattr(args, "num_gpus", 1))
        self.depth = args.im_depth
        self.height = args.im_height
        self.width = args.im_width
        self.channel = args.im_channel
        self.use_spatial = getattr(args, "use_spatial", False)
        self.ctx = None

    def bind_context(self, ctx, world: int = 1):
        self.ctx = ctx
        self.world = world
        return self

    def _build_network(self, *args, **kwargs):
        if getattr(self.args, "img_grad", False):
            raise NotImplementedError("--img_grad is outside the accelerated path")
        if self.ctx is None:
            from ..device import Context
            self.ctx = Context(0)
            self.world = 1
        w_rate, b_rate = self._get_regularizer()
        cfg = UNet3DConfig(
            batch=self.bs, depth=int(self.depth), height=int(self.height), width=int(self.width), channel=self.channel,
            classes=tuple(self.classes), init_channels=kwargs.get("init_channels", 30),
            max_channels=kwargs.get("max_channels", 320), num_pool_layers=kwargs.get("num_pool_layers", 4),
            use_spatial=self.use_spatial, guide_channel=getattr(self.args, "guide_channel", 2),
            normalizer=self._get_normalization(), weight_decay_rate=w_rate or 0.0,
            bias_decay=(w_rate is not None and b_rate is None), loss_type=getattr(self.args, "loss_type", "xentropy"),
            loss_weight_type=getattr(self.args, "loss_weight_type", "none"),
            loss_numeric_w=tuple(getattr(self.args, "loss_numeric_w", None) or ()),
            loss_proportion_decay=getattr(self.args, "loss_proportion_decay", 1000.0),
            **engine_optimizer_kwargs(self.args), weight_init=self._get_initializer(), training=self.mode == ModeKeys.TRAIN,
            world=getattr(self, "world", 1))
        if self.engine is None or self.engine.cfg != cfg:
            if self.engine is not None:
                self.engine.close()
            self.engine = UNet3DEngine(self.ctx, cfg)
            self.engine.init_weights(seed=getattr(self.args, "seed", 0))
        self.ret_prob = kwargs.get("ret_prob", False)
        self.ret_pred = kwargs.get("ret_pred", False)
        self._layers["logits"] = self.engine.logits

    def _build_loss(self):
        if "xentropy" not in self.args.loss_type:
            raise ValueError("Not supported loss_type: {}".format(self.args.loss_type))   # UNet3D.py:198-199
        self._loss = LossHandle(self)
        return self._loss

    def _build_metrics(self):
        if self.ret_pred:
            self._want_metrics = True

    def feed(self, images, labels=None, sp_guide=None):
        self.engine.set_inputs(images, labels, sp_guide)

    def run_forward(self):
        eng = self.engine
        eng.forward(self.is_training)
        want_counts = getattr(self, "_want_metrics", False) and "labels" in self._inputs
        eng.predict_outputs(want_counts)
        c = eng.cfg
        shp = (c.batch, c.depth, c.height, c.width)
        self.probability = eng.prob.download(np.float32, shp + (c.num_classes,))
        masks = eng.masks.download(np.uint8, (c.num_classes - 1,) + shp)
        for i in range(1, c.num_classes):
            self.predictions[self.classes[i] + "Pred"] = masks[i - 1][..., None]
        if want_counts:
            self.metrics_dict = metrics_from_counts(eng.read_counts(), self.classes,
                                                    getattr(self.args, "metrics_train", ["Dice"]))
        return self.probability
