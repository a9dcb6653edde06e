"""Mirror of /root/reference/core/models.py: model zoo, model flags, YAML kwargs, model_fn."""
from __future__ import annotations

from pathlib import Path

import yaml

from .networks import GUNet, UNet, UNet3D, UNetInter
from .networks.base import ModeKeys

MODEL_ZOO = [UNet, GUNet, UNet3D, UNetInter]


def add_arguments(parser):
    """core/models.py:41-89 -- same flags, same defaults."""
    g = parser.add_argument_group(title="Model Arguments")
    g.add_argument("--model", type=str, choices=[c.__name__ for c in MODEL_ZOO], required=True)
    g.add_argument("--model_config", type=str, required=False)
    g.add_argument("--classes", type=str, nargs="+", required=True)
    g.add_argument("--batch_size", type=int, default=8)
    g.add_argument("--weight_init", type=str, default="xavier", choices=["trunc_norm", "xavier"])
    g.add_argument("--normalizer", type=str, default="batch_norm", choices=["batch_norm", "instance_norm"])
    g.add_argument("--cls_branch", action="store_true")
    g.add_argument("--load_weights", type=str)
    g.add_argument("--load_weights_version", type=str, default="checkpoint")
    g.add_argument("--weights_scope", type=str)
    g.add_argument("--without_norm", action="store_true")
    g.add_argument("--batches_per_epoch", type=int, default=2000)
    g.add_argument("--eval_per_epoch", action="store_true")
    g.add_argument("--dropout", type=float)
    g.add_argument("--img_grad", action="store_true")
    g.add_argument("--mid_cat", action="store_true")


def get_model_params(args, build_metrics=False, build_summaries=False):
    """core/models.py:92-118: model class + kwargs from <model>.yml (or --model_config)."""
    params = {}
    for cls in MODEL_ZOO:
        if cls.__name__ == args.model:
            params["model"] = cls
            break
    else:
        raise ValueError("Not supported model: " + args.model)
    cfg = Path(args.model_config) if getattr(args, "model_config", None) else \
        Path(__file__).parent / "networks" / (args.model + ".yml")
    if not cfg.exists():
        raise FileNotFoundError(str(cfg))
    with cfg.open() as f:
        params["model_kwargs"] = yaml.safe_load(f) or {}
    params["model_kwargs"]["build_metrics"] = build_metrics
    params["model_kwargs"]["build_summaries"] = build_summaries
    return params


class EstimatorSpec:
    def __init__(self, mode, loss=None, train_op=None, predictions=None, model=None):
        self.mode, self.loss, self.train_op, self.predictions, self.model = mode, loss, train_op, predictions, model


def model_fn(features, labels, mode, params, config=None):
    """core/models.py:224-281: build the model, (TRAIN) its loss and train op; return an EstimatorSpec."""
    inputs = dict(features) if isinstance(features, dict) else {"images": features}
    if labels is not None:
        inputs["labels"] = labels
    model = params.get("model_instance")
    if model is None:
        model = params["model"](params["args"])
        if params.get("ctx") is not None:
            model.bind_context(params["ctx"], params.get("world", 1))
        params["model_instance"] = model
    loss = model(inputs, mode, **params.get("model_kwargs", {}))
    train_op = None
    if mode == ModeKeys.TRAIN:
        solver = params.get("solver_instance")
        if solver is None:
            solver = params["solver"](params["args"])
            params["solver_instance"] = solver
        train_op = solver(loss)
    return EstimatorSpec(mode, loss, train_op, model.predictions, model)


def init_model(model, args):
    """core/models.py:161-185: restore `--load_weights` (a checkpoint prefix, or a directory next to model_dir read
    through its `--load_weights_version` CheckpointState file) into the built model, renaming the model's root scope
    to `--weights_scope` (or the scope found in the checkpoint). Returns the checkpoint's global_step, or None when
    --load_weights is not set."""
    if not getattr(args, "load_weights", None):
        return None
    from . import checkpoint
    weights_dir = Path(getattr(args, "model_dir", ".")).parent / args.load_weights
    target = weights_dir if weights_dir.is_dir() else Path(args.load_weights)
    return checkpoint.restore_engine(model.engine, target, weights_scope=getattr(args, "weights_scope", None),
                                     latest_filename=getattr(args, "load_weights_version", "checkpoint"))
