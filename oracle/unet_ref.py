"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/tf_ops.py header; parity unpinned by the reference).

numpy restatement of one training / inference step of the reference's 2-D U-Net:
  graph      /root/reference/NetworksV2/UNet.py:58-117   (encoder, bridge, decoder, logits, softmax, masks)
  arg scope  /root/reference/NetworksV2/UNet.py:41-56    (slim.conv2d gets the normaliser, conv2d_transpose does not)
  loss       /root/reference/NetworksV2/UNet.py:120-135  + /root/reference/loss_metrics.py:115-226
  metrics    /root/reference/NetworksV2/UNet.py:137-155  + /root/reference/loss_metrics.py:261-339
  optimizer  /root/reference/core/solver.py:204-243      (Adam beta2 = 0.99, UPDATE_OPS before minimize)

Variables are named as TF-slim names them (UNet/Encode1/Repeat/convolution2d_1/weights, ...), so a
dict of numpy arrays here is interchangeable with a TF checkpoint's variable map.

`rnd` is the storage-rounding hook: identity gives the plain fp32/fp64 reference; `round_bf16`
rounds exactly where the B200 engine stores bf16 (conv outputs, activations, gradients wrt
activations, the bf16 weight shadows), which isolates implementation errors from precision.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import tf_ops as O


@dataclass
class UNetCfg:
    height: int = 256
    width: int = 256
    channel: int = 3
    classes: tuple = ("Background", "Liver", "Tumor")  # BaseNet.classes starts with Background (base.py:44)
    init_channels: int = 64          # NetworksV2/UNet.yml
    num_down_samples: int = 4
    normalizer: str = "batch_norm"   # or "instance_norm"  (core/models.py --normalizer)
    weight_decay_rate: float = 1e-5  # loss_metrics.py:28-31
    bias_decay: bool = False         # inverted flag: biases ARE regularised unless it is set (base.py:131)
    loss_type: str = "xentropy"      # "xentropy" | "dice"  (UNet.py:123-132)
    loss_weight_type: str = "none"
    loss_numeric_w: tuple = ()
    loss_proportion_decay: float = 1000.0
    bn_decay: float = 0.999          # slim.batch_norm default
    bn_eps: float = 1e-3
    in_eps: float = 1e-6             # slim.instance_norm default

    @property
    def num_classes(self):
        return len(self.classes)


def conv_specs(cfg: UNetCfg):
    """Ordered layer list of UNet._build_network: (kind, scope, cin, cout, level)."""
    specs = []
    c = cfg.init_channels
    cin = cfg.channel
    for i in range(cfg.num_down_samples):
        for j in (1, 2):
            specs.append(("conv", f"UNet/Encode{i + 1}/Repeat/convolution2d_{j}", cin, c, i))
            cin = c
        c *= 2
    for j in (1, 2):
        specs.append(("conv", f"UNet/ED-Bridge/ED-Bridge_{j}", cin, c, cfg.num_down_samples))
        cin = c
    for i in reversed(range(cfg.num_down_samples)):
        c //= 2
        specs.append(("convT", f"UNet/Decode{i + 1}/Conv2d_transpose", cin, cin // 2, i))
        cin_cat = c + cin // 2
        for j in (1, 2):
            specs.append(("conv", f"UNet/Decode{i + 1}/Repeat/convolution2d_{j}", cin_cat if j == 1 else c, c, i))
        cin = c
    specs.append(("logits", "UNet/AdjustChannels", cin, cfg.num_classes, 0))
    return specs


def norm_scope(cfg: UNetCfg) -> str:
    return "BatchNorm" if cfg.normalizer == "batch_norm" else "InstanceNorm"


def init_params(cfg: UNetCfg, seed: int = 0, dtype=np.float32) -> dict:
    """Seeded slim.xavier_initializer() weights, zero biases, gamma 1 / beta 0, moving mean 0 / variance 1."""
    rng = np.random.default_rng(seed)
    p = {}
    ns = norm_scope(cfg)
    for kind, scope, cin, cout, _ in conv_specs(cfg):
        if kind == "conv":
            p[f"{scope}/weights"] = O.xavier_uniform(rng, (3, 3, cin, cout), 9 * cin, 9 * cout, dtype)
            p[f"{scope}/{ns}/gamma"] = np.ones(cout, dtype)
            p[f"{scope}/{ns}/beta"] = np.zeros(cout, dtype)
            if cfg.normalizer == "batch_norm":
                p[f"{scope}/{ns}/moving_mean"] = np.zeros(cout, dtype)
                p[f"{scope}/{ns}/moving_variance"] = np.ones(cout, dtype)
        elif kind == "convT":
            p[f"{scope}/weights"] = O.xavier_uniform(rng, (2, 2, cout, cin), 4 * cout, 4 * cin, dtype)
            p[f"{scope}/biases"] = np.zeros(cout, dtype)
        else:
            p[f"{scope}/weights"] = O.xavier_uniform(rng, (1, 1, cin, cout), cin, cout, dtype)
            p[f"{scope}/biases"] = np.zeros(cout, dtype)
    return p


def trainable_names(cfg: UNetCfg, params: dict):
    return [k for k in params if not k.endswith(("moving_mean", "moving_variance"))]


def regularized_names(cfg: UNetCfg, params: dict):
    """Variables that carry slim.l2_regularizer: conv / convT / logits weights, and biases unless --bias_decay."""
    out = []
    for k in trainable_names(cfg, params):
        if k.endswith("/weights") or (k.endswith("/biases") and not cfg.bias_decay):
            out.append(k)
    return out


def _identity(a):
    return a


@dataclass
class Tape:
    logits: np.ndarray = None
    prob: np.ndarray = None
    acts: dict = field(default_factory=dict)
    new_moving: dict = field(default_factory=dict)
    layers: list = field(default_factory=list)


def forward(params: dict, images: np.ndarray, cfg: UNetCfg, is_training: bool, rnd=_identity, wrnd=None,
            stem_fp32: bool = True) -> Tape:
    """UNet._build_network. `wrnd` rounds the weights the convolutions read (bf16 shadows); the stem and the
    logits layer read the fp32 masters in the engine, hence `stem_fp32`."""
    wrnd = wrnd or rnd
    ns = norm_scope(cfg)
    tape = Tape()
    dt = images.dtype

    def W(scope, master=False):
        w = params[f"{scope}/weights"].astype(dt)
        return w if master else wrnd(w).astype(dt)

    def conv_block(x, scope, first=False):
        w = W(scope, master=first and stem_fp32)
        y = rnd(O.conv2d(x, w)).astype(dt)
        g, b = params[f"{scope}/{ns}/gamma"].astype(dt), params[f"{scope}/{ns}/beta"].astype(dt)
        if cfg.normalizer == "batch_norm":
            mm, mv = params[f"{scope}/{ns}/moving_mean"].astype(dt), params[f"{scope}/{ns}/moving_variance"].astype(dt)
            if is_training:
                z, cache, nmm, nmv = O.batch_norm_train(y, g, b, mm, mv, cfg.bn_eps, cfg.bn_decay)
                tape.new_moving[f"{scope}/{ns}/moving_mean"] = nmm
                tape.new_moving[f"{scope}/{ns}/moving_variance"] = nmv
            else:
                z, cache = O.batch_norm_infer(y, g, b, mm, mv, cfg.bn_eps), None
        else:
            z, cache = O.instance_norm(y, g, b, cfg.in_eps)
        a = rnd(O.relu(z)).astype(dt)
        tape.layers.append(dict(kind="conv", scope=scope, x=x, w=w, z=z, a=a, cache=cache, first=first))
        return a

    x = images
    skips = []
    first = True
    for i in range(cfg.num_down_samples):
        for j in (1, 2):
            x = conv_block(x, f"UNet/Encode{i + 1}/Repeat/convolution2d_{j}", first)
            first = False
        skips.append(x)
        pooled = O.max_pool_2x2(x)
        tape.layers.append(dict(kind="pool", x=x))
        x = pooled
    for j in (1, 2):
        x = conv_block(x, f"UNet/ED-Bridge/ED-Bridge_{j}")
    for i in reversed(range(cfg.num_down_samples)):
        scope = f"UNet/Decode{i + 1}/Conv2d_transpose"
        w = W(scope)
        up = rnd(O.relu(O.conv2d_transpose(x, w) + params[f"{scope}/biases"].astype(dt))).astype(dt)
        tape.layers.append(dict(kind="convT", scope=scope, x=x, w=w, a=up))
        x = np.concatenate((skips[i], up), axis=-1)
        tape.layers.append(dict(kind="concat", split=skips[i].shape[-1], level=i))
        for j in (1, 2):
            x = conv_block(x, f"UNet/Decode{i + 1}/Repeat/convolution2d_{j}")
    scope = "UNet/AdjustChannels"
    w = W(scope, master=True)
    logits = O.conv2d(x, w) + params[f"{scope}/biases"].astype(dt)
    tape.layers.append(dict(kind="logits", scope=scope, x=x, w=w))
    tape.logits = logits
    tape.prob = O.softmax(logits)
    return tape


def tape_from_stored(params: dict, images: np.ndarray, stored: dict, logits: np.ndarray, cfg: UNetCfg,
                     wrnd=_identity) -> Tape:
    """Rebuilds the backward tape from forward tensors that were STORED by another implementation
    (`stored[scope] = {"y": pre-norm conv output, "a": activation}`), so that a backward pass can be
    compared op by op on identical inputs: ReLU masks and max-pool arg-maxes then come from the same
    bits on both sides instead of diverging chaotically with the last-ulp differences of the forward."""
    ns = norm_scope(cfg)
    tape = Tape()
    dt = images.dtype

    def block(x, scope, first=False):
        w = params[f"{scope}/weights"].astype(dt)
        w = w if first else wrnd(w).astype(dt)
        y = stored[scope]["y"].astype(dt)
        g, b = params[f"{scope}/{ns}/gamma"].astype(dt), params[f"{scope}/{ns}/beta"].astype(dt)
        if cfg.normalizer == "batch_norm":
            z, cache, _, _ = O.batch_norm_train(y, g, b, np.zeros_like(g), np.ones_like(g), cfg.bn_eps, cfg.bn_decay)
        else:
            z, cache = O.instance_norm(y, g, b, cfg.in_eps)
        a = stored[scope]["a"].astype(dt)
        tape.layers.append(dict(kind="conv", scope=scope, x=x, w=w, z=z, a=a, cache=cache, first=first))
        return a

    x = images
    skips = []
    first = True
    for i in range(cfg.num_down_samples):
        for j in (1, 2):
            x = block(x, f"UNet/Encode{i + 1}/Repeat/convolution2d_{j}", first)
            first = False
        skips.append(x)
        tape.layers.append(dict(kind="pool", x=x))
        x = O.max_pool_2x2(x)
    for j in (1, 2):
        x = block(x, f"UNet/ED-Bridge/ED-Bridge_{j}")
    for i in reversed(range(cfg.num_down_samples)):
        scope = f"UNet/Decode{i + 1}/Conv2d_transpose"
        w = wrnd(params[f"{scope}/weights"].astype(dt)).astype(dt)
        up = stored[scope]["a"].astype(dt)
        tape.layers.append(dict(kind="convT", scope=scope, x=x, w=w, a=up))
        x = np.concatenate((skips[i], up), axis=-1)
        tape.layers.append(dict(kind="concat", split=skips[i].shape[-1], level=i))
        for j in (1, 2):
            x = block(x, f"UNet/Decode{i + 1}/Repeat/convolution2d_{j}")
    scope = "UNet/AdjustChannels"
    tape.layers.append(dict(kind="logits", scope=scope, x=x, w=params[f"{scope}/weights"].astype(dt)))
    tape.logits = logits.astype(dt)
    tape.prob = O.softmax(tape.logits)
    return tape


def layerwise_forward_errors(params: dict, images: np.ndarray, stored: dict, logits: np.ndarray, cfg: UNetCfg,
                             is_training: bool, wrnd=_identity) -> dict:
    """Op-by-op forward parity on identical inputs: every layer of UNet._build_network is evaluated by the
    oracle on the INPUT TENSOR THE OTHER IMPLEMENTATION STORED, and its result is compared with what that
    implementation stored as the output. Returns {"<scope>:y" | "<scope>:a" | "logits": relative L2 error}.
    (A free-running comparison of the final logits measures the chaos of 23 stacked bf16 roundings, not the
    kernels; this is the per-op statement the bf16 tolerance of 1e-2 applies to.)"""
    ns = norm_scope(cfg)
    dt = np.float64
    errs = {}

    def rel(a, b):
        a, b = np.asarray(a, dt), np.asarray(b, dt)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    def block(x, scope, first=False):
        w = params[f"{scope}/weights"].astype(dt)
        w = w if first else wrnd(w).astype(dt)
        y_ref = O.conv2d(x.astype(dt), w)
        y = stored[scope]["y"].astype(dt)
        errs[f"{scope}:y"] = rel(y, y_ref)
        g, b = params[f"{scope}/{ns}/gamma"].astype(dt), params[f"{scope}/{ns}/beta"].astype(dt)
        if cfg.normalizer == "batch_norm":
            mm, mv = params[f"{scope}/{ns}/moving_mean"].astype(dt), params[f"{scope}/{ns}/moving_variance"].astype(dt)
            if is_training:
                z = O.batch_norm_train(y, g, b, mm, mv, cfg.bn_eps, cfg.bn_decay)[0]
            else:
                z = O.batch_norm_infer(y, g, b, mm, mv, cfg.bn_eps)
        else:
            z = O.instance_norm(y, g, b, cfg.in_eps)[0]
        a = stored[scope]["a"].astype(dt)
        errs[f"{scope}:a"] = rel(a, O.relu(z))
        return a

    x = images.astype(dt)
    skips = []
    first = True
    for i in range(cfg.num_down_samples):
        for j in (1, 2):
            x = block(x, f"UNet/Encode{i + 1}/Repeat/convolution2d_{j}", first)
            first = False
        skips.append(x)
        x = O.max_pool_2x2(x)
    for j in (1, 2):
        x = block(x, f"UNet/ED-Bridge/ED-Bridge_{j}")
    for i in reversed(range(cfg.num_down_samples)):
        scope = f"UNet/Decode{i + 1}/Conv2d_transpose"
        w = wrnd(params[f"{scope}/weights"].astype(dt)).astype(dt)
        up = stored[scope]["a"].astype(dt)
        errs[f"{scope}:a"] = rel(up, O.relu(O.conv2d_transpose(x, w) + params[f"{scope}/biases"].astype(dt)))
        x = np.concatenate((skips[i], up), axis=-1)
        for j in (1, 2):
            x = block(x, f"UNet/Decode{i + 1}/Repeat/convolution2d_{j}")
    scope = "UNet/AdjustChannels"
    ref = O.conv2d(x, params[f"{scope}/weights"].astype(dt)) + params[f"{scope}/biases"].astype(dt)
    errs["logits"] = rel(logits, ref)
    return errs


def loss_and_dlogits(tape: Tape, labels: np.ndarray, cfg: UNetCfg, loss_scale: float = 1.0):
    """UNet._build_loss (data term only). Returns (loss, dlogits * loss_scale)."""
    kw = {}
    if cfg.loss_weight_type == "numerical":
        kw["numeric_w"] = cfg.loss_numeric_w
    elif cfg.loss_weight_type == "proportion" and cfg.loss_proportion_decay > 0:
        kw["proportion_decay"] = cfg.loss_proportion_decay
    if cfg.loss_type == "xentropy":
        loss, dl = O.weighted_sparse_softmax_cross_entropy(tape.logits, labels, cfg.loss_weight_type, **kw)
    elif cfg.loss_type == "dice":
        loss, dp = O.sparse_dice_loss(tape.prob, labels)
        dl = O.softmax_grad(dp, tape.prob)
    else:
        raise ValueError("Not supported loss_type: {}".format(cfg.loss_type))
    return loss, dl * tape.logits.dtype.type(loss_scale)


def regularization_loss(params: dict, cfg: UNetCfg) -> float:
    if cfg.weight_decay_rate <= 0:
        return 0.0
    return sum(O.l2_regularizer(params[k], cfg.weight_decay_rate) for k in regularized_names(cfg, params))


def backward(tape: Tape, dlogits: np.ndarray, cfg: UNetCfg, rnd=_identity) -> dict:
    """tf.gradients of the data loss w.r.t. every trainable variable (no L2 term: see total_grads)."""
    ns = norm_scope(cfg)
    grads = {}
    dt = dlogits.dtype
    d = dlogits
    skip_grads = {}
    for L in reversed(tape.layers):
        k = L["kind"]
        if k == "logits":
            grads[f"{L['scope']}/weights"] = O.conv2d_backprop_filter(L["x"], L["w"].shape, d)
            grads[f"{L['scope']}/biases"] = d.sum(axis=(0, 1, 2))
            d = rnd(O.conv2d_backprop_input(L["x"].shape, L["w"], d)).astype(dt)
        elif k == "conv":
            dz = O.relu_grad(d, L["a"])   # a > 0 <=> z > 0; with a stored tape the mask is the other side's bits
            if cfg.normalizer == "batch_norm":
                dy, dg, db = O.batch_norm_grad(dz, L["cache"])
            else:
                dy, dg, db = O.instance_norm_grad(dz, L["cache"])
            dy = rnd(dy).astype(dt)
            grads[f"{L['scope']}/{ns}/gamma"] = dg
            grads[f"{L['scope']}/{ns}/beta"] = db
            grads[f"{L['scope']}/weights"] = O.conv2d_backprop_filter(L["x"], L["w"].shape, dy)
            d = None if L["first"] else rnd(O.conv2d_backprop_input(L["x"].shape, L["w"], dy)).astype(dt)
        elif k == "concat":
            skip_grads[L["level"]] = d[..., :L["split"]]
            d = d[..., L["split"]:]
        elif k == "convT":
            dyr = rnd(O.relu_grad(d, L["a"])).astype(dt)
            dx, dw = O.conv2d_transpose_grad(L["x"], L["w"], dyr)
            grads[f"{L['scope']}/weights"] = dw
            grads[f"{L['scope']}/biases"] = dyr.sum(axis=(0, 1, 2))
            d = rnd(dx).astype(dt)
        elif k == "pool":
            level = max(skip_grads)  # deepest pending skip belongs to this pool's input
            d = rnd(O.max_pool_2x2_grad(L["x"], d) + skip_grads.pop(level)).astype(dt)
    return grads


def total_grads(params: dict, data_grads: dict, cfg: UNetCfg) -> dict:
    """Gradient of total_loss = data loss + slim L2 terms: adds rate * w on regularised variables."""
    g = dict(data_grads)
    if cfg.weight_decay_rate > 0:
        for k in regularized_names(cfg, params):
            g[k] = g[k] + cfg.weight_decay_rate * params[k].astype(g[k].dtype)
    return g


def train_step(params: dict, slots: dict, step: int, images, labels, cfg: UNetCfg, lr: float, rnd=_identity,
               wrnd=None, optimizer: str = "adam"):
    """One `sess.run([train_op, loss])` (/root/reference/core/estimator.py:756-757). Mutates params / slots.

    Returns (total_loss, tape, grads). `step` is the 1-based Adam step (global_step + 1)."""
    tape = forward(params, images, cfg, True, rnd, wrnd)
    data_loss, dl = loss_and_dlogits(tape, labels, cfg)
    total = float(data_loss) + regularization_loss(params, cfg)
    grads = total_grads(params, backward(tape, dl, cfg, rnd), cfg)
    for k, g in grads.items():
        w = params[k].astype(np.float64)
        if optimizer == "adam":
            m, v = slots.setdefault(k, (np.zeros_like(w), np.zeros_like(w)))
            w, m, v = O.adam_step(w, g.astype(np.float64), m, v, step, lr)
            slots[k] = (m, v)
        else:
            (acc,) = slots.setdefault(k, (np.zeros_like(w),))
            w, acc = O.momentum_step(w, g.astype(np.float64), acc, lr)
            slots[k] = (acc,)
        params[k] = w.astype(params[k].dtype)
    for k, v in tape.new_moving.items():  # UPDATE_OPS run under control_dependencies (solver.py:236-239)
        params[k] = v.astype(params[k].dtype)
    return total, tape, grads


def replica_grads(params: dict, images, labels, cfg: UNetCfg, replicas: int, rnd=_identity, wrnd=None):
    """What ONE replica contributes under MirroredStrategy (/root/reference/core/estimator.py:570-613,
    utils/distribution_utils.py:85-98): gradients of (data loss / R) on its own per-device batch -- batch-norm
    statistics are per replica, there is no sync-BN -- plus its new moving statistics and its data loss."""
    tape = forward(params, images, cfg, True, rnd, wrnd)
    data_loss, dl = loss_and_dlogits(tape, labels, cfg, loss_scale=1.0 / replicas)
    return backward(tape, dl, cfg, rnd), tape.new_moving, float(data_loss)


def mirrored_apply(params: dict, slots: dict, step: int, grad_sum: dict, moving_mean: dict, cfg: UNetCfg, lr: float):
    """The identical update every mirror applies after the all-reduce: SUM of the replica gradients (= gradient of
    the mean loss) + the L2 term, Adam; moving statistics are the cross-replica MEAN of the replica updates."""
    grads = total_grads(params, grad_sum, cfg)
    for k, g in grads.items():
        w = params[k].astype(np.float64)
        m, v = slots.setdefault(k, (np.zeros_like(w), np.zeros_like(w)))
        w, m, v = O.adam_step(w, g.astype(np.float64), m, v, step, lr)
        slots[k] = (m, v)
        params[k] = w.astype(params[k].dtype)
    for k, v in moving_mean.items():
        params[k] = np.asarray(v).astype(params[k].dtype)


def mirrored_train_step(params: dict, slots: dict, step: int, shards, cfg: UNetCfg, lr: float):
    """R virtual replicas in one process, gradients summed in rank order. `shards` = [(images, labels)] per replica.
    Returns the MEAN of the replica data losses + the regularisation loss (estimator.py:576-577)."""
    r = len(shards)
    reg = regularization_loss(params, cfg)
    outs = [replica_grads(params, im, lb, cfg, r) for im, lb in shards]
    gsum = {k: sum(o[0][k].astype(np.float64) for o in outs) for k in outs[0][0]}
    mmean = {k: sum(np.asarray(o[1][k], np.float64) for o in outs) / r for k in outs[0][1]}
    mirrored_apply(params, slots, step, gsum, mmean, cfg, lr)
    return sum(o[2] for o in outs) / r + reg


def predictions(tape: Tape, cfg: UNetCfg):
    """`<Cls>Pred` uint8 masks (UNet.py:112-117)."""
    masks = O.threshold_masks(tape.prob)
    return {cfg.classes[i + 1] + "Pred": m for i, m in enumerate(masks)}


def metrics(tape: Tape, labels, cfg: UNetCfg, names=("Dice",)):
    """UNet._build_metrics: {"<Cls>/<Metric>": value} on the threshold masks."""
    out = {}
    preds = predictions(tape, cfg)
    for i in range(1, cfg.num_classes):
        p = preds[cfg.classes[i] + "Pred"]
        for m in names:
            fn = {"dice": O.metric_dice, "voe": O.metric_voe, "vd": O.metric_vd}[m.lower()]
            out[f"{cfg.classes[i]}/{m}"] = fn(p, labels, i)
    return out
