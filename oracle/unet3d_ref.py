"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/tf_ops.py header; parity unpinned by the reference).

numpy restatement of one training / inference step of the reference's 3-D U-Net:
  layer table  /root/reference/NetworksV2/UNet3D.py:31-91    (_ModelConfig: kernels and strides per block)
  graph        /root/reference/NetworksV2/UNet3D.py:123-186  (conv3d + norm + ReLU, strided-conv down-sampling,
               conv3d_transpose WITHOUT bias + ReLU, skip concat [encoder, up], 1x1x1 logits with bias)
  loss         /root/reference/NetworksV2/UNet3D.py:188-202  (weighted cross-entropy only)
Variable names are slim's (UNet3D/conv_e0/conv1/weights, UNet3D/conv_d3/up/weights, UNet3D/logits/biases ...).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import tf_ops as O


def model_config(num_pool_layers: int):
    """UNet3D._ModelConfig.config[num_pool_layers] as an ordered list of (block, layer, kernel, stride)."""
    if num_pool_layers not in (4, 5):
        raise KeyError(num_pool_layers)
    np_ = num_pool_layers
    out = []
    for i in range(np_):
        k = (1, 3, 3) if i < 2 else (3, 3, 3)
        out.append((f"conv_e{i}", "conv1", k, (1, 1, 1) if i == 0 else (1, 2, 2)))
        out.append((f"conv_e{i}", "conv2", k, (1, 1, 1)))
    out.append(("bridge", "conv1", (3, 3, 3), (2, 2, 2)))
    out.append(("bridge", "conv2", (3, 3, 3), (1, 1, 1)))
    for i in reversed(range(np_)):
        up = (2, 2, 2) if i == np_ - 1 else (1, 2, 2)
        k = (1, 3, 3) if i < 2 else (3, 3, 3)
        out.append((f"conv_d{i}", "up", up, up))
        out.append((f"conv_d{i}", "conv1", k, (1, 1, 1)))
        out.append((f"conv_d{i}", "conv2", k, (1, 1, 1)))
    return out


@dataclass
class UNet3DCfg:
    depth: int = 64
    height: int = 128
    width: int = 128
    channel: int = 1
    classes: tuple = ("Background", "NF")
    init_channels: int = 30          # NetworksV2/UNet3D.yml
    max_channels: int = 320
    num_pool_layers: int = 4
    use_spatial: bool = False        # concat sp_guide to the images (UNet3D.py:142-144)
    guide_channel: int = 2
    normalizer: str = "instance_norm"
    weight_decay_rate: float = 3e-5
    bias_decay: bool = False
    loss_type: str = "xentropy"
    loss_weight_type: str = "numerical"
    loss_numeric_w: tuple = (1.0, 1.0)
    loss_proportion_decay: float = 1000.0
    in_eps: float = 1e-6

    @property
    def num_classes(self):
        return len(self.classes)

    @property
    def in_channels(self):
        return self.channel + (self.guide_channel if self.use_spatial else 0)


def layer_specs(cfg: UNet3DCfg):
    """[(kind, scope, cin, cout, kernel, stride, in_dhw)] in graph order; kind in conv | convT | logits. Decoder conv1
    layers read the concat [encoder features (c), up-sampled (c)]."""
    if cfg.normalizer != "instance_norm":
        raise NotImplementedError("UNet3D oracle: instance_norm (what the shipped 3-D scripts use)")
    specs = []
    c = cfg.init_channels
    cin = cfg.in_channels
    dhw = (cfg.depth, cfg.height, cfg.width)
    enc = {}
    for block, layer, k, s in model_config(cfg.num_pool_layers):
        scope = f"UNet3D/{block}/{layer}"
        if block.startswith("conv_e") or block == "bridge":
            specs.append(dict(kind="conv", scope=scope, cin=cin, cout=c, k=k, s=s, dhw=dhw, block=block, layer=layer))
            dhw = tuple(-(-dhw[i] // s[i]) for i in range(3))
            cin = c
            if layer == "conv2":
                enc[block] = dict(c=c, dhw=dhw)
                c = min(c * 2, cfg.max_channels)
        elif layer == "up":
            e = enc[block.replace("d", "e")]
            c = e["c"]
            specs.append(dict(kind="convT", scope=scope, cin=cin, cout=c, k=k, s=s, dhw=dhw, block=block, layer=layer))
            dhw = tuple(dhw[i] * s[i] for i in range(3))
            assert dhw == e["dhw"], (dhw, e["dhw"])
            cin = 2 * c
        else:
            specs.append(dict(kind="conv", scope=scope, cin=cin, cout=c, k=k, s=s, dhw=dhw, block=block, layer=layer))
            cin = c
    specs.append(dict(kind="logits", scope="UNet3D/logits", cin=cin, cout=cfg.num_classes, k=(1, 1, 1), s=(1, 1, 1),
                      dhw=dhw, block="logits", layer="logits"))
    return specs


def init_params(cfg: UNet3DCfg, seed: int = 0, dtype=np.float32) -> dict:
    rng = np.random.default_rng(seed)
    p = {}
    for s in layer_specs(cfg):
        sc, cin, cout, k = s["scope"], s["cin"], s["cout"], s["k"]
        rf = k[0] * k[1] * k[2]
        if s["kind"] == "conv":
            p[f"{sc}/weights"] = O.xavier_uniform(rng, k + (cin, cout), rf * cin, rf * cout, dtype)
            p[f"{sc}/InstanceNorm/gamma"] = np.ones(cout, dtype)
            p[f"{sc}/InstanceNorm/beta"] = np.zeros(cout, dtype)
        elif s["kind"] == "convT":
            p[f"{sc}/weights"] = O.xavier_uniform(rng, k + (cout, cin), rf * cout, rf * cin, dtype)   # no biases
        else:
            p[f"{sc}/weights"] = O.xavier_uniform(rng, k + (cin, cout), cin, cout, dtype)
            p[f"{sc}/biases"] = np.zeros(cout, dtype)
    return p


def regularized_names(cfg: UNet3DCfg, params: dict):
    return [k for k in params if k.endswith("/weights") or (k.endswith("/biases") and not cfg.bias_decay)]


def regularization_loss(params: dict, cfg: UNet3DCfg) -> float:
    if cfg.weight_decay_rate <= 0:
        return 0.0
    return sum(O.l2_regularizer(params[k], cfg.weight_decay_rate) for k in regularized_names(cfg, params))


def _identity(a):
    return a


@dataclass
class Tape:
    logits: np.ndarray = None
    prob: np.ndarray = None
    layers: list = field(default_factory=list)
    errs: dict = field(default_factory=dict)


def forward(params: dict, inputs: dict, cfg: UNet3DCfg, rnd=_identity, wrnd=None, stored=None) -> Tape:
    """UNet3D._build_network. inputs: images [n,d,h,w,c] (+ sp_guide [n,d,h,w,g] when cfg.use_spatial).
    `stored` as in oracle/gunet_ref.forward: per-layer errors on the stored inputs + a tape of the stored tensors."""
    wrnd = wrnd or rnd
    tape = Tape()
    x = inputs["images"]
    if cfg.use_spatial:
        x = np.concatenate((inputs["images"], inputs["sp_guide"]), axis=-1)
    dt = x.dtype

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    skips = {}
    first = True
    for s in layer_specs(cfg):
        sc = s["scope"]
        if s["kind"] == "conv":
            w = params[f"{sc}/weights"].astype(dt)
            w = w if first else wrnd(w).astype(dt)
            y = rnd(O.conv3d(x, w, s["s"])).astype(dt)
            if stored is not None:
                tape.errs[f"{sc}:y"] = rel(stored[sc]["y"], y)
                y = stored[sc]["y"].astype(dt)
            z, cache = O.instance_norm(y, params[f"{sc}/InstanceNorm/gamma"].astype(dt),
                                       params[f"{sc}/InstanceNorm/beta"].astype(dt), cfg.in_eps)
            a = rnd(O.relu(z)).astype(dt)
            if stored is not None:
                tape.errs[f"{sc}:a"] = rel(stored[sc]["a"], a)
                a = stored[sc]["a"].astype(dt)
            tape.layers.append(dict(kind="conv", spec=s, x=x, w=w, z=z, a=a, cache=cache, first=first))
            first = False
            x = a
            if s["layer"] == "conv2" and s["block"].startswith("conv_e"):
                skips[s["block"]] = a
                tape.layers.append(dict(kind="fork", block=s["block"]))
        elif s["kind"] == "convT":
            w = wrnd(params[f"{sc}/weights"].astype(dt)).astype(dt)
            up = rnd(O.relu(O.conv3d_transpose(x, w, s["s"]))).astype(dt)
            if stored is not None:
                tape.errs[f"{sc}:a"] = rel(stored[sc]["a"], up)
                up = stored[sc]["a"].astype(dt)
            tape.layers.append(dict(kind="convT", spec=s, x=x, w=w, a=up))
            skip = skips[s["block"].replace("d", "e")]
            x = np.concatenate((skip, up), axis=-1)
            tape.layers.append(dict(kind="concat", split=skip.shape[-1], block=s["block"].replace("d", "e")))
        else:
            w = params[f"{sc}/weights"].astype(dt)
            logits = O.conv3d(x, w) + params[f"{sc}/biases"].astype(dt)
            if stored is not None:
                tape.errs["logits"] = rel(stored["logits"], logits)
                logits = stored["logits"].astype(dt)
            tape.layers.append(dict(kind="logits", spec=s, x=x, w=w))
            tape.logits = logits
            tape.prob = O.softmax(logits)
    return tape


def loss_and_dlogits(tape: Tape, labels: np.ndarray, cfg: UNet3DCfg, loss_scale: float = 1.0):
    if "xentropy" not in cfg.loss_type:
        raise ValueError("Not supported loss_type: {}".format(cfg.loss_type))   # UNet3D.py:198-199
    kw = {}
    if cfg.loss_weight_type == "numerical":
        kw["numeric_w"] = cfg.loss_numeric_w
    elif cfg.loss_weight_type == "proportion" and cfg.loss_proportion_decay > 0:
        kw["proportion_decay"] = cfg.loss_proportion_decay
    loss, dl = O.weighted_sparse_softmax_cross_entropy(tape.logits, labels, cfg.loss_weight_type, **kw)
    return float(loss), dl * tape.logits.dtype.type(loss_scale)


def backward(tape: Tape, dlogits: np.ndarray, cfg: UNet3DCfg, rnd=_identity) -> dict:
    grads = {}
    dt = dlogits.dtype
    d = dlogits
    skip_grads = {}
    for L in reversed(tape.layers):
        k = L["kind"]
        if k == "logits":
            sc = L["spec"]["scope"]
            grads[f"{sc}/weights"] = O.conv3d_backprop_filter(L["x"], L["w"].shape, d)
            grads[f"{sc}/biases"] = d.sum(axis=(0, 1, 2, 3))
            d = rnd(O.conv3d_backprop_input(L["x"].shape, L["w"], d)).astype(dt)
        elif k == "conv":
            s = L["spec"]
            sc = s["scope"]
            dz = O.relu_grad(d, L["a"])   # a > 0 <=> z > 0; with a stored tape the mask is the other side's bits
            dy, dg, db = O.instance_norm_grad(dz, L["cache"])
            dy = rnd(dy).astype(dt)
            grads[f"{sc}/InstanceNorm/gamma"] = dg
            grads[f"{sc}/InstanceNorm/beta"] = db
            grads[f"{sc}/weights"] = O.conv3d_backprop_filter(L["x"], L["w"].shape, dy, s["s"])
            d = None if L["first"] else rnd(O.conv3d_backprop_input(L["x"].shape, L["w"], dy, s["s"])).astype(dt)
        elif k == "concat":
            skip_grads[L["block"]] = d[..., :L["split"]]
            d = d[..., L["split"]:]
        elif k == "convT":
            s = L["spec"]
            dyr = rnd(O.relu_grad(d, L["a"])).astype(dt)
            dx, dw = O.conv3d_transpose_grad(L["x"], L["w"], dyr, s["s"])
            grads[f"{s['scope']}/weights"] = dw
            d = rnd(dx).astype(dt)
        elif k == "fork":   # the encoder output feeds the skip AND the next block: AddN of the two gradients
            d = rnd(d + skip_grads.pop(L["block"])).astype(dt)
    return grads


def total_grads(params: dict, data_grads: dict, cfg: UNet3DCfg) -> dict:
    g = dict(data_grads)
    if cfg.weight_decay_rate > 0:
        for k in regularized_names(cfg, params):
            g[k] = g[k] + cfg.weight_decay_rate * params[k].astype(g[k].dtype)
    return g


def train_step(params: dict, slots: dict, step: int, inputs: dict, labels, cfg: UNet3DCfg, lr: float, rnd=_identity,
               wrnd=None):
    tape = forward(params, inputs, cfg, rnd, wrnd)
    data_loss, dl = loss_and_dlogits(tape, labels, cfg)
    total = float(data_loss) + regularization_loss(params, cfg)
    grads = total_grads(params, backward(tape, dl, cfg, rnd), cfg)
    for k, g in grads.items():
        w = params[k].astype(np.float64)
        m, v = slots.setdefault(k, (np.zeros_like(w), np.zeros_like(w)))
        w, m, v = O.adam_step(w, g.astype(np.float64), m, v, step, lr)
        slots[k] = (m, v)
        params[k] = w.astype(params[k].dtype)
    return total, tape, grads
