"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/tf_ops.py header; parity unpinned by the reference).

numpy restatement of one training / inference step of the reference's guided U-Net:
  graph       /root/reference/NetworksV2/GUNet.py:259-392   (GUNet._build_network)
  modulation  /root/reference/NetworksV2/GUNet.py:162-217   (modulated_conv_block: conv -> norm -> *gamma -> +s -> relu)
  context MLP /root/reference/NetworksV2/GUNet.py:31-59     + Backbone/slim_nets.py:34-57 (fc = mlp, dropout after
              every hidden layer with keep_prob = 1 - side_dropout, final layer he_normal and linear)
  guide convs /root/reference/NetworksV2/GUNet.py:136-159   (1x1 convs on the 2x2-average-pooled guide pyramid)
  loss        /root/reference/NetworksV2/GUNet.py:394-413   ("xentropy" and / or "dice": both are added when both appear)

Scope: normalizer = instance_norm (every shipped GUNet script passes --normalizer instance_norm), context_model "fc",
no --use_se / --fix / --without_norm / --dropout / after_affine / --img_grad / ct_conv: those raise NotImplementedError.
Variable names are slim's, so a dict here is interchangeable with a TF checkpoint's variable map.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import tf_ops as O


@dataclass
class GUNetCfg:
    height: int = 512
    width: int = 512
    channel: int = 3
    classes: tuple = ("Background", "Liver", "Tumor")
    init_channels: int = 64              # NetworksV2/GUNet.yml
    num_down_samples: int = 4
    mod_layers: tuple = (1, 2, 3, 4)
    context_fc_channels: tuple = (256, 256)
    norm_with_center: bool = True        # GUNet.yml; ext_config/GUNet_BOTH.yml has False
    norm_with_scale: bool = False
    after_affine: bool = False           # ext_config/GUNet*_AA.yml, GUNetV2.yml, GUNet_BOTHV2.yml: slim_nets.affine
                                         # (channel_wise_affine, Backbone/slim_nets.py:152-212) before every encoder ReLU
    use_context: bool = True             # --use_context
    use_spatial: bool = True             # --use_spatial
    guide_channel: int = 1               # --guide_channel
    context_dim: int = 200               # 2 x 100-bin histograms (DataLoader/Liver/input_pipeline_g.py:374-394)
    side_dropout: float = 0.5            # --side_dropout
    dropout_seed: int = 0
    normalizer: str = "instance_norm"
    weight_decay_rate: float = 1e-5
    bias_decay: bool = False
    loss_type: str = "xentropy"          # any string containing "xentropy" and / or "dice"
    loss_weight_type: str = "none"
    loss_numeric_w: tuple = ()
    loss_proportion_decay: float = 1000.0
    in_eps: float = 1e-6
    bn_eps: float = 1e-3                 # slim.batch_norm default epsilon
    prefix: str = "GUNet"                # "UNetInter": same variable layout, no modulation, guide concatenated to the
                                         # images at the input (/root/reference/NetworksV2/UNetInter.py:89-92,118-146)
    dropout: float = 0.0                 # backbone --dropout (GUNet.py:189-190): slim.dropout behind the normaliser of the
                                         # FIRST conv of every encoder block, in front of the modulation; training only
    img_grad: bool = False               # --img_grad (GUNet.py:333-337): the network input is concat(images, dy, dx) of
                                         # tf.image.image_gradients, 3 x channel input channels
    mid_cat: bool = False                # UNetInter --mid_cat (UNetInter.py:87-92,124-125): the guide is NOT an input channel;
                                         # it is concatenated to the first block's output in front of the first max-pool

    @property
    def num_classes(self):
        return len(self.classes)

    @property
    def n_modulator_param(self):
        return self.init_channels * sum(2 ** i for i in range(self.num_down_samples + 1) if i in self.mod_layers) * 2


def unetinter_cfg(channel: int = 3, guide_channel: int = 2, mid_cat: bool = False, **kw) -> GUNetCfg:
    """UNetInter (/root/reference/NetworksV2/UNetInter.py:73-146) expressed on the GUNet restatement: scope root
    "UNetInter", no modulated block, every conv followed by norm(center, scale) + ReLU (encoder_arg_scope, :100-117),
    and `channel + guide_channel` network input channels (the guide is concatenated to the images, :89-90) -- or, with
    --mid_cat, `channel` input channels and the guide concatenated in front of the first max-pool (:124-125)."""
    return GUNetCfg(channel=channel if mid_cat else channel + guide_channel, guide_channel=guide_channel,
                    prefix="UNetInter", use_context=False, use_spatial=False, mod_layers=(), mid_cat=mid_cat, **kw)


def unetinter_inputs(images: np.ndarray, sp_guide: np.ndarray, mid_cat: bool = False) -> dict:
    """tf.concat((images, sp_guide), axis=-1) -- UNetInter.py:90; with --mid_cat the guide stays a separate input."""
    if mid_cat:
        return dict(images=images, sp_guide=sp_guide)
    return dict(images=np.concatenate((images, sp_guide), axis=-1))


def layer_specs(cfg: GUNetCfg):
    """Ordered conv-type layers of GUNet._build_network as dicts (kind, scope, cin, cout, level, and for convs:
    center / scale of the normaliser, mod (modulated block), mod_off (column in the context vector), sp_off)."""
    if cfg.normalizer not in ("instance_norm", "batch_norm"):
        raise ValueError("Not supported normalization function: " + cfg.normalizer)
    bn = cfg.normalizer == "batch_norm"
    # batch_norm (GUNet.py:301,321-325; UNetInter.py:107-111): the encoder arg scope gives its convs decay 0.99; GUNet's
    # un-modulated encoder blocks override the params with {"scale": True, "is_training"} (slim default decay 0.999), as
    # does every decoder conv (BaseNet._get_normalization, base.py:153-162)
    enc_decay_all = cfg.prefix == "UNetInter"
    specs = []
    c, cin = cfg.init_channels, cfg.channel * (3 if cfg.img_grad and cfg.prefix == "GUNet" else 1)
    off = 0
    for i in range(cfg.num_down_samples + 1):
        mod = i in cfg.mod_layers and (cfg.use_context or cfg.use_spatial)
        for j in (1, 2):
            # encoder_arg_scope (GUNet.py:313-330): the modulated blocks' normaliser takes center / scale from the YAML,
            # both forced off by after_affine; un-modulated blocks pass normalizer_params={} (slim defaults: both on)
            aa = cfg.after_affine
            s = dict(kind="conv", scope=f"{cfg.prefix}/Encode/down_conv{i + 1}/mod_conv{j}/Conv", cin=cin, cout=c, level=i,
                     role=f"enc{j}", mod=mod, center=(cfg.norm_with_center and not aa) if mod else True,
                     scale=(cfg.norm_with_scale and not aa) if mod else True, mod_off=None, sp_off=None,
                     affine=f"{cfg.prefix}/Encode/down_conv{i + 1}/mod_conv{j}/ChannelWiseAffine" if aa else None,
                     decay=0.99 if (mod or enc_decay_all) else 0.999, drop=(j == 1))
            if mod and cfg.use_context:
                s["mod_off"] = off
                off += c
            if mod and cfg.use_spatial:
                s["sp_off"] = (j - 1) * c
            specs.append(s)
            cin = c
        if i == 0 and cfg.mid_cat:
            cin = c + cfg.guide_channel         # the pooled concat(block output, sp_guide) feeds the second block
        if i < cfg.num_down_samples:
            c *= 2
    for i in reversed(range(cfg.num_down_samples)):
        c //= 2
        specs.append(dict(kind="convT", scope=f"{cfg.prefix}/Decode/up{i + 1}", cin=cin, cout=cin // 2, level=i))
        for j in (1, 2):
            specs.append(dict(kind="conv", scope=f"{cfg.prefix}/Decode/up_conv{i + 1}/up_conv{i + 1}_{j}",
                              cin=c + cin // 2 if j == 1 else c, cout=c, level=i, role=f"dec{j}", mod=False, center=True,
                              scale=True, mod_off=None, sp_off=None, affine=None, decay=0.999))
        cin = c
    specs.append(dict(kind="logits", scope=f"{cfg.prefix}/AdjustChannels", cin=cin, cout=cfg.num_classes, level=0))
    return specs


def fc_specs(cfg: GUNetCfg):
    """[(scope, cin, cout, hidden)] of slim_nets.fc(context, context_fc_channels + [n_modulator_param])."""
    chans = list(cfg.context_fc_channels) + [cfg.n_modulator_param]
    out, cin = [], cfg.context_dim
    for k, co in enumerate(chans):
        out.append((f"GUNet/context/fc{k + 1}", cin, co, k < len(chans) - 1))
        cin = co
    return out


def norm_scope(cfg: GUNetCfg) -> str:
    return "BatchNorm" if cfg.normalizer == "batch_norm" else "InstanceNorm"


def init_params(cfg: GUNetCfg, seed: int = 0, dtype=np.float32) -> dict:
    rng = np.random.default_rng(seed)
    p = {}
    ns = norm_scope(cfg)
    for s in layer_specs(cfg):
        sc, cin, cout = s["scope"], s["cin"], s["cout"]
        if s["kind"] == "conv":
            p[f"{sc}/weights"] = O.xavier_uniform(rng, (3, 3, cin, cout), 9 * cin, 9 * cout, dtype)
            if s["scale"]:
                p[f"{sc}/{ns}/gamma"] = np.ones(cout, dtype)
            if s["center"]:
                p[f"{sc}/{ns}/beta"] = np.zeros(cout, dtype)
            if cfg.normalizer == "batch_norm":
                p[f"{sc}/{ns}/moving_mean"] = np.zeros(cout, dtype)
                p[f"{sc}/{ns}/moving_variance"] = np.ones(cout, dtype)
            if s["affine"]:
                p[f"{s['affine']}/gamma"] = np.ones(cout, dtype)
                p[f"{s['affine']}/beta"] = np.zeros(cout, dtype)
        elif s["kind"] == "convT":
            p[f"{sc}/weights"] = O.xavier_uniform(rng, (2, 2, cout, cin), 4 * cout, 4 * cin, dtype)
            p[f"{sc}/biases"] = np.zeros(cout, dtype)
        else:
            p[f"{sc}/weights"] = O.xavier_uniform(rng, (1, 1, cin, cout), cin, cout, dtype)
            p[f"{sc}/biases"] = np.zeros(cout, dtype)
    if cfg.use_context:
        for sc, cin, cout, hidden in fc_specs(cfg):
            p[f"{sc}/weights"] = (O.xavier_uniform(rng, (cin, cout), cin, cout, dtype) if hidden
                                  else O.he_normal(rng, (cin, cout), cin, dtype))
            p[f"{sc}/biases"] = np.zeros(cout, dtype)
    if cfg.use_spatial:
        for i in range(cfg.num_down_samples + 1):
            if i in cfg.mod_layers:
                co = cfg.init_channels * 2 ** (i + 1)
                sc = f"GUNet/spatial/conv{i + 1}"
                p[f"{sc}/weights"] = O.xavier_uniform(rng, (1, 1, cfg.guide_channel, co), cfg.guide_channel, co, dtype)
                p[f"{sc}/biases"] = np.zeros(co, dtype)
    return p


def regularized_names(cfg: GUNetCfg, params: dict):
    """slim.l2_regularizer reaches conv2d / conv2d_transpose variables only (GUNet._net_arg_scope): trunk and guide
    conv weights, and their biases unless --bias_decay. The context MLP (fully_connected) is not regularised."""
    out = []
    for k in params:
        if "/context/" in k or "InstanceNorm" in k or "BatchNorm" in k:
            continue
        if k.endswith("/weights") or (k.endswith("/biases") and not cfg.bias_decay):
            out.append(k)
    return out


def regularization_loss(params: dict, cfg: GUNetCfg) -> float:
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
    fc: list = field(default_factory=list)
    ctx_params: np.ndarray = None
    guides: list = field(default_factory=list)
    errs: dict = field(default_factory=dict)
    new_moving: dict = field(default_factory=dict)


def dropout_offset(step: int, layer: int) -> int:
    """Philox stream offset of hidden layer `layer` (0-based) at 1-based training step `step`."""
    return step * 16 + layer


def forward(params: dict, inputs: dict, cfg: GUNetCfg, is_training: bool, rnd=_identity, wrnd=None, stored=None,
            step: int = 1) -> Tape:
    """GUNet._build_network. inputs: images [n,h,w,c], context [n,context_dim], sp_guide [n,h,w,guide_channel].

    `rnd` / `wrnd` round where the engine stores bf16 (conv outputs and activations / bf16 weight shadows).
    With `stored` ({scope: {"y", "a"}} + "logits", tensors another implementation kept), every layer is evaluated on
    the stored INPUT and its result compared with the stored output (tape.errs, relative L2), and the tape is built
    from the stored tensors so a backward pass runs over identical ReLU masks / pool arg-maxes."""
    wrnd = wrnd or rnd
    tape = Tape()
    images = inputs["images"]
    dt = images.dtype

    def rel(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    # ---- context MLP (fp32 / dt throughout; no bf16 storage on this path)
    ctx_params = None
    if cfg.use_context:
        x = inputs["context"].astype(dt)
        for k, (sc, cin, cout, hidden) in enumerate(fc_specs(cfg)):
            w, b = params[f"{sc}/weights"].astype(dt), params[f"{sc}/biases"].astype(dt)
            pre = O.fully_connected(x, w, b)
            mult = None
            if hidden:
                y = O.relu(pre)
                if is_training and cfg.side_dropout:
                    mult = O.dropout_multipliers(y.size, 1.0 - cfg.side_dropout, cfg.dropout_seed,
                                                 dropout_offset(step, k)).reshape(y.shape).astype(dt)
                    y = y * mult
            else:
                y = pre
            tape.fc.append(dict(scope=sc, x=x, w=w, pre=pre, mult=mult, hidden=hidden))
            x = y
        ctx_params = x
    tape.ctx_params = ctx_params
    # ---- guide pyramid
    if cfg.use_spatial:
        g = inputs["sp_guide"].astype(dt)
        for i in range(cfg.num_down_samples + 1):
            tape.guides.append(g)
            if i < cfg.num_down_samples:
                g = O.avg_pool_2x2(g)

    def conv_block(x, s, first):
        sc, cout = s["scope"], s["cout"]
        w = params[f"{sc}/weights"].astype(dt)
        w = w if first else wrnd(w).astype(dt)
        y = rnd(O.conv2d(x, w)).astype(dt)
        if stored is not None:
            tape.errs[f"{sc}:y"] = rel(stored[sc]["y"], y)
            y = stored[sc]["y"].astype(dt)
        ns = norm_scope(cfg)
        gamma = params[f"{sc}/{ns}/gamma"].astype(dt) if s["scale"] else np.ones(cout, dt)
        beta = params[f"{sc}/{ns}/beta"].astype(dt) if s["center"] else np.zeros(cout, dt)
        if cfg.normalizer == "batch_norm":
            mm, mv = params[f"{sc}/{ns}/moving_mean"].astype(dt), params[f"{sc}/{ns}/moving_variance"].astype(dt)
            if is_training:
                z, cache, nmm, nmv = O.batch_norm_train(y, gamma, beta, mm, mv, cfg.bn_eps, s["decay"])
                tape.new_moving[f"{sc}/{ns}/moving_mean"], tape.new_moving[f"{sc}/{ns}/moving_variance"] = nmm, nmv
            else:
                z, cache = O.batch_norm_infer(y, gamma, beta, mm, mv, cfg.bn_eps), None
        else:
            z, cache = O.instance_norm(y, gamma, beta, cfg.in_eps)
        gm = sp = mult = None
        if cfg.dropout and is_training and s.get("drop") and cfg.prefix == "GUNet":
            # modulated_conv_block: `if i != repeat - 1 and dropout: net = slim.dropout(net, keep_prob=1 - dropout, ...)`.
            # The device keeps the normalised tensor in bf16 and multiplies it in place (two roundings, mirrored by rnd)
            mult = O.dropout_multipliers(z.size, 1.0 - cfg.dropout, cfg.dropout_seed,
                                         dropout_offset(step, 8 + s["level"])).reshape(z.shape).astype(dt)
            z = rnd(rnd(z).astype(dt) * mult).astype(dt)
        zn = z
        if s["mod_off"] is not None:
            gm = ctx_params[:, s["mod_off"]:s["mod_off"] + cout]
            z = z * gm[:, None, None, :]
        if s["sp_off"] is not None:
            ssc = f"GUNet/spatial/conv{s['level'] + 1}"
            wsp = params[f"{ssc}/weights"].astype(dt)[0, 0][:, s["sp_off"]:s["sp_off"] + cout]
            bsp = params[f"{ssc}/biases"].astype(dt)[s["sp_off"]:s["sp_off"] + cout]
            sp = dict(scope=ssc, guide=tape.guides[s["level"]], w=wsp)
            z = z + (sp["guide"] @ wsp + bsp)
        u = z
        if s["affine"]:                      # modulated_conv_block: `if after_affine: net = slim_nets.affine(net)`
            z = z * params[f"{s['affine']}/gamma"].astype(dt) + params[f"{s['affine']}/beta"].astype(dt)
        a = rnd(O.relu(z)).astype(dt)
        if stored is not None:
            tape.errs[f"{sc}:a"] = rel(stored[sc]["a"], a)
            a = stored[sc]["a"].astype(dt)
        tape.layers.append(dict(kind="conv", spec=s, x=x, w=w, z=z, zn=zn, u=u, a=a, cache=cache, first=first, gm=gm,
                                mult=mult, sp=sp, ga=params[f"{s['affine']}/gamma"].astype(dt) if s["affine"] else None))
        return a

    specs = layer_specs(cfg)
    it = iter(specs)
    x = images
    if cfg.img_grad and cfg.prefix == "GUNet":
        # the device differences the fp32 images and stores the 3 x channel input in bf16 (one rounding per value)
        dy, dx = O.image_gradients(images)
        x = rnd(np.concatenate((images, dy, dx), axis=-1)).astype(dt)
    skips = []
    first = True
    for i in range(cfg.num_down_samples + 1):
        for _ in (1, 2):
            x = conv_block(x, next(it), first)
            first = False
        if i < cfg.num_down_samples:
            skips.append(x)
            keep = x.shape[-1]
            if cfg.mid_cat and i == 0:      # UNetInter.py:124-125: only the pooled branch sees the guide
                x = np.concatenate((x, rnd(inputs["sp_guide"].astype(dt)).astype(dt)), axis=-1)
            tape.layers.append(dict(kind="pool", x=x, keep=keep))
            x = O.max_pool_2x2(x)
    for i in reversed(range(cfg.num_down_samples)):
        s = next(it)
        w = wrnd(params[f"{s['scope']}/weights"].astype(dt)).astype(dt)
        up = rnd(O.relu(O.conv2d_transpose(x, w) + params[f"{s['scope']}/biases"].astype(dt))).astype(dt)
        if stored is not None:
            tape.errs[f"{s['scope']}:a"] = rel(stored[s["scope"]]["a"], up)
            up = stored[s["scope"]]["a"].astype(dt)
        tape.layers.append(dict(kind="convT", spec=s, x=x, w=w, a=up))
        x = np.concatenate((skips[i], up), axis=-1)
        tape.layers.append(dict(kind="concat", split=skips[i].shape[-1], level=i))
        for _ in (1, 2):
            x = conv_block(x, next(it), False)
    s = next(it)
    w = params[f"{s['scope']}/weights"].astype(dt)
    logits = O.conv2d(x, w) + params[f"{s['scope']}/biases"].astype(dt)
    if stored is not None:
        tape.errs["logits"] = rel(stored["logits"], logits)
        logits = stored["logits"].astype(dt)
    tape.layers.append(dict(kind="logits", spec=s, x=x, w=w))
    tape.logits = logits
    tape.prob = O.softmax(logits)
    return tape


def loss_and_dlogits(tape: Tape, labels: np.ndarray, cfg: GUNetCfg, loss_scale: float = 1.0):
    """GUNet._build_loss (data terms). Returns (loss, dlogits * loss_scale)."""
    kw = {}
    if cfg.loss_weight_type == "numerical":
        kw["numeric_w"] = cfg.loss_numeric_w
    elif cfg.loss_weight_type == "proportion" and cfg.loss_proportion_decay > 0:
        kw["proportion_decay"] = cfg.loss_proportion_decay
    loss, dl, has = 0.0, 0.0, False
    if "xentropy" in cfg.loss_type:
        l1, d1 = O.weighted_sparse_softmax_cross_entropy(tape.logits, labels, cfg.loss_weight_type, **kw)
        loss, dl, has = loss + float(l1), dl + d1, True
    if "dice" in cfg.loss_type:
        l2, dp = O.sparse_dice_loss(tape.prob, labels)
        loss, dl, has = loss + float(l2), dl + O.softmax_grad(dp, tape.prob), True
    if not has:
        raise ValueError("Not supported loss_type: {}".format(cfg.loss_type))
    return loss, dl * tape.logits.dtype.type(loss_scale)


def backward(tape: Tape, dlogits: np.ndarray, cfg: GUNetCfg, rnd=_identity) -> dict:
    """tf.gradients of the data loss w.r.t. every trainable variable (L2 terms: total_grads)."""
    grads = {}
    dt = dlogits.dtype
    d = dlogits
    skip_grads = {}
    dctx = np.zeros_like(tape.ctx_params) if tape.ctx_params is not None else None
    for L in reversed(tape.layers):
        k = L["kind"]
        if k == "logits":
            sc = L["spec"]["scope"]
            grads[f"{sc}/weights"] = O.conv2d_backprop_filter(L["x"], L["w"].shape, d)
            grads[f"{sc}/biases"] = d.sum(axis=(0, 1, 2))
            d = rnd(O.conv2d_backprop_input(L["x"].shape, L["w"], d)).astype(dt)
        elif k == "conv":
            s = L["spec"]
            sc, cout = s["scope"], s["cout"]
            dz = O.relu_grad(d, L["a"])   # a > 0 <=> z > 0; with a stored tape the mask is the other side's bits
            if L["ga"] is not None:
                grads[f"{s['affine']}/gamma"] = (dz * L["u"]).sum(axis=(0, 1, 2))
                grads[f"{s['affine']}/beta"] = dz.sum(axis=(0, 1, 2))
                dz = dz * L["ga"]
            if L["sp"] is not None:
                ssc, off = L["sp"]["scope"], s["sp_off"]
                gw = np.einsum("nhwg,nhwc->gc", L["sp"]["guide"], dz)
                gfull = grads.setdefault(f"{ssc}/weights", np.zeros((1, 1, gw.shape[0], 2 * cout), dt))
                gfull[0, 0, :, off:off + cout] = gw
                bfull = grads.setdefault(f"{ssc}/biases", np.zeros(2 * cout, dt))
                bfull[off:off + cout] = dz.sum(axis=(0, 1, 2))
            if L["gm"] is not None:
                dctx[:, s["mod_off"]:s["mod_off"] + cout] = (dz * L["zn"]).sum(axis=(1, 2))
                dz = dz * L["gm"][:, None, None, :]
            if L.get("mult") is not None:    # gradient of slim.dropout: the same multipliers (bf16 in, bf16 out on the device)
                dz = rnd(rnd(dz).astype(dt) * L["mult"]).astype(dt)
            ns = norm_scope(cfg)
            if cfg.normalizer == "batch_norm":
                dy, dg, db = O.batch_norm_grad(dz, L["cache"])
            else:
                dy, dg, db = O.instance_norm_grad(dz, L["cache"])
            dy = rnd(dy).astype(dt)
            if s["scale"]:
                grads[f"{sc}/{ns}/gamma"] = dg
            if s["center"]:
                grads[f"{sc}/{ns}/beta"] = db
            grads[f"{sc}/weights"] = O.conv2d_backprop_filter(L["x"], L["w"].shape, dy)
            d = None if L["first"] else rnd(O.conv2d_backprop_input(L["x"].shape, L["w"], dy)).astype(dt)
        elif k == "concat":
            skip_grads[L["level"]] = d[..., :L["split"]]
            d = d[..., L["split"]:]
        elif k == "convT":
            sc = L["spec"]["scope"]
            dyr = rnd(O.relu_grad(d, L["a"])).astype(dt)
            dx, dw = O.conv2d_transpose_grad(L["x"], L["w"], dyr)
            grads[f"{sc}/weights"] = dw
            grads[f"{sc}/biases"] = dyr.sum(axis=(0, 1, 2))
            d = rnd(dx).astype(dt)
        elif k == "pool":
            level = max(skip_grads)
            d = rnd(O.max_pool_2x2_grad(L["x"], d)[..., :L["keep"]] + skip_grads.pop(level)).astype(dt)
    if dctx is not None:
        grads.update(fc_backward(tape, dctx))
    tape.dctx = dctx
    return grads


def fc_backward(tape: Tape, dctx: np.ndarray) -> dict:
    """Gradients of the context MLP (slim_nets.fc, Backbone/slim_nets.py:34-57) given dctx, the gradient w.r.t. its
    output (the concatenated gamma_mod slices of every modulated layer)."""
    grads = {}
    g = dctx
    for F in reversed(tape.fc):
        if F["hidden"]:
            if F["mult"] is not None:
                g = g * F["mult"]
            g = g * (F["pre"] > 0)
        dx, dw, db = O.fully_connected_grad(F["x"], F["w"], g)
        grads[f"{F['scope']}/weights"] = dw
        grads[f"{F['scope']}/biases"] = db
        g = dx
    return grads


def total_grads(params: dict, data_grads: dict, cfg: GUNetCfg) -> dict:
    g = dict(data_grads)
    if cfg.weight_decay_rate > 0:
        for k in regularized_names(cfg, params):
            g[k] = g[k] + cfg.weight_decay_rate * params[k].astype(g[k].dtype)
    return g


def train_step(params: dict, slots: dict, step: int, inputs: dict, labels, cfg: GUNetCfg, lr: float, rnd=_identity,
               wrnd=None):
    """One `sess.run([train_op, loss])` with Adam (core/solver.py:204-207). Mutates params / slots."""
    tape = forward(params, inputs, cfg, True, rnd, wrnd, step=step)
    data_loss, dl = loss_and_dlogits(tape, labels, cfg)
    total = float(data_loss) + regularization_loss(params, cfg)
    grads = total_grads(params, backward(tape, dl, cfg, rnd), cfg)
    for k, g in grads.items():
        w = params[k].astype(np.float64)
        m, v = slots.setdefault(k, (np.zeros_like(w), np.zeros_like(w)))
        w, m, v = O.adam_step(w, g.astype(np.float64), m, v, step, lr)
        slots[k] = (m, v)
        params[k] = w.astype(params[k].dtype)
    for k, v in tape.new_moving.items():          # UPDATE_OPS control dependency of the train op (core/solver.py:236-239)
        params[k] = v.astype(params[k].dtype)
    return total, tape, grads
