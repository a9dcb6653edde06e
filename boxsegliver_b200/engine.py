"""U-Net 2-D training / inference engine on the sm_100a kernels (host side, Python + ctypes only).

What TF's executor does for the reference's graph (/root/reference/NetworksV2/UNet.py:58-155 built by
BaseNet.__call__, differentiated and updated by Solver, /root/reference/core/solver.py:221-243) this
module does explicitly: it plans every buffer once, then a step is a fixed sequence of C-ABI enqueues
on one compute stream (plus a side stream for the gradient all-reduce), replayable as a CUDA graph.

Memory plan (all NHWC, bf16 unless noted):
  * per conv layer: `y` (pre-norm conv output, kept for backward) and `a` (post norm+ReLU);
  * skip-concat is zero-copy: encoder level i writes `a` into channels [0,C) of cat_i, the
    transposed conv writes into [C,2C) (UNet.py:93 order: skip first, up second);
  * parameters live in flat fp32 arenas W / G / M / V with a bf16 shadow of W for the tensor-core
    convs; region A (L2-regularised: weights, biases) precedes region B (gamma, beta);
  * two ping-pong gradient buffers + one dcat_i per level carry activation gradients.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .device import Context, DeviceBuffer, f32_to_bf16_bits

F32 = 4
BF16 = 2


@dataclass
class EngineConfig:
    batch: int
    height: int = 256
    width: int = 256
    channel: int = 3
    classes: tuple = ("Background", "Liver", "Tumor")
    init_channels: int = 64
    num_down_samples: int = 4
    normalizer: str = "batch_norm"
    weight_decay_rate: float = 1e-5
    bias_decay: bool = False
    loss_type: str = "xentropy"
    loss_weight_type: str = "none"
    loss_numeric_w: tuple = ()
    loss_proportion_decay: float = 1000.0
    optimizer: str = "adam"            # adam | momentum | adamw (core/solver.py:204-219)
    adam_beta1: float = 0.9            # reference defaults {"beta1": 0.9, "beta2": 0.99} (solver.py:206)
    adam_beta2: float = 0.99
    adam_eps: float = 1e-8
    momentum: float = 0.9              # solver.py:209
    use_nesterov: bool = False
    adamw_weight_decay: float = None   # AdamW decoupled decay; None -> weight_decay_rate (solver.py:212)
    weight_init: str = "xavier"        # xavier | trunc_norm (base.py:137-151)
    bn_decay: float = 0.999
    bn_eps: float = 1e-3
    in_eps: float = 1e-6
    training: bool = True            # allocate backward / optimizer state
    world: int = 1                   # data-parallel replicas (gradient mean over `world`)

    @property
    def num_classes(self):
        return len(self.classes)


class View:
    """A strided NHWC window into a device buffer: channels [c0, c0+c) of a tensor with stride ld."""

    def __init__(self, buf: DeviceBuffer, n, h, w, c, ld=None, c0=0, esize=BF16):
        self.buf, self.n, self.h, self.w, self.c = buf, n, h, w, c
        self.ld = ld or c
        self.c0 = c0
        self.esize = esize

    @property
    def p(self) -> C.c_void_p:
        return C.c_void_p(self.buf.ptr + self.c0 * self.esize)

    @property
    def pixels(self):
        return self.n * self.h * self.w

    def slice(self, c0, c):
        return View(self.buf, self.n, self.h, self.w, c, self.ld, self.c0 + c0, self.esize)


@dataclass
class Param:
    name: str
    shape: tuple
    offset: int = 0          # element offset in the arena
    size: int = 0
    alloc: int = 0           # elements reserved in the arena (>= size; the stem filter is padded to 64 x 64)
    region: str = "A"        # "A" regularised, "B" not, "S" moving statistics (not trained)
    dev_shape: tuple = None  # stored shape when the input-channel axis is zero-padded to a 64-channel block (else `shape`)

    def to_dev(self, a: np.ndarray) -> np.ndarray:
        """TF-shaped variable -> flat stored layout (zero rows for the padded input channels)."""
        if self.dev_shape is None:
            return a.ravel()
        out = np.zeros(self.dev_shape, np.float32)
        out[:, :, :self.shape[2], :] = a
        return out.ravel()

    def from_dev(self, flat: np.ndarray) -> np.ndarray:
        a = flat[self.offset:self.offset + self.size]
        if self.dev_shape is None:
            return a.reshape(self.shape).copy()
        return a.reshape(self.dev_shape)[:, :, :self.shape[2], :].copy()


@dataclass
class ConvL:
    kind: str                # "stem" | "conv" | "convT" | "logits"
    scope: str
    cin: int
    cout: int
    h: int
    w: int                   # INPUT spatial size
    level: int
    cin_dev: int = 0         # stored input channels when cin is not a multiple of 64 (UNetInter --mid_cat: 66 -> 128); 0 = cin
    role: str = ""           # enc1 | enc2 (pooled afterwards) | bridge1 | bridge2 | dec1 (reads the concat) | dec2
    center: bool = True      # normaliser has beta / gamma (GUNet's modulated blocks take them from the YAML)
    scale: bool = True
    mod_off: int = None      # GUNet: first column of this layer's slice of the context-MLP output
    sp_off: int = None       # GUNet: first column of this layer's slice of the level's 1x1 guide conv
    affine: str = None       # GUNet after_affine: scope of the ChannelWiseAffine variables of this encoder layer
    bn_decay: float = None   # batch-norm moving-average decay of this layer (None: cfg.bn_decay; GUNet encoder: 0.99)
    x: View = None
    y: View = None           # pre-norm conv output (conv/stem), or output (convT)
    a: View = None           # post-activation
    pooled: View = None
    params: dict = field(default_factory=dict)
    norm: dict = field(default_factory=dict)


def _align(n, a=64):
    return (n + a - 1) // a * a


def truncated_normal(rng, shape, stddev: float, mean: float = 0.0) -> np.ndarray:
    """tf.truncated_normal semantics: N(mean, stddev) samples, values beyond 2 stddev are dropped and re-drawn."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (mean + stddev * out).astype(np.float32)


OPTIMIZERS = ("adam", "momentum", "adamw")


def enqueue_optimizer(eng, lr: float):
    """One fused update over the flat arenas of `eng` (2-D and 3-D engines share the layout: region A = L2-regularised
    [0, n_reg), region B = the rest). Hyper-parameters come from the config (core/solver.py:86-97,204-219)."""
    ctx, s, cfg = eng.ctx, eng.stream, eng.cfg
    l2 = cfg.weight_decay_rate if cfg.weight_decay_rate > 0 else 0.0
    regions = [(0, eng.n_reg, l2, eng.sumsq.p), (eng.n_reg, eng.n_train - eng.n_reg, 0.0, None)]
    if cfg.optimizer not in OPTIMIZERS:
        raise ValueError("Not supported optimizer: " + cfg.optimizer)
    decay = 0.0
    if cfg.optimizer == "adamw":       # DecoupledWeightDecayExtension decays EVERY variable (decay_var_list=None)
        decay = cfg.weight_decay_rate if cfg.adamw_weight_decay is None else cfg.adamw_weight_decay
    for off, n, rate, sq in regions:
        if n <= 0:
            continue
        w = C.c_void_p(eng.W.ptr + off * F32)
        g = C.c_void_p(eng.G.ptr + off * F32)
        m = C.c_void_p(eng.M.ptr + off * F32)
        wb = C.c_void_p(eng.Wbf.ptr + off * BF16)
        if cfg.optimizer in ("adam", "adamw"):
            v = C.c_void_p(eng.V.ptr + off * F32)
            d = _lib.AdamDesc(lr, cfg.adam_beta1, cfg.adam_beta2, cfg.adam_eps, rate, 1.0, eng.step_count, decay)
            ctx.call("bsl_adam_step", C.byref(d), w, g, m, v, wb, C.c_size_t(n), sq, s)
        else:
            ctx.call("bsl_momentum_step", C.c_float(lr), C.c_float(cfg.momentum), C.c_int(int(cfg.use_nesterov)),
                     C.c_float(rate), C.c_float(1.0), w, g, m, wb, C.c_size_t(n), sq, s)


class UNetEngine:
    def __init__(self, ctx: Context, cfg: EngineConfig):
        self.ctx, self.cfg = ctx, cfg
        if 9 * cfg.channel > 64:
            raise ValueError("im_channel must be <= 7 (the stem's im2col row holds 9 * channels <= 64 columns)")
        if cfg.init_channels % 64:
            raise ValueError("init_channels must be a multiple of 64 for the tcgen05 conv path")
        ds = 2 ** cfg.num_down_samples
        if cfg.height % ds or cfg.width % ds:
            raise ValueError(f"height/width must be multiples of {ds}")
        if not self._loss_terms():
            raise ValueError("Not supported loss_type: {}".format(cfg.loss_type))  # UNet.py:132
        if cfg.loss_weight_type not in ("none", "numerical", "proportion"):
            raise ValueError("Not supported weight type: " + cfg.loss_weight_type)
        if cfg.loss_weight_type == "numerical" and len(cfg.loss_numeric_w) != cfg.num_classes:
            raise KeyError("w_type `numerical` need keyword argument `numeric_w` (one value per class)")
        if cfg.optimizer not in OPTIMIZERS:
            raise ValueError("Not supported optimizer: " + cfg.optimizer)       # solver.py:217
        if cfg.weight_init not in ("xavier", "trunc_norm"):
            raise ValueError("Not supported weight initializer: " + cfg.weight_init)   # base.py:147
        self.step_count = 0
        self._bufs = []
        self._skip_exchange = False   # bench.py's data-parallel check runs one backward without the gradient exchange
        self._plan_params()
        self._plan_activations()
        self.stream = ctx.stream
        self.comm_stream = None
        self._graph = None
        self._grad_stream = self.stream
        self._overlap_wgrad = cfg.training and os.environ.get("BSL_WGRAD_OVERLAP", "1") != "0"
        self._fork_pre = os.environ.get("BSL_WGRAD_FORK", "pre") == "pre"
        # Programmatic dependent launch (csrc/internal.h bsl_launch): 0 off, 1 every launch (default), 2 only in the
        # phases where the compute stream has the SMs to itself (forward, loss head, optimizer). Measured on B200
        # (tools/ab.sh, profiles/r01_ab_experiments.md): the launch attribute alone, every kernel waiting with griddepcontrol.wait before its first
        # global access, is worth 0.15-0.25 ms per step; an explicit early trigger (BSL_PDL_TRIGGER in ptx.cuh) makes
        # the step 0.6-0.9 ms SLOWER with or without the filter-gradient stream, so it is compiled out.
        self._pdl_mode = int(os.environ.get("BSL_PDL", "1"))
        self._ev_ring, self._ev_ring_i = [ctx.new_event() for _ in range(160)], 0
        if cfg.training:
            self.wg_stream = ctx.new_stream()
            self._dy_events = [ctx.new_event(), ctx.new_event()]
            self._dy_busy = [None, None]
        # Image-slice pipelining (bsl_pipe, include/bsl_b200.h): the HBM-bound normalisation apply of layer L runs on
        # `aux_stream` beside the tensor-core kernel that consumes its output (fprop of layer L + 1 / dgrad of layer L),
        # which loads the tiles of an image only after the apply pass has published that image's slice.
        # Measured on B200 (profiles/r01_pipe_experiment.md): beside a tcgen05 kernel the apply passes run 2-4x slower
        # (the UMMA operand fetch saturates the SM's L1 / shared-memory data path that their loads share), so the
        # step gets SLOWER (23.8 -> 25.1 ms at cfg2); the schedule stays available for experiments, off by default.
        self._pipe_on = os.environ.get("BSL_PIPE", "0") != "0"
        # measured: the mask loads in the epilogue cost dgrad +0.45 ms, more than the 0.49 ms relu_bwd pass they replace
        self._fuse_relu_bwd = os.environ.get("BSL_FUSE_RELU_BWD", "0") != "0"
        self._fuse_head = os.environ.get("BSL_FUSE_HEAD", "1") != "0"
        # instance-norm statistics out of the conv epilogue instead of a pass over the conv output (bit-identical outputs)
        self._fuse_inst_stats = os.environ.get("BSL_FUSE_INST_STATS", "1") != "0"
        # backward of the last normalised layer with the logits-layer dgrad recomputed per pixel from dlogits
        # (bsl_norm_bwd_reduce_head / _apply_head) instead of a 128-byte-per-pixel gradient tensor; bit-identical
        self._fuse_head_bwd = os.environ.get("BSL_FUSE_HEAD_BWD", "1") != "0"
        self._head_grad = None
        self.aux_stream = ctx.new_stream()
        self._pipe_rows = 2 * len(self.layers)
        self.pipe_buf = self._alloc(2 * self._pipe_rows * 64 * 4).zero()
        self._pipe_epoch = 0
        cap = int(os.environ.get("BSL_PIPE_SLICES", "8")) if cfg.normalizer == "batch_norm" else 64
        self._pipe_slices = max(d for d in range(1, min(cfg.batch, cap) + 1) if cfg.batch % d == 0)
        self._pending = None

    # ------------------------------------------------------------------ per-kernel timing (bench roofline)
    def enable_conv_timing(self, on: bool = True):
        """Bracket every tcgen05 conv launch with CUDA events on the compute stream (bench.py roofline)."""
        self._timing = on
        self._timed = []          # (name, flops, ev0, ev1) in launch order for the current step(s)
        self._ev_pool = getattr(self, "_ev_pool", [])
        self._ev_next = 0

    def _next_event(self):
        """Round-robin pool for fork / join events (a step records far fewer than the ring holds)."""
        e = self._ev_ring[self._ev_ring_i]
        self._ev_ring_i = (self._ev_ring_i + 1) % len(self._ev_ring)
        return e

    def _ev(self):
        if self._ev_next == len(self._ev_pool):
            self._ev_pool.append(self.ctx.new_event())
        e = self._ev_pool[self._ev_next]
        self._ev_next += 1
        return e

    def _tc(self, name: str, flops: float, fn_name: str, *args, stream=None):
        """Enqueue a tensor-core conv call, optionally bracketed by events on the stream it is launched on."""
        if getattr(self, "_timing", False):
            e0, e1 = self._ev(), self._ev()
            st = stream if stream is not None else self.stream
            self.ctx.record(e0, st)
            self.ctx.call(fn_name, *args)
            self.ctx.record(e1, st)
            self._timed.append((name, flops, e0, e1))
        else:
            self.ctx.call(fn_name, *args)

    def conv_timing_report(self):
        """Sum of device time and algorithmic FLOPs over the bracketed conv launches since enable_conv_timing."""
        self.ctx.sync(self.stream)
        tot_ms, tot_fl, per = 0.0, 0.0, {}
        for name, fl, e0, e1 in self._timed:
            ms = self.ctx.elapsed_ms(e0, e1)
            tot_ms += ms
            tot_fl += fl
            a = per.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += ms
            a[2] += fl
        return {"launches": len(self._timed), "ms": tot_ms, "flops": tot_fl, "per_kind": per}

    # ------------------------------------------------------------------ data parallel
    def attach_comm(self, rank: int, world: int, unique_id: bytes):
        """Join the NCCL communicator (one process per GPU). cfg.world must equal `world`."""
        assert world == self.cfg.world, (world, self.cfg.world)
        self.ctx.attach_comm(rank, world, unique_id)
        self.rank = rank
        self._plan_buckets()

    # Gradient exchange overlapped with backward (replaces AllReduceCrossDeviceOps("nccl", num_packs=2),
    # /root/reference/utils/distribution_utils.py:91-98): the regularised region of the gradient arena is laid out in
    # forward layer order and backward fills it from the end, so a bucket is a suffix range that is complete as soon
    # as its first layer's wgrad has been enqueued. Each bucket is all-reduced on a side stream behind an event.
    BUCKET_ELEMS = 6 << 20

    def _plan_buckets(self):
        self.comm_stream = self.ctx.new_stream()
        self._ev_ready = [self.ctx.new_event() for _ in range(16)]
        self._ev_done = self.ctx.new_event()
        self._ev_stats = self.ctx.new_event()
        self._bucket_at = {}          # scope of the layer that closes a bucket -> (offset, count)
        # parameters outside the conv trunk (GUNet guide convs) sit after the last layer in region A and get their
        # gradients late in backward: they travel with the tail
        lastp = [p for p in self.params.values() if p.region == "A" and p.name.startswith(self.layers[-1].scope + "/")]
        self._trunk_end = max(p.offset + _align(max(p.size, p.alloc)) for p in lastp)
        end, acc, last = self._trunk_end, 0, None
        # the final bucket closes at the stem's filter gradient, the last kernel of backward, and overlaps nothing: a
        # boundary at the first conv of the second level keeps it down to the first level's ~40 K values
        early = next((L.scope for L in self.layers if L.kind == "conv" and L.level == 1), None)
        split_early = os.environ.get("BSL_BUCKET_EARLY_SPLIT", "1") != "0"
        for L in reversed(self.layers):
            p = self.params[f"{L.scope}/weights"]
            acc = end - p.offset
            last = L
            full = acc >= self.BUCKET_ELEMS or (split_early and L.scope == early and acc > 0)
            if full and len(self._bucket_at) < len(self._ev_ready) - 1:
                self._bucket_at[L.scope] = (p.offset, acc)
                end, acc = p.offset, 0
        if end > 0:
            self._bucket_at[last.scope] = (0, end)
        self._bucket_i = 0

    def _after_grad(self, L: ConvL):
        """Called when layer L's parameter gradients have been enqueued (backward order)."""
        if self.cfg.world <= 1 or self._skip_exchange or L.scope not in getattr(self, "_bucket_at", {}):
            return
        off, cnt = self._bucket_at[L.scope]
        ev = self._ev_ready[self._bucket_i]
        self._bucket_i += 1
        self.ctx.record(ev, self._grad_stream)
        self.ctx.call("bsl_stream_wait_event", self.comm_stream, ev)
        self.ctx.call("bsl_allreduce_sum_f32", C.c_void_p(self.G.ptr + off * F32), C.c_size_t(cnt), self.comm_stream)

    def _allreduce_grads(self):
        """Tail of the exchange: the un-regularised region (gamma / beta / FC, tiny) follows the buckets on the side
        stream -- every gradient has been enqueued once loss_backward returns -- then the compute stream joins it, so
        the only collective work the optimizer waits for is the last bucket and this tail."""
        if self.cfg.world <= 1 or self._skip_exchange:
            return
        if self.comm_stream is None:       # attach_comm not called: single all-reduce on the compute stream
            self.ctx.call("bsl_allreduce_sum_f32", self.G.p, C.c_size_t(self.n_train), self.stream)
            return
        nb = self.n_train - self._trunk_end
        if nb > 0:
            ev = self._ev_ready[-1]
            self.ctx.record(ev, self.stream)
            self.ctx.call("bsl_stream_wait_event", self.comm_stream, ev)
            self.ctx.call("bsl_allreduce_sum_f32", C.c_void_p(self.G.ptr + self._trunk_end * F32), C.c_size_t(nb),
                          self.comm_stream)
        self.ctx.record(self._ev_done, self.comm_stream)
        self.ctx.call("bsl_stream_wait_event", self.stream, self._ev_done)
        self._bucket_i = 0

    def _allreduce_moving_stats(self):
        # MirroredStrategy aggregates the moving-average updates with MEAN across replicas
        # (/root/reference/core/estimator.py:570-613); batch statistics themselves stay per replica. Nothing reads the
        # moving statistics before the next step's forward, so the exchange rides the side stream (ahead of the
        # gradient buckets) and is ordered before the optimizer by the same join as the gradients.
        if self.cfg.world > 1 and self.n_stats > 0 and not self._skip_exchange:
            st = self.comm_stream if self.comm_stream is not None else self.stream
            if st is not self.stream:
                self.ctx.record(self._ev_stats, self.stream)
                self.ctx.call("bsl_stream_wait_event", st, self._ev_stats)
            self.ctx.call("bsl_allreduce_sum_f32", self.S.p, C.c_size_t(self.n_stats), st)
            self.ctx.call("bsl_scale_f32", self.S.p, C.c_size_t(self.n_stats), C.c_float(1.0 / self.cfg.world), st)

    # ------------------------------------------------------------------ planning
    def _alloc(self, nbytes) -> DeviceBuffer:
        b = self.ctx.alloc(max(int(nbytes), 16))
        self._bufs.append(b)
        return b

    def _layer_specs(self):
        cfg = self.cfg
        specs = []
        c, cin = cfg.init_channels, cfg.channel
        h, w = cfg.height, cfg.width
        for i in range(cfg.num_down_samples):
            for j in (1, 2):
                kind = "stem" if (i == 0 and j == 1) else "conv"
                specs.append(ConvL(kind, f"UNet/Encode{i + 1}/Repeat/convolution2d_{j}", cin, c, h, w, i, role=f"enc{j}"))
                cin = c
            c *= 2
            h //= 2
            w //= 2
        for j in (1, 2):
            specs.append(ConvL("conv", f"UNet/ED-Bridge/ED-Bridge_{j}", cin, c, h, w, cfg.num_down_samples,
                               role=f"bridge{j}"))
            cin = c
        for i in reversed(range(cfg.num_down_samples)):
            c //= 2
            specs.append(ConvL("convT", f"UNet/Decode{i + 1}/Conv2d_transpose", cin, cin // 2, h, w, i))
            h *= 2
            w *= 2
            for j in (1, 2):
                specs.append(ConvL("conv", f"UNet/Decode{i + 1}/Repeat/convolution2d_{j}",
                                   c + cin // 2 if j == 1 else c, c, h, w, i, role=f"dec{j}"))
            cin = c
        specs.append(ConvL("logits", "UNet/AdjustChannels", cin, cfg.num_classes, h, w, 0))
        return specs

    def _plan_params(self):
        cfg = self.cfg
        self.layers = self._layer_specs()
        ns = "BatchNorm" if cfg.normalizer == "batch_norm" else "InstanceNorm"
        self.norm_scope = ns
        plist = []
        for L in self.layers:
            if L.kind in ("stem", "conv"):
                plist.append(Param(f"{L.scope}/weights", (3, 3, L.cin, L.cout),
                                   alloc=64 * L.cout if L.kind == "stem" else 0,
                                   dev_shape=(3, 3, L.cin_dev, L.cout) if L.cin_dev else None))
                if L.scale:
                    plist.append(Param(f"{L.scope}/{ns}/gamma", (L.cout,), region="B"))
                if L.center:
                    plist.append(Param(f"{L.scope}/{ns}/beta", (L.cout,), region="B"))
                if cfg.normalizer == "batch_norm":
                    plist.append(Param(f"{L.scope}/{ns}/moving_mean", (L.cout,), region="S"))
                    plist.append(Param(f"{L.scope}/{ns}/moving_variance", (L.cout,), region="S"))
            elif L.kind == "convT":
                plist.append(Param(f"{L.scope}/weights", (2, 2, L.cout, L.cin)))
                plist.append(Param(f"{L.scope}/biases", (L.cout,), region="B" if cfg.bias_decay else "A"))
            else:
                plist.append(Param(f"{L.scope}/weights", (1, 1, L.cin, L.cout)))
                plist.append(Param(f"{L.scope}/biases", (L.cout,), region="B" if cfg.bias_decay else "A"))
        plist += self._extra_params()
        off = 0
        for region in ("A", "B"):
            for p in plist:
                if p.region == region:
                    p.size = int(np.prod(p.dev_shape or p.shape))
                    p.offset = off
                    off += _align(max(p.size, p.alloc))
            if region == "A":
                self.n_reg = off
        self.n_train = off
        soff = 0
        for p in plist:
            if p.region == "S":
                p.size = int(np.prod(p.shape))
                p.offset = soff
                soff += _align(p.size)
        self.n_stats = soff
        self.params = {p.name: p for p in plist}
        self.W = self._alloc(self.n_train * F32)
        self.Wbf = self._alloc(self.n_train * BF16)
        self.S = self._alloc(max(self.n_stats, 1) * F32)
        if cfg.training:
            self.G = self._alloc(self.n_train * F32).zero()
            self.M = self._alloc(self.n_train * F32).zero()
            self.V = self._alloc(self.n_train * F32).zero() if cfg.optimizer in ("adam", "adamw") else None
        self.sumsq = self._alloc(16)

    def _extra_params(self):
        """Parameters outside the conv trunk (GUNet: context MLP, guide convs)."""
        return []

    def _pp(self, arena: DeviceBuffer, name: str, esize=F32, off: int = 0) -> C.c_void_p:
        if name not in self.params:
            return None
        return C.c_void_p(arena.ptr + (self.params[name].offset + off) * esize)

    def _guide_channels(self, L: ConvL) -> int:
        return 0

    def _first_input(self, L: ConvL) -> View:
        raise NotImplementedError("the first layer of this engine is the im2col stem")

    def _pooled_lanes(self, L: ConvL) -> int:
        """Channel stride of the pooled tensor behind an encoder block."""
        return L.cout

    def _after_pool(self, L: ConvL, stream):
        """Hook behind the fused norm + ReLU + pool pass of an encoder block (UNetInter --mid_cat adds the guide lanes)."""

    def _norm_groups(self, L: ConvL, default: int) -> int:
        """Entries per channel of the layer's normalisation scalars (1 batch norm, n instance norm / modulated)."""
        return default

    def _apply_desc(self, L: ConvL, nd):
        """Descriptor of the apply / backward passes (GUNet: the per-sample view of a modulated batch-norm layer)."""
        return nd

    def _plan_activations(self):
        cfg = self.cfg
        n = cfg.batch
        cat = {}
        max_act = 0
        prev_a = None
        groups_max = n if cfg.normalizer == "instance_norm" else 1
        small = 0  # floats for per-layer normalisation scalars
        for L in self.layers:
            if L.kind in ("stem", "conv"):
                L.y = View(self._alloc(n * L.h * L.w * L.cout * BF16), n, L.h, L.w, L.cout)
                is_enc2 = L.role == "enc2"
                if is_enc2:
                    cbuf = View(self._alloc(n * L.h * L.w * 2 * L.cout * BF16), n, L.h, L.w, 2 * L.cout)
                    cat[L.level] = cbuf
                    L.a = cbuf.slice(0, L.cout)
                    pl = self._pooled_lanes(L)      # > cout: extra lanes behind the activation's (UNetInter --mid_cat)
                    pbuf = self._alloc(n * (L.h // 2) * (L.w // 2) * pl * BF16)
                    if pl != L.cout:
                        pbuf.zero()
                    L.pooled = View(pbuf, n, L.h // 2, L.w // 2, L.cout, pl)
                else:
                    L.a = View(self._alloc(n * L.h * L.w * L.cout * BF16), n, L.h, L.w, L.cout)
                if L.kind == "conv":
                    L.x = cat[L.level] if L.role == "dec1" else prev_a
                    if L.x is None:      # a first layer that is not the im2col stem (GUNet --img_grad: 9 input channels)
                        L.x = self._first_input(L)
                prev_a = L.pooled if is_enc2 else L.a
                if is_enc2 and L.pooled.ld != L.cout:
                    prev_a = View(L.pooled.buf, n, L.h // 2, L.w // 2, L.pooled.ld)
                g = self._norm_groups(L, groups_max)
                k = 2 + self._guide_channels(L)
                L.norm = dict(groups=g, off=small, k=k)
                small += (2 * k + 6) * _align(g * L.cout, 16)  # sums (k x f64), mean, rstd, scale, shift, c1, c2
                max_act = max(max_act, L.h * L.w * L.cout)
            elif L.kind == "convT":
                L.x = prev_a
                L.a = cat[L.level].slice(L.cout, L.cout)   # upper channel half of the concat buffer
                L.y = L.a
                prev_a = cat[L.level]
            else:
                L.x = prev_a
        self.cat = cat
        self.images = self._alloc(n * cfg.height * cfg.width * cfg.channel * F32)
        # im2col rows of the stem: 9 * channels live columns in a pitch of 32 (3 input channels) or 64 columns; the conv
        # kernels read a 32-column matrix as cin = 64 with the upper half zero-filled by TMA (bsl_stem_im2col_ld)
        sc = int(os.environ.get("BSL_STEM_COLS", "0")) or (32 if 9 * cfg.channel <= 32 else 64)
        self.stem_cols = sc
        self.stem_col = View(self._alloc(n * cfg.height * cfg.width * sc * BF16), n, cfg.height, cfg.width, sc)
        self.labels = self._alloc(n * cfg.height * cfg.width * 4)
        npx = n * cfg.height * cfg.width
        self.logits = self._alloc(npx * cfg.num_classes * F32)
        self.prob = self._alloc(npx * cfg.num_classes * F32)
        self.masks = self._alloc(npx * (cfg.num_classes - 1))
        self.argmax = self._alloc(npx)
        self.ilr = self._alloc(n * (cfg.num_classes - 1) * 3 * 4)
        self.counts = self._alloc(n * cfg.num_classes * 4)
        self.loss_dev = self._alloc(16)
        self.small = self._alloc(max(small, 16) * F32)
        ld = self._loss_desc()
        self.loss_ws_bytes = self.ctx.lib.bsl_loss_workspace(self.ctx.h, C.byref(ld))
        self.loss_ws = self._alloc(self.loss_ws_bytes)
        if cfg.training:
            self.dlogits = self._alloc(npx * cfg.num_classes * F32)
            self.g1 = self._alloc(n * max_act * BF16)
            self.g2 = self._alloc(n * max_act * BF16)
            self.dyb = [self._alloc(n * max_act * BF16), self._alloc(n * max_act * BF16)]
            self.dcat = {i: View(self._alloc(v.pixels * v.c * BF16), n, v.h, v.w, v.c) for i, v in cat.items()}
            ws = 0
            for L in self.layers:
                if L.kind == "stem":
                    d = _lib.Conv2dDesc(n, L.h, L.w, 64, L.cout, 1, 1, self.stem_cols, L.cout)
                    ws = max(ws, self.ctx.lib.bsl_conv2d_wgrad_workspace(self.ctx.h, C.byref(d)))
                elif L.kind == "conv":
                    d = self._conv_desc(L)
                    ws = max(ws, self.ctx.lib.bsl_conv2d_wgrad_workspace(self.ctx.h, C.byref(d)))
                elif L.kind == "convT":
                    d = self._convT_desc(L)
                    ws = max(ws, self.ctx.lib.bsl_convT2d_bwd_filter_workspace(self.ctx.h, C.byref(d)))
            self.wgrad_ws_bytes = int(ws)
            self.wgrad_ws = self._alloc(max(ws, 16))

    def _flops(self, L: ConvL) -> float:
        """Algorithmic FLOPs of ONE pass (fprop, dgrad or wgrad) of layer L: 2 * MACs, un-padded."""
        n = self.cfg.batch
        if L.kind == "convT":
            return 2.0 * n * (2 * L.h) * (2 * L.w) * L.cin * L.cout   # one tap per output pixel
        k = 1 if L.kind == "logits" else 9
        return 2.0 * n * L.h * L.w * k * L.cin * L.cout

    def _loss_terms(self):
        """UNet accepts exactly "xentropy" or "dice" (UNet.py:123-132); GUNet overrides with substring matching."""
        return [self.cfg.loss_type] if self.cfg.loss_type in ("xentropy", "dice") else []

    def step_flops(self) -> dict:
        """Algorithmic fwd / bwd FLOPs of one training step (bwd = dgrad + wgrad, no dgrad for the stem)."""
        fwd = sum(self._flops(L) for L in self.layers)
        bwd = sum(self._flops(L) * (1 if L.kind == "stem" else 2) for L in self.layers)
        tc = sum(self._flops(L) * (2 if L.kind == "stem" else 3) for L in self.layers if L.kind != "logits")
        return {"fwd": fwd, "bwd": bwd, "total": fwd + bwd, "tensor_core": tc}

    # ------------------------------------------------------------------ descriptors
    def _conv_desc(self, L: ConvL):
        k = 1 if L.kind == "logits" else 3
        x_ld = L.x.ld if L.x is not None else L.cin
        y_ld = L.y.ld if L.y is not None else L.cout
        return _lib.Conv2dDesc(self.cfg.batch, L.h, L.w, L.cin_dev or L.cin, L.cout, k, k, x_ld, y_ld)

    def _convT_desc(self, L: ConvL):
        return _lib.ConvT2dDesc(self.cfg.batch, L.h, L.w, L.cin, L.cout, L.x.ld, L.a.ld, 1)

    def _norm_desc(self, L: ConvL):
        cfg = self.cfg
        bn = cfg.normalizer == "batch_norm"
        return _lib.NormDesc(0 if bn else 1, cfg.batch, L.h * L.w, L.cout, L.y.ld, L.a.ld,
                             cfg.bn_eps if bn else cfg.in_eps, L.bn_decay if L.bn_decay is not None else cfg.bn_decay, 1,
                             int(L.center), int(L.scale))

    def _loss_desc(self):
        cfg = self.cfg
        wt = {"none": 0, "numerical": 1, "proportion": 2}[cfg.loss_weight_type]
        nw = (C.c_float * 8)(*([float(x) for x in cfg.loss_numeric_w] + [0.0] * (8 - len(cfg.loss_numeric_w))))
        return _lib.LossDesc(cfg.batch, cfg.height * cfg.width, cfg.num_classes, wt, nw,
                             float(cfg.loss_proportion_decay), 1.0 / cfg.world)

    def _norm_ptrs(self, L: ConvL):
        """sums (f64 x 2), mean, rstd, scale, shift, c1, c2 carved from the small-scalar arena."""
        g = L.norm["groups"]
        n = _align(g * L.cout, 16)
        base = self.small.ptr + L.norm["off"] * F32
        k2 = 2 * L.norm["k"]
        names = ["mean", "rstd", "scale", "shift", "c1", "c2"]
        out = {nm: C.c_void_p(base + (k2 + i) * n * F32) for i, nm in enumerate(names)}
        out["sums"] = C.c_void_p(base)
        return out

    # ------------------------------------------------------------------ weights
    def set_weights(self, weights: dict):
        """Load a {tf variable name: numpy array} map (TF layouts: HWIO, [k,k,Cout,Cin], vectors)."""
        hostW = np.zeros(self.n_train, np.float32)
        hostS = np.zeros(max(self.n_stats, 1), np.float32)
        for name, p in self.params.items():
            if name not in weights:
                raise KeyError(f"missing variable {name}")
            a = np.asarray(weights[name], np.float32)
            if tuple(a.shape) != tuple(p.shape):
                raise ValueError(f"{name}: shape {a.shape} != {p.shape}")
            (hostS if p.region == "S" else hostW)[p.offset:p.offset + p.size] = p.to_dev(a)
        self.W.upload(hostW)
        self.Wbf.upload(f32_to_bf16_bits(hostW))
        self.S.upload(hostS)

    def get_weights(self) -> dict:
        hostW = self.W.download(np.float32, (self.n_train,))
        hostS = self.S.download(np.float32, (max(self.n_stats, 1),))
        out = {}
        for name, p in self.params.items():
            out[name] = p.from_dev(hostS if p.region == "S" else hostW)
        return out

    def get_stored_forward(self) -> dict:
        """{scope: {"y": pre-norm conv output, "a": activation}} as fp32 numpy (what backward will read)."""
        out = {}
        for L in self.layers:
            if L.kind == "logits":
                continue
            d = {}
            for key, v in (("y", L.y), ("a", L.a)):
                if v is None:
                    continue
                full = self.ctx.bf16_to_f32(v.buf, (v.n, v.h, v.w, v.ld))
                d[key] = full[..., v.c0:v.c0 + v.c].copy()
            out[L.scope] = d
        return out

    def get_grads(self) -> dict:
        hostG = self.G.download(np.float32, (self.n_train,))
        return {name: p.from_dev(hostG) for name, p in self.params.items() if p.region != "S"}

    def get_slots(self) -> dict:
        """Optimizer slots in TF variable shapes: {variable name: (m, v)} for Adam / AdamW, {name: (acc,)} for Momentum."""
        m = self.M.download(np.float32, (self.n_train,))
        v = self.V.download(np.float32, (self.n_train,)) if self.V is not None else None
        out = {}
        for name, p in self.params.items():
            if p.region == "S":
                continue
            out[name] = tuple(p.from_dev(a) for a in ((m, v) if v is not None else (m,)))
        return out

    def set_slots(self, slots: dict):
        """Inverse of get_slots (every trainable variable must be present)."""
        m = np.zeros(self.n_train, np.float32)
        v = np.zeros(self.n_train, np.float32) if self.V is not None else None
        for name, p in self.params.items():
            if p.region == "S":
                continue
            if name not in slots:
                raise KeyError(f"missing optimizer slots of {name}")
            for dst, a in zip((m, v), slots[name]):
                a = np.asarray(a, np.float32)
                if tuple(a.shape) != tuple(p.shape):
                    raise ValueError(f"{name}: slot shape {a.shape} != {p.shape}")
                dst[p.offset:p.offset + p.size] = p.to_dev(a)
        self.M.upload(m)
        if v is not None:
            self.V.upload(v)

    def _draw_weight(self, rng, shp, fan_in, fan_out):
        """--weight_init (/root/reference/NetworksV2/base.py:137-151): slim.xavier_initializer() (uniform) or
        tf.truncated_normal_initializer(stddev=0.01) -- normal samples re-drawn while |x| > 2 sigma."""
        if getattr(self.cfg, "weight_init", "xavier") == "trunc_norm":
            return truncated_normal(rng, shp, 0.01)
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shp).astype(np.float32)

    def init_weights(self, seed: int = 0):
        """--weight_init weights (seeded numpy stream), zero biases, gamma 1, beta 0, moving 0 / 1."""
        rng = np.random.default_rng(seed)
        w = {}
        for name, p in self.params.items():
            if name.endswith("/weights"):
                shp = p.shape
                rf = shp[0] * shp[1]
                w[name] = self._draw_weight(rng, shp, rf * shp[2], rf * shp[3])
            elif name.endswith(("gamma", "moving_variance")):
                w[name] = np.ones(p.shape, np.float32)
            else:
                w[name] = np.zeros(p.shape, np.float32)
        self.set_weights(w)
        return w

    # ------------------------------------------------------------------ inputs
    def set_inputs(self, images: np.ndarray, labels: np.ndarray | None = None, stream=None):
        cfg = self.cfg
        assert images.shape == (cfg.batch, cfg.height, cfg.width, cfg.channel), images.shape
        self.images.upload(np.ascontiguousarray(images, np.float32), stream)
        if labels is not None:
            assert labels.shape == (cfg.batch, cfg.height, cfg.width), labels.shape
            self.labels.upload(np.ascontiguousarray(labels, np.int32), stream)

    # ------------------------------------------------------------------ forward
    # ------------------------------------------------------------------ image-slice pipelining
    def _pipe(self, row: int) -> _lib.Pipe:
        return _lib.Pipe(self.pipe_buf.ptr + row * 256, self.pipe_buf.ptr + (self._pipe_rows + row) * 256,
                         self._pipe_slices, self._pipe_epoch)

    def _pipe_active(self) -> bool:
        return self._pipe_on and self.ctx._prof is None and not getattr(self, "_timing", False)

    @staticmethod
    def _pipe_shape_ok(L: ConvL) -> bool:
        """The layer's tensor-core kernel is the halo-tile one (the only kernel that can wait on slices)."""
        return L.kind in ("conv", "convT") and L.h % 16 == 0 and L.w % 8 == 0

    def _take_pending(self, s):
        """(pipe to wait on or None, completion callback): the previous layer's apply pass still running on the
        auxiliary stream. The callback orders the main stream after that pass once the consumer is enqueued."""
        pend, self._pending = self._pending, None
        if pend is None:
            return None, lambda: None
        pipe, ev = pend
        return pipe, lambda: self.ctx.call("bsl_stream_wait_event", s, ev)

    def _pdl(self, on: bool):
        if self._pdl_mode == 2:
            self.ctx.call("bsl_debug_set", C.c_int(4), C.c_int(1 if on else 0))

    def forward(self, is_training: bool):
        ctx, s = self.ctx, self.stream
        call = ctx.call
        self._pdl(True)
        self._pipe_epoch += 1
        piping = self._pipe_active()
        self._head_done = False
        for idx, L in enumerate(self.layers):
            ctx.tag = L.scope
            if L.kind in ("stem", "conv"):
                d = self._conv_desc(L)
                nd = self._norm_desc(L)
                q = self._norm_ptrs(L)
                ns = self.norm_scope
                bn = self.cfg.normalizer == "batch_norm"
                # batch-norm training: the per-channel sums come out of the conv epilogue (no pass over y); instance norm
                # (training and inference): per-(sample, channel) sums from the same epilogue (bsl_conv2d_fprop_group_stats)
                fused = bn and is_training
                inst = (not bn) and self._fuse_inst_stats
                fn = "bsl_conv2d_fprop_stats" if fused else "bsl_conv2d_fprop"
                extra = (q["sums"],) if fused else ()
                if inst:
                    fn, extra = "bsl_conv2d_fprop_group_stats", (C.c_int(1), q["sums"])
                wbf = self._pp(self.Wbf, f"{L.scope}/weights", BF16)
                if L.kind == "stem":
                    # im2col (27 -> 64 columns, bf16) + 1x1 conv on the tensor cores
                    call("bsl_stem_im2col_ld", C.byref(d), self.images.p, self.stem_col.p, C.c_int(self.stem_cols), s)
                    d1 = _lib.Conv2dDesc(self.cfg.batch, L.h, L.w, 64, L.cout, 1, 1, self.stem_cols, L.y.ld)
                    self._tc("fprop", self._flops(L), fn, C.byref(d1), self.stem_col.p, wbf, L.y.p, *extra, s)
                else:
                    pw, done = self._take_pending(s)
                    if pw is not None:
                        inst = False
                        self._tc("fprop", self._flops(L), "bsl_conv2d_fprop_pipe", C.byref(d), L.x.p, wbf, L.y.p,
                                 q["sums"] if fused else None, C.byref(pw), s)
                        done()
                    else:
                        self._tc("fprop", self._flops(L), fn, C.byref(d), L.x.p, wbf, L.y.p, *extra, s)
                if not fused and not inst and (not bn or is_training):
                    call("bsl_norm_stats", C.byref(nd), L.y.p, q["sums"], s)
                mm = C.c_void_p(self.S.ptr + self.params[f"{L.scope}/{ns}/moving_mean"].offset * F32) if bn else None
                mv = C.c_void_p(self.S.ptr + self.params[f"{L.scope}/{ns}/moving_variance"].offset * F32) if bn else None
                call("bsl_norm_finalize", C.byref(nd), C.c_int(1 if is_training else 0), q["sums"],
                     self._pp(self.W, f"{L.scope}/{ns}/gamma"), self._pp(self.W, f"{L.scope}/{ns}/beta"), mm, mv,
                     q["mean"], q["rstd"], q["scale"], q["shift"], s)
                src = self._dropout_stage(L, nd, q, is_training)   # GUNet --dropout: a dropped copy of the normalised tensor
                guide = self._modulate(L, nd, q)   # GUNet: folds gamma_mod / guide bias into scale, shift
                nd = self._apply_desc(L, nd)
                gp = C.byref(guide) if guide is not None else None
                nxt = self.layers[idx + 1] if idx + 1 < len(self.layers) else None
                sig, st = None, s
                if piping and nxt is not None and self._pipe_shape_ok(nxt):
                    # the apply pass goes to the auxiliary stream and publishes image slices; the next layer's
                    # tensor-core kernel starts right away on the main stream and follows it slice by slice
                    sig, st = self._pipe(idx), self.aux_stream
                    ev = self._next_event()
                    ctx.record(ev, s)
                    call("bsl_stream_wait_event", st, ev)
                sp = C.byref(sig) if sig is not None else None
                cg = L.cout // 8
                if (self._fuse_head and nxt is not None and nxt.kind == "logits" and gp is None and L.pooled is None
                        and cg <= 32 and cg & (cg - 1) == 0 and 2 <= self.cfg.num_classes <= 4):
                    # last normalised layer: the logits come out of the same pass (no second read of the activation)
                    call("bsl_norm_apply_head", C.byref(nd), src, q["scale"], q["shift"], L.a.p,
                         self._pp(self.W, f"{nxt.scope}/weights"), self._pp(self.W, f"{nxt.scope}/biases"),
                         C.c_int(self.cfg.num_classes), self.logits.p, s)
                    self._head_done = True
                elif L.pooled is not None:
                    call("bsl_norm_apply_pool_mod_pipe", C.byref(nd), C.c_int(L.h), C.c_int(L.w), src, q["scale"],
                         q["shift"], gp, L.a.p, L.pooled.p, C.c_int(L.pooled.ld), sp, st)
                    self._after_pool(L, st)
                else:
                    call("bsl_norm_apply_mod_pipe", C.byref(nd), src, q["scale"], q["shift"], gp, L.a.p, sp, st)
                if sig is not None:
                    ev = self._next_event()
                    ctx.record(ev, st)
                    self._pending = (sig, ev)
            elif L.kind == "convT":
                d = self._convT_desc(L)
                pw, done = self._take_pending(s)
                self._tc("convT_fwd", self._flops(L), "bsl_convT2d_fwd_pipe", C.byref(d), L.x.p,
                         self._pp(self.Wbf, f"{L.scope}/weights", BF16), self._pp(self.W, f"{L.scope}/biases"), L.a.p,
                         C.byref(pw) if pw is not None else None, s)
                done()
            elif not self._head_done:
                d = self._conv_desc(L)
                call("bsl_conv2d_head_fprop", C.byref(d), L.x.p, self._pp(self.W, f"{L.scope}/weights"),
                     self._pp(self.W, f"{L.scope}/biases"), self.logits.p, s)

    def _modulate(self, L: ConvL, nd, q):
        """Hook between norm_finalize and norm_apply; returns the bsl_guide of the layer (or None)."""
        return None

    def _dropout_stage(self, L: ConvL, nd, q, is_training: bool):
        """Hook behind norm_finalize: the tensor the apply pass reads (GUNet's backbone --dropout substitutes a dropped
        copy of the normalised conv output and turns (scale, shift) into the identity)."""
        return L.y.p

    def _is_modulated(self, L: ConvL) -> bool:
        return False

    def _norm_backward_reduce(self, L: ConvL, nd, q, cur):
        """First half of the gradient through ReLU + normalisation of layer L: the per-channel sums over `cur` (the
        gradient w.r.t. the activation) and the gradients of the normaliser's own parameters (into G)."""
        call, s, ns = self.ctx.call, self.stream, self.norm_scope
        if self._head_grad is not None:
            dl, wh, classes = self._head_grad
            call("bsl_norm_bwd_reduce_head", C.byref(nd), L.y.p, dl, wh, C.c_int(classes), q["mean"], q["rstd"],
                 q["scale"], q["shift"], q["sums"], s)
        else:
            call("bsl_norm_bwd_reduce", C.byref(nd), L.y.p, cur.p, C.c_int(L.cout), q["mean"], q["rstd"],
                 q["scale"], q["shift"], q["sums"], s)
        call("bsl_norm_bwd_finalize", C.byref(nd), q["sums"], q["c1"], q["c2"],
             self._pp(self.G, f"{L.scope}/{ns}/gamma"), self._pp(self.G, f"{L.scope}/{ns}/beta"), s)

    def _norm_backward_apply(self, L: ConvL, nd, q, cur, oth, stream, sig=None):
        """Second half: `cur` -> `oth` (gradient w.r.t. the conv output), optionally publishing image slices."""
        if self._head_grad is not None:
            dl, wh, classes = self._head_grad
            self._head_grad = None
            assert sig is None
            self.ctx.call("bsl_norm_bwd_apply_head", C.byref(nd), L.y.p, dl, wh, C.c_int(classes), q["mean"], q["rstd"],
                          q["scale"], q["shift"], q["c1"], q["c2"], oth.p, C.c_int(L.cout), stream)
            return
        self.ctx.call("bsl_norm_bwd_apply_mod_pipe", C.byref(nd), L.y.p, cur.p, C.c_int(L.cout), q["mean"], q["rstd"],
                      q["scale"], q["shift"], q["c1"], q["c2"], None, oth.p, C.c_int(L.cout),
                      C.byref(sig) if sig is not None else None, stream)

    def _norm_backward(self, L: ConvL, nd, q, cur, oth):
        self._norm_backward_reduce(L, nd, q, cur)
        self._norm_backward_apply(L, nd, q, cur, oth, self.stream)

    def predict_outputs(self, with_counts: bool):
        """softmax, `<Cls>Pred` masks, argmax and (optionally) the integer Dice sums, one pass over the logits."""
        ld = self._loss_desc()
        self.ctx.call("bsl_softmax_threshold", C.byref(ld), self.logits.p, self.labels.p if with_counts else None,
                      self.prob.p, self.masks.p, self.argmax.p, self.ilr.p if with_counts else None, self.stream)

    # ------------------------------------------------------------------ loss + backward
    def loss_backward(self):
        ctx, s, cfg = self.ctx, self.stream, self.cfg
        call = ctx.call
        ld = self._loss_desc()
        terms = self._loss_terms()
        if "xentropy" in terms:
            call("bsl_label_counts", C.byref(ld), self.labels.p, self.counts.p, s)
            call("bsl_wxent_fwd_bwd", C.byref(ld), self.logits.p, self.labels.p, self.counts.p, self.loss_dev.p,
                 self.dlogits.p, self.loss_ws.p, C.c_size_t(self.loss_ws_bytes), s)
        if "dice" in terms:   # second scalar slot; dlogits accumulate when both terms are present (GUNet.py:399-408)
            call("bsl_dice_fwd_bwd", C.byref(ld), self.logits.p, self.labels.p,
                 self.loss_dev.at(4 * terms.index("dice")), self.dlogits.p, C.c_int(1 if len(terms) > 1 else 0),
                 self.loss_ws.p, C.c_size_t(self.loss_ws_bytes), s)
        n = cfg.batch
        # `cur` holds the gradient w.r.t. the current activation; dY (w.r.t. the conv output) alternates between two
        # buffers so that the filter gradient of layer L can run on the side stream while the main stream goes on
        # with dgrad(L) and the HBM-bound normalisation backward of layer L-1 (tensor-bound and memory-bound work
        # overlap); a dY buffer is reused two conv layers later, behind an event.
        cur, alt = self.g1, self.g2
        overlap = self._overlap_wgrad and ctx._prof is None
        if overlap:
            self._pdl(False)
        piping = self._pipe_active()
        self._pipe_epoch += 1
        relu_fused = set()      # levels whose transposed-conv ReluGrad was applied by the decoder dgrad's epilogue
        ws = self.wg_stream if overlap else s
        self._grad_stream = ws
        k = 0
        wsb = C.c_size_t(self.wgrad_ws_bytes)

        def fork():
            """side stream waits for everything enqueued on the main stream so far"""
            if overlap:
                ev = self._next_event()
                ctx.record(ev, s)
                call("bsl_stream_wait_event", ws, ev)

        for idx in range(len(self.layers) - 1, -1, -1):
            L = self.layers[idx]
            ctx.tag = L.scope
            if L.kind == "logits":
                d = self._conv_desc(L)
                fork()
                call("bsl_conv2d_head_wgrad", C.byref(d), L.x.p, self.dlogits.p, self._pp(self.G, f"{L.scope}/weights"),
                     self._pp(self.G, f"{L.scope}/biases"), ws)
                prev = self.layers[idx - 1]
                if (self._fuse_head_bwd and not piping and prev.kind == "conv" and prev.pooled is None
                        and not self._is_modulated(prev) and prev.a is L.x
                        and ctx.lib.bsl_norm_bwd_head_ok(ctx.h, C.byref(self._norm_desc(prev)), C.c_int(L.cout))):
                    # no Conv2DBackpropInput launch: the two normalisation-backward passes of `prev` recompute it
                    self._head_grad = (self.dlogits.p, self._pp(self.W, f"{L.scope}/weights"), L.cout)
                else:
                    call("bsl_conv2d_head_dgrad", C.byref(d), self.dlogits.p, self._pp(self.W, f"{L.scope}/weights"),
                         cur.p, s)
                self._after_grad(L)
                continue
            if L.kind in ("stem", "conv"):
                if L.pooled is not None:
                    # `cur` is the gradient w.r.t. the POOLED tensor; merge MaxPoolGrad with the skip gradient
                    dc = self.dcat[L.level]
                    call("bsl_maxpool2x2_bwd_add", C.c_int(n), C.c_int(L.h), C.c_int(L.w), C.c_int(L.cout), L.a.p,
                         C.c_int(L.a.ld), cur.p, C.c_int(L.pooled.ld), dc.p, C.c_int(dc.ld), alt.p, C.c_int(L.cout), s)
                    cur, alt = alt, cur
                dyb = self.dyb[k]
                nd = self._norm_desc(L)
                q = self._norm_ptrs(L)
                # pipelined: the apply half runs on the auxiliary stream beside dgrad(L), which follows it image slice
                # by image slice and writes into `alt` (its output must not overwrite the gradient apply still reads)
                pipe_b = overlap and piping and L.kind == "conv" and self._pipe_shape_ok(L)
                self._norm_backward_reduce(L, nd, q, cur)
                sig, ev_apply = None, None
                if pipe_b:
                    sig, aux = self._pipe(len(self.layers) + idx), self.aux_stream
                    ev = self._next_event()
                    ctx.record(ev, s)
                    call("bsl_stream_wait_event", aux, ev)
                    if self._dy_busy[k] is not None:
                        # on the MAIN stream too: dgrad(L) must not occupy the SMs (spinning on slices) while the
                        # filter-gradient kernel that still reads this dY buffer waits for shared memory
                        call("bsl_stream_wait_event", aux, self._dy_busy[k])
                        call("bsl_stream_wait_event", s, self._dy_busy[k])
                    self._norm_backward_apply(L, nd, q, cur, dyb, aux, sig)
                    ev_apply = self._next_event()
                    ctx.record(ev_apply, aux)
                else:
                    if overlap and self._dy_busy[k] is not None:   # the wgrad that read this buffer two layers ago
                        call("bsl_stream_wait_event", s, self._dy_busy[k])
                    self._norm_backward_apply(L, nd, q, cur, dyb, s)
                # dyb = dY (gradient w.r.t. the conv output), dense with ld = cout
                d = self._conv_desc(L)
                d.y_ld = L.cout
                gw = self._pp(self.G, f"{L.scope}/weights")
                # dgrad first, on the main stream; the side stream forks AFTER it, so that the filter gradient of this
                # layer runs beside the HBM-bound normalisation backward of the next one instead of beside dgrad
                if self._fork_pre:
                    fork()
                if L.kind != "stem" and idx > 0:     # no gradient w.r.t. the network input
                    wbf = self._pp(self.Wbf, f"{L.scope}/weights", BF16)
                    dd = self._conv_desc(L)
                    dd.y_ld = L.cout
                    wp = C.byref(sig) if sig is not None else None
                    if L.role == "dec1":
                        dc = self.dcat[L.level]
                        dd.x_ld = dc.ld
                        if self._fuse_relu_bwd and self._pipe_shape_ok(L):
                            # ReluGrad of the transposed conv (upper channel half of the concat) in the epilogue
                            up = self.layers[idx - 1]
                            assert up.kind == "convT" and up.level == L.level
                            self._tc("dgrad", self._flops(L), "bsl_conv2d_dgrad_relu", C.byref(dd), dyb.p, wbf, dc.p,
                                     self.cat[L.level].p, C.c_int(L.cin - up.cout), wp, s)
                            relu_fused.add(L.level)
                        else:
                            self._tc("dgrad", self._flops(L), "bsl_conv2d_dgrad_pipe", C.byref(dd), dyb.p, wbf, dc.p, wp, s)
                    else:
                        dd.x_ld = L.cin_dev or L.cin
                        dst = alt if pipe_b else cur
                        self._tc("dgrad", self._flops(L), "bsl_conv2d_dgrad_pipe", C.byref(dd), dyb.p, wbf, dst.p, wp, s)
                        if pipe_b:
                            cur, alt = alt, cur
                if ev_apply is not None:
                    call("bsl_stream_wait_event", s, ev_apply)
                    call("bsl_stream_wait_event", ws, ev_apply)
                if not self._fork_pre:
                    fork()
                if L.kind == "stem":
                    d1 = _lib.Conv2dDesc(n, L.h, L.w, 64, L.cout, 1, 1, self.stem_cols, L.cout)
                    self._tc("wgrad", self._flops(L), "bsl_conv2d_wgrad", C.byref(d1), self.stem_col.p, dyb.p, gw,
                             self.wgrad_ws.p, wsb, ws, stream=ws)
                else:
                    self._tc("wgrad", self._flops(L), "bsl_conv2d_wgrad", C.byref(d), L.x.p, dyb.p, gw,
                             self.wgrad_ws.p, wsb, ws, stream=ws)
                if overlap:
                    self._dy_busy[k] = self._dy_events[k]
                    ctx.record(self._dy_busy[k], ws)
                k ^= 1
                self._after_grad(L)
            elif L.kind == "convT":
                dc = self.dcat[L.level]
                dup = dc.slice(L.cout, L.cout)
                # ReluGrad in place on the upper half of dcat (unless dgrad's epilogue already applied it)
                gb = self._pp(self.G, f"{L.scope}/biases")
                if L.level not in relu_fused:
                    # ... and the bias gradient (sum over pixels of the masked gradient) from the same pass
                    call("bsl_relu_bwd_bias", C.c_longlong(L.a.pixels), C.c_int(L.cout), L.a.p, C.c_int(L.a.ld), dup.p,
                         C.c_int(dup.ld), dup.p, C.c_int(dup.ld), gb, s)
                    gb = None
                d = self._convT_desc(L)
                d.y_ld = dup.ld
                d.x_ld = L.cin
                self._tc("convT_dgrad", self._flops(L), "bsl_convT2d_bwd_data", C.byref(d), dup.p,
                         self._pp(self.Wbf, f"{L.scope}/weights", BF16), cur.p, s)
                fork()
                d.x_ld = L.x.ld
                self._tc("convT_wgrad", self._flops(L), "bsl_convT2d_bwd_filter", C.byref(d), L.x.p, dup.p,
                         self._pp(self.G, f"{L.scope}/weights"), gb,
                         self.wgrad_ws.p, wsb, ws, stream=ws)
                self._after_grad(L)
        if overlap:     # join: everything downstream (optimizer, host reads of G) is ordered after the filter gradients
            ev = self._next_event()
            ctx.record(ev, ws)
            call("bsl_stream_wait_event", s, ev)
            self._dy_busy = [None, None]

    # ------------------------------------------------------------------ optimizer
    def optimizer_step(self, lr: float):
        self._pdl(True)
        self.step_count += 1
        enqueue_optimizer(self, lr)

    # ------------------------------------------------------------------ results
    def read_loss(self):
        """(data loss, L2 regularisation loss) of the step just run; the only D2H read of a train step."""
        data = float(self.loss_dev.download(np.float32, (len(self._loss_terms()),)).sum())
        sq = self.sumsq.download(np.float64, (1,))[0]
        return float(data), float(self.cfg.weight_decay_rate * 0.5 * sq) if self.cfg.weight_decay_rate > 0 else 0.0

    def read_counts(self):
        k = self.cfg.num_classes - 1
        return self.ilr.download(np.uint32, (self.cfg.batch, k, 3))

    def train_step(self, lr: float, with_metrics: bool = False):
        """One `sess.run(train_op)` on device-resident inputs (/root/reference/core/estimator.py:756-757)."""
        self.forward(True)
        if self.cfg.normalizer == "batch_norm":
            self._allreduce_moving_stats()
        if with_metrics:
            self.predict_outputs(True)
        self.loss_backward()
        self._allreduce_grads()
        self.optimizer_step(lr)

    # ------------------------------------------------------------------ host-fed step (end-to-end path)
    def pinned_inputs(self):
        """numpy views of page-locked staging buffers (images fp32 [N,H,W,C], labels int32 [N,H,W])."""
        if not hasattr(self, "_pin"):
            cfg = self.cfg
            shp_i = (cfg.batch, cfg.height, cfg.width, cfg.channel)
            shp_l = (cfg.batch, cfg.height, cfg.width)
            pi, pl, po = C.c_void_p(), C.c_void_p(), C.c_void_p()
            self.ctx.call("bsl_host_alloc", C.c_size_t(int(np.prod(shp_i)) * 4), C.byref(pi))
            self.ctx.call("bsl_host_alloc", C.c_size_t(int(np.prod(shp_l)) * 4), C.byref(pl))
            self.ctx.call("bsl_host_alloc", C.c_size_t(64), C.byref(po))
            img = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_float)), shape=shp_i)
            lab = np.ctypeslib.as_array(C.cast(pl, C.POINTER(C.c_int32)), shape=shp_l)
            out = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_double)), shape=(8,))
            self._pin = (img, lab, out, pi, pl, po)
        return self._pin[0], self._pin[1]

    def train_step_host(self, lr: float, with_metrics: bool = False):
        """H2D copy of the staged batch, one training step, D2H read of (loss, sum w^2). Returns total loss.

        This is the call a user of the reference makes per step (feed a batch, fetch the loss)."""
        img, lab, out, pi, pl, po = self._pin
        s = self.stream
        self.ctx.call("bsl_memcpy_h2d", self.images.p, pi, C.c_size_t(img.nbytes), s)
        self.ctx.call("bsl_memcpy_h2d", self.labels.p, pl, C.c_size_t(lab.nbytes), s)
        self.train_step(lr, with_metrics)
        nt = len(self._loss_terms())
        self.ctx.call("bsl_memcpy_d2h", po, self.loss_dev.p, C.c_size_t(4 * nt), s)
        self.ctx.call("bsl_memcpy_d2h", C.c_void_p(po.value + 8), self.sumsq.p, C.c_size_t(8), s)
        self.ctx.sync(s)
        data = float(np.frombuffer(out[:1].tobytes(), np.float32)[:nt].sum())
        reg = float(self.cfg.weight_decay_rate * 0.5 * out[1]) if self.cfg.weight_decay_rate > 0 else 0.0
        return data + reg

    # ---- prefetching variant: a 2-deep ring of pinned staging slots + device input buffers. The H2D copy of batch
    # i + 1 runs on a copy stream while step i computes (what tf.data prefetch / StagingArea does for the reference's
    # feed path); every step still pays one H2D of its own inputs and one D2H of its loss.
    def staging_slot(self, j: int):
        """numpy views (images, labels) of pinned slot j in {0, 1}. Do not refill a slot between submit_staged(j) and
        the train_step_prefetched() that consumes it."""
        if not hasattr(self, "_ring"):
            cfg = self.cfg
            shp_i = (cfg.batch, cfg.height, cfg.width, cfg.channel)
            shp_l = (cfg.batch, cfg.height, cfg.width)
            self._ring = []
            for k in range(2):
                pi, pl = C.c_void_p(), C.c_void_p()
                self.ctx.call("bsl_host_alloc", C.c_size_t(int(np.prod(shp_i)) * 4), C.byref(pi))
                self.ctx.call("bsl_host_alloc", C.c_size_t(int(np.prod(shp_l)) * 4), C.byref(pl))
                img = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_float)), shape=shp_i)
                lab = np.ctypeslib.as_array(C.cast(pl, C.POINTER(C.c_int32)), shape=shp_l)
                dimg = self.images if k == 0 else self._alloc(img.nbytes)
                dlab = self.labels if k == 0 else self._alloc(lab.nbytes)
                self._ring.append(dict(img=img, lab=lab, pi=pi, pl=pl, dimg=dimg, dlab=dlab, ev=self.ctx.new_event()))
            self.copy_stream = self.ctx.new_stream()
            self._submitted = []
            self.pinned_inputs()   # the D2H slot for the loss
        return self._ring[j]["img"], self._ring[j]["lab"]

    def submit_staged(self, j: int):
        """Enqueue the H2D copy of pinned slot j into device input buffer j on the copy stream."""
        r = self._ring[j]
        cs = self.copy_stream
        self.ctx.call("bsl_memcpy_h2d", r["dimg"].p, r["pi"], C.c_size_t(r["img"].nbytes), cs)
        self.ctx.call("bsl_memcpy_h2d", r["dlab"].p, r["pl"], C.c_size_t(r["lab"].nbytes), cs)
        self.ctx.record(r["ev"], cs)
        self._submitted.append(j)

    def train_step_prefetched(self, lr: float, with_metrics: bool = False):
        """One training step on the oldest submitted slot; returns the total loss (D2H read, host sync)."""
        j = self._submitted.pop(0)
        r = self._ring[j]
        s = self.stream
        self.ctx.call("bsl_stream_wait_event", s, r["ev"])
        self.images, self.labels = r["dimg"], r["dlab"]
        self.train_step(lr, with_metrics)
        out, po = self._pin[2], self._pin[5]
        nt = len(self._loss_terms())
        self.ctx.call("bsl_memcpy_d2h", po, self.loss_dev.p, C.c_size_t(4 * nt), s)
        self.ctx.call("bsl_memcpy_d2h", C.c_void_p(po.value + 8), self.sumsq.p, C.c_size_t(8), s)
        self.ctx.sync(s)
        data = float(np.frombuffer(out[:1].tobytes(), np.float32)[:nt].sum())
        reg = float(self.cfg.weight_decay_rate * 0.5 * out[1]) if self.cfg.weight_decay_rate > 0 else 0.0
        return data + reg

    def h2d_bytes_per_step(self):
        cfg = self.cfg
        return cfg.batch * cfg.height * cfg.width * (cfg.channel * 4 + 4)

    def close(self):
        for b in self._bufs:
            b.free()
        self._bufs = []
