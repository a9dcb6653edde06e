"""Mirror of /root/reference/core/solver.py: learning-rate policies and the optimizer choice.

`Solver(args)(loss)` returns a TrainOp; running it is one `sess.run(train_op)`: forward, loss,
backward, gradient all-reduce, fused optimizer step (BN moving statistics are updated inside the
forward pass, which is the UPDATE_OPS control dependency of solver.py:236-239)."""
from __future__ import annotations


def add_arguments(parser):
    """solver.py:23-82 -- same flags, same defaults."""
    g = parser.add_argument_group(title="Training Arguments")
    g.add_argument("--learning_rate", type=float, default=1e-3)
    g.add_argument("--learning_policy", type=str, default="period_step",
                   choices=["custom_step", "period_step", "poly", "plateau"])
    g.add_argument("--num_of_steps", type=int, default=0)
    g.add_argument("--num_of_total_steps", type=int, default=1000)
    g.add_argument("--lr_decay_boundaries", type=int, nargs="*")
    g.add_argument("--lr_custom_values", type=float, nargs="+")
    g.add_argument("--lr_decay_step", type=int, default=int(1e5))
    g.add_argument("--lr_decay_rate", type=float, default=0.1)
    g.add_argument("--lr_power", type=float, default=0.9)
    g.add_argument("--lr_end", type=float, default=1e-6)
    g.add_argument("--optimizer", type=str, default="Adam", choices=["Adam", "Momentum", "AdamW"])
    g.add_argument("--lr_warm_up", action="store_true")
    g.add_argument("--slow_start_step", type=int, default=1000)
    g.add_argument("--slow_start_lr", type=float, default=1e-4)
    g.add_argument("--adam_beta1", type=float)
    g.add_argument("--adam_beta2", type=float)
    g.add_argument("--adam_eps", type=float)
    g.add_argument("--mm_mm", type=float)
    g.add_argument("--mm_nesterov", action="store_true")
    g.add_argument("--lr_patience", type=int, default=30)


def optimizer_params_from_args(args):
    """get_solver_params (solver.py:84-96): the dict handed to Solver(optimizer_params=...), or None.
    Falsy values (0, None) are skipped exactly as the reference's `if args.adam_beta1:` tests do."""
    p = {}
    if getattr(args, "adam_beta1", None):
        p["beta1"] = args.adam_beta1
    if getattr(args, "adam_beta2", None):
        p["beta2"] = args.adam_beta2
    if getattr(args, "adam_eps", None):
        p["epsilon"] = args.adam_eps
    if getattr(args, "mm_mm", None):
        p["momentum"] = args.mm_mm
    if getattr(args, "mm_nesterov", False):
        p["use_nesterov"] = True
    return p or None


def engine_optimizer_kwargs(args) -> dict:
    """Engine-config fields equivalent to Solver._get_model_optimizer (solver.py:204-219).

    The reference's `self.optimizer_params or {defaults}` means ANY explicit flag replaces the whole default dict: with
    only --adam_beta1 given, beta2 falls back to TensorFlow's 0.999 (not the repo's 0.99). Flags that do not belong to
    the chosen optimizer reach its constructor as unexpected keywords, and a Momentum / AdamW optimizer built from a
    flag dict lacks its required `momentum` / `weight_decay` argument: both are TypeErrors in the reference and here."""
    name = getattr(args, "optimizer", "Adam").lower()
    params = optimizer_params_from_args(args)
    allowed = {"adam": {"beta1", "beta2", "epsilon"}, "momentum": {"momentum", "use_nesterov"},
               "adamw": {"beta1", "beta2", "epsilon", "weight_decay", "learning_rate"}}
    ctor = {"adam": "AdamOptimizer", "momentum": "MomentumOptimizer", "adamw": "AdamWOptimizer"}
    if name not in allowed:
        raise ValueError("Not supported optimizer: " + name)
    out = {"optimizer": name}
    if params is None:
        if name in ("adam", "adamw"):
            out.update(adam_beta1=0.9, adam_beta2=0.99, adam_eps=1e-8)       # solver.py:206,212-214
        if name == "momentum":
            out.update(momentum=0.9, use_nesterov=False)                      # solver.py:209
        if name == "adamw":
            out.update(adamw_weight_decay=getattr(args, "weight_decay_rate", 0.0) or 0.0)
        return out
    bad = sorted(set(params) - allowed[name])
    if bad:
        raise TypeError("{}.__init__() got an unexpected keyword argument '{}'".format(ctor[name], bad[0]))
    if name == "momentum":
        if "momentum" not in params:
            raise TypeError("MomentumOptimizer.__init__() missing 1 required positional argument: 'momentum'")
        out.update(momentum=params["momentum"], use_nesterov=params.get("use_nesterov", False))
    elif name == "adam":     # TensorFlow's own defaults for what the dict leaves out
        out.update(adam_beta1=params.get("beta1", 0.9), adam_beta2=params.get("beta2", 0.999),
                   adam_eps=params.get("epsilon", 1e-8))
    else:
        raise TypeError("AdamWOptimizer.__init__() missing 1 required positional argument: 'weight_decay'")
    return out


class TrainOp:
    def __init__(self, solver, loss):
        self.solver, self.loss = solver, loss

    def run(self, with_metrics: bool = False, host_fed: bool = False):
        """Enqueue one training step; returns nothing (read `loss.value` via Solver.fetch_loss)."""
        s = self.solver
        eng = self.loss.model.engine
        lr = s.learning_rate(s.global_step)
        if host_fed:
            self.loss.value = eng.train_step_host(lr, with_metrics)
        else:
            eng.train_step(lr, with_metrics)
            self.loss.value = None
        s.global_step += 1
        return lr

    def fetch_loss(self):
        if self.loss.value is None:
            self.loss.value = sum(self.loss.model.engine.read_loss())
        return self.loss.value


class Solver(object):
    def __init__(self, args, name=None):
        self._args = args
        self.name = name or "Optimizer"
        self.learning_policy = args.learning_policy
        self.base_learning_rate = args.learning_rate
        self.learning_rate_decay_step = args.lr_decay_step
        self.learning_rate_decay_rate = args.lr_decay_rate
        self.num_of_total_steps = args.num_of_total_steps
        self.learning_power = args.lr_power
        self.end_learning_rate = args.lr_end
        self.learning_rate_decay_boundaries = args.lr_decay_boundaries
        self.learning_rate_custom_values = args.lr_custom_values
        if self.learning_policy == "custom_step":
            b, v = self.learning_rate_decay_boundaries or [], self.learning_rate_custom_values or []
            if len(v) - len(b) != 1:
                raise ValueError("Make sure len(lr_custom_values) - len(lr_decay_boundaries) = 1")
        self.optimizer = args.optimizer.lower()
        self.optimizer_params = optimizer_params_from_args(args)
        self.engine_kwargs = engine_optimizer_kwargs(args)     # raises where the reference's constructor would
        self.slow_start_step = args.slow_start_step if getattr(args, "lr_warm_up", False) else 0
        self.slow_start_lr = args.slow_start_lr
        self.global_step = 0
        self._plateau_lr = float(self.base_learning_rate)   # plateau_decay variable (solver.py:246-254)

    @property
    def args(self):
        return self._args

    def learning_rate(self, step: int) -> float:
        """_get_model_learning_rate (solver.py:138-202) evaluated at `step`."""
        p = self.learning_policy
        if p == "period_step":   # tf.train.exponential_decay(staircase=True)
            lr = self.base_learning_rate * self.learning_rate_decay_rate ** (step // int(self.learning_rate_decay_step))
        elif p == "custom_step":  # tf.train.piecewise_constant: values[i] while x <= boundaries[i]
            lr = self.learning_rate_custom_values[-1]
            for b, v in zip(self.learning_rate_decay_boundaries, self.learning_rate_custom_values):
                if step <= b:
                    lr = v
                    break
        elif p == "poly":         # tf.train.polynomial_decay(cycle=False)
            s = min(step, self.num_of_total_steps)
            lr = (self.base_learning_rate - self.end_learning_rate) * \
                 (1 - s / self.num_of_total_steps) ** self.learning_power + self.end_learning_rate
        elif p == "plateau":
            lr = self._plateau_lr
        else:
            raise ValueError("Not supported learning policy.")
        if self.slow_start_step > 0 and step < self.slow_start_step:
            lr = self.slow_start_lr
        return float(lr)

    def plateau_update(self):
        """update_lr_op: lr <- max(lr * factor, min_lr) (solver.py:251), run by ReduceLROnPlateauHook."""
        self._plateau_lr = max(self._plateau_lr * self.learning_rate_decay_rate, self.end_learning_rate)
        return self._plateau_lr

    def __call__(self, loss, *args, **kwargs):
        if "slow_start_step" in kwargs:
            self.slow_start_step = kwargs.pop("slow_start_step")
        if "slow_start_learning_rate" in kwargs:
            self.slow_start_lr = kwargs.pop("slow_start_learning_rate")
        if self.optimizer not in ("adam", "momentum", "adamw"):
            raise ValueError("Not supported optimizer: " + self.optimizer)
        ecfg = loss.model.engine.cfg
        for k, v in self.engine_kwargs.items():
            if getattr(ecfg, k) != v:
                raise ValueError("engine was planned with %s=%r, the solver asks for %r" % (k, getattr(ecfg, k), v))
        return TrainOp(self, loss)
