"""Does replaying the step as a CUDA graph shorten it? Forward + loss + backward (both streams) of the cfg2 UNet, eager vs
captured (bsl_graph_*), CUDA events over `--steps` repetitions. The optimizer stays outside (its step-dependent scalars are
kernel arguments). Usage: python tools/graph_probe.py [--batch 64] [--hw 256] [--steps 10]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxsegliver_b200 import synthetic  # noqa: E402
from boxsegliver_b200.device import Context  # noqa: E402
from boxsegliver_b200.engine import EngineConfig, UNetEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--hw", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
ctx = Context(0)
cfg = EngineConfig(batch=a.batch, height=a.hw, width=a.hw, loss_weight_type="numerical", loss_numeric_w=(0.2, 0.4, 4.4),
                   weight_decay_rate=1e-5)
eng = UNetEngine(ctx, cfg)
eng.init_weights(0)
im, lb = synthetic.make_batch(a.batch, a.hw, a.hw, 3)
eng.set_inputs(im, lb)


def body():
    eng.forward(True)
    eng.loss_backward()


def timed(fn):
    for _ in range(3):
        fn()
    ctx.sync()
    e0, e1 = ctx.new_event(), ctx.new_event()
    ctx.record(e0)
    for _ in range(a.steps):
        fn()
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1) / a.steps
    ctx.check_device()
    return ms


eager = timed(body)
g0 = eng.get_grads()
ctx.call("bsl_graph_begin", ctx.stream)
body()
ge = C.c_void_p()
ctx.call("bsl_graph_end", ctx.stream, C.byref(ge))
graph = timed(lambda: ctx.call("bsl_graph_launch", ge, ctx.stream))
g1 = eng.get_grads()
same = all((g0[k] == g1[k]).all() for k in g0)
print(f"fwd+bwd eager {eager:.3f} ms, graph {graph:.3f} ms, gradients bit-identical: {same}")
eager2 = timed(body)
print(f"eager again {eager2:.3f} ms")
