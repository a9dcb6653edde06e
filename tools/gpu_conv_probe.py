"""GPU probe for the tcgen05 implicit-GEMM kernels: parity vs the numpy oracle + first timings.

Run on the B200 box: `python tools/gpu_conv_probe.py [--time]`. Prints one line per case:
  <op> <shape> rel=<||a-b||/||b||> max=<max abs err> OK|FAIL
Exit code 0 iff every parity case passed. Used to bring the kernels up; the pytest -m gpu suite
covers the same ground for the driver.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from boxsegliver_b200 import _lib  # noqa: E402
from boxsegliver_b200.device import Context, round_bf16  # noqa: E402
from oracle import tf_ops as O  # noqa: E402

TOL = 1e-2  # rel L2 for bf16 operands / fp32 accumulation / bf16 output (north_star: rel <= 1e-2 bf16)


def rel(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)), float(np.abs(a - b).max())


def conv_case(ctx, n, h, w, cin, cout, k=3, x_ld=None, y_ld=None, seed=0):
    rng = np.random.default_rng(seed)
    x_ld = x_ld or cin
    y_ld = y_ld or cout
    x = round_bf16(rng.standard_normal((n, h, w, cin), dtype=np.float32))
    wt = round_bf16(rng.standard_normal((k, k, cin, cout), dtype=np.float32) * 0.05)
    dy = round_bf16(rng.standard_normal((n, h, w, cout), dtype=np.float32))
    xbuf = np.zeros((n, h, w, x_ld), np.float32); xbuf[..., :cin] = x
    dybuf = np.zeros((n, h, w, y_ld), np.float32); dybuf[..., :cout] = dy
    dx_ = ctx.bf16_from_f32(xbuf)
    dw_ = ctx.bf16_from_f32(wt)
    ddy = ctx.bf16_from_f32(dybuf)
    dyo = ctx.alloc(n * h * w * y_ld * 2).zero()
    ddx = ctx.alloc(n * h * w * x_ld * 2).zero()
    desc = _lib.Conv2dDesc(n, h, w, cin, cout, k, k, x_ld, y_ld)
    res = {}
    # fprop
    ctx.call("bsl_conv2d_fprop", C.byref(desc), dx_.p, dw_.p, dyo.p, ctx.stream)
    ctx.check_device()
    got = ctx.bf16_to_f32(dyo, (n, h, w, y_ld))[..., :cout]
    res["fprop"] = rel(got, O.conv2d(x.astype(np.float64), wt.astype(np.float64)))
    # dgrad
    ctx.call("bsl_conv2d_dgrad", C.byref(desc), ddy.p, dw_.p, ddx.p, ctx.stream)
    ctx.check_device()
    got = ctx.bf16_to_f32(ddx, (n, h, w, x_ld))[..., :cin]
    res["dgrad"] = rel(got, O.conv2d_backprop_input(x.shape, wt.astype(np.float64), dy.astype(np.float64)))
    # wgrad
    ws_bytes = ctx.lib.bsl_conv2d_wgrad_workspace(ctx.h, C.byref(desc))
    ws = ctx.alloc(max(ws_bytes, 16))
    dwo = ctx.alloc(k * k * cin * cout * 4).zero()
    ctx.call("bsl_conv2d_wgrad", C.byref(desc), dx_.p, ddy.p, dwo.p, ws.p, C.c_size_t(ws_bytes), ctx.stream)
    ctx.check_device()
    got = dwo.download(np.float32, (k, k, cin, cout))
    res["wgrad"] = rel(got, O.conv2d_backprop_filter(x.astype(np.float64), wt.shape, dy.astype(np.float64)))
    for b in (dx_, dw_, ddy, dyo, ddx, ws, dwo):
        b.free()
    return res


def convT_case(ctx, n, h, w, cin, cout, y_ld=None, seed=1):
    rng = np.random.default_rng(seed)
    y_ld = y_ld or cout
    x = round_bf16(rng.standard_normal((n, h, w, cin), dtype=np.float32))
    wt = round_bf16(rng.standard_normal((2, 2, cout, cin), dtype=np.float32) * 0.05)
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    dy = round_bf16(rng.standard_normal((n, 2 * h, 2 * w, cout), dtype=np.float32))
    dybuf = np.zeros((n, 2 * h, 2 * w, y_ld), np.float32); dybuf[..., :cout] = dy
    dx_ = ctx.bf16_from_f32(x)
    dw_ = ctx.bf16_from_f32(wt)
    db_ = ctx.from_numpy(bias)
    ddy = ctx.bf16_from_f32(dybuf)
    dyo = ctx.alloc(n * 4 * h * w * y_ld * 2).zero()
    ddx = ctx.alloc(n * h * w * cin * 2).zero()
    desc = _lib.ConvT2dDesc(n, h, w, cin, cout, cin, y_ld, 1)
    res = {}
    ctx.call("bsl_convT2d_fwd", C.byref(desc), dx_.p, dw_.p, db_.p, dyo.p, ctx.stream)
    ctx.check_device()
    got = ctx.bf16_to_f32(dyo, (n, 2 * h, 2 * w, y_ld))[..., :cout]
    ref = O.relu(O.conv2d_transpose(x.astype(np.float64), wt.astype(np.float64)) + bias)
    res["convT_fwd"] = rel(got, ref)
    ctx.call("bsl_convT2d_bwd_data", C.byref(desc), ddy.p, dw_.p, ddx.p, ctx.stream)
    ctx.check_device()
    got = ctx.bf16_to_f32(ddx, (n, h, w, cin))
    rdx, rdw = O.conv2d_transpose_grad(x.astype(np.float64), wt.astype(np.float64), dy.astype(np.float64))
    res["convT_bwd_data"] = rel(got, rdx)
    ws_bytes = ctx.lib.bsl_convT2d_bwd_filter_workspace(ctx.h, C.byref(desc))
    ws = ctx.alloc(max(ws_bytes, 16))
    dwo = ctx.alloc(4 * cout * cin * 4).zero()
    dbo = ctx.alloc(cout * 4).zero()
    ctx.call("bsl_convT2d_bwd_filter", C.byref(desc), dx_.p, ddy.p, dwo.p, dbo.p, ws.p, C.c_size_t(ws_bytes),
             ctx.stream)
    ctx.check_device()
    res["convT_bwd_filter"] = rel(dwo.download(np.float32, (2, 2, cout, cin)), rdw)
    res["convT_dbias"] = rel(dbo.download(np.float32, (cout,)), dy.astype(np.float64).sum(axis=(0, 1, 2)))
    for b in (dx_, dw_, db_, ddy, dyo, ddx, ws, dwo, dbo):
        b.free()
    return res


def time_conv(ctx, n, h, w, cin, cout, iters=20):
    """Device-timed fprop/dgrad/wgrad at a bench-size layer; returns TFLOP/s per op (algorithmic FLOPs)."""
    rng = np.random.default_rng(3)
    x = ctx.bf16_from_f32(rng.standard_normal((n, h, w, cin), dtype=np.float32))
    wt = ctx.bf16_from_f32(rng.standard_normal((3, 3, cin, cout), dtype=np.float32) * 0.05)
    dy = ctx.bf16_from_f32(rng.standard_normal((n, h, w, cout), dtype=np.float32))
    y = ctx.alloc(n * h * w * cout * 2)
    dx = ctx.alloc(n * h * w * cin * 2)
    desc = _lib.Conv2dDesc(n, h, w, cin, cout, 3, 3, cin, cout)
    ws_bytes = ctx.lib.bsl_conv2d_wgrad_workspace(ctx.h, C.byref(desc))
    ws = ctx.alloc(max(ws_bytes, 16))
    dw = ctx.alloc(9 * cin * cout * 4)
    flops = 2.0 * n * h * w * 9 * cin * cout
    out = {}
    ops = {
        "fprop": lambda: ctx.call("bsl_conv2d_fprop", C.byref(desc), x.p, wt.p, y.p, ctx.stream),
        "dgrad": lambda: ctx.call("bsl_conv2d_dgrad", C.byref(desc), dy.p, wt.p, dx.p, ctx.stream),
        "wgrad": lambda: ctx.call("bsl_conv2d_wgrad", C.byref(desc), x.p, dy.p, dw.p, ws.p,
                                  C.c_size_t(ws_bytes), ctx.stream),
    }
    e0, e1 = ctx.new_event(), ctx.new_event()
    for name, fn in ops.items():
        for _ in range(3):
            fn()
        ctx.sync()
        ctx.record(e0)
        for _ in range(iters):
            fn()
        ctx.record(e1)
        ms = ctx.elapsed_ms(e0, e1) / iters
        ctx.check_device()
        out[name] = {"ms": ms, "tflops": flops / ms / 1e9}
    for b in (x, wt, dy, y, dx, ws, dw):
        b.free()
    return out


def main():
    do_time = "--time" in sys.argv
    ctx = Context(0)
    print(ctx.lib.bsl_version().decode())
    failures = 0
    report = {}

    def run(name, fn, *a, **kw):
        nonlocal failures
        t0 = time.time()
        try:
            res = fn(ctx, *a, **kw)
        except Exception as e:  # noqa: BLE001
            print(f"{name} {a} EXC {e!r}")
            failures += 1
            report[f"{name}{a}"] = {"exc": repr(e)}
            return
        for op, (r, m) in res.items():
            ok = r <= TOL
            failures += 0 if ok else 1
            print(f"{op:18s} {str(a):34s} {kw or ''} rel={r:.3e} max={m:.3e} {'OK' if ok else 'FAIL'}"
                  f" ({time.time() - t0:.1f}s)")
            report[f"{op}{a}{kw}"] = {"rel": r, "max": m, "ok": ok}

    cases = [
        (2, 16, 16, 64, 64),      # BN=64, single k-block per tap
        (1, 16, 16, 128, 128),    # BN=128, two k-blocks
        (1, 8, 16, 64, 256),      # BN=256
        (1, 16, 16, 256, 64),
        (2, 12, 20, 64, 128),     # ragged spatial size: OOB rows masked
        (3, 4, 4, 128, 64),       # tiny spatial: box spans several images
        (1, 32, 32, 64, 64),
    ]
    for c in cases:
        run("conv", conv_case, *c)
    run("conv", conv_case, 2, 16, 16, 64, 64, x_ld=128, y_ld=192)   # concat-slice views
    run("conv1x1", conv_case, 2, 16, 16, 128, 64, k=1)
    for c in [(2, 8, 8, 128, 64), (1, 16, 16, 64, 64), (2, 4, 8, 256, 128), (1, 6, 10, 128, 64)]:
        run("convT", convT_case, *c)
    run("convT", convT_case, 2, 8, 8, 128, 64, y_ld=128)

    if failures:
        # Which MN-major descriptor convention does the hardware want? Try the alternatives once.
        for lbo, sbo, kadv in [(1024, 8192, 2048), (8192, 1024, 1024), (1024, 8192, 1024)]:
            print(f"--- retry with mn_lbo={lbo} mn_sbo={sbo} mn_kadv={kadv}")
            ctx.call("bsl_debug_set", C.c_int(0), C.c_int(lbo))
            ctx.call("bsl_debug_set", C.c_int(1), C.c_int(sbo))
            ctx.call("bsl_debug_set", C.c_int(2), C.c_int(kadv))
            try:
                for k_, (r, m) in conv_case(ctx, 1, 16, 16, 128, 128).items():
                    print(f"    {k_} rel={r:.3e}")
            except Exception as e:  # noqa: BLE001
                print("    EXC", e)
        ctx.call("bsl_debug_set", C.c_int(0), C.c_int(8192))
        ctx.call("bsl_debug_set", C.c_int(1), C.c_int(1024))
        ctx.call("bsl_debug_set", C.c_int(2), C.c_int(2048))

    if do_time:
        for shp in [(16, 128, 128, 128, 128), (16, 64, 64, 256, 256), (64, 16, 16, 1024, 1024),
                    (8, 256, 256, 64, 64), (16, 32, 32, 512, 512), (8, 256, 256, 128, 64)]:
            try:
                t = time_conv(ctx, *shp)
                print("time", shp, json.dumps(t))
                report[f"time{shp}"] = t
            except Exception as e:  # noqa: BLE001
                print("time", shp, "EXC", e)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/conv_probe.json", "w") as f:
        json.dump(report, f, indent=1)
    print("FAILURES", failures)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
