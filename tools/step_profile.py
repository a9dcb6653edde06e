"""Runs W warm-up + K training steps of one engine at the BASELINE shapes (for ncu launch lists / --set full captures).
The K measured steps sit between cudaProfilerStart / cudaProfilerStop, so `ncu --profile-from-start off` sees only them.

  python tools/step_profile.py [--model unet|unet3d|gunet] [--batch N] [--hw 256] [--steps 1] [--warmup 2]
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxsegliver_b200 import synthetic  # noqa: E402
from boxsegliver_b200.device import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="unet")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--hw", type=int, default=0)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
ctx = Context(0)
if a.model == "unet3d":
    from boxsegliver_b200.unet3d_engine import UNet3DConfig, UNet3DEngine
    n = a.batch or 4
    eng = UNet3DEngine(ctx, UNet3DConfig(batch=n, depth=64, height=128, width=128, loss_numeric_w=(1.0, 1.0)))
    eng.init_weights(0)
    eng.set_inputs(*synthetic.make_volume_batch(n, 64, 128, 128))
    lr = 3e-4
elif a.model == "gunet":
    import numpy as np
    from boxsegliver_b200.gunet_engine import GUNetConfig, GUNetEngine
    n, hw = a.batch or 32, a.hw or 512
    eng = GUNetEngine(ctx, GUNetConfig(batch=n, height=hw, width=hw, loss_type="xentropy+dice", loss_weight_type="numerical",
                                       loss_numeric_w=(0.2, 0.4, 4.4), weight_decay_rate=1e-5, guide_channel=1))
    eng.init_weights(0)
    im, lb = synthetic.make_batch(4, hw, hw, 3)
    im, lb = np.tile(im, (n // 4, 1, 1, 1)), np.tile(lb, (n // 4, 1, 1))
    eng.set_inputs(im, lb)
    eng.set_guides(*synthetic.make_guides(im, lb, 200, 1))
    lr = 1e-3
else:
    from boxsegliver_b200.engine import EngineConfig, UNetEngine
    n, hw = a.batch or 64, a.hw or 256
    eng = UNetEngine(ctx, EngineConfig(batch=n, height=hw, width=hw, loss_weight_type="numerical",
                                       loss_numeric_w=(0.2, 0.4, 4.4)))
    eng.init_weights(0)
    eng.set_inputs(*synthetic.make_batch(n, hw, hw, 3))
    lr = 1e-3
for _ in range(a.warmup):
    eng.train_step(lr)
ctx.sync()
try:
    rt = ctypes.CDLL("libcudart.so.12")
except OSError:
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
rt.cudaProfilerStart()
for _ in range(a.steps):
    eng.train_step(lr)
ctx.sync()
rt.cudaProfilerStop()
ctx.check_device()
print("loss", eng.read_loss())
