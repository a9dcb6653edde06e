"""Per-tensor gradient errors of the UNet3D engine (pixel-pair packing on/off) against the oracle's backward pass over
the device's stored forward tape. Usage: python tools/pair_debug.py [n d hw]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from boxsegliver_b200 import synthetic
from boxsegliver_b200.device import Context, round_bf16
from boxsegliver_b200.unet3d_engine import UNet3DConfig, UNet3DEngine
from oracle import unet3d_ref as U
from tests.gpu_util import rel

n, d, hw = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 4, 32)
ctx = Context(0)
base = dict(depth=d, height=hw, width=hw, channel=1, weight_decay_rate=3e-5)
ecfg, rcfg = UNet3DConfig(batch=n, **base), U.UNet3DCfg(**base)
images, labels = synthetic.make_volume_batch(n, d, hw, hw, seed=9)
params = U.init_params(rcfg, seed=6)
eng = UNet3DEngine(ctx, ecfg)
print("pair", eng._pair)
eng.set_weights(params)
eng.set_inputs(images, labels, None)
eng.forward(True)
eng.loss_backward()
ctx.check_device()
logits = eng.logits.download(np.float32, (n, d, hw, hw, 2))
grads = eng.get_grads()
stored = eng.get_stored_forward()
stored["logits"] = logits
p64 = {k: v.astype(np.float64) for k, v in params.items()}
tft = U.forward(p64, dict(images=round_bf16(images).astype(np.float64)), rcfg, wrnd=round_bf16, stored=stored)
print("fwd worst", max(tft.errs.items(), key=lambda t: t[1]))
_, dl = U.loss_and_dlogits(tft, labels, rcfg)
g_ref = U.backward(tft, dl, rcfg, rnd=round_bf16)
for name, g in g_ref.items():
    print(f"{rel(grads[name], g):10.3e}  |g| {np.linalg.norm(grads[name]):10.3e} ref {np.linalg.norm(g):10.3e}  {name}")
