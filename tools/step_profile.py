"""Runs W warm-up + K training steps of the U-Net engine at a given batch (for ncu launch lists)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxsegliver_b200.device import Context
from boxsegliver_b200.engine import EngineConfig, UNetEngine
from boxsegliver_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--hw", type=int, default=256)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
ctx = Context(0)
eng = UNetEngine(ctx, EngineConfig(batch=a.batch, height=a.hw, width=a.hw, loss_weight_type="numerical",
                                   loss_numeric_w=(0.2, 0.4, 4.4)))
eng.init_weights(0)
im, lb = synthetic.make_batch(a.batch, a.hw, a.hw, 3)
eng.set_inputs(im, lb)
for _ in range(a.warmup + a.steps):
    eng.train_step(1e-3)
ctx.check_device()
print("loss", eng.read_loss())
