"""Two-stream timeline of one U-Net training step: start / end of every C-ABI enqueue relative to the step start,
taken from CUDA events recorded on the stream each call runs on (nothing is serialised, the filter-gradient overlap
stays on). Shows where tensor-bound and HBM-bound kernels actually run side by side.

  python tools/timeline.py [--batch 64] [--hw 256] [--out gpurun_out/timeline.txt]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
      tools/timeline.py --out gpurun_out/timeline_n8.txt      # data parallel: rank 0 writes its timeline, with the NCCL
                                                               # calls of the side stream and the exposed communication
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxsegliver_b200 import synthetic  # noqa: E402
from boxsegliver_b200.device import Context  # noqa: E402
from boxsegliver_b200.engine import EngineConfig, UNetEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--hw", type=int, default=256)
ap.add_argument("--normalizer", default="batch_norm")
ap.add_argument("--out", default="")
a = ap.parse_args()

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
dist = None
if world > 1:
    import ctypes as C
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
ctx = Context(local)
eng = UNetEngine(ctx, EngineConfig(batch=a.batch, height=a.hw, width=a.hw, normalizer=a.normalizer,
                                   weight_decay_rate=1e-6, loss_weight_type="numerical",
                                   loss_numeric_w=(0.2, 0.4, 4.4), world=world))
eng.init_weights(0)
if world > 1:
    uid = (C.c_char * 128)()
    if rank == 0:
        ctx.call("bsl_comm_unique_id", uid)
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)
    eng.attach_comm(rank, world, box[0])
im, lb = synthetic.make_batch(a.batch, a.hw, a.hw, 3, seed=1357 + rank)
eng.set_inputs(im, lb)
for _ in range(4):
    eng.train_step(1e-3)
ctx.sync()
if dist is not None:
    dist.barrier()
ctx.timeline_begin()
eng.train_step(1e-3)
rec = ctx.timeline_end()
ctx.check_device()
if dist is not None:
    dist.barrier()
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)

TC = ("bsl_conv2d_fprop", "bsl_conv2d_dgrad", "bsl_conv2d_wgrad", "bsl_convT2d")
streams = {}
lines = []
for fn, tag, st, t0, t1 in rec:
    sid = streams.setdefault(st, len(streams))
    lines.append((t0, t1, sid, fn, tag))
lines.sort()
end = max(l[1] for l in lines)
out = [f"step span {end:.3f} ms, {len(lines)} calls, {len(streams)} streams"]
for sid in range(len(streams)):
    busy = sum(t1 - t0 for t0, t1, s_, _, _ in lines if s_ == sid)
    out.append(f"stream {sid}: busy {busy:.3f} ms")
# union of busy intervals and pairwise overlap between the streams
def union(iv):
    iv = sorted(iv)
    tot, cur0, cur1 = 0.0, None, None
    for a0, a1 in iv:
        if cur1 is None or a0 > cur1:
            if cur1 is not None:
                tot += cur1 - cur0
            cur0, cur1 = a0, a1
        else:
            cur1 = max(cur1, a1)
    return tot + (cur1 - cur0 if cur1 is not None else 0.0)
out.append(f"union busy {union([(l[0], l[1]) for l in lines]):.3f} ms; "
           f"sum busy {sum(l[1] - l[0] for l in lines):.3f} ms")
if world > 1:
    # exposed communication = time between the end of the last backward kernel and the start of the optimizer on the
    # compute stream (the join with the side stream), and what the collectives themselves took
    nccl = [(t0, t1, tag) for t0, t1, sid, fn, tag in lines if fn == "bsl_allreduce_sum_f32"]
    adam = [t0 for t0, t1, sid, fn, tag in lines if fn in ("bsl_adam_step", "bsl_momentum_step")]
    before = [t1 for t0, t1, sid, fn, tag in lines if fn not in ("bsl_allreduce_sum_f32", "bsl_scale_f32", "bsl_adam_step",
                                                                  "bsl_momentum_step") and t1 <= min(adam)]
    out.append(f"data parallel x{world}: {len(nccl)} all-reduce calls, {sum(t1 - t0 for t0, t1, _ in nccl):.3f} ms on the side "
               f"stream (incl. waiting for peers); last backward kernel ends {max(before):.3f} ms, optimizer starts "
               f"{min(adam):.3f} ms -> exposed communication {min(adam) - max(before):.3f} ms per step")
    for t0, t1, tag in nccl:
        out.append(f"    all-reduce {t0:8.3f} .. {t1:8.3f} ms ({t1 - t0:6.3f})  closed by {tag}")
out.append(f"{'start':>8s} {'end':>8s} {'dur':>7s} st  call")
for t0, t1, sid, fn, tag in lines:
    out.append(f"{t0:8.3f} {t1:8.3f} {t1 - t0:7.3f} {sid:2d}  {'    ' * sid}{fn[4:]:26s} {tag}")
txt = "\n".join(out)
print(txt)
if a.out:
    open(a.out, "w").write(txt + "\n")
