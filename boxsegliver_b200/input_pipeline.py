"""Device-side input stage: the per-sample map function of the reference's training pipeline
(/root/reference/DataLoader/Liver/input_pipeline.py:243-284 `data_processing_train`, and the guide variant
DataLoader/Liver/input_pipeline_g.py:357-412) run as one kernel over a staged batch of decoded slices, writing
straight into an engine's input buffers. The random decisions stay with the caller (bbox jitter, flips), exactly the
arguments the reference's generator passes to the map function; the noise stream is keyed by (seed, step).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class DeviceInputStage:
    def __init__(self, ctx, batch: int, channels: int, src_hw, out_hw, noise_scale: float = 0.0, seed: int = 0,
                 max_centers: int = 8, min_std: float = 1.0, with_guide: bool = False):
        self.ctx, self.n, self.c = ctx, batch, channels
        self.src_hw, self.out_hw = tuple(src_hw), tuple(out_hw)
        self.noise_scale, self.seed, self.min_std = float(noise_scale), int(seed), float(min_std)
        self.max_centers, self.with_guide = max_centers, with_guide
        sh, sw = self.src_hw
        self.slices = ctx.alloc(batch * channels * sh * sw * 2)
        self.seg = ctx.alloc(batch * sh * sw)
        # per-sample parameter rows, packed into one small device buffer
        self._rows = dict(bbox=(np.int32, 4), clip=(np.float32, 2), lab_scale=(np.int32, 1),
                          present=(np.uint8, channels), flips=(np.int32, 1), centers=(np.float32, 2 * max_centers),
                          stddevs=(np.float32, 2 * max_centers), n_centers=(np.int32, 1))
        self._off, off = {}, 0
        for k, (dt, w) in self._rows.items():
            self._off[k] = off
            off += (batch * w * np.dtype(dt).itemsize + 255) // 256 * 256
        self.params = ctx.alloc(off)
        self.step = 0

    def stage(self, slices_u16: np.ndarray, seg_u8: np.ndarray | None, bbox, clip, lab_scale, present=None, flips=None,
              centers=None, stddevs=None, stream=None):
        """Upload one batch: slices uint16 [n, C, src_h, src_w], seg uint8 [n, src_h, src_w], per-sample rows."""
        n, c = self.n, self.c
        assert slices_u16.shape == (n, c) + self.src_hw and slices_u16.dtype == np.uint16, slices_u16.shape
        self.slices.upload(slices_u16, stream)
        self._has_seg = seg_u8 is not None
        if self._has_seg:
            assert seg_u8.shape == (n,) + self.src_hw and seg_u8.dtype == np.uint8
            self.seg.upload(seg_u8, stream)
        bbox = np.asarray(bbox, np.int32).reshape(n, 4)
        sh, sw = self.src_hw
        if (bbox[:, 0] < 0).any() or (bbox[:, 1] < 0).any() or (bbox[:, 2] < 1).any() or (bbox[:, 3] < 1).any() or \
                (bbox[:, 0] + bbox[:, 2] > sh).any() or (bbox[:, 1] + bbox[:, 3] > sw).any():
            raise ValueError("bounding box outside the image (tf.image.crop_to_bounding_box would raise)")
        rows = dict(bbox=bbox, clip=np.asarray(clip, np.float32).reshape(n, 2),
                    lab_scale=np.broadcast_to(np.asarray(lab_scale, np.int32), (n,)).reshape(n, 1),
                    present=np.ones((n, c), np.uint8) if present is None else np.asarray(present, np.uint8).reshape(n, c),
                    flips=np.zeros((n, 1), np.int32) if flips is None else np.asarray(flips, np.int32).reshape(n, 1))
        k = self.max_centers
        ctr, sdv, cnt = np.full((n, k, 2), -1, np.float32), np.ones((n, k, 2), np.float32), np.zeros((n, 1), np.int32)
        if centers is not None:
            for i, (ci, si) in enumerate(zip(centers, stddevs)):
                m = len(ci)
                if m > k:
                    raise ValueError(f"sample {i}: {m} guide centres > max_centers={k}")
                if m:
                    ctr[i, :m], sdv[i, :m] = np.asarray(ci, np.float32), np.asarray(si, np.float32)
                cnt[i, 0] = m
        rows.update(centers=ctr.reshape(n, -1), stddevs=sdv.reshape(n, -1), n_centers=cnt)
        for name, arr in rows.items():
            self.params.upload(np.ascontiguousarray(arr), stream, byte_offset=self._off[name])

    def run(self, images, labels=None, sp_guide=None, stream=None):
        """Enqueue the pass. images / labels / sp_guide are DeviceBuffers (an engine's `images`, `labels`, guide)."""
        oh, ow = self.out_hw
        d = _lib.InputDesc(self.n, self.c, self.src_hw[0], self.src_hw[1], oh, ow, self.max_centers, self.noise_scale,
                           self.min_std, self.seed, self.step)
        pa = lambda k: self.params.ptr + self._off[k]   # noqa: E731
        p = _lib.InputParams(pa("bbox"), pa("clip"), pa("lab_scale"), pa("present"), pa("flips"), pa("centers"),
                             pa("stddevs"), pa("n_centers"))
        if labels is not None and not self._has_seg:
            raise ValueError("labels requested but no segmentation was staged")
        self.ctx.call("bsl_input_stage", C.byref(d), C.byref(p), self.slices.p, self.seg.p if labels is not None else None,
                      images.p, labels.p if labels is not None else None,
                      sp_guide.p if sp_guide is not None else None, self.ctx.stream_arg(stream))
        self.step += 1

    def close(self):
        for b in (self.slices, self.seg, self.params):
            b.free()
