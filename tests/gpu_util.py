"""Helpers shared by the -m gpu parity tests (all device work goes through the C ABI)."""
import ctypes as C

import numpy as np

from boxsegliver_b200 import _lib
from boxsegliver_b200.device import round_bf16

TOL_BF16 = 1e-2   # north_star: rel <= 1e-2 for the bf16 path (relative L2 norm per tensor)
TOL_F32 = 1e-4    # fp32-accumulated outputs (wgrad partials, reductions)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def bf16_randn(rng, shape, scale=1.0):
    return round_bf16(rng.standard_normal(shape, dtype=np.float32) * np.float32(scale))


def padded(a, ld):
    """Embed [..., c] into a zero buffer with channel stride ld."""
    out = np.zeros(a.shape[:-1] + (ld,), np.float32)
    out[..., :a.shape[-1]] = a
    return out


def fptr(buf, off_elems=0, esize=4):
    return C.c_void_p(buf.ptr + off_elems * esize)
