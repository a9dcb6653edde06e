"""Device context, buffers and bf16 host helpers on top of the C ABI (no PyTorch involved)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BslError


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns (matches cvt.rn.bf16.f32)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u >> 16) & 1) + 0x7FFF
    out = ((u + r) >> 16).astype(np.uint16)
    out[nan] = 0x7FC0
    return out


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def round_bf16(a: np.ndarray) -> np.ndarray:
    """fp32 values rounded to the nearest bf16-representable value (still stored as fp32)."""
    return bf16_bits_to_f32(f32_to_bf16_bits(a)).reshape(np.shape(a))


class DeviceBuffer:
    """A caller-owned device allocation (what TF's allocator would hand the op kernels)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        ctx.call("bsl_malloc", C.c_size_t(self.nbytes), C.byref(p))
        self.ptr = p.value or 0

    def __int__(self):
        return self.ptr

    def at(self, byte_offset: int) -> C.c_void_p:
        return C.c_void_p(self.ptr + int(byte_offset))

    @property
    def p(self) -> C.c_void_p:
        return C.c_void_p(self.ptr)

    def free(self):
        if self.ptr:
            self.ctx.call("bsl_free", C.c_void_p(self.ptr))
            self.ptr = 0

    def upload(self, arr: np.ndarray, stream=None, byte_offset: int = 0):
        arr = np.ascontiguousarray(arr)
        assert byte_offset + arr.nbytes <= self.nbytes, (arr.nbytes, self.nbytes)
        self.ctx.call("bsl_memcpy_h2d", self.at(byte_offset), arr.ctypes.data_as(C.c_void_p),
                      C.c_size_t(arr.nbytes), self.ctx.stream_arg(stream))
        self.ctx.sync(stream)
        return self

    def download(self, dtype, shape, stream=None, byte_offset: int = 0) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert byte_offset + out.nbytes <= self.nbytes, (out.nbytes, self.nbytes)
        self.ctx.sync(stream)
        self.ctx.call("bsl_memcpy_d2h", out.ctypes.data_as(C.c_void_p), self.at(byte_offset),
                      C.c_size_t(out.nbytes), self.ctx.stream_arg(stream))
        self.ctx.sync(stream)
        return out

    def zero(self, stream=None):
        self.ctx.call("bsl_memset", self.p, C.c_int(0), C.c_size_t(self.nbytes), self.ctx.stream_arg(stream))
        return self


class Context:
    """One per process / GPU. Owns the library context and a default compute stream."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.bsl_init(C.c_int(device), C.byref(h))
        if rc != 0:
            raise BslError(rc, "bsl_init failed (is an sm_100 GPU visible?)")
        self.h = h
        self.device = device
        self.tag = ""            # label attached to profiled calls (the engine sets it per layer)
        self._prof = None
        self._timeline = None
        self._trace = None        # list of C-ABI entry-point names in call order (tests: launch-sequence equality)
        s = C.c_void_p()
        self.call("bsl_stream_create", C.byref(s))
        self.stream = s

    # -- plumbing
    _NO_PROF = ("bsl_malloc", "bsl_free", "bsl_mem_info", "bsl_stream_", "bsl_event_", "bsl_host_", "bsl_device_status",
                "bsl_launch_count", "bsl_graph_", "bsl_comm_", "bsl_debug_set")

    def call(self, name: str, *args) -> int:
        prof = self._prof is not None and not name.startswith(self._NO_PROF)
        if prof:
            e0, e1 = self._prof_event(), self._prof_event()
            self.lib.bsl_event_record(self.h, e0, self.stream)
        tl = self._timeline is not None and not name.startswith(self._NO_PROF)
        if tl:   # every enqueue-only entry point takes its stream as the last argument
            st = args[-1] if args and isinstance(args[-1], C.c_void_p) and args[-1].value else self.stream
            e0, e1 = self._prof_event(), self._prof_event()
            self.lib.bsl_event_record(self.h, e0, st)
        if self._trace is not None and not name.startswith(self._NO_PROF):
            self._trace.append(name)
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            raise BslError(rc, (self.lib.bsl_last_error(self.h) or b"").decode())
        if prof:
            self.lib.bsl_event_record(self.h, e1, self.stream)
            self._prof.append((name, self.tag, e0, e1))
        if tl:
            self.lib.bsl_event_record(self.h, e1, st)
            self._timeline.append((name, self.tag, st.value, e0, e1))
        return rc

    # -- multi-stream timeline (tools/timeline.py): events on the stream each call is enqueued on, no serialisation
    def timeline_begin(self):
        self._prof_pool = getattr(self, "_prof_pool", [])
        self._prof_next = 0
        self._timeline = []
        self._t0 = self.new_event()
        self.record(self._t0, self.stream)

    def timeline_end(self):
        """[(function, tag, stream, start_ms, end_ms)] relative to timeline_begin."""
        rec, self._timeline = self._timeline, None
        self.sync()
        out = []
        for n, t, st, a, b in rec:
            self.call("bsl_event_sync", b)
            out.append((n, t, st, self.elapsed_ms(self._t0, a), self.elapsed_ms(self._t0, b)))
        return out

    # -- per-call device timing (tools/step_breakdown.py): brackets every enqueue on the compute stream
    def _prof_event(self):
        if self._prof_next == len(self._prof_pool):
            self._prof_pool.append(self.new_event())
        e = self._prof_pool[self._prof_next]
        self._prof_next += 1
        return e

    def profile_begin(self):
        self._prof_pool = getattr(self, "_prof_pool", [])
        self._prof_next = 0
        self._prof = []

    def profile_end(self):
        """[(function, tag, ms)] for every call since profile_begin, in enqueue order."""
        rec, self._prof = self._prof, None
        self.sync()
        return [(n, t, self.elapsed_ms(a, b)) for n, t, a, b in rec]

    def stream_arg(self, stream=None) -> C.c_void_p:
        return self.stream if stream is None else stream

    def sync(self, stream=None):
        self.call("bsl_stream_sync", self.stream_arg(stream))

    def check_device(self):
        b, s = C.c_int(), C.c_int()
        self.sync()
        self.call("bsl_device_status", C.byref(b), C.byref(s))

    def new_stream(self) -> C.c_void_p:
        s = C.c_void_p()
        self.call("bsl_stream_create", C.byref(s))
        return s

    def new_event(self) -> C.c_void_p:
        e = C.c_void_p()
        self.call("bsl_event_create", C.byref(e))
        return e

    def record(self, ev, stream=None):
        self.call("bsl_event_record", ev, self.stream_arg(stream))

    def elapsed_ms(self, e0, e1) -> float:
        self.call("bsl_event_sync", e1)
        ms = C.c_float()
        self.call("bsl_event_elapsed_ms", e0, e1, C.byref(ms))
        return ms.value

    # -- memory
    def mem_info(self):
        """(free bytes, total bytes) of the device."""
        f, t = C.c_size_t(), C.c_size_t()
        self.call("bsl_mem_info", C.byref(f), C.byref(t))
        return f.value, t.value

    def attach_comm(self, rank: int, world: int, unique_id: bytes):
        """Join the NCCL communicator once per context; engines built later on this context share it."""
        if getattr(self, "_comm", None) is None:
            buf = (C.c_char * 128).from_buffer_copy(unique_id)
            self.call("bsl_comm_init", buf, C.c_int(rank), C.c_int(world))
            self._comm = (rank, world)
        assert self._comm == (rank, world), (self._comm, rank, world)

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def from_numpy(self, arr: np.ndarray) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        return DeviceBuffer(self, arr.nbytes).upload(arr)

    def bf16_from_f32(self, arr: np.ndarray) -> DeviceBuffer:
        return self.from_numpy(f32_to_bf16_bits(arr))

    def bf16_to_f32(self, buf: DeviceBuffer, shape, byte_offset: int = 0) -> np.ndarray:
        return bf16_bits_to_f32(buf.download(np.uint16, shape, byte_offset=byte_offset)).reshape(shape)

    def close(self):
        if self.h:
            self.lib.bsl_destroy(self.h)
            self.h = None
