"""TF-1.x `tf.train.Saver` V2 checkpoint wire format, read and written without TensorFlow, so
that a reference `model_dir` loads into the engines and an engine's state is saved under the names the reference's
tools expect.

Reference call sites: core/estimator.py:694-703 (Saver / Scaffold), core/models.py:151-185 (`init_model`,
`_find_root_scope`: --load_weights with scope renaming), core/hooks.py:193-228 (best-checkpoint saver),
utils/ckpt_kits.py:21-59 (`list_vars_shape`, `ckpt_vars_rename`, `find_checkpoint`).

The format lives in the un-vendored tensorflow-gpu==1.13 (requirements.txt:2); its published layout, restated:
  <prefix>.index                 an immutable sorted string table (LevelDB table format): data blocks of
                                 prefix-compressed (key, value) entries + restart array, each followed by a 5-byte
                                 trailer (compression type 0, masked CRC-32C of block + type); a metaindex block; an
                                 index block mapping separator keys to block handles; a 48-byte footer ending in the
                                 magic 0xdb4775248b80fb57. Key "" holds BundleHeaderProto (num_shards, endianness,
                                 version); every other key is a variable name holding BundleEntryProto
                                 (dtype, shape, shard_id, offset, size, masked crc32c of the tensor bytes).
  <prefix>.data-00000-of-00001   the raw little-endian tensor bytes at those offsets.
  checkpoint                     CheckpointState text proto (model_checkpoint_path, all_model_checkpoint_paths).
PARITY UNPINNED by the reference (no checkpoint ships with it and TF cannot run here): the module is pinned by the
CRC-32C / masking known answers, by byte-level checks of the footer and block trailers, and by write -> read round
trips (tests/test_checkpoint.py).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from pathlib import Path

import numpy as np

from . import _lib

TABLE_MAGIC = 0xDB4775248B80FB57
BLOCK_SIZE = 4096            # target uncompressed data-block size (any value is valid for readers)
RESTART_INTERVAL = 16
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
      17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DT_OF = {np.dtype(v): k for k, v in DT.items()}


_CRC_TABLES = None


def _crc32c_tables():
    """Slicing-by-8 tables of the Castagnoli polynomial (reflected 0x82F63B78), built once with numpy."""
    global _CRC_TABLES
    if _CRC_TABLES is None:
        t0 = np.arange(256, dtype=np.uint32)
        for _ in range(8):
            t0 = np.where(t0 & 1, (t0 >> 1) ^ np.uint32(0x82F63B78), t0 >> 1).astype(np.uint32)
        tabs = [t0]
        for _ in range(7):
            prev = tabs[-1]
            tabs.append((t0[prev & 0xFF] ^ (prev >> 8)).astype(np.uint32))
        _CRC_TABLES = tabs
    return _CRC_TABLES


def crc32c_host(data: bytes | np.ndarray, crc: int = 0) -> int:
    """Pure numpy / Python CRC-32C (RFC 3720): the checkpoint tools (list / rename / read) must run on machines without
    the CUDA toolchain, like the reference's utils/ckpt_kits.py. 8 bytes per Python-level iteration over a whole
    column of the message would need carried state, so the loop is over 8-byte words with table look-ups."""
    buf = np.ascontiguousarray(data).view(np.uint8).ravel() if isinstance(data, np.ndarray) else np.frombuffer(data, np.uint8)
    T = [t.tolist() for t in _crc32c_tables()]
    t0, t1, t2, t3, t4, t5, t6, t7 = T
    c = (~crc) & 0xFFFFFFFF
    n8 = buf.size // 8
    if n8:
        words = buf[:n8 * 8].view("<u4").reshape(n8, 2)
        lo, hi = words[:, 0].tolist(), words[:, 1].tolist()
        for a, b in zip(lo, hi):
            a ^= c
            c = (t7[a & 0xFF] ^ t6[(a >> 8) & 0xFF] ^ t5[(a >> 16) & 0xFF] ^ t4[a >> 24]
                 ^ t3[b & 0xFF] ^ t2[(b >> 8) & 0xFF] ^ t1[(b >> 16) & 0xFF] ^ t0[b >> 24])
    for x in buf[n8 * 8:].tolist():
        c = t0[(c ^ x) & 0xFF] ^ (c >> 8)
    return (~c) & 0xFFFFFFFF


def crc32c(data: bytes | np.ndarray, crc: int = 0) -> int:
    """CRC-32C through libbsl_b200.so's bsl_crc32c when the library is built (fast path for multi-hundred-MB tensors),
    else the host implementation above (identical results, pinned by the RFC 3720 vectors in tests/test_checkpoint.py)."""
    if _lib.LIB_PATH.exists() and not os.environ.get("BSL_CRC_HOST"):
        buf = np.ascontiguousarray(data).view(np.uint8) if isinstance(data, np.ndarray) else np.frombuffer(data, np.uint8)
        return int(_lib.load().bsl_crc32c(crc, buf.ctypes.data_as(C.c_void_p), buf.size))
    return crc32c_host(data, crc)


def mask_crc(crc: int) -> int:
    """crc32c::Mask: rotate right by 15 bits and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / minimal protobuf
def _varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _read_varint(b: bytes, pos: int):
    shift = v = 0
    while True:
        c = b[pos]
        pos += 1
        v |= (c & 0x7F) << shift
        if c < 0x80:
            return v, pos
        shift += 7


def _pb_fields(b: bytes):
    """Yield (field number, wire type, value) of a serialised message (varint, fixed64, bytes, fixed32)."""
    pos = 0
    while pos < len(b):
        tag, pos = _read_varint(b, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(b, pos)
        elif wt == 1:
            v, pos = b[pos:pos + 8], pos + 8
        elif wt == 2:
            n, pos = _read_varint(b, pos)
            v, pos = b[pos:pos + n], pos + n
        elif wt == 5:
            v, pos = b[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield f, wt, v


def _pb(field: int, wt: int, payload) -> bytes:
    tag = _varint((field << 3) | wt)
    if wt == 0:
        return tag + _varint(payload)
    if wt == 2:
        return tag + _varint(len(payload)) + payload
    if wt == 5:
        return tag + struct.pack("<I", payload)
    raise ValueError(wt)


def _encode_shape(shape) -> bytes:
    """TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }."""
    return b"".join(_pb(2, 2, _pb(1, 0, int(s))) for s in shape)


def _decode_shape(b: bytes):
    dims = []
    for f, _, v in _pb_fields(b):
        if f == 2:
            size = 0
            for g, _, w in _pb_fields(v):
                if g == 1:
                    size = w
            dims.append(size)
    return tuple(dims)


def _encode_entry(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    """BundleEntryProto: dtype = 1, shape = 2, shard_id = 3 (0 omitted), offset = 4, size = 5, crc32c = 6 (fixed32)."""
    out = _pb(1, 0, dtype) + _pb(2, 2, _encode_shape(shape))
    if offset:
        out += _pb(4, 0, offset)
    if size:
        out += _pb(5, 0, size)
    return out + _pb(6, 5, crc)


def _decode_entry(b: bytes) -> dict:
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=0, sliced=False)
    for f, _, v in _pb_fields(b):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _decode_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif f == 7:
            e["sliced"] = True
    return e


# ------------------------------------------------------------------ table (LevelDB format)
class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count % RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def size(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4


def _write_table(path: Path, items):
    """items: sorted [(key bytes, value bytes)]."""
    out = bytearray()

    def emit(block: bytes):
        handle = _varint(len(out)) + _varint(len(block))
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return handle

    index = _BlockBuilder()
    blk = _BlockBuilder()
    for key, value in items:
        blk.add(key, value)
        if blk.size() >= BLOCK_SIZE:
            index.add(blk.last, emit(blk.finish()))      # the block's last key is a valid separator
            blk = _BlockBuilder()
    if blk.count:
        index.add(blk.last, emit(blk.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    path.write_bytes(bytes(out))


def _read_block(data: bytes, offset: int, size: int, verify: bool = True):
    block, trailer = data[offset:offset + size], data[offset + size:offset + size + 5]
    if len(trailer) != 5:
        raise ValueError("truncated table block")
    if trailer[0] != 0:
        raise NotImplementedError("compressed table blocks (tensor bundles are written uncompressed)")
    if verify and unmask_crc(struct.unpack("<I", trailer[1:])[0]) != crc32c(block + trailer[:1]):
        raise ValueError("table block checksum mismatch")
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _read_table(path: Path):
    data = path.read_bytes()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TF checkpoint index (bad table magic)")
    footer = data[-48:-8]
    _, pos = _read_varint(footer, 0)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    items = []
    for _, handle in _read_block(data, ioff, isize):
        off, p2 = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p2)
        items += _read_block(data, off, size)
    return items


# ------------------------------------------------------------------ bundle
def _data_path(prefix, shard=0, shards=1) -> Path:
    return Path(f"{prefix}.data-{shard:05d}-of-{shards:05d}")


def save_checkpoint(prefix, tensors: dict):
    """Write {variable name: numpy array} as <prefix>.index + <prefix>.data-00000-of-00001."""
    prefix = Path(prefix)
    prefix.parent.mkdir(parents=True, exist_ok=True)
    header = _pb(1, 0, 1) + _pb(3, 2, _pb(1, 0, 1))          # num_shards = 1, endianness LITTLE (default), version.producer = 1
    items, offset = [(b"", header)], 0
    with open(_data_path(prefix), "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in DT_OF:
                raise TypeError(f"{name}: dtype {a.dtype} has no DataType mapping")
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            items.append((name.encode(), _encode_entry(DT_OF[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    _write_table(Path(str(prefix) + ".index"), items)
    return str(prefix)


class CheckpointReader:
    """What `tf.train.load_checkpoint(path)` / `pywrap_tensorflow.NewCheckpointReader` return (utils/ckpt_kits.py:21-33,
    core/models.py:151-158): get_variable_to_shape_map, get_variable_to_dtype_map, has_tensor, get_tensor."""

    def __init__(self, prefix, verify: bool = True):
        self.prefix, self.verify = str(prefix), verify
        items = _read_table(Path(self.prefix + ".index"))
        if not items or items[0][0] != b"":
            raise ValueError("checkpoint index has no bundle header")
        self.num_shards = 1
        for f, _, v in _pb_fields(items[0][1]):
            if f == 1:
                self.num_shards = v
            elif f == 2 and v != 0:
                raise NotImplementedError("big-endian tensor bundle")
        self.entries = {k.decode(): _decode_entry(v) for k, v in items[1:]}

    def get_variable_to_shape_map(self) -> dict:
        return {k: list(e["shape"]) for k, e in self.entries.items()}

    def get_variable_to_dtype_map(self) -> dict:
        return {k: np.dtype(DT[e["dtype"]]) for k, e in self.entries.items()}

    def has_tensor(self, name: str) -> bool:
        return name in self.entries

    def get_tensor(self, name: str) -> np.ndarray:
        if name not in self.entries:
            raise KeyError(f"Key {name} not found in checkpoint")      # NotFoundError text of the TF reader
        e = self.entries[name]
        if e["sliced"]:
            raise NotImplementedError(f"{name}: partitioned variables are not used by the reference's models")
        if e["dtype"] not in DT:
            raise TypeError(f"{name}: unsupported DataType {e['dtype']}")
        if not 0 <= e["shard_id"] < self.num_shards:
            raise ValueError(f"{name}: shard_id {e['shard_id']} outside the bundle's {self.num_shards} shard(s)")
        want = int(np.prod(e["shape"], dtype=np.int64)) * np.dtype(DT[e["dtype"]]).itemsize
        if want != e["size"]:
            raise ValueError(f"{name}: entry size {e['size']} != prod(shape {list(e['shape'])}) * itemsize = {want}")
        with open(_data_path(self.prefix, e["shard_id"], self.num_shards), "rb") as f:
            f.seek(e["offset"])
            raw = f.read(e["size"])
        if len(raw) != e["size"]:
            raise ValueError(f"{name}: data file truncated")
        if self.verify and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError(f"{name}: tensor checksum mismatch")
        return np.frombuffer(raw, dtype=np.dtype(DT[e["dtype"]]).newbyteorder("<")).reshape(e["shape"]).copy()


def load_checkpoint(prefix, verify: bool = True) -> CheckpointReader:
    return CheckpointReader(prefix, verify)


def checkpoint_exists(prefix) -> bool:
    return Path(str(prefix) + ".index").exists()


# ------------------------------------------------------------------ CheckpointState text proto
def update_checkpoint_state(save_dir, model_checkpoint_path, all_model_checkpoint_paths=None,
                            latest_filename: str = "checkpoint"):
    """tf.train.update_checkpoint_state: paths inside save_dir are stored relative to it (for absolute AND relative
    inputs -- TF relativises with os.path.relpath(path, save_dir) when save_dir itself is relative), so that
    get_checkpoint_state, which joins every non-absolute entry onto the directory, finds them again."""
    save_dir = Path(save_dir)
    root = os.path.abspath(save_dir)

    def rel(p):
        ap = os.path.abspath(str(p))
        if ap == root or ap.startswith(root + os.sep):
            return os.path.relpath(ap, root)
        if not save_dir.is_absolute() and not os.path.isabs(str(p)):
            return os.path.relpath(ap, root)       # TF: relative paths are rewritten relative to a relative save_dir
        return str(p)

    paths = [str(p) for p in (all_model_checkpoint_paths or [])]
    if not paths or os.path.abspath(paths[-1]) != os.path.abspath(str(model_checkpoint_path)):
        paths.append(str(model_checkpoint_path))
    lines = [f'model_checkpoint_path: "{rel(model_checkpoint_path)}"']
    lines += [f'all_model_checkpoint_paths: "{rel(p)}"' for p in paths]
    (save_dir / latest_filename).write_text("\n".join(lines) + "\n")


class CheckpointState:
    def __init__(self, model_checkpoint_path, all_model_checkpoint_paths):
        self.model_checkpoint_path = model_checkpoint_path
        self.all_model_checkpoint_paths = all_model_checkpoint_paths


def get_checkpoint_state(checkpoint_dir, latest_filename: str = "checkpoint"):
    f = Path(checkpoint_dir) / (latest_filename or "checkpoint")
    if not f.exists():
        return None
    model, allp = None, []
    for line in f.read_text().splitlines():
        key, _, val = line.partition(":")
        val = val.strip().strip('"')
        if not val:
            continue
        full = val if Path(val).is_absolute() else str(Path(checkpoint_dir) / val)
        if key.strip() == "model_checkpoint_path":
            model = full
        elif key.strip() == "all_model_checkpoint_paths":
            allp.append(full)
    return CheckpointState(model, allp) if model else None


def latest_checkpoint(checkpoint_dir, latest_filename: str = "checkpoint"):
    st = get_checkpoint_state(checkpoint_dir, latest_filename)
    return st.model_checkpoint_path if st and checkpoint_exists(st.model_checkpoint_path) else None


# ------------------------------------------------------------------ engines <-> checkpoints
def _slot_suffixes(engine):
    """Slot names the TF optimizers create (Optimizer._zeros_slot(var, slot, self._name)): Adam -> Adam, Adam_1;
    AdamW -> AdamW, AdamW_1; Momentum -> Momentum."""
    opt = engine.cfg.optimizer
    return {"adam": ("Adam", "Adam_1"), "adamw": ("AdamW", "AdamW_1"), "momentum": ("Momentum",)}[opt]


def _slot_arrays(engine) -> dict:
    """Adam / AdamW / Momentum slots under the names `optimizer.minimize` gives them inside variable_scope("Optimizer")
    (core/solver.py:232-239): Optimizer/<variable>/Adam, /Adam_1 (or /Momentum), plus beta1_power / beta2_power.
    Goes through engine.get_slots(), which un-pads / un-permutes the engine's own arena layout."""
    out = {}
    suffixes = _slot_suffixes(engine)
    for name, arrs in engine.get_slots().items():
        for sfx, a in zip(suffixes, arrs):
            out[f"Optimizer/{name}/{sfx}"] = a
    if len(suffixes) == 2:
        t = engine.step_count
        cfg = engine.cfg                                             # TF stores beta^(t+1) after t steps
        out["Optimizer/beta1_power"] = np.float32(cfg.adam_beta1) ** np.float32(t + 1)
        out["Optimizer/beta2_power"] = np.float32(cfg.adam_beta2) ** np.float32(t + 1)
    return out


def save_engine(engine, prefix, global_step: int | None = None, save_dir_state: bool = True):
    """Saver.save(sess, prefix, global_step): variables, moving statistics, optimizer slots, global_step."""
    tensors = dict(engine.get_weights())
    if engine.cfg.training:
        tensors.update(_slot_arrays(engine))
    step = engine.step_count if global_step is None else global_step
    tensors["global_step"] = np.array(step, np.int64)
    path = f"{prefix}-{step}"
    save_checkpoint(path, tensors)
    if save_dir_state:
        d = Path(path).parent
        st = get_checkpoint_state(d)
        update_checkpoint_state(d, path, (st.all_model_checkpoint_paths if st else []))
    return path


def find_root_scope(reader: CheckpointReader):
    """core/models.py:151-158: the scope under which the checkpoint's model variables live."""
    for var in reader.get_variable_to_shape_map():
        if var.startswith("Optimizer") and not var.endswith("power"):
            return var.split("/")[1]
    return None


def restore_engine(engine, path, weights_scope: str | None = None, latest_filename: str = "checkpoint",
                   with_slots: bool = False):
    """core/models.py:161-185 `init_model`: `path` is a checkpoint prefix or a directory holding a CheckpointState;
    variables are looked up with the model's root scope replaced by `weights_scope` (or the scope found in the file).
    Returns the checkpoint's global_step (or None)."""
    p = Path(path)
    ckpt = str(p)
    if p.is_dir():
        st = get_checkpoint_state(p, latest_filename)
        if st and st.model_checkpoint_path:
            ckpt = st.model_checkpoint_path
    if not checkpoint_exists(ckpt):
        raise FileNotFoundError("ckpt_filename {} doesn't exist".format(ckpt))
    reader = load_checkpoint(ckpt)
    model_root = next(iter(engine.params)).split("/")[0]
    root = weights_scope or find_root_scope(reader) or model_root
    ren = lambda name: root + name[len(model_root):] if name.startswith(model_root) else name  # noqa: E731
    weights = {}
    for name, prm in engine.params.items():
        a = reader.get_tensor(ren(name))
        if tuple(a.shape) != tuple(prm.shape):
            raise ValueError(f"{name}: checkpoint shape {a.shape} != model shape {tuple(prm.shape)}")
        weights[name] = a
    engine.set_weights(weights)
    step = int(reader.get_tensor("global_step")) if reader.has_tensor("global_step") else None
    if with_slots and engine.cfg.training:
        suffixes = _slot_suffixes(engine)
        slots = {}
        for name, prm in engine.params.items():
            if getattr(prm, "region", "A") == "S":
                continue
            base = f"Optimizer/{ren(name)}"
            slots[name] = tuple(reader.get_tensor(f"{base}/{sfx}") for sfx in suffixes)
        engine.set_slots(slots)
        if step is not None:
            engine.step_count = step
    return step


def ckpt_vars_rename(input_, output=None, replace_from=(), replace_to=(), add_prefix=None):
    """utils/ckpt_kits.py:36-62: rename variables of a checkpoint (returns {old: new}; writes `output` if given)."""
    assert len(replace_from) == len(replace_to), (len(replace_from), len(replace_to))
    reader = load_checkpoint(input_)
    replace_to = ["" if x == "empty" else x for x in replace_to]
    mapping, tensors = {}, {}
    for k in sorted(reader.get_variable_to_shape_map()):
        new = k
        for f, t in zip(replace_from, replace_to):
            new = new.replace(f, t)
        if add_prefix:
            new = add_prefix + new
        mapping[k] = new
        if output:
            tensors[new] = reader.get_tensor(k)
    if output:
        save_checkpoint(output, tensors)
    return mapping
