"""TF Saver V2 wire format without TensorFlow (boxsegliver_b200/checkpoint.py), CPU only: CRC-32C / masking known
answers, byte-level layout of the index table (footer magic, block trailers, prefix compression), round trips over
every dtype the reference saves, corruption detection, CheckpointState files and variable renaming
(/root/reference/utils/ckpt_kits.py:21-59, core/models.py:151-185)."""
import struct

import numpy as np
import pytest

from boxsegliver_b200 import checkpoint as K


def _crc32c_bitwise(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 & -(c & 1))
    return c ^ 0xFFFFFFFF


def test_crc32c_known_answers():
    assert K.crc32c(b"123456789") == 0xE3069283                      # CRC-32C check value (RFC 3720 B.4)
    assert K.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720: 32 bytes of zeros
    assert K.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                 # RFC 3720: 32 bytes of ones
    assert K.crc32c(bytes(range(32))) == 0x46DD794E                   # RFC 3720: incrementing bytes
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 1000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert K.crc32c(d) == _crc32c_bitwise(d)
    d = rng.integers(0, 256, 4099, dtype=np.uint8).tobytes()
    assert K.crc32c(d[1000:], K.crc32c(d[:1000])) == K.crc32c(d)      # running form
    # leveldb crc32c_test "Mask": masking is not idempotent and is inverted by unmask
    crc = K.crc32c(b"foo")
    assert K.mask_crc(crc) != crc and K.mask_crc(K.mask_crc(crc)) != crc
    assert K.unmask_crc(K.mask_crc(crc)) == crc and K.unmask_crc(K.unmask_crc(K.mask_crc(K.mask_crc(crc)))) == crc


def _tensors(rng):
    t = {
        "UNet/Encode1/Repeat/convolution2d_1/weights": rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
        "UNet/Encode1/Repeat/convolution2d_1/BatchNorm/gamma": np.ones(64, np.float32),
        "UNet/Encode1/Repeat/convolution2d_1/BatchNorm/moving_variance": rng.uniform(0.5, 2, 64).astype(np.float32),
        "UNet/AdjustChannels/biases": np.zeros(3, np.float32),
        "Optimizer/UNet/AdjustChannels/biases/Adam": rng.standard_normal(3).astype(np.float32),
        "Optimizer/beta1_power": np.float32(0.81),
        "global_step": np.array(123456789012, np.int64),
        "misc/f64": rng.standard_normal((2, 5)),
        "misc/i32": rng.integers(-5, 5, (4,)).astype(np.int32),
        "misc/empty": np.zeros((0, 7), np.float32),
    }
    for i in range(400):                                               # several data blocks, shared key prefixes
        t[f"UNet/Decode{i % 4 + 1}/Repeat/convolution2d_{i}/weights"] = rng.standard_normal((2, 3)).astype(np.float32)
    return t


def test_round_trip_and_byte_layout(tmp_path):
    rng = np.random.default_rng(1)
    t = _tensors(rng)
    prefix = tmp_path / "model.ckpt-5"
    K.save_checkpoint(prefix, t)
    index = (tmp_path / "model.ckpt-5.index").read_bytes()
    data = (tmp_path / "model.ckpt-5.data-00000-of-00001").read_bytes()
    assert struct.unpack("<Q", index[-8:])[0] == 0xDB4775248B80FB57 and len(index) > 48
    assert len(data) == sum(np.asarray(a).nbytes for a in t.values())
    # first data block starts at 0: entry 0 is key "" (shared 0, non_shared 0) holding the bundle header
    assert index[0] == 0 and index[1] == 0
    r = K.load_checkpoint(prefix)
    assert set(r.get_variable_to_shape_map()) == set(t)
    assert r.get_variable_to_shape_map()["UNet/Encode1/Repeat/convolution2d_1/weights"] == [3, 3, 3, 64]
    assert r.get_variable_to_dtype_map()["global_step"] == np.int64
    for k, a in t.items():
        b = r.get_tensor(k)
        assert b.dtype == np.asarray(a).dtype and b.shape == np.asarray(a).shape and np.array_equal(b, a), k
    assert int(r.get_tensor("global_step")) == 123456789012
    assert K.find_root_scope(r) == "UNet"
    with pytest.raises(KeyError, match="not found in checkpoint"):
        r.get_tensor("UNet/nope")
    # tensors are laid out in key order; offsets are what the entries say
    e = r.entries["UNet/AdjustChannels/biases"]
    assert data[e["offset"]:e["offset"] + e["size"]] == t["UNet/AdjustChannels/biases"].tobytes()
    assert K.unmask_crc(e["crc32c"]) == K.crc32c(t["UNet/AdjustChannels/biases"].tobytes())


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(2)
    prefix = tmp_path / "m"
    K.save_checkpoint(prefix, {"a/w": rng.standard_normal((8, 8)).astype(np.float32), "b": np.float32(1)})
    dpath = tmp_path / "m.data-00000-of-00001"
    raw = bytearray(dpath.read_bytes())
    raw[17] ^= 0x40
    dpath.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        K.load_checkpoint(prefix).get_tensor("a/w")
    assert K.load_checkpoint(prefix, verify=False).get_tensor("a/w").shape == (8, 8)
    ipath = tmp_path / "m.index"
    raw = bytearray(ipath.read_bytes())
    raw[5] ^= 0x01
    ipath.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        K.load_checkpoint(prefix)
    ipath.write_bytes(b"not a table" * 10)
    with pytest.raises(ValueError, match="magic"):
        K.load_checkpoint(prefix)


def test_checkpoint_state_and_rename(tmp_path):
    rng = np.random.default_rng(3)
    w = {"UNet/x/weights": rng.standard_normal((2, 2)).astype(np.float32),
         "Optimizer/UNet/x/weights/Adam": np.zeros((2, 2), np.float32)}
    p1 = K.save_checkpoint(tmp_path / "model.ckpt-10", w)
    K.update_checkpoint_state(tmp_path, p1)
    p2 = K.save_checkpoint(tmp_path / "model.ckpt-20", w)
    K.update_checkpoint_state(tmp_path, p2, [p1])
    txt = (tmp_path / "checkpoint").read_text()
    assert txt.splitlines()[0] == 'model_checkpoint_path: "model.ckpt-20"'
    assert txt.count("all_model_checkpoint_paths") == 2
    assert K.latest_checkpoint(tmp_path) == str(tmp_path / "model.ckpt-20")
    st = K.get_checkpoint_state(tmp_path)
    assert st.all_model_checkpoint_paths == [str(tmp_path / "model.ckpt-10"), str(tmp_path / "model.ckpt-20")]
    assert K.get_checkpoint_state(tmp_path, "checkpoint_best") is None and K.latest_checkpoint(tmp_path / "nope") is None
    m = K.ckpt_vars_rename(p2, str(tmp_path / "renamed"), replace_from=["UNet"], replace_to=["GUNet"], add_prefix="pre/")
    assert m["UNet/x/weights"] == "pre/GUNet/x/weights"
    r = K.load_checkpoint(tmp_path / "renamed")
    assert np.array_equal(r.get_tensor("pre/GUNet/x/weights"), w["UNet/x/weights"])
    assert K.ckpt_vars_rename(p2, replace_from=["x"], replace_to=["empty"])["UNet/x/weights"] == "UNet//weights"
