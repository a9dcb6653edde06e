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


def test_crc32c_host_fallback_matches(monkeypatch):
    """The checkpoint tools must work without the CUDA toolchain (the reference's utils/ckpt_kits.py is CPU-only):
    the numpy/Python CRC-32C gives the RFC 3720 answers and the library's results."""
    assert K.crc32c_host(b"123456789") == 0xE3069283 and K.crc32c_host(bytes(32)) == 0x8A9136AA
    assert K.crc32c_host(bytes([0xFF] * 32)) == 0x62A8AB43 and K.crc32c_host(bytes(range(32))) == 0x46DD794E
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 8, 9, 15, 16, 17, 1000, 4099):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert K.crc32c_host(d) == _crc32c_bitwise(d), n
    d = rng.integers(0, 256, 999, dtype=np.uint8).tobytes()
    assert K.crc32c_host(d[100:], K.crc32c_host(d[:100])) == K.crc32c_host(d)
    monkeypatch.setenv("BSL_CRC_HOST", "1")          # route crc32c() itself through the fallback
    assert K.crc32c(d) == K.crc32c_host(d)
    a = rng.standard_normal((5, 3)).astype(np.float32)
    assert K.crc32c(a) == K.crc32c_host(a.tobytes())


def test_relative_model_dir_round_trip(tmp_path, monkeypatch):
    """A relative model_dir (the reference's default `--model_dir model_dir/<tag>`): entries inside the directory are
    stored relative to it and found again; re-feeding the state never doubles the directory."""
    monkeypatch.chdir(tmp_path)
    w = {"UNet/x/weights": np.arange(4, dtype=np.float32).reshape(2, 2)}
    p1 = K.save_checkpoint("out/model-5", w)
    K.update_checkpoint_state("out", p1)
    assert (tmp_path / "out" / "checkpoint").read_text().splitlines()[0] == 'model_checkpoint_path: "model-5"'
    assert K.latest_checkpoint("out") is not None and K.checkpoint_exists(K.latest_checkpoint("out"))
    st = K.get_checkpoint_state("out")
    p2 = K.save_checkpoint("out/model-9", w)
    K.update_checkpoint_state("out", p2, st.all_model_checkpoint_paths)
    txt = (tmp_path / "out" / "checkpoint").read_text()
    assert "out/" not in txt and txt.count("all_model_checkpoint_paths") == 2
    st = K.get_checkpoint_state("out")
    assert [K.checkpoint_exists(p) for p in st.all_model_checkpoint_paths] == [True, True]
    assert np.array_equal(K.load_checkpoint(K.latest_checkpoint("out")).get_tensor("UNet/x/weights"), w["UNet/x/weights"])
    # a checkpoint outside the directory keeps the path it was given
    p3 = K.save_checkpoint("elsewhere/m", w)
    K.update_checkpoint_state("out", p3)
    assert (tmp_path / "out" / "checkpoint").read_text().splitlines()[0] == 'model_checkpoint_path: "../elsewhere/m"'
    assert K.checkpoint_exists(K.get_checkpoint_state("out").model_checkpoint_path)


def test_reader_validates_entries(tmp_path):
    """Inconsistent BundleEntryProtos give a clear error instead of an opaque numpy reshape failure."""
    K.save_checkpoint(tmp_path / "m", {"a": np.zeros((4, 4), np.float32)})
    r = K.load_checkpoint(tmp_path / "m")
    r.entries["a"]["shard_id"] = 3
    with pytest.raises(ValueError, match="shard_id 3 outside"):
        r.get_tensor("a")
    r = K.load_checkpoint(tmp_path / "m")
    r.entries["a"]["shape"] = (4, 5)
    with pytest.raises(ValueError, match="entry size 64 != prod"):
        r.get_tensor("a")
