"""TensorFlow Saver-V2 checkpoint reader / writer (no TensorFlow here: known answers of the format + round trips)."""
import struct

import numpy as np
import pytest

from imagecaptionlearn_py_b200 import tf_checkpoint as T


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                      # the standard CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                  # RFC 3720 B.4: 32 bytes of ones
    c = T.crc32c(b"abc")
    assert T.masked_crc(b"abc") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF     # leveldb crc32c::Mask
    big = np.arange(100000, dtype=np.uint8).tobytes()                # > 4 KB: the C-ABI library's slice-by-8 implementation
    ref = 0
    for i in range(0, len(big), 1000):                               # chained small (pure-Python) pieces
        ref = T.crc32c(big[i:i + 1000], ref)
    assert T.crc32c(big) == ref


def test_bundle_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"bidirectional_lstm/bidirectional_rnn/fw/basic_lstm_cell/kernel": rng.standard_normal((40, 80)).astype(np.float32),
               "bidirectional_lstm/bidirectional_rnn/fw/basic_lstm_cell/bias": rng.standard_normal((80,)).astype(np.float32),
               "hdn_1/Variable": rng.standard_normal((90, 16)).astype(np.float32), "hdn_1/Variable_1": np.zeros((1, 16), np.float32),
               "beta1_power": np.float32(0.9 ** 7), "global_step": np.int64(7)}
    tensors.update({"v%03d" % i: rng.standard_normal((3, i + 1)).astype(np.float32) for i in range(150)})    # several data blocks
    prefix = str(tmp_path / "m.model")
    T.write_bundle(prefix, tensors)
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xDB4775248B80FB57 and len(idx) > 48              # leveldb table magic, 48-byte footer
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert len(data) == sum(np.asarray(v).nbytes for v in tensors.values())
    back = T.read_bundle(prefix)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape and np.array_equal(back[k], v), k
    # a flipped data byte must be caught by the per-tensor CRC
    corrupt = bytearray(data)
    corrupt[10] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(corrupt))
    with pytest.raises(ValueError):
        T.read_bundle(prefix)


def test_state_dict_mapping_uses_tf_slot_names():
    st = {"hdn_1/Variable": np.ones((2, 2), np.float32), "adam_m/hdn_1/Variable": np.full((2, 2), 2, np.float32),
          "adam_v/hdn_1/Variable": np.full((2, 2), 3, np.float32), "adam_step": np.int64(12)}
    tf_vars = T.from_state_dict(st)
    assert set(tf_vars) == {"hdn_1/Variable", "hdn_1/Variable/Adam", "hdn_1/Variable/Adam_1", "beta1_power", "beta2_power"}
    back = T.to_state_dict(tf_vars)
    assert int(back["adam_step"]) == 12
    for k in ("hdn_1/Variable", "adam_m/hdn_1/Variable", "adam_v/hdn_1/Variable"):
        assert np.array_equal(back[k], st[k])


def test_adam_step_round_trips_through_tf_beta_powers():
    """TF 1.x stores beta ** (t + 1) after t steps (initial value beta, multiplied once per apply).  t = 0 must give beta1 --
    not 1.0, which would make a TensorFlow resume divide by 1 - 1 -- and t > 1000 must survive the float32 underflow of 0.9 ** t."""
    for t in (0, 1, 12, 979, 1500, 50000):
        tf_vars = T.from_state_dict({"adam_step": np.int64(t)})
        assert tf_vars["beta1_power"] == np.float32(0.9 ** (t + 1)) and tf_vars["beta2_power"] == np.float32(0.999 ** (t + 1))
        assert int(T.to_state_dict(tf_vars)["adam_step"]) == t, t
    assert T.from_state_dict({"adam_step": np.int64(0)})["beta1_power"] == np.float32(0.9)
    # only beta1_power present (an optimizer saved without beta2): still recovered while representable
    assert int(T.to_state_dict({"beta1_power": np.float32(0.9 ** 13)})["adam_step"]) == 12
    with pytest.warns(UserWarning):
        assert int(T.to_state_dict({"beta1_power": np.float32(0.0), "beta2_power": np.float32(0.0)})["adam_step"]) == 0
