"""TensorBundle (TF checkpoint V2) reader: known-answer checks of the primitives and a round trip through the
in-tree writer with the variable names Keras gives the reference model (clair3_rna/model.py:126-156).  No file
written by TensorFlow is available in this image: parity of the reader with real checkpoints is unpinned."""
import os
import struct

import numpy as np
import pytest

from clair3_rna_b200 import tf_bundle, weights

S = tf_bundle.SUFFIX


def keras_keys(w):
    out = {}
    for name, arr in w.items():
        parts = name.split("/")
        if parts[0].startswith("LSTM"):
            key = "%s/%s_layer/cell/%s" % (parts[0], parts[1], parts[2])
        else:
            key = name
        out[key + S] = arr
    return out


def test_crc32c_and_varint_known_answers():
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283          # CRC-32C check value
    assert tf_bundle.crc32c(b"") == 0
    assert tf_bundle._put_varint(300) == b"\xac\x02" and tf_bundle._varint(b"\xac\x02", 0) == (300, 2)
    # snappy raw format: literal "abcd", copy (offset 4, length 8) -> "abcdabcdabcd"
    assert tf_bundle._snappy_decompress(bytes([12, 3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])) == b"abcd" * 3


@pytest.mark.parametrize("channels", [18, 30])
def test_round_trip_with_keras_names(tmp_path, channels):
    w = weights.synthetic(channels, seed=5)
    t = keras_keys(w)
    # what else such a checkpoint holds: the save counter, optimizer slots, a differently typed tensor
    t["save_counter" + S] = np.array(3, np.int64)
    t["LSTM1/forward_layer/cell/kernel/.OPTIMIZER_SLOT/optimizer/m" + S] = np.zeros((channels, 512), np.float32)
    t["optimizer/iter" + S] = np.array(77, np.int64)
    t["half_precision_thing" + S] = np.arange(6, dtype=np.float16).reshape(2, 3)
    prefix = str(tmp_path / "pileup")
    tf_bundle.write_bundle(prefix, t)
    assert os.path.getsize(prefix + ".data-00000-of-00001") >= sum(a.nbytes for a in w.values())
    idx = tf_bundle.read_index(prefix)
    assert idx[""]["num_shards"] == 1 and len(idx) == len(t) + 1
    e = idx["L4/kernel" + S]
    assert e["dtype"] == 1 and e["shape"] == (33 * 320, 128) and e["size"] == 33 * 320 * 128 * 4
    got = tf_bundle.read_tensors(prefix, verify_data=True)
    assert set(got) == set(t)
    for k in t:
        assert got[k].dtype == t[k].dtype and np.array_equal(got[k], t[k]), k
    back = weights.load(prefix)                                  # the path --chkpnt_fn takes
    assert set(back) == set(w)
    for k in w:
        assert np.array_equal(back[k], w[k]), k
    assert tf_bundle.load_keras_checkpoint(prefix)["LSTM1/forward/kernel"].shape[0] == channels


def test_damage_is_detected(tmp_path):
    w = weights.synthetic(18, seed=6)
    prefix = str(tmp_path / "m")
    tf_bundle.write_bundle(prefix, keras_keys(w))
    raw = bytearray(open(prefix + ".index", "rb").read())
    bad = bytearray(raw)
    bad[10] ^= 0x40                                              # inside the first data block
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        tf_bundle.read_index(prefix)
    bad = bytearray(raw)
    bad[-1] ^= 1                                                 # magic
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        tf_bundle.read_index(prefix)
    open(prefix + ".index", "wb").write(bytes(raw))
    # a variable missing from the checkpoint
    t = keras_keys(w)
    del t["L5_2/bias" + S]
    tf_bundle.write_bundle(prefix, t)
    with pytest.raises(ValueError, match="L5_2/bias"):
        tf_bundle.load_keras_checkpoint(prefix)
    # data file damaged
    tf_bundle.write_bundle(prefix, keras_keys(w))
    with open(prefix + ".data-00000-of-00001", "r+b") as fp:
        fp.seek(100)
        fp.write(struct.pack("<f", 123.0))
    with pytest.raises(ValueError, match="checksum"):
        tf_bundle.read_tensors(prefix, verify_data=True)
