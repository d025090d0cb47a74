"""TensorFlow checkpoint (TensorBundle, "checkpoint V2") reader and writer without TensorFlow.

The reference loads its network with `m.load_weights(args.chkpnt_fn)` (clair3_rna/call_variants.py:1472), i.e. a
Keras object-graph checkpoint `<prefix>.index` + `<prefix>.data-00000-of-00001`.  TensorFlow is not in this
image and no checkpoint is on disk, so this module restates the two published on-disk formats:

  * `<prefix>.index` is a LevelDB-style sorted string table (tensorflow/core/lib/io/table_builder.cc, format.cc):
    data blocks of prefix-compressed (key, value) records with a restart array, each followed by a 5-byte trailer
    (compression type, masked CRC32C), a meta-index block, an index block and a 48-byte footer ending in the
    magic 0xdb4775248b80fb57.  The value of key "" is a BundleHeaderProto, every other value a BundleEntryProto
    (tensorflow/core/protobuf/tensor_bundle.proto: dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6).
  * `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at [offset, offset + size).

PARITY UNPINNED for real files: the reader is checked against the writer below (round trip, tests/
test_tf_bundle_cpu.py), not against a file written by TensorFlow.  Keras names the variables of the reference model
`<attr>/.../.ATTRIBUTES/VARIABLE_VALUE` with attr = LSTM1, LSTM2, L4, L5_1, L5_2, Y_gt21_logits, Y_genotype_logits
(clair3_rna/model.py:126-156), Bidirectional sublayers `forward_layer` / `backward_layer`, LSTM variables under `cell`.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Tuple

import numpy as np

MAGIC = 0xdb4775248b80fb57
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 19: np.dtype("<f2")}
DT_OF = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9, np.dtype("float16"): 19}

# ----------------------------------------------------------------------------- crc32c (Castagnoli), masked like TF
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.arange(256, dtype=np.uint32)
        for _ in range(8):
            t = np.where(t & 1, (t >> 1) ^ np.uint32(0x82F63B78), t >> 1).astype(np.uint32)
        _CRC_TABLE = [int(x) for x in t]
    return _CRC_TABLE


def crc32c(data: bytes, crc: int = 0) -> int:
    t = _crc_table()
    c = crc ^ 0xffffffff
    for b in data:
        c = t[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


# ----------------------------------------------------------------------------- varints / protobuf subset
def _varint(buf: bytes, i: int) -> Tuple[int, int]:
    v = s = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7f) << s
        if b < 0x80:
            return v, i
        s += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _proto_fields(buf: bytes):
    """yield (field number, wire type, value) of a serialized message (varint, fixed32/64, length-delimited)"""
    i = 0
    while i < len(buf):
        tag, i = _varint(buf, i)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, i)[0]
            i += 8
        elif wt == 2:
            n, i = _varint(buf, i)
            v = buf[i:i + n]
            i += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, i)[0]
            i += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _parse_entry(buf: bytes) -> dict:
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=0, sliced=False)
    for f, _wt, v in _proto_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            dims = []
            for f2, _w2, v2 in _proto_fields(v):
                if f2 == 2:                           # TensorShapeProto.dim
                    size = 0
                    for f3, _w3, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


def _entry_bytes(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(n)) for n in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ----------------------------------------------------------------------------- snappy (raw format), decompress only
def _snappy_decompress(buf: bytes) -> bytes:
    n, i = _varint(buf, 0)
    out = bytearray()
    while i < len(buf):
        tag = buf[i]
        i += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[i:i + nb], "little")
                i += nb
            ln += 1
            out += buf[i:i + ln]
            i += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[i]
            i += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[i] | (buf[i + 1] << 8)
            i += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[i:i + 4], "little")
            i += 4
        for _ in range(ln):                           # overlapping copies are allowed
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy: length mismatch")
    return bytes(out)


# ----------------------------------------------------------------------------- table reader
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    body, ctype = buf[offset:offset + size], buf[offset + size]
    if verify:
        want = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if masked_crc(buf[offset:offset + size + 1]) != want:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return body
    if ctype == 1:
        return _snappy_decompress(body)
    raise ValueError("unsupported block compression %d" % ctype)


def _block_records(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    i, key = 0, b""
    while i < end:
        shared, i = _varint(block, i)
        non_shared, i = _varint(block, i)
        vlen, i = _varint(block, i)
        key = key[:shared] + block[i:i + non_shared]
        i += non_shared
        yield key, block[i:i + vlen]
        i += vlen


def read_index(prefix: str, verify: bool = True) -> Dict[str, dict]:
    """{tensor key: entry dict} of `<prefix>.index` (the header record "" is returned under "")."""
    with open(prefix + ".index", "rb") as fp:
        buf = fp.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s.index is not a TensorBundle index (bad magic)" % prefix)
    foot = buf[len(buf) - 48:]
    _mo, i = _varint(foot, 0)
    _ms, i = _varint(foot, i)
    io, i = _varint(foot, i)
    isz, i = _varint(foot, i)
    out = {}
    for _last_key, handle in _block_records(_read_block(buf, io, isz, verify)):
        bo, j = _varint(handle, 0)
        bs, j = _varint(handle, j)
        for key, val in _block_records(_read_block(buf, bo, bs, verify)):
            if key == b"":
                hdr = dict(num_shards=1, endianness=0)
                for f, _wt, v in _proto_fields(val):
                    if f == 1:
                        hdr["num_shards"] = v
                    elif f == 2:
                        hdr["endianness"] = v
                out[""] = hdr
            else:
                out[key.decode("utf-8")] = _parse_entry(val)
    return out


def read_tensors(prefix: str, verify_index: bool = True, verify_data: bool = False) -> Dict[str, np.ndarray]:
    """every numeric tensor of the checkpoint, by key"""
    idx = read_index(prefix, verify_index)
    hdr = idx.pop("", dict(num_shards=1, endianness=0))
    if hdr["endianness"] != 0:
        raise ValueError("big-endian bundles are not supported")
    out, files = {}, {}
    for key, e in idx.items():
        if e["dtype"] not in DTYPES or e["sliced"]:
            continue                                  # strings (the object graph), partitioned variables
        fn = "%s.data-%05d-of-%05d" % (prefix, e["shard_id"], hdr["num_shards"])
        if fn not in files:
            files[fn] = open(fn, "rb")
        fp = files[fn]
        fp.seek(e["offset"])
        raw = fp.read(e["size"])
        dt = DTYPES[e["dtype"]]
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if len(raw) != e["size"] or n * dt.itemsize != e["size"]:
            raise ValueError("tensor %s: %d bytes on disk, shape %s" % (key, len(raw), (e["shape"],)))
        if verify_data and masked_crc(raw) != e["crc32c"]:
            raise ValueError("tensor %s: checksum mismatch" % key)
        out[key] = np.frombuffer(raw, dt).reshape(e["shape"]).copy()
    for fp in files.values():
        fp.close()
    return out


# ----------------------------------------------------------------------------- Keras object-graph keys -> our names
_PATTERNS = [
    (r"^(LSTM[12])/(forward|backward)_layer/(?:cell/)?(kernel|recurrent_kernel|bias)$", r"\1/\2/\3"),
    (r"^(L4|L5_1|L5_2|Y_gt21_logits|Y_genotype_logits)/(kernel|bias)$", r"\1/\2"),
]


def load_keras_checkpoint(prefix: str, channels: int = None) -> Dict[str, np.ndarray]:
    """The pileup network's weights from a Keras TF-format checkpoint, under the names of weights.shapes()."""
    from . import weights as W
    got = {}
    for key, arr in read_tensors(prefix).items():
        if not key.endswith(SUFFIX) or ".OPTIMIZER_SLOT" in key or key.startswith("optimizer"):
            continue
        name = key[:-len(SUFFIX)]
        for pat, rep in _PATTERNS:
            if re.match(pat, name):
                got[re.sub(pat, rep, name)] = np.ascontiguousarray(arr, np.float32)
                break
    if channels is None:
        k = got.get("LSTM1/forward/kernel")
        if k is None:
            raise ValueError("%s: no LSTM1/forward_layer kernel among its keys" % prefix)
        channels = int(k.shape[0])
    want = W.shapes(channels)
    missing = [n for n in want if n not in got]
    if missing:
        raise ValueError("%s: variables not found: %s" % (prefix, ", ".join(missing)))
    for n, shp in want.items():
        if tuple(got[n].shape) != tuple(shp):
            raise ValueError("%s: %s has shape %s, expected %s" % (prefix, n, got[n].shape, shp))
    return {n: got[n] for n in want}


def is_checkpoint_prefix(path: str) -> bool:
    return os.path.exists(path + ".index")


# ----------------------------------------------------------------------------- writer (fixtures, export)
def _build_block(records, restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for n, (key, val) in enumerate(records):
        shared = 0
        if n % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(prev), len(key))
            while shared < m and prev[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_records: int = 24) -> None:
    """`<prefix>.index` + `<prefix>.data-00000-of-00001` holding `tensors` (key -> array), keys sorted as the table
    requires; uncompressed blocks like BundleWriter writes them."""
    keys = sorted(tensors, key=lambda k: k.encode("utf-8"))
    records, offset = [(b"", b"\x08\x01\x1a\x02\x08\x01")], 0   # header: num_shards 1, little endian, version {producer 1}
    with open("%s.data-00000-of-00001" % prefix, "wb") as fp:
        for k in keys:
            a = np.asarray(tensors[k])
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            fp.write(raw)
            records.append((k.encode("utf-8"), _entry_bytes(DT_OF[a.dtype], a.shape, offset, len(raw), masked_crc(raw))))
            offset += len(raw)
    out, index = bytearray(), []

    def emit(block: bytes):
        off = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", masked_crc(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))
    for i in range(0, len(records), block_records):
        chunk = records[i:i + block_records]
        index.append((chunk[-1][0], emit(_build_block(chunk))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    foot = meta + idx
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC))
    with open(prefix + ".index", "wb") as fp:
        fp.write(bytes(out))
