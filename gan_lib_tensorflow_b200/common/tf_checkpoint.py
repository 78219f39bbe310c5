"""Reader (and a small writer) for TensorFlow-1 "tensor bundle" checkpoints -- the files `tf.train.Saver` of the
reference scripts writes (SNGAN/gan_cifar_resnet.py:588, 651-656: `model.ckpt-<step>.index` +
`model.ckpt-<step>.data-00000-of-00001`) and `optimistic_restore` reads through `tf.train.NewCheckpointReader`
(common/misc.py:275-307).  SURVEY 8(f) rank 2: lets a published TF-1 checkpoint (SNGAN/README.md:75-79) be loaded into
the variable store by name + shape without TensorFlow.

Format, restated from the public TensorFlow sources (tensorflow/core/util/tensor_bundle, core/lib/io/table -- the
LevelDB table format; TensorFlow is not installable here, so this is NOT validated against a TensorFlow-produced file,
only by round trips through the independent writer below and hand-built blocks in tests/test_edges.py):

  <prefix>.index   an SSTable: data blocks of prefix-compressed (key, value) entries
                   [varint shared | varint non_shared | varint value_len | key suffix | value]* + uint32 restart
                   offsets + uint32 restart count, each block followed by a 5-byte trailer (compression type, masked
                   crc32c); an index block mapping separator keys to block handles (varint offset, varint size); a
                   48-byte footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).
                   key ""    -> BundleHeaderProto  {1: num_shards, 2: endianness, 3: version}
                   key name  -> BundleEntryProto   {1: dtype, 2: shape {2: dim {1: size}}, 3: shard_id, 4: offset,
                                                    5: size, 6: fixed32 masked crc32c of the bytes}
  <prefix>.data-SSSSS-of-NNNNN   raw little-endian tensor bytes at [offset, offset + size) of shard shard_id.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           14: None, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items() if v is not None}


# ------------------------------------------------------------------------------------------------ varints / protos
def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """Yields (field number, wire type, value) of one serialized message (varint, fixed64, bytes, fixed32)."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
        elif wire == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            val = bytes(buf[pos:pos + ln])
            pos += ln
        elif wire == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        yield field, wire, val


def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for field, _, val in _proto_fields(buf):
        if field == 1:
            e["dtype"] = val
        elif field == 2:       # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
            for f2, _, dim in _proto_fields(val):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _proto_fields(dim):
                        if f3 == 1:
                            size = v3
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = val
        elif field == 4:
            e["offset"] = val
        elif field == 5:
            e["size"] = val
        elif field == 6:
            e["crc32c"] = val
        elif field == 7:
            e["sliced"] = True
    return e


# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli)
_CRC_TABLES = None


def _crc_tables():
    global _CRC_TABLES
    if _CRC_TABLES is None:
        t0 = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t0.append(c)
        tables = [t0]
        for k in range(1, 8):
            prev = tables[k - 1]
            tables.append([(prev[i] >> 8) ^ t0[prev[i] & 0xFF] for i in range(256)])
        _CRC_TABLES = tables
    return _CRC_TABLES


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli) in pure Python, slicing-by-8: a few MB per second -- fine for tests and small exports."""
    t0, t1, t2, t3, t4, t5, t6, t7 = _crc_tables()
    data = bytes(data)
    c = 0xFFFFFFFF
    n8 = len(data) & ~7
    for lo, hi in struct.iter_unpack("<II", data[:n8]):
        lo ^= c
        c = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24]
             ^ t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    for b in data[n8:]:
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    """crc32c::Mask: rotate right by 15 and add a constant (both the table trailers and the bundle entries store this)."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ SSTable reading
def _block_entries(block: bytes):
    """(key, value) pairs of one table block (prefix-compressed keys, restart array at the end)."""
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    block, ctype = buf[offset:offset + size], buf[offset + size]
    if ctype != 0:
        raise NotImplementedError("compressed table block (type %d): TensorFlow writes bundle indices uncompressed" % ctype)
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if masked_crc32c(buf[offset:offset + size + 1]) != stored:
            raise ValueError("table block at %d fails its crc32c" % offset)
    return block


def read_index(index_path: str, verify: bool = False):
    """{tensor name: entry dict} and the header dict of a `<prefix>.index` file."""
    with open(index_path, "rb") as fh:
        buf = fh.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow bundle index (bad table magic)" % index_path)
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _varint(footer, pos)       # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)    # index handle
    isize, pos = _varint(footer, pos)
    entries, header = {}, None
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p2 = _varint(handle, 0)
        bsize, _ = _varint(handle, p2)
        for key, value in _block_entries(_read_block(buf, boff, bsize, verify)):
            if key == b"":
                header = {f: v for f, _, v in _proto_fields(value) if f in (1, 2)}
            else:
                entries[key.decode("utf-8")] = _parse_entry(value)
    if header is not None and header.get(2, 0) != 0:
        raise NotImplementedError("big-endian bundle")
    return entries, {"num_shards": (header or {}).get(1, 1)}


class CheckpointReader:
    """tf.train.NewCheckpointReader(prefix) for tensor-bundle checkpoints: get_variable_to_shape_map(), has_tensor(),
    get_tensor().  `prefix` is what saver.save returned, e.g. './checkpoint/model.ckpt-99999'."""

    def __init__(self, prefix: str, verify: bool = False):
        self.prefix = prefix
        self.verify = verify
        self.entries, hdr = read_index(prefix + ".index", verify=verify)
        self.num_shards = hdr["num_shards"]

    def get_variable_to_shape_map(self):
        return {k: list(e["shape"]) for k, e in self.entries.items()}

    def has_tensor(self, name: str) -> bool:
        return name in self.entries

    def get_tensor(self, name: str) -> np.ndarray:
        e = self.entries[name]
        dtype = _DTYPES.get(e["dtype"])
        if dtype is None or e["sliced"]:
            raise NotImplementedError("tensor %s: dtype code %d / partitioned variables are not supported" % (name, e["dtype"]))
        shard = "%s.data-%05d-of-%05d" % (self.prefix, e["shard_id"], self.num_shards)
        with open(shard, "rb") as fh:
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
        if len(raw) != e["size"]:
            raise ValueError("tensor %s: shard %s is truncated" % (name, shard))
        if self.verify and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise ValueError("tensor %s fails its crc32c" % name)
        return np.frombuffer(raw, dtype=dtype).reshape(e["shape"]).copy()

    def state_dict(self, names=None):
        return {k: self.get_tensor(k) for k in (names if names is not None else self.entries)
                if _DTYPES.get(self.entries[k]["dtype"]) is not None and not self.entries[k]["sliced"]}


def latest_checkpoint(checkpoint_dir: str):
    """tf.train.latest_checkpoint: the prefix named in `<dir>/checkpoint` (model_checkpoint_path: "..."), else the
    `*.index` file with the largest global step."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if os.path.exists(state):
        with open(state) as fh:
            for line in fh:
                if line.startswith("model_checkpoint_path:"):
                    p = line.split(":", 1)[1].strip().strip('"')
                    return p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)
    best = None
    for f in os.listdir(checkpoint_dir) if os.path.isdir(checkpoint_dir) else []:
        if f.endswith(".index"):
            stem = f[:-len(".index")]
            try:
                step = int(stem.rsplit("-", 1)[1])
            except (IndexError, ValueError):
                step = -1
            if best is None or step > best[0]:
                best = (step, os.path.join(checkpoint_dir, stem))
    return best[1] if best else None


# ------------------------------------------------------------------------------------------------ writer (one shard)
def _block(pairs, restart_interval: int = 16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (key, value) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _msg(fields) -> bytes:
    out = bytearray()
    for field, wire, val in fields:
        out += _put_varint((field << 3) | wire)
        if wire == 0:
            out += _put_varint(val)
        elif wire == 2:
            out += _put_varint(len(val)) + val
        elif wire == 5:
            out += struct.pack("<I", val)
    return bytes(out)


def write_checkpoint(prefix: str, tensors: dict, block_entries: int = 64) -> None:
    """Writes {name: ndarray} as a one-shard tensor bundle (`prefix.index`, `prefix.data-00000-of-00001`).  The crc32c
    values are computed in pure Python: meant for exporting small models and for the reader's tests."""
    names = sorted(tensors)
    data, entries = bytearray(), []
    for name in names:
        arr = np.asarray(tensors[name])        # (ascontiguousarray would turn a scalar into shape [1])
        code = _DTYPE_CODES.get(arr.dtype)
        if code is None:
            raise NotImplementedError("dtype %s" % arr.dtype)
        raw = arr.tobytes()
        shape = _msg([(2, 2, _msg([(1, 0, int(d))])) for d in arr.shape])
        entry = _msg([(1, 0, code), (2, 2, shape), (4, 0, len(data)), (5, 0, len(raw)), (6, 5, masked_crc32c(raw))])
        entries.append((name.encode("utf-8"), entry))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))
    header = _msg([(1, 0, 1), (3, 2, _msg([(1, 0, 1)]))])          # num_shards = 1, little-endian, version.producer = 1
    pairs = [(b"", header)] + entries
    out, index_pairs = bytearray(), []

    def emit(block: bytes):
        off = len(out)
        out.extend(block + b"\x00")
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    for i in range(0, len(pairs), block_entries):
        chunk = pairs[i:i + block_entries]
        index_pairs.append((chunk[-1][0] + b"\x00" if i + block_entries < len(pairs) else chunk[-1][0] + b"\xff",
                            emit(_block(chunk))))
    meta = emit(_block([]))
    index = emit(_block(index_pairs, restart_interval=1))
    footer = meta + index
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(out))
    with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as fh:
        fh.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n'
                 % (os.path.basename(prefix), os.path.basename(prefix)))
