"""TensorFlow Saver-V2 ("tensor bundle") checkpoints without TensorFlow: `<prefix>.index` + `<prefix>.data-00000-of-00001`.

The reference saves and restores its models with `tf.train.Saver` (icl_core_lstm.py:107,155,393-394;
icl_multitask_lstm.py:262,756-757), i.e. in this format.  Variables keep their TF names here (`core.Session.state_dict`), so a
reference checkpoint maps 1:1 onto a session:

    vars = tf_checkpoint.read_bundle("/path/model")             # {name: np.ndarray}
    sess.load_state(tf_checkpoint.to_state_dict(vars))          # + Adam slots ('<var>/Adam', '<var>/Adam_1', beta powers)
    tf_checkpoint.write_bundle("/path/model", tf_checkpoint.from_state_dict(sess.state_dict()))

Format (tensorflow/core/util/tensor_bundle + the LevelDB table format it embeds), restated from its published description:
  * `.data-*`: the raw little-endian bytes of every tensor, back to back.
  * `.index`: an SSTable.  Blocks of prefix-compressed entries `varint(shared) varint(non_shared) varint(value_len) key_delta
    value`, then a restart array (uint32 offsets + uint32 count), then a 5-byte trailer (compression type, masked CRC32C).  The
    file ends with a 48-byte footer: BlockHandle(metaindex) BlockHandle(index) (varint64 offset, varint64 size each), zero padding
    to 40 bytes, magic 0xdb4775248b80fb57.  The index block maps a separator key >= the last key of a data block to its handle.
  * key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: VersionDef{1: producer}};
    key <tensor name> -> BundleEntryProto {1: dtype, 2: TensorShapeProto{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size,
    6: fixed32 crc32c (masked) of the tensor bytes}.
The `.meta` graph file a `tf.train.import_meta_graph` call needs (icl_multitask_lstm.py:756) cannot be produced without TF.
Only uncompressed tables (what TF writes) are read.
"""
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
# tensorflow DataType enum values of the dtypes that occur in these checkpoints
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}
_DT_INV = {np.dtype(v): k for k, v in _DT.items()}


# ----------------------------------------------------------------------------------------------------------- crc32c (Castagnoli)
def _make_table():
    poly, tab = 0x82F63B78, []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab.append(c)
    return np.array(tab, dtype=np.uint32)


_TAB = _make_table()


def crc32c(data, crc=0):
    data = bytes(data)
    if len(data) > 4096:                              # the C-ABI library has a fast host implementation
        try:
            import ctypes
            from . import _cabi
            return int(_cabi.lib().icl_crc32c(ctypes.c_char_p(data), len(data), crc))
        except Exception:
            pass
    crc ^= 0xFFFFFFFF
    tab = _TAB
    for b in bytes(data):
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------------------- varints / protobuf
def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _parse_proto(buf):
    """Minimal protobuf reader: {field: [values]} with varints as int, length-delimited as bytes, fixed32/64 as int."""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _entry_proto(dtype, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(int(s))) for s in shape))
    msg = _field(1, 0, _put_varint(dtype)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _put_varint(offset))
    msg += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return msg


# ----------------------------------------------------------------------------------------------------------- SSTable
def _read_block(buf, offset, size):
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable block (type %d): TensorFlow writes checkpoint indices uncompressed" % ctype)
    block = buf[offset:offset + size]
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _build_block(entries, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _emit_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b"\x00" + struct.pack("<I", masked_crc(block + b"\x00")))        # type 0 = no compression, crc covers block + type
    return off, len(block)


# ----------------------------------------------------------------------------------------------------------- public API
def read_bundle(prefix):
    """{tensor name: ndarray} of a Saver-V2 checkpoint `<prefix>.index` / `<prefix>.data-00000-of-00001` (CRCs verified)."""
    idx = open(prefix + ".index", "rb").read()
    if len(idx) < 48 or struct.unpack_from("<Q", idx, len(idx) - 8)[0] != _MAGIC:
        raise ValueError("%s.index is not an SSTable (bad magic)" % prefix)
    footer = idx[-48:]
    _, pos = _get_varint(footer, 0)                   # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    entries = []
    for _, handle in _read_block(idx, ioff, isize):
        boff, p = _get_varint(handle, 0)
        bsize, p = _get_varint(handle, p)
        entries.extend(_read_block(idx, boff, bsize))
    header = _parse_proto(dict(entries).get(b"", b""))
    n_shards = header.get(1, [1])[0]
    if header.get(2, [0])[0] != 0:
        raise ValueError("big-endian checkpoint")
    shards = {}
    out = {}
    for key, val in entries:
        if key == b"":
            continue
        e = _parse_proto(val)
        dtype = e.get(1, [0])[0]
        if dtype not in _DT:
            continue                                  # strings etc. are not variables of this model
        shape = []
        if 2 in e:
            for dim in _parse_proto(e[2][0]).get(2, []):
                shape.append(_parse_proto(dim).get(1, [0])[0])
        shard, off, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        if shard not in shards:
            shards[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, n_shards), "rb").read()
        raw = shards[shard][off:off + size]
        if 6 in e and masked_crc(raw) != e[6][0]:
            raise ValueError("checksum mismatch for tensor %r" % key.decode())
        out[key.decode()] = np.frombuffer(raw, dtype=_DT[dtype]).reshape(shape).copy()
    return out


def write_bundle(prefix, tensors):
    """Write {name: ndarray} as a single-shard Saver-V2 checkpoint (what `tf.train.Saver.restore` reads)."""
    names = sorted(tensors.keys(), key=lambda s: s.encode())
    items, off = [], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.asarray(tensors[n])                                 # 0-d stays 0-d (TF scalars: empty TensorShapeProto)
            if a.dtype not in _DT_INV:
                a = a.astype(np.float32 if a.dtype.kind == "f" else np.int64)
            raw = np.ascontiguousarray(a).tobytes()
            f.write(raw)
            items.append((n.encode(), _entry_proto(_DT_INV[a.dtype], a.shape, off, len(raw), masked_crc(raw))))
            off += len(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))   # 1 shard, little-endian, producer 1
    entries = [(b"", header)] + items
    with open(prefix + ".index", "wb") as f:
        handles = []
        for i in range(0, len(entries), 64):                             # a few data blocks, like a real table
            chunk = entries[i:i + 64]
            o, s = _emit_block(f, _build_block(chunk))
            handles.append((chunk[-1][0], _put_varint(o) + _put_varint(s)))
        mo, ms = _emit_block(f, _build_block([]))
        io, isz = _emit_block(f, _build_block(handles, restart_interval=1))
        footer = _put_varint(mo) + _put_varint(ms) + _put_varint(io) + _put_varint(isz)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    return prefix


# names of the Adam slots tf.train.AdamOptimizer creates next to each variable, and of its two accumulators.
# TF 1.x initialises beta1_power / beta2_power to beta1 / beta2 and multiplies them by beta at the END of every apply
# (AdamOptimizer._create_slots / _finish), so after t update steps a checkpoint holds beta ** (t + 1).
def from_state_dict(st, beta1=0.9, beta2=0.999):
    """core.Session.state_dict() -> TF variable names (Adam slots as '<var>/Adam', '<var>/Adam_1', 'beta1_power', 'beta2_power')."""
    out = {}
    for k, v in st.items():
        if k.startswith("adam_m/"):
            out[k[7:] + "/Adam"] = v
        elif k.startswith("adam_v/"):
            out[k[7:] + "/Adam_1"] = v
        elif k == "adam_step":
            t = int(v)
            out["beta1_power"] = np.float32(beta1 ** (t + 1))
            out["beta2_power"] = np.float32(beta2 ** (t + 1))
        elif "@" not in k.split("/")[0]:
            out[k] = v
    return out


def _steps_from_power(b, beta):
    """t with beta ** (t + 1) == b, or None when b carries no information (0 after float32 underflow, or outside (0, 1))."""
    if not (0.0 < b < 1.0):
        return None
    return max(0, int(round(np.log(b) / np.log(beta))) - 1)


def to_state_dict(tensors, beta1=0.9, beta2=0.999):
    """TF checkpoint variables -> the keys core.Session.load_state understands.  The Adam step count is recovered from
    beta2_power (0.999 ** t stays a normal float32 for ~87 000 steps; 0.9 ** t underflows to 0 near t = 980), falling back to
    beta1_power; when neither is usable the moments are restored with step 0 and a warning (bias correction restarts)."""
    st = {}
    t2 = t1 = None
    have_power = False
    for k, v in tensors.items():
        if k.endswith("/Adam"):
            st["adam_m/" + k[:-5]] = v
        elif k.endswith("/Adam_1"):
            st["adam_v/" + k[:-7]] = v
        elif k == "beta1_power":
            have_power = True
            t1 = _steps_from_power(float(np.asarray(v).reshape(-1)[0]), beta1)
        elif k == "beta2_power":
            have_power = True
            t2 = _steps_from_power(float(np.asarray(v).reshape(-1)[0]), beta2)
        else:
            st[k] = v
    if have_power:
        t = t2 if t2 is not None else t1
        if t is None:
            import warnings
            warnings.warn("TF checkpoint: beta1_power / beta2_power are not in (0, 1); Adam step count set to 0")
            t = 0
        st["adam_step"] = np.int64(t)
    return st
