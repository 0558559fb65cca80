"""TensorFlow TensorBundle (``tf.train.Checkpoint``) reader / writer for the
exact key set ENERO ships (``modelsSP_3top_15_B_NEW/ckpt_ACT-*``), with no
TensorFlow dependency.

Layout:

* ``<prefix>.index`` -- LevelDB-format SSTable (footer magic 0xdb4775248b80fb57),
  uncompressed blocks with prefix-compressed keys; values are ``BundleEntryProto``
  (1 dtype, 2 shape, 3 shard_id, 4 offset, 5 size, 6 masked crc32c); key ``""`` holds
  ``BundleHeaderProto``.
* ``<prefix>.data-00000-of-00001`` -- raw little-endian tensors at those offsets.

Replaces ``tf.train.Checkpoint(model=actor, optimizer=opt).restore/save`` at
``script_eval_on_single_topology.py:607-609`` and
``train_Enero_3top_script.py:449-457,714-715``.
"""
from __future__ import annotations

import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

DT_FLOAT, DT_STRING, DT_INT64 = 1, 7, 9
_NP_OF = {DT_FLOAT: np.dtype("<f4"), DT_INT64: np.dtype("<i8"), 3: np.dtype("<i4"),
          2: np.dtype("<f8")}
_DT_OF = {np.dtype("float32"): DT_FLOAT, np.dtype("int64"): DT_INT64,
          np.dtype("int32"): 3, np.dtype("float64"): 2}

# Keras variable order (train_Enero_3top_script.py:238-248) -> checkpoint key stems
MODEL_KEYS = (
    "model/Message/layer_with_weights-0/kernel",
    "model/Message/layer_with_weights-0/bias",
    "model/Update/kernel",
    "model/Update/recurrent_kernel",
    "model/Update/bias",
    "model/Readout/layer_with_weights-0/kernel",
    "model/Readout/layer_with_weights-0/bias",
    "model/Readout/layer_with_weights-1/kernel",
    "model/Readout/layer_with_weights-1/bias",
    "model/Readout/layer_with_weights-2/kernel",
    "model/Readout/layer_with_weights-2/bias",
)
MODEL_SHAPES = ((40, 20), (20,), (20, 60), (20, 60), (2, 60), (20, 20), (20,), (20, 20),
                (20,), (20, 1), (1,))
_VAL = "/.ATTRIBUTES/VARIABLE_VALUE"
_SLOT = "/.OPTIMIZER_SLOT/optimizer/%s" + _VAL
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"


# --------------------------------------------------------------------------
# crc32c (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h
# --------------------------------------------------------------------------
def _make_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_TBL = _make_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    tbl = _TBL
    for b in data:
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# --------------------------------------------------------------------------
# varint / protobuf helpers (just the wire types the bundle uses)
# --------------------------------------------------------------------------
def _get_varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
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


def _parse_proto(buf):
    """-> list of (field, wire_type, value)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)
        out.append((f, wt, v))
    return out


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
    for f, _, v in _parse_proto(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            dims = []
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:                       # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            size = v3
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
    return e


def _encode_entry(dtype, shape, offset, size, crc):
    shp = b"".join(b"\x12" + _put_varint(len(d)) + d
                   for d in (b"\x08" + _put_varint(s) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(shp)) + shp
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# --------------------------------------------------------------------------
# SSTable
# --------------------------------------------------------------------------
def _read_block(buf, offset, size, verify=True):
    blk = buf[offset:offset + size]
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable blocks are not supported")
    if verify:
        want = struct.unpack_from("<I", buf, offset + size + 1)[0]
        got = mask_crc(crc32c(bytes(buf[offset:offset + size + 1])))
        if want != got:
            raise ValueError("SSTable block crc mismatch")
    n_restarts = struct.unpack_from("<I", blk, len(blk) - 4)[0]
    end = len(blk) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(blk, pos)
        non_shared, pos = _get_varint(blk, pos)
        vlen, pos = _get_varint(blk, pos)
        key = key[:shared] + bytes(blk[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(blk[pos:pos + vlen])))
        pos += vlen
    return out


def read_index(path: str):
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a TensorBundle index (bad magic)" % path)
    foot = buf[-48:]
    pos = 0
    _, pos = _get_varint(foot, pos)          # metaindex handle
    _, pos = _get_varint(foot, pos)
    ioff, pos = _get_varint(foot, pos)
    isize, pos = _get_varint(foot, pos)
    entries = {}
    for _, handle in _read_block(buf, ioff, isize):
        boff, p = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p)
        for k, v in _read_block(buf, boff, bsize):
            entries[k.decode()] = v
    return entries


def _block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            n = min(len(prev), len(k))
            while shared < n and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


def _short_successor(key: bytes) -> bytes:
    """leveldb BytewiseComparator::FindShortSuccessor."""
    for i, c in enumerate(key):
        if c != 0xFF:
            return key[:i] + bytes([c + 1])
    return key


def _write_table(items) -> bytes:
    data = _block(items)
    out = bytearray(_with_trailer(data))
    meta_off = len(out)
    meta = _block([])
    out += _with_trailer(meta)
    idx_off = len(out)
    handle = _put_varint(0) + _put_varint(len(data))
    idx = _block([(_short_successor(items[-1][0]), handle)])
    out += _with_trailer(idx)
    foot = _put_varint(meta_off) + _put_varint(len(meta)) + _put_varint(idx_off) + _put_varint(len(idx))
    foot += b"\x00" * (40 - len(foot)) + struct.pack("<Q", _MAGIC)
    return bytes(out + foot)


class Bundle:
    """In-memory view of one checkpoint: ``tensors[key] -> np.ndarray | bytes``."""

    def __init__(self):
        self.tensors = {}
        self.order = []

    # ---- reading ---------------------------------------------------------
    @classmethod
    def load(cls, prefix: str, verify: bool = True) -> "Bundle":
        idx = read_index(prefix + ".index")
        hdr = idx.pop("")
        num_shards = 1
        for f, _, v in _parse_proto(hdr):
            if f == 1:
                num_shards = v
        if num_shards != 1:
            raise ValueError("only single-shard bundles are supported")
        data = open(prefix + ".data-00000-of-00001", "rb").read()
        b = cls()
        for key in sorted(idx):
            e = _parse_entry(idx[key])
            raw = data[e["offset"]:e["offset"] + e["size"]]
            if e["dtype"] == DT_STRING:
                # scalar string: varint length, masked crc32c of the length (as uint32), bytes;
                # entry crc = crc32c(uint32 length | masked length crc | bytes), masked
                ln, p = _get_varint(raw, 0)
                val = bytes(raw[p + 4:p + 4 + ln])
                if verify:
                    lcrc = mask_crc(crc32c(struct.pack("<I", ln)))
                    if struct.unpack_from("<I", raw, p)[0] != lcrc:
                        raise ValueError("string length crc mismatch for %s" % key)
                    c = crc32c(val, crc32c(struct.pack("<I", lcrc), crc32c(struct.pack("<I", ln))))
                    if mask_crc(c) != e["crc32c"]:
                        raise ValueError("crc32c mismatch for %s" % key)
            else:
                if verify and mask_crc(crc32c(raw)) != e["crc32c"]:
                    raise ValueError("crc32c mismatch for %s" % key)
                val = np.frombuffer(raw, dtype=_NP_OF[e["dtype"]]).reshape(e["shape"]).copy()
            b.tensors[key] = val
            b.order.append(key)
        return b

    # ---- writing -----------------------------------------------------------
    def save(self, prefix: str) -> None:
        """Writes <prefix>.index and <prefix>.data-00000-of-00001 (single shard).  Data is laid
        out in ``self.order`` (the shipped files use object-graph traversal order); the index
        is a one-data-block SSTable with LevelDB prefix compression (restart interval 16)."""
        data = bytearray()
        entries = {}
        for key in self.order:
            val = self.tensors[key]
            off = len(data)
            if isinstance(val, (bytes, bytearray)):
                ln = len(val)
                lcrc = mask_crc(crc32c(struct.pack("<I", ln)))
                blob = _put_varint(ln) + struct.pack("<I", lcrc) + bytes(val)
                c = crc32c(bytes(val), crc32c(struct.pack("<I", lcrc), crc32c(struct.pack("<I", ln))))
                entries[key] = _encode_entry(DT_STRING, (), off, len(blob), mask_crc(c))
            else:
                arr = np.asarray(val)           # (ascontiguousarray would turn scalars into shape (1,))
                blob = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
                entries[key] = _encode_entry(_DT_OF[np.dtype(arr.dtype.name)], arr.shape, off, len(blob),
                                             mask_crc(crc32c(blob)))
            data += blob
        items = [(b"", b"\x08\x01\x1a\x02\x08\x01")] + [(k.encode(), entries[k]) for k in sorted(entries)]
        index = _write_table(items)
        d = os.path.dirname(prefix)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(prefix + ".data-00000-of-00001", "wb") as fp:
            fp.write(bytes(data))
        with open(prefix + ".index", "wb") as fp:
            fp.write(index)

    # ---- ENERO-specific accessors -----------------------------------------
    def model_weights(self):
        """The 11 model tensors in Keras variable order (fp32)."""
        return [np.ascontiguousarray(self.tensors[k + _VAL], dtype=np.float32) for k in MODEL_KEYS]

    def adam_slots(self, which: str):
        return [np.ascontiguousarray(self.tensors[k + _SLOT % which], dtype=np.float32)
                for k in MODEL_KEYS]

    def optimizer_scalars(self):
        g = self.tensors
        return {
            "iter": int(g["optimizer/iter" + _VAL]),
            "beta_1": float(g["optimizer/beta_1" + _VAL]),
            "beta_2": float(g["optimizer/beta_2" + _VAL]),
            "decay": float(g["optimizer/decay" + _VAL]),
            "learning_rate": float(g["optimizer/learning_rate" + _VAL]),
            "save_counter": int(g["save_counter" + _VAL]),
        }


def load_actor_weights(prefix: str):
    """``checkpoint.restore(dir + '/ckpt_ACT-' + id)`` for the model part."""
    if not os.path.isfile(prefix + ".index"):
        raise FileNotFoundError(prefix + ".index")
    return Bundle.load(prefix).model_weights()


# --------------------------------------------------------------------------
# tf.train.Checkpoint look-alike for (model=myModel, optimizer=Adam)
# --------------------------------------------------------------------------
# order of the tensors inside the shipped .data file (object-graph traversal order)
_DATA_MODEL_ORDER = (2, 3, 4, 0, 1, 5, 6, 7, 8, 9, 10)      # indices into MODEL_KEYS
_SCALARS = (("save_counter", np.int64), ("optimizer/iter", np.int64), ("optimizer/beta_1", np.float32),
            ("optimizer/beta_2", np.float32), ("optimizer/decay", np.float32),
            ("optimizer/learning_rate", np.float32))


def _object_graph_blob() -> bytes:
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "data", "object_graph_model_optimizer.bin"), "rb") as fp:
        return fp.read()


class Checkpoint:
    """``tf.train.Checkpoint(model=..., optimizer=...)`` for ENERO's actor / critic
    (train_Enero_3top_script.py:449-457,714-715; script_eval_on_single_topology.py:607-609).

    ``model`` needs get_weights()/set_weights() (11 tensors, Keras variable order);
    ``optimizer`` is an enero_b200.optim.Adam (slots keyed by model)."""

    def __init__(self, model=None, optimizer=None):
        self.model, self.optimizer = model, optimizer
        self.save_counter = 0

    def restore(self, save_path: str):
        b = Bundle.load(save_path)
        if self.model is not None:
            self.model.set_weights(b.model_weights())
        sc = b.optimizer_scalars()
        self.save_counter = sc["save_counter"]
        if self.optimizer is not None:
            # restore wins over the constructor's hyper-parameters
            self.optimizer.iterations = sc["iter"]
            self.optimizer.learning_rate = np.float32(sc["learning_rate"])
            self.optimizer.beta_1 = np.float32(sc["beta_1"])
            self.optimizer.beta_2 = np.float32(sc["beta_2"])
            self.optimizer.decay = np.float32(sc["decay"])
            if self.model is not None:
                self.optimizer.set_slots(self.model, b.adam_slots("m"), b.adam_slots("v"))
        return self

    def save(self, file_prefix: str) -> str:
        self.save_counter += 1
        path = "%s-%d" % (file_prefix, self.save_counter)
        b = Bundle()
        opt = self.optimizer
        vals = {"save_counter": self.save_counter, "optimizer/iter": opt.iterations,
                "optimizer/beta_1": opt.beta_1, "optimizer/beta_2": opt.beta_2,
                "optimizer/decay": opt.decay, "optimizer/learning_rate": opt.learning_rate}
        for name, dt in _SCALARS:
            b.tensors[name + _VAL] = np.array(vals[name], dtype=dt)
            b.order.append(name + _VAL)
        w = self.model.get_weights()
        m, v = opt.get_slots(self.model)
        for suffix, tensors in ((_VAL, w), (_SLOT % "m", m), (_SLOT % "v", v)):
            for i in _DATA_MODEL_ORDER:
                key = MODEL_KEYS[i] + suffix
                b.tensors[key] = np.asarray(tensors[i], dtype=np.float32).reshape(MODEL_SHAPES[i])
                b.order.append(key)
        b.tensors[OBJECT_GRAPH_KEY] = _object_graph_blob()
        b.order.append(OBJECT_GRAPH_KEY)
        b.save(path)
        d = os.path.dirname(path)
        name = os.path.basename(path)
        with open(os.path.join(d, "checkpoint"), "w") as fp:
            fp.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))
        return path
