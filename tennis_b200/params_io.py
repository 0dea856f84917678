"""Checkpoint container for `save_parameters` / `load_parameters` (train.py:227,294,345,497).

Writes/reads the MXNet NDArray-list ".params" layout recalled in SURVEY.md §8f-1 ([UPSTREAM], unpinned: no real
MXNet file is available offline to validate against):
  u64 magic 0x112, u64 reserved, u64 count, then per array
  { u32 magic 0xF993FAC9 (or ...CA), i32 stype(0 = dense), u32 ndim, i64 dims[ndim], i32 dev_type(1 = cpu), i32 dev_id,
    i32 dtype flag (0 = float32, 4 = int32, 6 = int64), raw little-endian data },
  then u64 name count and per name { u64 length, bytes }.
Keys are the structural names Gluon's save_parameters uses ("backbone.conv0.weight", "rnn.l0_i2h_weight", ...).
"""
import struct

import numpy as np

_LIST_MAGIC = 0x112
_ND_MAGIC_V2 = 0xF993FAC9
_ND_MAGIC_V3 = 0xF993FACA  # same record layout, written by MXNet >= 1.6 when numpy-shape semantics are on
_DTYPE_FLAG = {np.dtype("float32"): 0, np.dtype("float64"): 1, np.dtype("float16"): 2, np.dtype("uint8"): 3,
               np.dtype("int32"): 4, np.dtype("int8"): 5, np.dtype("int64"): 6}
_FLAG_DTYPE = {v: k for k, v in _DTYPE_FLAG.items()}


def save(filename, arrays):
    with open(filename, "wb") as f:
        f.write(struct.pack("<QQQ", _LIST_MAGIC, 0, len(arrays)))
        for arr in arrays.values():
            a = np.ascontiguousarray(arr)
            f.write(struct.pack("<Ii", _ND_MAGIC_V2, 0))
            f.write(struct.pack("<I", a.ndim))
            f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(struct.pack("<iii", 1, 0, _DTYPE_FLAG[a.dtype]))
            f.write(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
        f.write(struct.pack("<Q", len(arrays)))
        for name in arrays.keys():
            b = name.encode("utf-8")
            f.write(struct.pack("<Q", len(b)))
            f.write(b)


def load(filename):
    with open(filename, "rb") as f:
        buf = f.read()
    off = 0

    def rd(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, buf, off)
        off += struct.calcsize(fmt)
        return v

    magic, _, count = rd("<QQQ")
    if magic != _LIST_MAGIC:
        raise ValueError("%s is not an NDArray-list file (magic 0x%x)" % (filename, magic))
    arrays = []
    for _ in range(count):
        nd_magic, stype = rd("<Ii")
        if nd_magic not in (_ND_MAGIC_V2, _ND_MAGIC_V3) or stype != 0:
            raise ValueError("unsupported NDArray record (magic 0x%x, stype %d)" % (nd_magic, stype))
        (ndim,) = rd("<I")
        shape = rd("<%dq" % ndim) if ndim else ()
        _, _, flag = rd("<iii")
        dt = _FLAG_DTYPE[flag]
        n = int(np.prod(shape)) if ndim else 1
        arrays.append(np.frombuffer(buf, dtype=dt.newbyteorder("<"), count=n, offset=off).reshape(shape).copy())
        off += n * dt.itemsize
    (ncount,) = rd("<Q")
    names = []
    for _ in range(ncount):
        (ln,) = rd("<Q")
        names.append(buf[off:off + ln].decode("utf-8"))
        off += ln
    if ncount != count:
        raise ValueError("name/array count mismatch in %s" % filename)
    out = {}
    for n, a in zip(names, arrays):
        if n.startswith("arg:") or n.startswith("aux:"):
            n = n[4:]
        out[n] = a
    return out
