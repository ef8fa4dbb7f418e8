"""`.svw` — the voice-weight blob read by libsonata_b200 (csrc/voice.cpp) and by the oracle.

The reference never touches weights itself: it hands a Piper ``<voice>.onnx`` file to
onnxruntime (``crates/sonata/models/piper/src/lib.rs:79-86``).  No Piper voice and no ONNX
reader exist in this sandbox, so the voice file here is a flat, named, fp32 tensor table
that uses Piper's *state-dict* names (``enc_p.emb.weight``, ``dec.ups.0.weight`` …) so a
real checkpoint can be dropped into the same container later.

Layout (little endian)::

    magic   8 bytes  b"SVW1\\0\\0\\0\\0"
    count   u32
    repeat count times:
        name_len u16, name bytes (utf-8)
        dtype    u8   (0 = f32, 1 = i32)
        ndim     u8
        dims     u32 * ndim
        pad      to a 16-byte boundary (relative to file start)
        data     prod(dims) * 4 bytes
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"SVW1\0\0\0\0"
_DT = {0: np.float32, 1: np.int32}


def write_svw(path, tensors: "OrderedDict[str, np.ndarray]") -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(tensors)))
        for name, arr in tensors.items():
            arr = np.asarray(arr, order="C")
            if arr.dtype == np.float32:
                dt = 0
            elif arr.dtype == np.int32:
                dt = 1
            else:
                raise TypeError(f"{name}: unsupported dtype {arr.dtype}")
            nb = name.encode("utf-8")
            f.write(struct.pack("<H", len(nb)))
            f.write(nb)
            f.write(struct.pack("<BB", dt, arr.ndim))
            f.write(struct.pack("<%dI" % arr.ndim, *arr.shape))
            pos = f.tell()
            f.write(b"\0" * ((-pos) % 16))
            f.write(arr.tobytes())


def read_svw(path) -> "OrderedDict[str, np.ndarray]":
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not an SVW1 file")
    (count,) = struct.unpack_from("<I", buf, 8)
    pos = 12
    for _ in range(count):
        (nl,) = struct.unpack_from("<H", buf, pos)
        pos += 2
        name = buf[pos:pos + nl].decode("utf-8")
        pos += nl
        dt, nd = struct.unpack_from("<BB", buf, pos)
        pos += 2
        dims = struct.unpack_from("<%dI" % nd, buf, pos)
        pos += 4 * nd
        pos += (-pos) % 16
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(buf, dtype=_DT[dt], count=n, offset=pos).reshape(dims).copy()
        pos += 4 * n
        out[name] = arr
    return out
