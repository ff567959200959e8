"""Reader / writer of the tiny binary container gen_ref_golden.cpp uses (magic "B2REF1\\0\\0", then records: char name[32], u32 dtype
(0 f32, 1 u8, 2 u32, 3 f64), u32 count, payload).  Test infrastructure."""
import struct

import numpy as np

_DT = {0: np.float32, 1: np.uint8, 2: np.uint32, 3: np.float64}
_ID = {np.dtype(v): k for k, v in _DT.items()}


def write(path, records):
    with open(path, "wb") as f:
        f.write(b"B2REF1\0\0")
        for name, arr in records.items():
            a = np.ascontiguousarray(arr)
            if a.dtype not in _ID:
                a = a.astype(np.float32)
            f.write(name.encode()[:31].ljust(32, b"\0"))
            f.write(struct.pack("<II", _ID[a.dtype], a.size))
            f.write(a.tobytes())


def read(path):
    buf = open(path, "rb").read()
    assert buf[:6] == b"B2REF1", path
    out, o = {}, 8
    while o + 40 <= len(buf):
        name = buf[o:o + 32].split(b"\0")[0].decode()
        dt, cnt = struct.unpack("<II", buf[o + 32:o + 40])
        t = np.dtype(_DT[dt])
        out[name] = np.frombuffer(buf, t, cnt, o + 40).copy()
        o += 40 + t.itemsize * cnt
    return out
