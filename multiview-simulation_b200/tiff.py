"""Minimal float32 TIFF stack IO (host side, no dependency).

Reads what the reference's Tools.open reads (S/Tools.java:162-238: float32 only, big or little
endian, one IFD per slice -- the PSF fixtures src/main/resources/Angle*.tif are big-endian,
uncompressed, 51 IFDs) and writes stacks ImageJ opens (Tools.save, S/Tools.java:88-105).
"""
import struct

import numpy as np


def read_float_stack(path):
    with open(path, "rb") as f:
        buf = f.read()
    bo = {b"II": "<", b"MM": ">"}.get(buf[:2])
    if bo is None or struct.unpack(bo + "H", buf[2:4])[0] != 42:
        raise ValueError(f"{path}: not a classic TIFF")
    off = struct.unpack(bo + "I", buf[4:8])[0]
    slices = []
    tsize = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 11: 4, 12: 8}
    while off:
        n = struct.unpack(bo + "H", buf[off:off + 2])[0]
        tags = {}
        for i in range(n):
            e = off + 2 + 12 * i
            tag, typ, cnt = struct.unpack(bo + "HHI", buf[e:e + 8])
            size = tsize.get(typ, 1) * cnt
            voff = e + 8 if size <= 4 else struct.unpack(bo + "I", buf[e + 8:e + 12])[0]
            fmt = {3: "H", 4: "I"}.get(typ)
            tags[tag] = struct.unpack(bo + fmt * cnt, buf[voff:voff + size]) if fmt else (voff, cnt)
        off = struct.unpack(bo + "I", buf[off + 2 + 12 * n:off + 6 + 12 * n])[0]
        w, h = tags[256][0], tags[257][0]
        if tags.get(258, (1,))[0] != 32 or tags.get(339, (1,))[0] != 3 or tags.get(259, (1,))[0] != 1:
            raise ValueError(f"{path}: only uncompressed 32-bit float TIFF is supported (like Tools.open)")
        strips, counts = tags[273], tags[279]
        data = b"".join(buf[o:o + c] for o, c in zip(strips, counts))
        slices.append(np.frombuffer(data, dtype=bo + "f4", count=w * h).reshape(h, w))
    return np.ascontiguousarray(np.stack(slices), dtype=np.float32)


def write_float_stack(path, vol):
    """Writes (Z, Y, X) float32 as a little-endian multi-IFD TIFF with an ImageJ stack description."""
    vol = np.ascontiguousarray(vol, dtype="<f4")
    if vol.ndim == 2:
        vol = vol[None]
    z, h, w = vol.shape
    desc = f"ImageJ=1.48p\nimages={z}\nslices={z}\nloop=false\n".encode() + b"\0"
    plane = w * h * 4
    ntags = 9
    ifd_size = 2 + 12 * ntags + 4
    data_off = 8 + len(desc) + (len(desc) & 1)
    ifd_off = data_off + plane * z
    out = bytearray(struct.pack("<2sHI", b"II", 42, ifd_off if z else 0))
    out += desc + (b"\0" if len(desc) & 1 else b"")
    out += vol.tobytes()
    for i in range(z):
        nxt = ifd_off + ifd_size * (i + 1) if i + 1 < z else 0
        ent = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 32), (259, 3, 1, 1), (262, 3, 1, 1),
               (270, 2, len(desc), 8), (273, 4, 1, data_off + plane * i), (279, 4, 1, plane), (339, 3, 1, 3)]
        out += struct.pack("<H", ntags)
        for tag, typ, cnt, val in ent:
            out += struct.pack("<HHII", tag, typ, cnt, val)
        out += struct.pack("<I", nxt)
    with open(path, "wb") as f:
        f.write(out)
