"""Tiny dependency-free image helpers (PPM reader, PNG writer) for tests and tools."""
import struct
import zlib

import numpy as np


def read_ppm(path):
    """Read a binary P6 PPM (with optional # comment lines) -> uint8 array (h, w, 3)."""
    b = open(path, "rb").read()
    toks, i = [], 0
    while len(toks) < 4:
        while b[i:i + 1].isspace():
            i += 1
        if b[i:i + 1] == b"#":
            while b[i:i + 1] != b"\n":
                i += 1
            continue
        j = i
        while not b[j:j + 1].isspace():
            j += 1
        toks.append(b[i:j])
        i = j
    i += 1
    assert toks[0] == b"P6", toks
    w, h, mx = int(toks[1]), int(toks[2]), int(toks[3])
    if mx < 256:
        return np.frombuffer(b[i:i + w * h * 3], dtype=np.uint8).reshape(h, w, 3).copy()
    return np.frombuffer(b[i:i + w * h * 6], dtype=">u2").reshape(h, w, 3).astype(np.uint16)


def write_png(path, img):
    """Write uint8 (h, w, 3|4) as PNG."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape
    ct = {3: 2, 4: 6}[c]
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return (struct.pack(">I", len(data)) + tag + data +
                struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF))

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ct, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))
