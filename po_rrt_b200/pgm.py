"""PNM gray-map reader/writer (P2 ASCII and P5 binary, 8 bit), the decode step of Map::open_image
(reference src/map_io.rs:98-105: image::open -> ImageLuma8, anything else panics)."""
import numpy as np


def _tokens(data, pos, n):
    out = []
    while len(out) < n:
        while pos < len(data) and data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while pos < len(data) and data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        start = pos
        while pos < len(data) and not data[pos:pos + 1].isspace():
            pos += 1
        out.append(data[start:pos])
    return out, pos


def read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    (magic, w, h, maxval), pos = _tokens(data, 0, 4)
    w, h, maxval = int(w), int(h), int(maxval)
    if magic not in (b"P2", b"P5") or not 0 < maxval < 256:
        raise ValueError("Wrong image format! (only 8-bit P2/P5 gray maps decode to ImageLuma8)")
    if magic == b"P5":
        pos += 1  # exactly one whitespace byte after maxval
        img = np.frombuffer(data, np.uint8, w * h, pos).reshape(h, w).copy()
    else:
        vals, _ = _tokens(data, pos, w * h)
        img = np.array([int(v) for v in vals], np.uint8).reshape(h, w)
    return img


def write_pgm(path, img, binary=True):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    with open(path, "wb") as f:
        if binary:
            f.write(b"P5\n%d %d\n255\n" % (w, h))
            f.write(img.tobytes())
        else:
            f.write(b"P2\n%d %d\n255\n" % (w, h))
            for row in img:
                f.write(b" ".join(b"%d" % v for v in row) + b"\n")
