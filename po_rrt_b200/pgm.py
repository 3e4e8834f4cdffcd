"""PNM gray-map reader/writer (P2 ASCII and P5 binary, 8 bit), the decode step of Map::open_image
(reference src/map_io.rs:98-105: image::open -> ImageLuma8, anything else panics).  Thin wrappers: the decoder is the library's
(csrc/formats.cu: porrt_pgm_read / porrt_pgm_write), the same code a Rust or C caller gets."""
import ctypes as C

import numpy as np

from . import _lib


def read_pgm(path):
    lib = _lib.load()
    h, w = C.c_int32(), C.c_int32()
    rc = lib.porrt_pgm_read(None, str(path).encode(), None, 0, C.byref(h), C.byref(w))
    if rc != 4:   # PORRT_ERR_CAPACITY = the size query answered; anything else is the reference's panic
        raise ValueError("Wrong image format! (only 8-bit P2/P5 gray maps decode to ImageLuma8) [status %d]" % rc)
    img = np.empty((h.value, w.value), np.uint8)
    rc = lib.porrt_pgm_read(None, str(path).encode(), img.ctypes.data_as(C.c_void_p), img.size, C.byref(h), C.byref(w))
    if rc != 0:
        raise ValueError("Impossible to open image: %s [status %d]" % (path, rc))
    return img


def write_pgm(path, img, binary=True):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    rc = _lib.load().porrt_pgm_write(None, str(path).encode(), img.ctypes.data_as(C.c_void_p), h, w, 1 if binary else 0)
    if rc != 0:
        raise OSError("porrt_pgm_write failed [status %d]" % rc)
