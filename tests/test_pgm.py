"""PGM reader/writer of the host-side mirror (Map::open's decode step, map_io.rs:98-105): P2 (ASCII, as data/map0.pgm's size
says it is) and P5 (binary, 200x200 => 40054 bytes like the reference's LFS pointers) decode to the same ImageLuma8 bytes."""
import os

import numpy as np
import pytest

from po_rrt_b200 import pgm, synth


def test_p5_size_matches_reference_pointer_sizes(tmp_path):
    occ, _ = synth.shelf_map(200, n_zones=2)
    p = tmp_path / "m.pgm"
    pgm.write_pgm(str(p), occ, binary=True)
    # the reference's 200x200 P5 blobs are 40054 B = 40000 pixels + a 54-byte header (ours has no comment line: 15 bytes)
    assert os.path.getsize(p) == 200 * 200 + len(b"P5\n200 200\n255\n")
    np.testing.assert_array_equal(pgm.read_pgm(str(p)), occ)


def test_p2_and_p5_decode_identically(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53), dtype=np.uint8)      # non-square: width/height order matters
    a, b = tmp_path / "a.pgm", tmp_path / "b.pgm"
    pgm.write_pgm(str(a), img, binary=True)
    pgm.write_pgm(str(b), img, binary=False)
    np.testing.assert_array_equal(pgm.read_pgm(str(a)), img)
    np.testing.assert_array_equal(pgm.read_pgm(str(b)), img)


def test_comments_and_odd_whitespace(tmp_path):
    p = tmp_path / "c.pgm"
    p.write_bytes(b"P2\n# made by hand\n3 2\n# maxval next\n255\n0 127\t255\n  1\n2 3\n")
    np.testing.assert_array_equal(pgm.read_pgm(str(p)), [[0, 127, 255], [1, 2, 3]])
    q = tmp_path / "d.pgm"
    q.write_bytes(b"P5 2 2 255\n" + bytes([10, 0, 255, 7]))   # 10 == '\\n' as a pixel value right after the header
    np.testing.assert_array_equal(pgm.read_pgm(str(q)), [[10, 0], [255, 7]])


def test_wrong_formats_are_rejected(tmp_path):  # map_io.rs:101-104 panics with "Wrong image format!"
    p = tmp_path / "e.pgm"
    p.write_bytes(b"P6\n1 1\n255\n\x00\x00\x00")
    with pytest.raises(ValueError):
        pgm.read_pgm(str(p))
    p.write_bytes(b"P5\n1 1\n65535\n\x00\x00")
    with pytest.raises(ValueError):
        pgm.read_pgm(str(p))
