"""Shared fixtures/helpers for the parity tests: build the same map in the oracle (CPU restatement) and in the product."""
import numpy as np

from oracle import pyoracle as O
import po_rrt_b200 as P
from po_rrt_b200 import synth

LOW, UP = [-1.0, -1.0], [1.0, 1.0]


def make_pair(ctx, occ, zones, kind, visibility=0.5, low=LOW, up=UP):
    """-> (oracle GridMap, product Map/MapShelfDomain) on the same images"""
    omap = O.GridMap(occ, zones, low, up, kind, visibility)
    cls = P.Map if kind == P.DOOR else P.MapShelfDomain
    pmap = cls(ctx, occ, low, up)
    if zones is None:
        pmap.init_without_zones()
    else:
        pmap.add_zones(zones, visibility)
    return omap, pmap


def small_door_map(size=512, n_zones=3, seed=11):
    """door map with obstacles; zones far enough apart for edges of length <= 0.1 (25 px at 512)"""
    return synth.door_map(size=size, n_rects=4096, n_zones=n_zones, seed=seed)


def planning_door_map(size=200, seed=3):
    """stand-in for the reference's map2-style maps: two rooms split by a wall with 2 doors (zones 0, 1)"""
    occ = np.full((size, size), 255, np.uint8)
    occ[0, :] = occ[-1, :] = 0
    occ[:, 0] = occ[:, -1] = 0
    w0, w1 = int(0.55 * size), int(0.58 * size)
    occ[:int(0.85 * size), w0:w1] = 0       # vertical wall, open at the bottom (long detour when both doors are closed)
    zones = np.full((size, size), 255, np.uint8)
    for z, (a, b) in enumerate([(0.2, 0.3), (0.65, 0.75)]):
        i0, i1 = int(a * size), int(b * size)
        occ[i0:i1, w0:w1] = 128             # door z
        zones[i0:i1, w0:w1] = z
    rng = np.random.default_rng(seed)
    for _ in range(6):
        h, w = rng.integers(size // 20, size // 8, 2)
        i, j = rng.integers(size // 10, size - h - size // 10), rng.integers(size // 10, w0 - w - 5)
        occ[i:i + h, j:j + w] = 0
    return occ, zones


def csr_from_oracle_graph(g, which=0):
    xy, nvid, rp, col, ev = g.export(which)
    return xy, nvid, rp, col, ev
