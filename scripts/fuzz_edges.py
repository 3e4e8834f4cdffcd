"""Differential fuzz of the edge / state / visibility kernels against the oracle: random map sizes (incl. sides that are not
multiples of 16), both domains, dense / touching zones, gray pixels without zone ids, long and degenerate edges, end points
outside the map.  usage: fuzz_edges.py [rounds] [seed] [edges per round]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import po_rrt_b200 as P
from oracle import pyoracle as O

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n_edges = int(sys.argv[3]) if len(sys.argv) > 3 else 150_000
rng = np.random.default_rng(seed)
ctx = P.Context(0)
bad = 0
for it in range(rounds):
    H, W = (int(rng.integers(40, 1500)), int(rng.integers(40, 1500))) if it % 3 else (int(rng.integers(2, 12)) * 128,) * 2
    kind = P.SHELF if it % 4 == 1 else P.DOOR
    occ = np.full((H, W), 255, np.uint8)
    for _ in range(int(rng.integers(0, 200))):
        h, w = int(rng.integers(1, max(2, H // 5))), int(rng.integers(1, max(2, W // 5)))
        i, j = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        occ[i:i + h, j:j + w] = 0 if (kind == P.DOOR or rng.random() < 0.6) else int(rng.integers(127, 255))
    zones = np.full((H, W), 255, np.uint8)
    nz = int(rng.integers(1, 9)) if kind == P.DOOR else int(rng.integers(1, 13))   # 7-8 door zones: 128 / 256 worlds, two / four mask words
    for z in range(nz):
        h, w = int(rng.integers(2, max(3, H // 8))), int(rng.integers(2, max(3, W // 8)))
        i, j = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        if kind == P.DOOR:
            occ[i:i + h, j:j + w] = 128
        else:
            occ[i:i + h, j:j + w] = 255
        zones[i:i + h, j:j + w] = z
    # every zone id must own at least one pixel (the reference divides by the pixel count)
    for z in range(nz):
        if not (zones == z).any():
            zones[z % H, (7 * z) % W] = z
            if kind == P.DOOR: occ[z % H, (7 * z) % W] = 128
    if kind == P.DOOR and it % 5 == 0:
        i, j = int(rng.integers(0, H - 3)), int(rng.integers(0, W - 3))
        occ[i:i + 3, j:j + 3] = 77                      # gray without zone id
    if kind == P.DOOR:
        zones[occ == 255] = 255; zones[occ == 0] = 255  # zone ids only matter on gray pixels; keep the images consistent
    low, up = [-1.0, -1.0], [1.0, 1.0]
    vis = float(rng.uniform(0.1, 1.0))
    omap = O.GridMap(occ, zones, low, up, kind, vis)
    pmap = (P.MapShelfDomain if kind == P.SHELF else P.Map)(ctx, occ, low, up)
    pmap.add_zones(zones, vis)
    n = n_edges
    a = rng.uniform(-1.05, 1.05, (n, 2))
    L = np.where(rng.random(n) < 0.1, rng.uniform(0, 2.5, n), rng.uniform(0, 0.3, n))
    t = rng.uniform(0, 2 * np.pi, n)
    b = a + np.stack([L * np.cos(t), L * np.sin(t)], 1)
    a[:200] = b[:200]                                    # zero-length
    b[200:400, 0] = a[200:400, 0]                        # axis-aligned
    b[400:600, 1] = a[400:600, 1]
    d = b[600:800] - a[600:800]; d[:, 1] = np.sign(d[:, 1]) * np.abs(d[:, 0]); b[600:800] = a[600:800] + d   # diagonals
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    want = omap.edge_validity(a, b)
    got, masks = pmap.transition_validator(a, b, want_masks=True)
    ok_e = np.array_equal(got.astype(np.int64), want)
    wv = pmap.world_validities_words()
    ok_m = np.array_equal(masks, np.where((want >= 0)[:, None], wv[np.clip(want, 0, None)], 0))
    ok_s = np.array_equal(pmap.state_validity(a).astype(np.int64), omap.state_validity(a))
    nv = min(20000, n); wm, wp = omap.visible_zones(a[:nv]); gm, gs = pmap.visible_zones(a[:nv])
    ok_v = np.array_equal(gm, wm) and np.array_equal(gs.astype(np.int64), wp)
    codes = {int(c): int((want == c).sum()) for c in np.unique(want)}
    print("round %2d %s %4dx%-4d zones %d: edges %s masks %s states %s visibility %s  %s" % (it, "SHELF" if kind == P.SHELF else "DOOR ", H, W, nz, ok_e, ok_m, ok_s, ok_v, codes), flush=True)
    if not (ok_e and ok_m and ok_s and ok_v):
        bad += 1
        if not ok_e:
            k = np.nonzero(got.astype(np.int64) != want)[0][:5]
            for kk in k: print("   edge", kk, a[kk], b[kk], "got", got[kk], "want", want[kk])
print("FUZZ", "FAILED (%d rounds)" % bad if bad else "ok")
sys.exit(1 if bad else 0)
