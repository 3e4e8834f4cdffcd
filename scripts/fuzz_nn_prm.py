"""Differential fuzz of the nearest-neighbour / PRM / multi-modal PRM paths against the oracle: vertex sets that are uniform,
clustered or on a lattice with exact duplicates; radii from 0 to "everything" (register, block-level and radix segment sorts);
prefix limits; PRM builds with random (max_step, search_radius) on random shelf maps; multi-modal PRMs with 2-4 shelves.
Dev tool like tests/: the oracle is the checker.  usage: fuzz_nn_prm.py [rounds] [seed]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import po_rrt_b200 as P
from po_rrt_b200 import synth
from oracle import pyoracle as O

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = P.Context(0)
LOW, UP = [-1.0, -1.0], [1.0, 1.0]
bad = 0


def vertex_set(n):
    kind = int(rng.integers(0, 3))
    if kind == 0:
        p = rng.uniform(-1, 1, (n, 2))
    elif kind == 1:                                   # clusters: crowded cells next to empty ones
        c = rng.uniform(-0.9, 0.9, (max(1, n // 400), 2))
        p = c[rng.integers(0, len(c), n)] + rng.normal(0, 0.01, (n, 2))
    else:                                             # lattice: exact ties in distance, exact duplicates
        g = int(max(2, np.sqrt(n) / 2))
        p = np.stack([rng.integers(0, g, n), rng.integers(0, g, n)], 1) / g * 1.8 - 0.9
    if n > 10:
        k = int(rng.integers(1, max(2, n // 10)))
        p[rng.integers(0, n, k)] = p[rng.integers(0, n, k)]          # duplicates of earlier / later points
    return np.ascontiguousarray(p)


t_start = time.time()
for it in range(rounds):
    # ---------------- radius / nearest
    n = int(rng.integers(1, 30_000))
    pts = vertex_set(n)
    m = int(rng.integers(1, 600)) if it % 4 else int(rng.integers(2048, 6000))   # every fourth round is large enough for the tile kernels
    q = np.ascontiguousarray(np.concatenate([rng.uniform(-1.2, 1.2, (m, 2)), pts[rng.integers(0, n, max(1, m // 4))]]))
    m = len(q)
    r = rng.choice([0.0, 1e-9, 0.003, 0.02, 0.1, 0.5, 3.0], m, p=[0.1, 0.05, 0.2, 0.3, 0.2, 0.1, 0.05]) * rng.uniform(0.5, 1.5, m)
    prefix = rng.integers(0, n + 1, m).astype(np.uint32) if it % 2 else None
    tree = P.KdTree(ctx, pts, cell_size=float(rng.choice([0.0, 0.01, 0.05, 0.3])))
    offs, ids = tree.nearest_neighbors(q, r, prefix_limit=prefix)
    otree = O.KdTree(pts[0], 0)
    if n > 1:
        otree.add_batch(pts[1:], 1)
    rank = tree.preorder_rank()
    for k in range(m):
        want = np.asarray(otree.nearest_neighbors(q[k], r[k]), np.int64)
        if prefix is not None:
            want = want[want < prefix[k]]
        got = ids[offs[k]:offs[k + 1]].astype(np.int64)
        if not (np.array_equal(got, np.sort(want)) and np.array_equal(got[np.argsort(rank[got], kind="stable")], want)):
            bad += 1
            print("RADIUS MISMATCH round", it, "query", k, "n", n, "r", r[k], len(got), len(want))
            break
    nid, nd, ties = tree.nearest_neighbor(q)
    onear = otree.nearest_batch(q)
    d_got = np.sqrt(((pts[nid] - q) ** 2).sum(1))
    d_want = np.sqrt(((pts[onear] - q) ** 2).sum(1))
    if not np.array_equal(d_got, d_want) or not np.array_equal(nid[ties == 1].astype(np.int64), onear[ties == 1]):
        bad += 1
        print("NEAREST MISMATCH round", it)
    # ---------------- PRM on a random shelf map
    Z = int(rng.integers(2, 5))
    size = int(rng.choice([120, 200, 333]))
    occ, zones = synth.shelf_map(size, n_rects=int(rng.integers(2, 14)), n_zones=Z, seed=int(rng.integers(0, 1 << 30)))
    vis = float(rng.uniform(0.2, 0.8))
    omap = O.GridMap(occ, zones, LOW, UP, O.SHELF, vis)
    pmap = P.MapShelfDomain(ctx, occ, LOW, UP)
    pmap.add_zones(zones, vis)
    n_iter = int(rng.integers(1, 5000))
    ms, sr = float(rng.choice([0.02, 0.05, 0.1, 0.3])), float(rng.choice([0.5, 2.0, 5.0, 9.0]))
    oprm = O.PRM(omap, LOW, UP, seed=0)
    oprm.init([0.0, 0.0])
    oprm.grow_graph(ms, sr, n_iter)
    xy, _, rp, col, _ = oprm.graph.export(0)
    prm = P.PRM(pmap)
    prm.init([0.0, 0.0])
    prm.grow_graph(xy[1:], ms, sr)
    if not (np.array_equal(prm.row_ptr, rp) and np.array_equal(prm.col, col)):
        bad += 1
        print("PRM MISMATCH round", it, n_iter, ms, sr)
    # ---------------- multi-modal PRM
    if it % 3 == 0:
        tamp = O.TampPRM(omap, LOW, UP)
        try:
            opol = tamp.plan((0.0, -0.9), [1.0 / Z] * Z, 0.1, 2.0, int(rng.integers(600, 1800)))
        except RuntimeError:
            opol = None                               # no policy on this map: nothing to compare
        if opol is not None:
            sch = tamp.schedule()
            dist, graph, pol, _ = P.mmprm_plan(pmap, sch)
            typ, bid, brp, bcol = tamp.belief_graph.export()
            if not (np.array_equal(dist, sch["expected_costs"]) and np.array_equal(graph.col.astype(np.int64), bcol) and
                    np.array_equal(pol[0].astype(np.int64), opol.original)):
                bad += 1
                print("MMPRM MISMATCH round", it, Z)
    print("round %d ok so far (bad=%d): n=%d m=%d hits=%d | prm %d nodes %d edges | %.0f s" % (it, bad, n, m, len(ids), len(xy), len(col), time.time() - t_start), flush=True)
print("fuzz_nn_prm: %d rounds, %d mismatches" % (rounds, bad))
sys.exit(1 if bad else 0)
