"""One pass over the non-headline kernels for ncu (launch list / --set full): radius query at the c5 NN shape (tiles + register
segment sort), PRM build V = 1e6 on the c5 map, belief-space planning at the c3 shape (implicit VI with work skipping) and the
explicit-graph conditional_dijkstra on the same problem materialised.  Dev tool (like tests/): the oracle grows the PTO roadmap
(sequential planner, stays on the CPU) and materialises the belief graph that is fed to the explicit entry point."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import po_rrt_b200 as P
from po_rrt_b200 import synth
from oracle import pyoracle as O

ctx = P.Context(0)
what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("all", "nn"):
    V = Q = 1_000_000
    pts, qs = synth.points(V, seed=3), synth.points(Q, seed=4)
    r = 2.0 * (np.log(V) / V) ** 0.5
    tree = P.KdTree(ctx, pts, cell_size=r)
    for rep in range(2):
        t0 = time.perf_counter(); offs, ids = tree.nearest_neighbors(qs, r); t1 = time.perf_counter()
        print("radius", t1 - t0, ctx.last_phase_ms())

if what in ("all", "prm"):
    occ, zones = synth.door_map(size=8192, n_rects=4096, n_zones=6, seed=1)
    pmap = P.Map(ctx, occ, [-1.0, -1.0], [1.0, 1.0])
    pmap.add_zones(zones, 0.3)
    pts = synth.points(1_000_000, seed=3)
    for rep in range(2):
        prm = P.PRM(pmap)
        t0 = time.perf_counter(); prm.grow_graph(pts, 0.1, 2.0); t1 = time.perf_counter()
        print("prm", t1 - t0, [round(float(x), 3) for x in prm.phase_ms[:7]])

if what in ("all", "belief"):
    Z = 8
    occ, zones = synth.shelf_map(200, n_rects=10, n_zones=Z, seed=5)
    low, up = [-1.0, -1.0], [1.0, 1.0]
    omap = O.GridMap(occ, zones, low, up, O.SHELF, 0.5)
    smap = P.MapShelfDomain(ctx, occ, low, up)
    smap.add_zones(zones, 0.5)
    zp = omap.zone_positions()
    goals = []
    for z in range(Z):
        m = [0] * Z
        m[z] = 1
        goals.append(((float(zp[z][0]) - 0.08, float(zp[z][1])), m))
    pto = O.PTO(omap, low, up, seed=0)
    assert pto.grow_graph((0.0, -0.9), O.SquareGoal(goals, 0.05), 0.1, 2.0, 5000, 100000) == 0
    b0 = [1.0 / Z] * Z
    xy, nvid, rp, col, ev = pto.graph.export(0)
    fin_ids, fin_bits = pto.reach.finals()
    for rep in range(2):
        t0 = time.perf_counter()
        plan = P.plan_belief_space(smap, rp, col, ev, xy, nvid, b0, fin_ids, P.words_from_bits(fin_bits))
        print("belief implicit", time.perf_counter() - t0, plan.sweeps, [round(float(x), 2) for x in plan.phase_ms])
    # the same problem as an explicit BeliefGraph (what PTO::build_belief_graph materialises) through porrt_conditional_dijkstra
    pto.build_belief_graph(b0)
    typ, bid, rp_b, col_b = pto.belief_graph.export()
    B = len(plan.beliefs)
    xy_b = np.repeat(xy, B, axis=0)
    finals = np.nonzero(plan.dist.reshape(-1) == 0.0)[0]
    g = P.BeliefGraph(ctx, rp_b, col_b, xy_b, typ, bid, plan.beliefs)
    for rep in range(2):
        t0 = time.perf_counter(); d = g.conditional_dijkstra(finals); t1 = time.perf_counter()
        print("belief explicit", t1 - t0, g.sweeps, "belief nodes", g.V, "edges", len(col_b))
    assert np.array_equal(d, plan.dist.reshape(-1)), "explicit and implicit value tables differ"
    print("explicit == implicit: bit-identical")
