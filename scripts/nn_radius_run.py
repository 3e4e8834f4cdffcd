"""radius query at the c5 shape (V = Q = 1e6, r = heuristic_radius(1e6)): the command the NN ncu captures are taken from"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po_rrt_b200 as P
from po_rrt_b200 import synth

ctx = P.Context(0)
V = Q = 1_000_000
pts, qs = synth.points(V, seed=3), synth.points(Q, seed=4)
r = 2.0 * (np.log(V) / V) ** 0.5
tree = P.KdTree(ctx, pts, cell_size=r)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    t0 = time.perf_counter(); offs, ids = tree.nearest_neighbors(qs, r, cap=64 * Q); t1 = time.perf_counter()
    print("radius: %.3f ms wall, phases %s, hits/query %.2f" % (1e3 * (t1 - t0), [round(x, 3) for x in ctx.last_phase_ms()], len(ids) / Q))
