"""radius / nearest / knn on the c5 NN shape (V = Q = 1e6), a few repetitions: for ncu launch lists"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import po_rrt_b200 as P
from po_rrt_b200 import synth
ctx = P.Context(0)
V = Q = 1_000_000
pts, qs = synth.points(V, seed=3), synth.points(Q, seed=4)
r = 2.0 * (np.log(V) / V) ** 0.5
tree = P.KdTree(ctx, pts, cell_size=r)
for rep in range(2):
    t0 = time.perf_counter(); offs, ids = tree.nearest_neighbors(qs, r); t1 = time.perf_counter()
    print("radius", t1 - t0, ctx.last_phase_ms())
    t0 = time.perf_counter(); tree.nearest_neighbor(qs); t1 = time.perf_counter()
    print("nearest", t1 - t0, ctx.last_phase_ms())
    t0 = time.perf_counter(); tree.knn(qs, 16); t1 = time.perf_counter()
    print("knn16", t1 - t0, ctx.last_phase_ms())
