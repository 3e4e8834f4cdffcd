import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import po_rrt_b200 as P
from po_rrt_b200 import synth
ctx = P.Context(0)
V = Q = 1_000_000
pts, qs = synth.points(V, seed=3), synth.points(Q, seed=4)
r = 2.0 * (np.log(V) / V) ** 0.5
t0=time.perf_counter(); tree = P.KdTree(ctx, pts, cell_size=r); ctx.synchronize(); print("vertices_set %.2f ms" % (1e3*(time.perf_counter()-t0)))
tree.nearest_neighbors(qs[:1000], r)
for rep in range(2):
    t0 = time.perf_counter(); offs, ids = tree.nearest_neighbors(qs, r, cap=64*Q); t1 = time.perf_counter()
    print("radius e2e %.2f ms  phases(ms) [count+scan+fill, sort, d2h]" % (1e3*(t1-t0)), ["%.3f"%x for x in ctx.last_phase_ms()], "hits/q", len(ids)/Q)
for rep in range(2):
    t0 = time.perf_counter(); tree.nearest_neighbor(qs); t1 = time.perf_counter()
    print("nearest e2e %.2f ms phases [kernel, d2h]" % (1e3*(t1-t0)), ["%.3f"%x for x in ctx.last_phase_ms()])
rng = np.random.default_rng(0)
reach = rng.integers(0, 2**63, V, dtype=np.uint64); world = rng.integers(0, 63, Q).astype(np.uint32)
t0 = time.perf_counter(); tree.nearest_neighbor(qs, reach, world); t1 = time.perf_counter()
print("nearest filtered e2e %.2f ms" % (1e3*(t1-t0)), ["%.3f"%x for x in ctx.last_phase_ms()])
for rep in range(2):
    t0 = time.perf_counter(); tree.knn(qs, 16); t1 = time.perf_counter()
    print("knn16 e2e %.2f ms" % (1e3*(t1-t0)), ["%.3f"%x for x in ctx.last_phase_ms()])
occ, zones = synth.door_map(size=8192, n_zones=6, seed=1)
pmap = P.Map(ctx, occ, [-1,-1],[1,1]); pmap.add_zones(zones, 0.3)
for n in (10_000, 100_000, 1_000_000):
    prm = P.PRM(pmap)
    t0 = time.perf_counter(); prm.grow_graph(pts[:n], 0.1, 2.0); t1 = time.perf_counter()
    print("PRM V=%d: %.1f ms edges %d phases" % (n, 1e3*(t1-t0), len(prm.col)), [round(float(x),2) for x in prm.phase_ms])
