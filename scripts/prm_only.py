import sys, time, numpy as np
sys.path.insert(0, '.')
import po_rrt_b200 as P
from po_rrt_b200 import synth
ctx = P.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pts = synth.points(n, seed=3)
occ, zones = synth.door_map(size=8192, n_zones=6, seed=1)
pmap = P.Map(ctx, occ, [-1,-1],[1,1]); pmap.add_zones(zones, 0.3)
for rep in range(2):
    prm = P.PRM(pmap)
    t0 = time.perf_counter(); prm.grow_graph(pts, 0.1, 2.0); t1 = time.perf_counter()
    print("PRM V=%d: %.1f ms edges %d phases" % (n, 1e3*(t1-t0), len(prm.col)), [round(float(x),2) for x in prm.phase_ms])
