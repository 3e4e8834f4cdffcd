"""plan_qmdp at PRM scale: 1e6-node roadmap built on the device, 64 worlds (porrt_sssp_worlds_prm)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import po_rrt_b200 as P
from po_rrt_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = P.Context(0)
occ, zones = synth.door_map(size=8192, n_zones=6, seed=1)
pmap = P.Map(ctx, occ, [-1, -1], [1, 1]); pmap.add_zones(zones, 0.3)
pts = synth.points(n, seed=3)
prm = P.PRM(pmap); prm.grow_graph(pts, 0.1, 2.0)
rng = np.random.default_rng(9)
finals = [rng.choice(n, 4, replace=False).tolist() for _ in range(64)]
for rep in range(2):
    t0 = time.perf_counter(); _, rounds = P.dijkstra_worlds_resident_prm(pmap, n, finals, want_dist=False); t1 = time.perf_counter()
    ph = ctx.last_phase_ms()[:2]
    print("qmdp V=%d: %.1f ms wall, device %.1f ms, rounds %d, pairs %.3g = %.2f full sweeps" % (n, 1e3 * (t1 - t0), ph[0], rounds, ph[1], ph[1] / (len(prm.col) * 64.0)))
