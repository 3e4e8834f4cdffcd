import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import po_rrt_b200 as P
from po_rrt_b200 import synth
ctx=P.Context(0)
occ, zones = synth.door_map(size=8192, n_rects=4096, n_zones=6, seed=1)
pmap = P.Map(ctx, occ, [-1.0,-1.0],[1.0,1.0]); pmap.add_zones(zones, 0.3)
pts = torch.from_numpy(synth.points(1_000_000, seed=3)).pin_memory().numpy()
for n in (100_000, 1_000_000):
    pin_col = torch.empty(64*n, dtype=torch.int32).pin_memory().numpy(); pin_row = torch.empty(n+1, dtype=torch.int64).pin_memory().numpy()
    best=None
    for _ in range(7):
        prm = P.PRM(pmap); t0=time.perf_counter(); prm.grow_graph(pts[:n], 0.1, 2.0, col_out=pin_col, row_ptr_out=pin_row); t=time.perf_counter()-t0
        if best is None or t<best[0]: best=(t,[round(float(x),3) for x in prm.phase_ms[:7]])
    print(n, round(best[0]*1e3,3), best[1])
