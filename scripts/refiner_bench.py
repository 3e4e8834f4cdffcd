import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import po_rrt_b200 as P
from po_rrt_b200 import synth
import bench
ctx = P.Context(0)
occ, zones = synth.door_map(size=8192, n_zones=6, seed=1)
pmap = P.Map(ctx, occ, [-1.0, -1.0], [1.0, 1.0]); pmap.add_zones(zones, 0.3)
for n_pieces, n_states, n_it in ((64, 30, 1000), (256, 60, 1000)):
    print(bench.refiner_measurement(ctx, pmap, n_pieces, n_states, n_it))
