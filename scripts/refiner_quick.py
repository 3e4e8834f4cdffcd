"""refiner timings: (a) refine_solution(PartialShortCut(1500)) on a config-3 shaped policy, (b) 64 pieces x 30 states x 1000 trials on the c5 map"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import po_rrt_b200 as P
from po_rrt_b200 import synth
ctx = P.Context(0)
prob = bench.belief_problem(8, 0.5, 0.1, 2.0)
plan, best, _ = bench.belief_run(ctx, prob)
for _ in range(3):
    t0 = time.perf_counter(); ref = P.refine_policy_shortcut(ctx, plan, 1500); t1 = time.perf_counter()
    print("refine_solution(1500): %.2f ms, commits %d, cost %.6f -> %.6f" % (1e3 * (t1 - t0), ref["commits"], plan.expected_cost, ref["expected_cost"]))
t0 = time.perf_counter(); oref = prob["pto"].build_belief_graph([1 / 8.0] * 8); prob["pto"].compute_expected_costs_to_goals(); t1 = time.perf_counter()
t0 = time.perf_counter(); oref = prob["pto"].refine_policy_shortcut(1500); t1 = time.perf_counter()
print("oracle refine: %.2f ms, equal %s" % (1e3 * (t1 - t0), ref["xy"].tobytes() == oref.xy.tobytes()))
occ, zones = synth.door_map(size=8192, n_zones=6, seed=1)
pmap = P.Map(ctx, occ, [-1, -1], [1, 1]); pmap.add_zones(zones, 0.3)
print(bench.refiner_measurement(ctx, pmap))
