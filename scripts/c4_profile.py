import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ctypes as C
import po_rrt_b200 as P
from po_rrt_b200 import synth, api
from oracle import pyoracle as O
Z=12
occ, zones = synth.shelf_map(200, n_rects=10, n_zones=Z, seed=5)
low, up = [-1.0, -1.0], [1.0, 1.0]
omap = O.GridMap(occ, zones, low, up, O.SHELF, 0.2)
ctx = P.Context(0)
pmap = P.MapShelfDomain(ctx, occ, low, up); pmap.add_zones(zones, 0.2)
zp = omap.zone_positions()
goals = []
for z in range(Z):
    m = [0] * Z; m[z] = 1
    goals.append(((float(zp[z][0]) - 0.08, float(zp[z][1])), m))
pto = O.PTO(omap, low, up, seed=0)
assert pto.grow_graph((0.0, -0.9), O.SquareGoal(goals, 0.05), 0.05, 5.0, 5000, 100000) == 0
b0 = [1.0 / Z] * Z
xy, nvid, rp, col, ev = pto.graph.export(0)
fin_ids, fin_bits = pto.reach.finals()
fm = P.words_from_bits(fin_bits)
for rep in range(4):
    t0=time.perf_counter(); bel = pmap.reachable_belief_states(b0); t1=time.perf_counter()
    vis, st = pmap.visible_zones(xy); t2=time.perf_counter()
    plan = P.plan_belief_space(pmap, rp, col, ev, xy, nvid, b0, fin_ids, fm, beliefs=bel, copy=(True if rep == 0 else False if rep == 1 else None)); t3=time.perf_counter()
    print("reachable %.1f ms, visible %.1f ms, plan(with beliefs given) %.1f ms, phases %s sweeps %d" % (1e3*(t1-t0),1e3*(t2-t1),1e3*(t3-t2),[round(float(x),1) for x in plan.phase_ms], plan.sweeps))
ctx.set_option(P.OPT_FORCE_GLOBAL_SWEEPS, 1)
for rep in range(2):
    t2 = time.perf_counter(); plan = P.plan_belief_space(pmap, rp, col, ev, xy, nvid, b0, fin_ids, fm, beliefs=bel, copy=None); t3 = time.perf_counter()
    print("global-memory frontier path: plan %.1f ms, phases %s rounds %d, device %.1f ms, pairs %.3g" % (1e3 * (t3 - t2), [round(float(x), 1) for x in plan.phase_ms], plan.sweeps, ctx.last_phase_ms()[0], ctx.last_phase_ms()[1]))
ctx.set_option(P.OPT_FORCE_GLOBAL_SWEEPS, 0)
