"""BASELINE configs c3/c4 shape: PTO roadmap growth (sequential, CPU side: the oracle plays the reference) on a shelf map
with Z goal zones, then belief-space planning -- oracle (materialised belief graph + conditional_dijkstra) vs the GPU
implicit-graph value iteration.  Prints timings and checks bit-exact parity."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import po_rrt_b200 as P
from po_rrt_b200 import synth
from oracle import pyoracle as O

def goal_map(Z, size=200, seed=5):
    occ, zones = synth.shelf_map(size, n_rects=10, n_zones=Z, seed=seed)
    return occ, zones

def run(ctx, Z, n_min, max_step, search_radius, visibility, skip_oracle_belief=False):
    occ, zones = goal_map(Z)
    low, up = [-1.0, -1.0], [1.0, 1.0]
    omap = O.GridMap(occ, zones, low, up, O.SHELF, visibility)
    pmap = P.MapShelfDomain(ctx, occ, low, up); pmap.add_zones(zones, visibility)
    zp = omap.zone_positions()
    goals = []
    for z in range(Z):
        m = [0] * Z; m[z] = 1
        goals.append(((float(zp[z][0]) - 0.08, float(zp[z][1])), m))
    goal = O.SquareGoal(goals, 0.05)
    pto = O.PTO(omap, low, up, seed=0)
    t0 = time.perf_counter(); rc = pto.grow_graph((0.0, -0.9), goal, max_step, search_radius, n_min, 100000); t_grow = time.perf_counter() - t0
    V, E = pto.graph.n_nodes(), pto.graph.n_edges()
    print("Z=%d grow rc=%d n_it=%d V=%d E=%d  %.1f ms" % (Z, rc, pto.n_it(), V, E, 1e3 * t_grow))
    if rc != 0: return
    b0 = [1.0 / Z] * Z
    xy, nvid, rp, col, ev = pto.graph.export(0)
    fin_ids, fin_bits = pto.reach.finals()
    for rep in range(2):
        t0 = time.perf_counter()
        plan = P.plan_belief_space(pmap, rp, col, ev, xy, nvid, b0, fin_ids, P.words_from_bits(fin_bits))
        t_gpu = time.perf_counter() - t0
    B = len(plan.beliefs)
    print("   GPU plan_belief_space: %.1f ms total, B=%d, V*B=%d, sweeps=%d, phases(ms)[host tables, upload+types, sweeps, download]=%s, policy nodes %d, cost %.6f"
          % (1e3 * t_gpu, B, V * B, plan.sweeps, [round(float(x), 2) for x in plan.phase_ms], len(plan.policy_node), plan.expected_cost))
    if skip_oracle_belief: return
    t0 = time.perf_counter(); pto.build_belief_graph(b0); t_build = time.perf_counter() - t0
    t0 = time.perf_counter(); want = pto.compute_expected_costs_to_goals(); t_dp = time.perf_counter() - t0
    t0 = time.perf_counter(); opol = pto.extract_policy(); t_pol = time.perf_counter() - t0
    print("   CPU oracle: build_belief_graph %.1f ms, conditional_dijkstra %.1f ms, extract_policy %.1f ms" % (1e3 * t_build, 1e3 * t_dp, 1e3 * t_pol))
    ok = np.array_equal(plan.dist.reshape(-1), want) and np.array_equal(plan.policy_node.astype(np.int64) * B + plan.policy_belief, opol.original)
    print("   parity (bit-exact dist + identical policy):", ok, " speedup belief planning: %.1fx" % ((t_build + t_dp + t_pol) / t_gpu))

ctx = P.Context(0)
run(ctx, 2, 2000, 0.05, 5.0, 0.5)
run(ctx, 8, 5000, 0.1, 2.0, 0.5)
run(ctx, 10, 5000, 0.05, 5.0, 0.3)
run(ctx, 12, 5000, 0.05, 5.0, 0.2, skip_oracle_belief=(len(sys.argv) > 1 and sys.argv[1] == "skip12"))
