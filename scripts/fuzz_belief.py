"""Differential fuzz of the value backups against the oracle: random shelf / door maps, random roadmap sizes and start beliefs;
belief-space planning (table, node types, policy), plan_qmdp through both backup paths (on chip / global frontier), the QMDP on
the same roadmap, refine_solution(PartialShortCut) and refine_solution(Reparent).  usage: fuzz_belief.py [rounds] [seed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import po_rrt_b200 as P
from po_rrt_b200 import synth
from oracle import pyoracle as O
import porrt_testutil as util

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = P.Context(0)
bad = 0
for it in range(rounds):
    shelf = it % 3 != 2
    if shelf:
        Z = int(rng.integers(2, 7))
        occ, zones = synth.shelf_map(200, n_rects=int(rng.integers(4, 14)), n_zones=Z, seed=int(rng.integers(1, 1000)))
        vis = float(rng.uniform(0.2, 0.8))
        omap, pmap = util.make_pair(ctx, occ, zones, P.SHELF, vis)
        zp = omap.zone_positions()
        goals = [((float(zp[z][0]) - 0.08, float(zp[z][1])), [1 if k == z else 0 for k in range(Z)]) for z in range(Z)]
        b0 = rng.uniform(0.05, 1.0, Z); b0 = list(b0 / b0.sum()) if it % 2 else [1.0 / Z] * Z
        start, ms, sr = (0.0, -0.9), float(rng.choice([0.05, 0.1, 0.15])), float(rng.choice([2.0, 5.0]))
    else:
        occ, zones = util.planning_door_map(200)
        vis = float(rng.uniform(0.2, 0.5))
        omap, pmap = util.make_pair(ctx, occ, zones, P.DOOR, vis)
        goals = [((0.8, 0.8), [1, 1, 1, 1])]
        b0 = rng.uniform(0.05, 1.0, 4); b0 = list(b0 / b0.sum())
        start, ms, sr = (-0.8, -0.8), 0.05, 5.0
    pto = O.PTO(omap, util.LOW, util.UP, seed=0)
    n_min = int(rng.integers(800, 3000))
    if pto.grow_graph(start, O.SquareGoal(goals, 0.05), ms, sr, n_min, 60000) != 0:
        print("round %2d: roadmap growth did not complete, skipped" % it); continue
    xy, nvid, rp, col, ev = pto.graph.export(0)
    fin_ids, fin_bits = pto.reach.finals()
    try:
        pto.build_belief_graph(b0)
        want = pto.compute_expected_costs_to_goals()
    except RuntimeError as e:
        print("round %2d: reference panic (%s), skipped" % (it, e)); continue
    typ, _, _, _ = pto.belief_graph.export()
    try:
        opol = pto.extract_policy()
    except RuntimeError as e:   # infinite costs at the root: the reference's walk would not return
        try:
            P.plan_belief_space(pmap, rp, col, ev, xy, nvid, b0, fin_ids, P.words_from_bits(fin_bits))
            refused = False
        except P.PorrtError:
            refused = True
        bad += 0 if refused else 1
        print("round %2d: the reference's extract_policy would not return; product refuses: %s" % (it, refused)); continue
    oks = {}
    plan = P.plan_belief_space(pmap, rp, col, ev, xy, nvid, b0, fin_ids, P.words_from_bits(fin_bits))
    B = len(plan.beliefs)
    oks["table"] = np.array_equal(plan.dist.reshape(-1), want) and np.array_equal(plan.type.reshape(-1).astype(np.int32), typ)
    oks["policy"] = (np.array_equal(plan.policy_node.astype(np.int64) * B + plan.policy_belief, opol.original) and
                     np.array_equal(plan.policy_parent.astype(np.int64), opol.parent) and plan.expected_cost == opol.expected_costs)
    ctx.set_option(P.OPT_FORCE_GLOBAL_SWEEPS, 1)
    try:
        plan2 = P.plan_belief_space(pmap, rp, col, ev, xy, nvid, b0, fin_ids, P.words_from_bits(fin_bits))
    finally:
        ctx.set_option(P.OPT_FORCE_GLOBAL_SWEEPS, 0)
    oks["sweeps"] = np.array_equal(plan2.dist, plan.dist)
    W = omap.world_validities().shape[1]
    finals = [pto.reach.get_final_nodes_for_world(w) for w in range(W)]
    if all(len(f) for f in finals):
        q1, _ = P.dijkstra_worlds(ctx, rp, col, xy, nvid, pmap.world_validities_words(), finals)
        ctx.set_option(P.OPT_FORCE_GLOBAL_SWEEPS, 1)
        try:
            q2, _ = P.dijkstra_worlds(ctx, rp, col, xy, nvid, pmap.world_validities_words(), finals)
        finally:
            ctx.set_option(P.OPT_FORCE_GLOBAL_SWEEPS, 0)
        oks["qmdp"] = np.array_equal(q1, pto.plan_qmdp()) and np.array_equal(q1, q2)
    n_it = int(rng.integers(0, 1200))
    P.plan_belief_space(pmap, rp, col, ev, xy, nvid, b0, fin_ids, P.words_from_bits(fin_bits), copy=None)   # (the refiner reads the ctx's last plan)
    got, ow = P.refine_policy_shortcut(ctx, plan, n_it), pto.refine_policy_shortcut(n_it)
    oks["refine"] = got["xy"].tobytes() == ow.xy.tobytes() and np.array_equal(got["parent"], ow.parent) and got["expected_cost"] == ow.expected_costs
    radius = float(rng.choice([0.05, 0.1, 0.2, 0.3, 0.4]))
    try:
        orp = pto.refine_policy_reparent(radius)
    except RuntimeError:
        orp = None   # a reference panic inside is_transition_valid: the product must refuse as well
    try:
        grp = P.refine_policy_reparent(ctx, plan, radius)
    except P.PorrtError:
        grp = None
    oks["reparent"] = (orp is None and grp is None) or (orp is not None and grp is not None and grp["xy"].tobytes() == orp.xy.tobytes() and
                                                        np.array_equal(grp["parent"], orp.parent) and np.array_equal(grp["belief"], orp.belief_id) and
                                                        grp["expected_cost"] == orp.expected_costs)
    ok = all(oks.values())
    bad += 0 if ok else 1
    print("round %2d %s V=%d B=%d policy %d nodes, refine %d trials, reparent r=%.2f: %s" % (it, "SHELF" if shelf else "DOOR ", len(xy), B, len(plan.policy_node), n_it, radius, oks), flush=True)
print("FUZZ", "FAILED (%d rounds)" % bad if bad else "ok")
sys.exit(1 if bad else 0)
