"""tests/golden/standin_v1.npz (scripts/make_golden.py): inputs and oracle outputs on the stand-in maps, committed so that
(a) the oracle cannot drift unnoticed (CPU test) and (b) the CUDA path is compared with fixed vectors as well as with the live
oracle (GPU test).  These are ORACLE outputs, not reference outputs: the reference's maps are Git-LFS pointers and its toolchain is
absent; the reference's own map-free golden vectors are in test_oracle_golden.py."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "standin_v1.npz"))
LOW, UP = [-1.0, -1.0], [1.0, 1.0]


def test_oracle_reproduces_golden():
    omap = O.GridMap(G["door_occ"], G["door_zones"], LOW, UP, O.DOOR, 0.6)
    np.testing.assert_array_equal(omap.edge_validity(G["door_a"], G["door_b"]), G["door_edge"])
    np.testing.assert_array_equal(omap.state_validity(G["door_a"]), G["door_state"])
    wm, ws = omap.visible_zones(G["door_a"][:1500])
    np.testing.assert_array_equal(wm, G["door_vis_mask"]); np.testing.assert_array_equal(ws, G["door_vis_status"])
    np.testing.assert_array_equal(omap.zone_positions(), G["door_zone_positions"])
    np.testing.assert_array_equal(omap.world_validities(), G["door_world_validities"])
    for code in (O.PANIC_OOB, O.PANIC_ZONE_UNWRAP, O.PANIC_MULTI_ZONE, O.NONE, 0, 1, 2):
        assert (G["door_edge"] == code).any(), code
    smap = O.GridMap(G["shelf_occ"], G["shelf_zones"], LOW, UP, O.SHELF, 0.7)
    np.testing.assert_array_equal(smap.edge_validity(G["shelf_a"], G["shelf_b"]), G["shelf_edge"])
    np.testing.assert_array_equal(smap.state_validity(G["shelf_a"]), G["shelf_state"])
    wm, ws = smap.visible_zones(G["shelf_a"][:1500])
    np.testing.assert_array_equal(wm, G["shelf_vis_mask"]); np.testing.assert_array_equal(ws, G["shelf_vis_status"])
    pts = G["kd_pts"]
    tree = O.KdTree(pts[0], 0); tree.add_batch(pts[1:], 1)
    offs, ids, tot = tree.radius_batch(G["kd_q"], G["kd_radius"], cap=200000)
    np.testing.assert_array_equal(offs, G["kd_offsets"]); np.testing.assert_array_equal(ids[:tot], G["kd_ids"])
    np.testing.assert_array_equal(tree.nearest_batch(G["kd_q"]), G["kd_nearest"])
    assert float(G["bel_policy_cost"]) > 0 and np.isfinite(G["bel_cost"][0])
    prm = O.PRM(smap, LOW, UP, seed=0); prm.init([0.0, 0.0]); prm.grow_graph(0.15, 3.0, 400)
    xy, nvid, rp, col, ev = prm.graph.export(0)
    np.testing.assert_array_equal(xy, G["prm_xy"]); np.testing.assert_array_equal(rp, G["prm_row_ptr"]); np.testing.assert_array_equal(col, G["prm_col"])
    np.testing.assert_array_equal(prm.graph.dijkstra([0]), G["prm_dijkstra"])


@pytest.mark.gpu
def test_cuda_reproduces_golden():
    import po_rrt_b200 as P
    ctx = P.Context(0)
    try:
        pmap = P.Map(ctx, G["door_occ"], LOW, UP); pmap.add_zones(G["door_zones"], 0.6)
        np.testing.assert_array_equal(pmap.transition_validator(G["door_a"], G["door_b"]).astype(np.int64), G["door_edge"])
        np.testing.assert_array_equal(pmap.state_validity(G["door_a"]).astype(np.int64), G["door_state"])
        gm, gs = pmap.visible_zones(G["door_a"][:1500])
        np.testing.assert_array_equal(gm, G["door_vis_mask"]); np.testing.assert_array_equal(gs.astype(np.int64), G["door_vis_status"])
        np.testing.assert_array_equal(pmap.zone_positions(), G["door_zone_positions"])
        np.testing.assert_array_equal(pmap.world_validities(), G["door_world_validities"])
        smap = P.MapShelfDomain(ctx, G["shelf_occ"], LOW, UP); smap.add_zones(G["shelf_zones"], 0.7)
        # PRM on the shelf map from the oracle's sample stream (prm.rs:38-109) and dijkstra towards node 0
        prm = P.PRM(smap); prm.init([0.0, 0.0]); prm.grow_graph(G["prm_xy"][1:], 0.15, 3.0)
        np.testing.assert_array_equal(prm.row_ptr, G["prm_row_ptr"]); np.testing.assert_array_equal(prm.col, G["prm_col"])
        dist, _ = P.dijkstra_worlds(ctx, prm.row_ptr, prm.col, prm.states, None, None, [0])
        np.testing.assert_array_equal(dist, G["prm_dijkstra"])
        np.testing.assert_array_equal(smap.transition_validator(G["shelf_a"], G["shelf_b"]).astype(np.int64), G["shelf_edge"])
        np.testing.assert_array_equal(smap.state_validity(G["shelf_a"]).astype(np.int64), G["shelf_state"])
        gm, gs = smap.visible_zones(G["shelf_a"][:1500])
        np.testing.assert_array_equal(gm, G["shelf_vis_mask"]); np.testing.assert_array_equal(gs.astype(np.int64), G["shelf_vis_status"])
        # kd-tree: radius sets, restored to the kd pre-order by rank; nearest
        tree = P.KdTree(ctx, G["kd_pts"])
        offs, ids = tree.nearest_neighbors(G["kd_q"], G["kd_radius"])
        np.testing.assert_array_equal(offs, G["kd_offsets"])
        rank = tree.preorder_rank()
        for k in range(len(offs) - 1):
            got = ids[offs[k]:offs[k + 1]]
            np.testing.assert_array_equal(got[np.argsort(rank[got], kind="stable")], G["kd_ids"][offs[k]:offs[k + 1]])
        nid, _, ties = tree.nearest_neighbor(G["kd_q"])
        ok = ties == 1                                              # exact duplicates tie: the kd order decides, on the host
        np.testing.assert_array_equal(nid[ok].astype(np.int64), G["kd_nearest"][ok])
        # belief-space planning + QMDP on the planning map's roadmap
        import porrt_testutil as util
        pocc, pzones = util.planning_door_map(200)
        bmap = P.Map(ctx, pocc, LOW, UP); bmap.add_zones(pzones, 0.5)
        plan = P.plan_belief_space(bmap, G["bel_row_ptr"], G["bel_col"], G["bel_ev"], G["bel_xy"], G["bel_nvid"], list(G["bel_b0"]),
                                   list(G["bel_fin_ids"]), P.words_from_bits(G["bel_fin_bits"]))
        B = len(plan.beliefs)
        np.testing.assert_array_equal(plan.dist.reshape(-1), G["bel_cost"])
        np.testing.assert_array_equal(plan.policy_node.astype(np.int64) * B + plan.policy_belief, G["bel_policy"])
        assert plan.expected_cost == float(G["bel_policy_cost"])
        st, commits, _ = bmap.partial_shortcut(G["ref_path"], [1, 1, 1], 300)
        np.testing.assert_array_equal(st, G["ref_states"]); assert commits == int(G["ref_commits"])
    finally:
        ctx.close()


G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "standin_v2.npz"))
_MM_KEYS = ("mode_node_ptr", "samples", "max_step", "search_radius", "mode_belief_id", "beliefs", "tr_from_mode", "tr_to_mode",
            "tr_pair_ptr", "tr_pairs", "mode_final_ptr", "mode_final_nodes", "expected_costs")


def test_oracle_reproduces_golden_mmprm():
    """multi-modal PRM: the schedule is decided by the restated RNG streams (Pcg64 seed_from_u64 / gen_range, the mode tree),
    so this pins them together with the PRMs, the belief graph and the value backups"""
    smap = O.GridMap(G["shelf_occ"], G["shelf_zones"], LOW, UP, O.SHELF, 0.7)
    tamp = O.TampPRM(smap, LOW, UP)
    pol = tamp.plan((0.0, -0.9), [1.0 / 3] * 3, 0.15, 3.0, 800)
    sch = tamp.schedule()
    for k in _MM_KEYS:
        np.testing.assert_array_equal(sch[k], G2["mm_" + k], err_msg=k)
    np.testing.assert_array_equal(pol.original, G2["mm_policy"])
    assert pol.expected_costs == float(G2["mm_policy_cost"]) and len(pol.leafs) == 3


@pytest.mark.gpu
def test_cuda_reproduces_golden_mmprm():
    import po_rrt_b200 as P
    ctx = P.Context(0)
    try:
        smap = P.MapShelfDomain(ctx, G["shelf_occ"], LOW, UP); smap.add_zones(G["shelf_zones"], 0.7)
        dist, graph, (node, parent, leaf, cost), _ = P.mmprm_plan(smap, {k: G2["mm_" + k] for k in _MM_KEYS})
        np.testing.assert_array_equal(dist, G2["mm_expected_costs"])
        np.testing.assert_array_equal(node.astype(np.int64), G2["mm_policy"])
        np.testing.assert_array_equal(parent.astype(np.int64), G2["mm_policy_parent"])
        assert cost == float(G2["mm_policy_cost"])
    finally:
        ctx.close()


G3 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "standin_v3.npz"))


def _refined_matches(prefix, xy, parent, belief, original, cost):
    np.testing.assert_array_equal(np.asarray(xy), G3[prefix + "_xy"])
    np.testing.assert_array_equal(np.asarray(parent, np.int64), G3[prefix + "_parent"])
    np.testing.assert_array_equal(np.asarray(belief, np.int64), G3[prefix + "_belief"])
    np.testing.assert_array_equal(np.asarray(original, np.int64), G3[prefix + "_original"])
    assert cost == float(G3[prefix + "_cost"])


def test_oracle_reproduces_golden_refiners():
    """refine_solution(PartialShortCut(300)) and refine_solution(Reparent(0.3)) on the planning map's belief-space policy: pins the
    refiners' restatement (sampler stream, decompose / recompose, the priority queue's pop order) against drift"""
    import porrt_testutil as util
    pocc, pzones = util.planning_door_map(200)
    pmap = O.GridMap(pocc, pzones, LOW, UP, O.DOOR, 0.5)
    pto = O.PTO(pmap, LOW, UP, seed=0)
    assert pto.grow_graph((-0.8, -0.8), O.SquareGoal([((0.8, 0.8), [1, 1, 1, 1])], 0.05), 0.05, 5.0, 1500, 100000) == 0
    pto.build_belief_graph(list(G["bel_b0"])); pto.compute_expected_costs_to_goals(); pto.extract_policy()
    sc = pto.refine_policy_shortcut(300)
    _refined_matches("sc", sc.xy, sc.parent, sc.belief_id, sc.original, sc.expected_costs)
    rp = pto.refine_policy_reparent(0.3)
    _refined_matches("rp", rp.xy, rp.parent, rp.belief_id, rp.original, rp.expected_costs)
    np.testing.assert_array_equal(rp.leafs, G3["rp_leafs"])
    assert len(rp.xy) < len(sc.xy)      # reparenting straightens the pieces: fewer policy nodes


@pytest.mark.gpu
def test_cuda_reproduces_golden_refiners():
    import po_rrt_b200 as P
    import porrt_testutil as util
    ctx = P.Context(0)
    try:
        pocc, pzones = util.planning_door_map(200)
        bmap = P.Map(ctx, pocc, LOW, UP); bmap.add_zones(pzones, 0.5)
        plan = P.plan_belief_space(bmap, G["bel_row_ptr"], G["bel_col"], G["bel_ev"], G["bel_xy"], G["bel_nvid"], list(G["bel_b0"]),
                                   list(G["bel_fin_ids"]), P.words_from_bits(G["bel_fin_bits"]))
        B = len(plan.beliefs)
        sc = P.refine_policy_shortcut(ctx, plan, 300)
        _refined_matches("sc", sc["xy"], sc["parent"], sc["belief"], sc["node"].astype(np.int64) * B + sc["belief"], sc["expected_cost"])
        rp = P.refine_policy_reparent(ctx, plan, 0.3)
        _refined_matches("rp", rp["xy"], rp["parent"], rp["belief"], rp["node"].astype(np.int64) * B + rp["belief"], rp["expected_cost"])
        np.testing.assert_array_equal(np.nonzero(rp["is_leaf"])[0], G3["rp_leafs"])
    finally:
        ctx.close()
