"""Regenerates tests/golden/standin_v1.npz: inputs and ORACLE outputs for the stand-in maps (the reference's own maps are
Git-LFS pointers and its toolchain is absent, so these vectors pin the oracle against drift -- they are not reference output;
the reference's own map-free golden vectors live in tests/test_oracle_golden.py).
usage: python scripts/make_golden.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import pyoracle as O
import porrt_testutil as util
from po_rrt_b200 import synth

out = {}
rng = np.random.default_rng(2024)
# ---- door map with two touching zones, a gray patch without zone id, obstacles; edges with every outcome
size = 96
occ = np.full((size, size), 255, np.uint8); zones = np.full((size, size), 255, np.uint8)
occ[10:30, 50:60] = 0; occ[60:70, 10:80] = 0; occ[40:44, 30:34] = 0
occ[35:50, 70:75] = 128; zones[35:50, 70:75] = 0
occ[35:50, 75:80] = 128; zones[35:50, 75:80] = 1
occ[80:84, 5:40] = 90
low, up = [-1.0, -1.0], [1.0, 1.0]
omap = O.GridMap(occ, zones, low, up, O.DOOR, 0.6)
a = rng.uniform(-1.15, 1.15, (4000, 2)); b = a + rng.uniform(-0.5, 0.5, (4000, 2))
a[:50] = b[:50]
out.update(door_occ=occ, door_zones=zones, door_a=a, door_b=b, door_edge=omap.edge_validity(a, b), door_state=omap.state_validity(a))
wm, ws = omap.visible_zones(a[:1500]); out.update(door_vis_mask=wm, door_vis_status=ws)
out["door_zone_positions"] = omap.zone_positions(); out["door_world_validities"] = omap.world_validities()
# ---- shelf map
socc, szones = synth.shelf_map(120, n_rects=8, n_zones=3, seed=9)
smap = O.GridMap(socc, szones, low, up, O.SHELF, 0.7)
sa, sb = synth.edges(3000, seed=5, max_len=0.6)
out.update(shelf_occ=socc, shelf_zones=szones, shelf_a=sa, shelf_b=sb, shelf_edge=smap.edge_validity(sa, sb), shelf_state=smap.state_validity(sa))
wm, ws = smap.visible_zones(sa[:1500]); out.update(shelf_vis_mask=wm, shelf_vis_status=ws)
# ---- kd-tree: radius sets in kd pre-order, nearest
pts = rng.uniform(-1, 1, (600, 2)); pts[100:120] = pts[0:20]
q = rng.uniform(-1.1, 1.1, (120, 2)); radius = rng.uniform(0.0, 0.4, 120)
tree = O.KdTree(pts[0], 0); tree.add_batch(pts[1:], 1)
offs, ids, tot = tree.radius_batch(q, radius, cap=200000)
out.update(kd_pts=pts, kd_q=q, kd_radius=radius, kd_offsets=offs, kd_ids=ids[:tot], kd_nearest=tree.nearest_batch(q))
# ---- PRM on the shelf map (prm.rs:38-109; the door map above has touching zones: the reference would panic on it) + dijkstra
prm = O.PRM(smap, low, up, seed=0); prm.init([0.0, 0.0]); prm.grow_graph(0.15, 3.0, 400)
xy, nvid, rp, col, ev = prm.graph.export(0)
out.update(prm_xy=xy, prm_row_ptr=rp, prm_col=col, prm_dijkstra=prm.graph.dijkstra([0]))
# ---- belief planning (pto.rs:185-283) on the two-door planning map
pocc, pzones = util.planning_door_map(200)
pmap = O.GridMap(pocc, pzones, low, up, O.DOOR, 0.5)
pto = O.PTO(pmap, low, up, seed=0)
assert pto.grow_graph((-0.8, -0.8), O.SquareGoal([((0.8, 0.8), [1, 1, 1, 1])], 0.05), 0.05, 5.0, 1500, 100000) == 0
b0 = [0.1, 0.1, 0.1, 0.7]
pto.build_belief_graph(b0); cost = pto.compute_expected_costs_to_goals(); pol = pto.extract_policy()
xy, nvid, rp, col, ev = pto.graph.export(0); fin_ids, fin_bits = pto.reach.finals()
out.update(bel_xy=xy, bel_nvid=nvid, bel_row_ptr=rp, bel_col=col, bel_ev=ev, bel_fin_ids=np.asarray(fin_ids), bel_fin_bits=np.asarray(fin_bits),
           bel_b0=np.asarray(b0), bel_cost=cost, bel_policy=np.asarray(pol.original), bel_policy_cost=np.float64(pol.expected_costs),
           bel_qmdp=pto.plan_qmdp())
# ---- policy refinement
path = np.array([[-0.8 + 0.05 * k, -0.8 + 0.03 * ((k * 7) % 5)] for k in range(25)])
st, commits = pmap.refiner_partial_shortcut(path, [1, 1, 1], 300)
out.update(ref_path=path, ref_states=st, ref_commits=np.int64(commits))
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "standin_v1.npz"), **out)
print("wrote tests/golden/standin_v1.npz:", {k: np.asarray(v).shape for k, v in out.items()})

# ---- v2: multi-modal PRM (map_shelves_tamp_prm.rs): the RNG-decided schedule + expected costs + policy on the 3-shelf map
tamp = O.TampPRM(smap, low, up)
tpol = tamp.plan((0.0, -0.9), [1.0 / 3] * 3, 0.15, 3.0, 800)
sch = tamp.schedule()
out2 = {"mm_" + k: v for k, v in sch.items()}
out2.update(mm_policy=np.asarray(tpol.original), mm_policy_parent=np.asarray(tpol.parent), mm_policy_cost=np.float64(tpol.expected_costs))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "standin_v2.npz"), **out2)
print("wrote tests/golden/standin_v2.npz:", {k: np.asarray(v).shape for k, v in out2.items()})

# ---- v3: both refinement strategies on the belief-space policy of the planning map (pto_policy_refiner.rs:85-133)
sc = pto.refine_policy_shortcut(300)
rp3 = pto.refine_policy_reparent(0.3)
out3 = dict(sc_xy=sc.xy, sc_parent=sc.parent, sc_belief=sc.belief_id, sc_original=sc.original, sc_cost=np.float64(sc.expected_costs),
            rp_xy=rp3.xy, rp_parent=rp3.parent, rp_belief=rp3.belief_id, rp_original=rp3.original, rp_leafs=rp3.leafs,
            rp_cost=np.float64(rp3.expected_costs))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "standin_v3.npz"), **out3)
print("wrote tests/golden/standin_v3.npz:", {k: np.asarray(v).shape for k, v in out3.items()})
