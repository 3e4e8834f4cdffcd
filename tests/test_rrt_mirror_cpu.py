"""The RRT* restatement used for BASELINE config 1 (tests/rrt_mirror.py, rrt.rs:102-181), run over the oracle backend only:
structural invariants of the grown tree, so that the mirror itself is exercised without a GPU (the GPU test compares the same
planner over the product's per-query wrappers)."""
import numpy as np

from oracle import pyoracle as O
from po_rrt_b200 import synth
import rrt_mirror as R


def test_rrt_star_tree_invariants():
    occ, zones = synth.shelf_map(200, n_zones=2)
    omap = O.GridMap(occ, zones, [-1.0, -1.0], [1.0, 1.0], O.SHELF, 0.5)
    start = [0.0, -0.8]
    goal = O.SquareGoal([((0.6, 0.36), [1])], 0.05)
    samples = O.Pcg64(0).sample_states([-1.0, -1.0], [1.0, 1.0], 1500)
    st, par, dist, fin = R.grow_tree(R.OracleBackend(omap, start), samples, start, goal, 0.1, 2.0, 500, 1500)
    assert par[0] == -1 and dist[0] == 0.0 and len(st) > 300 and len(fin) > 0
    assert (omap.state_validity(st) >= 0).all()                                   # only valid states enter the tree (rrt.rs:117)
    for k in range(1, len(st)):
        p = par[k]
        assert 0 <= p < len(st) and p != k
        # dist_from_root is set when a node is (re)parented and not refreshed when an ancestor is rewired later (rrt.rs:29-46):
        # it can only overestimate the current path through the parent
        assert dist[k] >= dist[p] + R.norm2(st[p], st[k]) - 1e-12
        assert R.norm1(st[p], st[k]) <= 2.0                                       # sanity: edges stay inside the domain
    assert (omap.edge_validity(st[par[1:]], st[1:]) >= 0).mean() > 0.99           # parents were validated when chosen
    for f in fin:
        assert goal.goal(st[f]) is not None                                       # norm1 diamond around the goal (common.rs:336-350)
    # determinism: same stream, same tree
    st2, par2, dist2, fin2 = R.grow_tree(R.OracleBackend(omap, start), samples, start, goal, 0.1, 2.0, 500, 1500)
    assert np.array_equal(st, st2) and np.array_equal(par, par2) and np.array_equal(dist, dist2) and fin == fin2
