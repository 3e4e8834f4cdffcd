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


def test_pto_growth_hooks_plumbing():
    """PTO.set_hooks (the per-query hook points of the oracle's PTO::grow_graph restatement): with the hooks answered by the
    oracle's own primitives the growth must reproduce itself -- the GPU test plugs the product in at the same five points"""
    import porrt_testutil as util
    occ, zones = util.planning_door_map(200)
    omap = O.GridMap(occ, zones, [-1.0, -1.0], [1.0, 1.0], O.DOOR, 0.3)
    goal = O.SquareGoal([((0.8, 0.8), [1, 1, 1, 1])], 0.05)
    ref = O.PTO(omap, [-1.0, -1.0], [1.0, 1.0], seed=0)
    assert ref.grow_graph((-0.8, -0.8), goal, 0.05, 5.0, 400, 100000) == 0

    class OwnAnswers:
        def __init__(self):
            self.tree, self.states = None, []

        def add_vertex(self, q, node_id):
            if self.tree is None:
                self.tree = O.KdTree(q, 0)
            else:
                self.tree.add(q, node_id)
            self.states.append(q)

        def nearest_filtered(self, q, world, reach_words):
            ok = [i for i in range(len(self.states)) if (int(reach_words[i, 0]) >> world) & 1]
            if not ok:
                return 0
            st = np.asarray(self.states)
            dx, dy = st[ok, 0] - q[0], st[ok, 1] - q[1]
            return ok[int(np.argmin(np.sqrt(dx * dx + dy * dy)))]

        def radius(self, q, r):
            return [int(i) for i in self.tree.nearest_neighbors(q, r)]

        def state_validity(self, q):
            return int(omap.state_validity([q])[0])

        def edges(self, frm, to):
            return [int(v) for v in omap.edge_validity(frm, to)]

    pto = O.PTO(omap, [-1.0, -1.0], [1.0, 1.0], seed=0)
    pto.set_hooks(OwnAnswers())
    assert pto.grow_graph((-0.8, -0.8), goal, 0.05, 5.0, 400, 100000) == 0
    assert pto.n_it() == ref.n_it()
    for a, b in zip(pto.graph.export(0), ref.graph.export(0)):
        np.testing.assert_array_equal(a, b)


def test_plan_empty_space():  # rrt.rs:254-268
    """the reference's own map-free RRT test: default RTTFuncs (everything valid), SquareGoal([0.9, 0.9], 0.05),
    plan([0, 0], 0.1, 1.0, 1000, 10000) finds a path of more than two states -- here over the oracle's kd-tree with an
    always-valid world, with get_best_solution (:183-193) restated on the mirror's tree"""
    class EmptySpace(R.OracleBackend):
        def __init__(self, start):
            self.tree = O.KdTree(start, 0)

        def state_valid(self, q):
            return True

        def edges_valid(self, froms, to):
            return [True] * len(froms)

    start = [0.0, 0.0]
    goal = O.SquareGoal([((0.9, 0.9), [1])], 0.05)
    samples = O.Pcg64(0).sample_states([-1.0, -1.0], [1.0, 1.0], 10000)
    st, par, dist, fin = R.grow_tree(EmptySpace(start), samples, start, goal, 0.1, 1.0, 1000, 10000)
    assert fin, "No path found!"

    def path_to(k):                        # RRTTree::get_path_to (:48-60)
        out = []
        while k >= 0:
            out.append(st[k])
            k = par[k]
        return out[::-1]
    best = min((sum(R.norm2(a, b) for a, b in zip(p[:-1], p[1:])), len(p)) for p in (path_to(f) for f in fin))
    assert best[1] > 2
    assert best[0] >= R.norm2(start, [0.9, 0.9]) - 0.05 * 2 ** 0.5 - 1e-9      # no path beats the straight line to the goal region
