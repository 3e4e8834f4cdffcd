"""The host-side rows of the C ABI (csrc/host_side.cu: steer, heuristic_radius, samplers, SquareGoal, Reachability, the QMDP walk)
against the reference's own map-free golden tests and against the oracle.  No GPU needed: these entry points do no device work."""
import math

import numpy as np

from oracle import pyoracle as O
import po_rrt_b200 as P
from po_rrt_b200 import synth
import porrt_testutil as util


# ------------------------------------------------------------------ pto_reachability.rs:109-230
def test_reachability():                                # :109-135
    r = P.Reachability()
    r.set_root([1, 1]); r.add_node([1, 0]); r.add_node([1, 0]); r.add_node([0, 1])
    r.add_edge(0, 1, [1, 0]); r.add_edge(1, 2, [1, 0]); r.add_edge(1, 3, [0, 1])
    assert [r.reachability(i) for i in range(4)] == [[1, 1], [1, 0], [1, 0], [0, 0]]


def test_reachability_diamond_shape():                  # :137-163
    r = P.Reachability()
    r.set_root([1, 1]); r.add_node([1, 0]); r.add_node([0, 1]); r.add_node([1, 1])
    r.add_edge(0, 1, [1, 0]); r.add_edge(0, 2, [0, 1]); r.add_edge(1, 3, [1, 1]); r.add_edge(2, 3, [1, 1])
    assert [r.reachability(i) for i in range(4)] == [[1, 1], [1, 0], [0, 1], [1, 1]]


def test_final_nodes_completness():                     # :165-196
    r = P.Reachability()
    r.set_root([1, 1]); r.add_node([1, 1]); r.add_node([1, 0]); r.add_node([0, 1])
    r.add_edge(0, 1, [1, 1]); r.add_edge(1, 2, [1, 0]); r.add_edge(1, 3, [0, 1])
    assert not r.is_final_set_complete()
    r.add_final_node(2, [1, 1])
    assert not r.is_final_set_complete()
    r.add_final_node(3, [1, 1])
    assert r.is_final_set_complete()
    assert r.get_final_nodes_for_world(0) == [2] and r.get_final_nodes_for_world(1) == [3]


def test_final_nodes_completness_2_goals_2_worlds():    # :198-230
    r = P.Reachability()
    r.set_root([1, 1]); r.add_node([1, 1]); r.add_node([1, 1]); r.add_node([1, 1])
    r.add_edge(0, 1, [1, 1]); r.add_edge(1, 2, [1, 1]); r.add_edge(1, 3, [1, 1])
    r.add_final_node(2, [1, 0])
    assert not r.is_final_set_complete()
    r.add_final_node(3, [0, 1])
    assert r.is_final_set_complete()
    assert r.get_final_nodes_for_world(0) == [2] and r.get_final_nodes_for_world(1) == [3]


def test_reachability_random_vs_oracle_wide_masks():
    """130 worlds (3 mask words): random graph insertions, the product's table == the oracle's after every phase"""
    rng = np.random.default_rng(5)
    W, n = 130, 300
    o, p = O.Reachability(n_worlds=W), P.Reachability()
    root = list(rng.integers(0, 2, W))
    o.set_root(root); p.set_root(root)
    for _ in range(n - 1):
        v = list(rng.integers(0, 2, W))
        o.add_node(v); p.add_node(v)
    for _ in range(4000):
        a, b = rng.integers(0, n, 2)
        m = list((rng.random(W) < 0.8).astype(int))
        o.add_edge(int(a), int(b), m); p.add_edge(int(a), int(b), m)
        if rng.random() < 0.01:
            f = list((rng.random(W) < 0.3).astype(int))
            k = int(rng.integers(0, n))
            o.add_final_node(k, f); p.add_final_node(k, f)
            assert o.is_final_set_complete() == p.is_final_set_complete()
    want = o.all(n)
    got = p.masks_words()
    np.testing.assert_array_equal(got, P.words_from_bits(want))
    for w in (0, 63, 64, 129):
        assert p.get_final_nodes_for_world(w) == [int(x) for x in o.get_final_nodes_for_world(w)]
    ids, masks = p.finals()
    oids, obits = o.finals()
    np.testing.assert_array_equal(ids, oids)
    np.testing.assert_array_equal(masks, P.words_from_bits(obits))


# ------------------------------------------------------------------ common.rs:401-523
def test_goal():                                        # :402-411
    g = P.SquareGoal([([0.1, 0.1], [1, 0]), ([0.9, 0.9], [0, 1])], 0.1)
    assert g.goal([0.11, 0.11]) == [1, 0]
    assert g.goal([0.5, 0.5]) is None
    assert g.goal([0.91, 0.91]) == [0, 1]
    assert list(g.goal_example(0)) == [0.1, 0.1] and list(g.goal_example(1)) == [0.9, 0.9]


def test_goal_is_an_l1_diamond_and_first_match_wins():
    og = O.SquareGoal([([0.0, 0.0], [1, 0, 0]), ([0.05, 0.0], [0, 1, 0])], 0.1)
    pg = P.SquareGoal([([0.0, 0.0], [1, 0, 0]), ([0.05, 0.0], [0, 1, 0])], 0.1)
    pts = np.random.default_rng(1).uniform(-0.2, 0.2, (4000, 2))
    idx = pg.goal_index(pts)
    for k, s in enumerate(pts):
        want = og.goal(s)
        assert (None if idx[k] < 0 else [int(b) for b in pg.bits[idx[k]]]) == (None if want is None else [int(b) for b in want])
    assert (idx == 0).any() and (idx == 1).any() and (idx == -1).any()
    assert list(pg.goal_example(2)) == [0.0, 0.0]        # a world without goal keeps the zero state
    try:
        P.SquareGoal([([0.0, 0.0], [1, 0]), ([0.5, 0.0], [1, 1])], 0.1)
        assert False, "overlapping validities must be refused (assert, common.rs:320)"
    except P.PorrtError as e:
        assert e.code == 5


def test_steer_and_radius():
    assert list(P.steer([[0.0, 0.0]], [[1.0, 1.0]], 0.5)[0]) == [0.25, 0.25]      # norm1 step = 2 -> lambda = 0.25
    rng = np.random.default_rng(2)
    f, t = rng.uniform(-1, 1, (5000, 2)), rng.uniform(-1, 1, (5000, 2))
    got = P.steer(f, t, 0.1)
    for k in range(0, 5000, 7):
        want = O.f64a(t[k]).copy()
        O.lib().orc_steer(O.P(O.f64a(f[k])), O.P(want), 0.1)
        assert list(got[k]) == list(want)
    for n in (1, 2, 10, 100, 1000, 10000, 1000000):
        assert P.heuristic_radius(n, 0.1, 2.0, 2) == O.lib().orc_heuristic_radius(n, 0.1, 2.0, 2)
        assert P.heuristic_radius(n, 0.05, 5.0, 2) == O.lib().orc_heuristic_radius(n, 0.05, 5.0, 2)
    assert P.heuristic_radius(1, 0.1, 2.0, 2) == 0.0
    # steer<N> for the other state dimensions (common.rs:215-225 is generic): plain Python floats in the reference's operation order
    for dim in (3, 7, 9):
        f, t = rng.uniform(-1, 1, (300, dim)), rng.uniform(-1, 1, (300, dim))
        got = P.steer(f, t, 0.3, dim=dim)
        for k in range(300):
            step = 0.0
            for a, b in zip(f[k].tolist(), t[k].tolist()):
                step += abs(b - a)
            want = [a + (b - a) * (0.3 / step) for a, b in zip(f[k].tolist(), t[k].tolist())] if step > 0.3 else t[k].tolist()
            assert got[k].tolist() == want


# ------------------------------------------------------------------ sample_space.rs
def test_samplers_follow_the_reference_streams():
    s = P.Sampler(0)
    got = s.sample_states([-1.0, -1.0], [1.0, 1.0], 3000)
    np.testing.assert_array_equal(got, O.Pcg64(0).sample_states([-1.0, -1.0], [1.0, 1.0], 3000))
    assert (got >= -1.0).all() and (got < 1.0).all()                              # sample_space.rs:62-90 (range checks)
    d, od = P.Sampler(0), O.Pcg64(0)
    for n in (2, 3, 12, 64, 1000):
        assert list(d.sample_discrete(n, 200)) == [od.gen_range_usize(n) for _ in range(200)]
    assert set(P.Sampler(0).sample_discrete(3, 500)) == {0, 1, 2}                  # :92-115
    # interleaved use of ONE stream, different seeds
    a, oa = P.Sampler(7), O.Pcg64(7)
    x = a.sample_states([0.0, -3.0], [1.0, 5.0], 10)
    np.testing.assert_array_equal(x, oa.sample_states([0.0, -3.0], [1.0, 5.0], 10))
    y = a.sample_states([0.0, -3.0, 2.0], [1.0, 5.0, 2.5], 4)                     # N = 3: one draw per dimension, in order
    want = [oa.gen_range_f64(lo, hi) for _ in range(4) for lo, hi in ((0.0, 1.0), (-3.0, 5.0), (2.0, 2.5))]
    assert list(y.reshape(-1)) == want
    assert int(a.sample_discrete(10, 1)[0]) == oa.gen_range_usize(10)


# ------------------------------------------------------------------ qmdp_policy_extractor.rs:38-123
def _grow(omap, start, goals, max_step, radius, n_min):
    pto = O.PTO(omap, util.LOW, util.UP, seed=0)
    assert pto.grow_graph(start, O.SquareGoal(goals, 0.05), max_step, radius, n_min, 100000) == 0
    return pto


def _paths_as_states(paths, xy):
    return [xy[p] for p in paths]


def test_react_qmdp_config2_literals():
    """BASELINE config 2 (qmdp_policy_extractor.rs:176-199): grow_graph((-0.8,-0.8), 0.05, 5.0, 2000, 100000) on a 2-shelf map,
    plan_qmdp, react_qmdp((-0.8,-0.8), [0.5, 0.5], 0.2): common path + per-world paths identical to the reference algorithm"""
    occ, zones = synth.shelf_map(200, n_zones=2)
    omap = O.GridMap(occ, zones, util.LOW, util.UP, O.SHELF, 0.5)
    zp = omap.zone_positions()
    goals = [((float(zp[0][0]) - 0.06, float(zp[0][1])), [1, 0]), ((float(zp[1][0]) - 0.06, float(zp[1][1])), [0, 1])]
    pto = _grow(omap, (-0.8, -0.8), goals, 0.05, 5.0, 2000)
    costs = pto.plan_qmdp()
    xy, nvid, rp, col, ev = pto.graph.export(0)
    for start, belief, horizon in (((-0.8, -0.8), [0.5, 0.5], 0.2), ((-0.8, -0.8), [0.9, 0.1], 1.0), ((0.3, 0.2), [0.0, 1.0], 0.0),
                                   ((0.0, 0.0), [0.5, 0.5], 50.0)):
        want = pto.react_qmdp(start, belief, horizon)
        start_node = pto.kdtree.nearest_neighbor(start)
        got, n_common = P.react_qmdp(None, rp, col, xy, costs, start_node, belief, horizon)
        assert len(got) == len(want) == 2
        for w in range(2):
            np.testing.assert_array_equal(xy[got[w]], want[w])
        assert n_common >= (1 if horizon > 0 else 0)
        assert all(list(got[w][:n_common]) == list(got[0][:n_common]) for w in range(2))
    assert len(got[0]) > n_common or len(got[1]) > n_common or True


def test_react_qmdp_door_map_and_errors():
    occ, zones = util.planning_door_map(200)
    omap = O.GridMap(occ, zones, util.LOW, util.UP, O.DOOR, 0.3)
    pto = _grow(omap, (-0.8, -0.8), [((0.8, 0.8), [1, 1, 1, 1])], 0.05, 5.0, 3000)
    costs = pto.plan_qmdp()
    xy, nvid, rp, col, ev = pto.graph.export(0)
    rng = np.random.default_rng(9)
    n_ok = 0
    for _ in range(12):
        start = tuple(rng.uniform(-0.9, 0.9, 2))
        b = rng.random(4); b /= b.sum()
        horizon = float(rng.uniform(0.0, 1.5))
        start_node = pto.kdtree.nearest_neighbor(start)
        try:
            want = pto.react_qmdp(start, list(b), horizon)
        except RuntimeError:
            want = None                                   # the reference would not terminate (unreachable in some world)
        try:
            got, _ = P.react_qmdp(None, rp, col, xy, costs, start_node, list(b), horizon)
        except P.PorrtError as e:
            assert want is None and e.code == 5
            continue
        assert want is not None
        n_ok += 1
        for w in range(4):
            np.testing.assert_array_equal(xy[got[w]], want[w])
    assert n_ok >= 1
    try:
        P.react_qmdp(None, rp, col, xy, costs, 0, [0.5, 0.5], 0.2)      # belief of the wrong length: Err(..).unwrap() panics
        assert False
    except P.PorrtError as e:
        assert e.code == 5
