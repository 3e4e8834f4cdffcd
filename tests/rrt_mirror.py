"""RRT* (reference src/rrt.rs:102-181, grow_tree + sample) restated over a per-query backend -- TEST INFRASTRUCTURE.

BASELINE config 1 is the sequential RRT planner: every iteration depends on the tree built so far, so it stays on the host and sees
the hot path only through per-query calls (1-NN, radius search in kd order, one state check, the candidate edges of one sample).
The same planner code runs over the oracle backend (reference restatement) and over the product backend (batches of one / of a
few through the C ABI); equal trees mean the per-query wrappers are drop-in for this caller."""
import math

import numpy as np

from oracle import pyoracle as O


def norm1(a, b):      # common.rs:192-201
    return abs(b[0] - a[0]) + abs(b[1] - a[1])


def norm2(a, b):      # common.rs:203-213
    dx, dy = b[0] - a[0], b[1] - a[1]
    return math.sqrt(dx * dx + dy * dy)


def steer(frm, to, max_step):      # common.rs:215-225 (the step is measured with norm1)
    step = norm1(frm, to)
    if step > max_step:
        lam = max_step / step
        return [frm[0] + (to[0] - frm[0]) * lam, frm[1] + (to[1] - frm[1]) * lam]
    return list(to)


def heuristic_radius(n_nodes, max_step, search_radius, dim=2):      # common.rs:357-369
    n = float(n_nodes)
    s = search_radius * math.pow(math.log(n) / n, 1.0 / dim)
    return s if s < max_step else max_step


class OracleBackend:
    def __init__(self, omap, start):
        self.map = omap
        self.tree = O.KdTree(start, 0)

    def add(self, state, node_id):
        self.tree.add(state, node_id)

    def nearest(self, q):
        return int(self.tree.nearest_neighbor(q))

    def neighbors(self, q, radius):          # kd pre-order, nearest_neighbor.rs:94-126
        return [int(i) for i in self.tree.nearest_neighbors(q, radius)]

    def state_valid(self, q):                # rrt.rs:274-286 wrapper: is_state_valid == Free
        return int(self.map.state_validity([q])[0]) >= 0

    def edges_valid(self, froms, to):
        if not len(froms):
            return []
        return [int(v) >= 0 for v in self.map.edge_validity(froms, [to] * len(froms))]


class ProductBackend:
    """the same five questions through libporrt_b200 (po_rrt_b200.api); a new node is appended to the device-resident vertex set
    (KdTree.add -> porrt_vertices_append: 16 bytes cross the bus, not the whole set)"""

    def __init__(self, pmap, start):
        import po_rrt_b200 as P
        self.P, self.map, self.ctx = P, pmap, pmap.ctx
        self.states = [list(start)]
        self.tree = P.KdTree(self.ctx, np.asarray(self.states, np.float64))
        self.rank = None

    def _sync(self):
        pass

    def add(self, state, node_id):
        assert node_id == len(self.states)
        self.states.append(list(state))
        self.tree.add([state])
        self.rank = None

    def nearest(self, q):
        self._sync()
        nid, _, ties = self.tree.nearest_neighbor([q])
        if ties[0] != 1:
            # several vertices at exactly the winning distance.  Exact duplicates of one point (the goal-biased samples repeat the
            # goal, rrt.rs:176-181) chain to the right in insertion order, so the reference's strict `d < dmin` keeps the lowest id
            # -- which is what the library returns.  Distinct equidistant points would need the kd visit order: not expected here.
            st = np.asarray(self.states)
            dx, dy = q[0] - st[:, 0], q[1] - st[:, 1]
            d2 = dx * dx + dy * dy
            tied = np.nonzero(d2 == d2[nid[0]])[0]
            assert len(tied) == ties[0] and (st[tied] == st[tied[0]]).all() and tied[0] == nid[0], "tie among distinct points"
        return int(nid[0])

    def neighbors(self, q, radius):
        self._sync()
        _, ids = self.tree.nearest_neighbors([q], radius)
        if len(ids) > 1:
            if self.rank is None:
                self.rank = self.tree.preorder_rank()
            ids = ids[np.argsort(self.rank[ids], kind="stable")]      # the reference's visit order
        return [int(i) for i in ids]

    def state_valid(self, q):
        return int(self.map.state_validity([q])[0]) >= 0

    def edges_valid(self, froms, to):
        if not len(froms):
            return []
        return [int(v) >= 0 for v in self.map.transition_validator(froms, [to] * len(froms))]


def grow_tree(backend, samples, start, goal, max_step, search_radius, n_iter_min, n_iter_max):
    """rrt.rs:102-174.  `samples`: the ContinuousSampler stream (consumed when iteration % 100 != 0, :176-181).
    Returns (states, parent ids, dist_from_root, final node ids)."""
    states, parent, dist = [list(start)], [-1], [0.0]
    finals = []
    k = 0
    i = 0
    while i < n_iter_min or (not finals and i < n_iter_max):
        i += 1
        if i % 100 == 0:
            new_state = list(goal.goal_example(0))
        else:
            new_state = list(samples[k])
            k += 1
        kd_from = backend.nearest(new_state)
        new_state = steer(states[kd_from], new_state, max_step)
        if not backend.state_valid(new_state):
            continue
        radius = heuristic_radius(len(states), max_step, search_radius)
        cand = backend.neighbors(new_state, radius)
        ok = backend.edges_valid([states[c] for c in cand], new_state)
        neigh = [c for c, v in zip(cand, ok) if v]
        if not neigh:
            neigh = [kd_from]
        from_parent = [norm2(states[c], new_state) for c in neigh]
        best = 0
        for j in range(1, len(neigh)):      # Iterator::min_by keeps the first of equal minima
            if dist[neigh[j]] + from_parent[j] < dist[neigh[best]] + from_parent[best]:
                best = j
        new_id = len(states)
        states.append(new_state)
        parent.append(neigh[best])
        dist.append(dist[neigh[best]] + from_parent[best])
        for c in neigh:
            if c == neigh[best]:
                continue
            d_new = norm2(new_state, states[c])
            if dist[new_id] + d_new < dist[c]:
                parent[c] = new_id
                dist[c] = dist[new_id] + d_new
        backend.add(new_state, new_id)
        if goal.goal(new_state) is not None:
            finals.append(new_id)
    return np.asarray(states), np.asarray(parent), np.asarray(dist), finals


class ProductPTOBackend:
    """the five per-query answers PTO::grow_graph needs (oracle/pyoracle.py: PTO.set_hooks), from libporrt_b200"""

    def __init__(self, pmap):
        import po_rrt_b200 as P
        self.P, self.map, self.ctx = P, pmap, pmap.ctx
        self.states, self.tree, self.rank = [], P.KdTree(self.ctx), None

    def _sync(self):
        pass

    def add_vertex(self, q, node_id):
        assert node_id == len(self.states)
        self.states.append(list(q))
        self.tree.add([q])             # porrt_vertices_append: only the new state is uploaded
        self.rank = None

    def nearest_filtered(self, q, world, reach_words):      # pto.rs:74-77: 1-NN among nodes reachable in `world`
        self._sync()
        nid, _, ties = self.tree.nearest_neighbor([q], reach_mask=reach_words[:, 0].copy(), world=[world])
        if nid[0] < 0:
            return 0                                         # nothing passes the filter: the root (nearest_neighbor.rs:89)
        if ties[0] != 1:      # exact duplicates (repeated goal samples): the lowest id wins in the reference too, see ProductBackend.nearest
            st = np.asarray(self.states)
            ok = np.nonzero((reach_words[:, 0] >> np.uint64(world)) & np.uint64(1))[0]
            dx, dy = q[0] - st[ok, 0], q[1] - st[ok, 1]
            d2 = dx * dx + dy * dy
            tied = ok[d2 == d2.min()]
            assert (st[tied] == st[tied[0]]).all() and tied[0] == nid[0], "tie among distinct points"
        return int(nid[0])

    def radius(self, q, r):
        self._sync()
        _, ids = self.tree.nearest_neighbors([q], r)
        if len(ids) > 1:
            if self.rank is None:
                self.rank = self.tree.preorder_rank()
            ids = ids[np.argsort(self.rank[ids], kind="stable")]
        return [int(i) for i in ids]

    def state_validity(self, q):
        return int(self.map.state_validity([q])[0])

    def edges(self, frm, to):
        return [int(v) for v in self.map.transition_validator(frm, to)]
