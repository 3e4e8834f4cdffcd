"""PTOGraph <-> JSON in the reference's on-disk format (src/pto_graph.rs:22-118: serde_json of SerializablePTOGraph), and
<-> the CSR arrays the C ABI takes.  A roadmap saved by the Rust planner (`pto_graph::save`) can be loaded here and handed to
porrt_sssp_worlds / porrt_belief_vi; a graph built on the GPU can be written back for `pto_graph::load`.

Format: {"nodes": [{"state": [x, y], "validity_id": n, "parents": [{"id": i, "validity_id": v}, ...], "children": [...]}, ...],
         "validities": [[bool, ...], ...]}          (validities[v][w]: validity v holds in world w, pto_graph.rs:95-101)
f64 values are written with Python's shortest round-trip repr (serde_json/ryu also writes shortest round-trip digits; the two may
differ in exponent spelling, which every JSON reader accepts)."""
import json

import numpy as np


class PTOGraphArrays:
    """xy[V,2], node_vid[V]; children / parents as CSR in stored order: row_ptr[V+1], col[E], edge_vid[E]; validities[n_val, n_worlds]"""

    def __init__(self, xy, node_vid, row_ptr, col, edge_vid, p_row_ptr, p_col, p_edge_vid, validities):
        self.xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        self.node_vid = np.ascontiguousarray(node_vid, np.int32)
        self.row_ptr, self.col, self.edge_vid = (np.ascontiguousarray(row_ptr, np.int64), np.ascontiguousarray(col, np.int32),
                                                 np.ascontiguousarray(edge_vid, np.int32))
        self.p_row_ptr, self.p_col, self.p_edge_vid = (np.ascontiguousarray(p_row_ptr, np.int64), np.ascontiguousarray(p_col, np.int32),
                                                       np.ascontiguousarray(p_edge_vid, np.int32))
        self.validities = np.ascontiguousarray(validities, np.uint8)

    @property
    def n_nodes(self):
        return len(self.xy)


def _csr(lists, key):
    rp = np.zeros(len(lists) + 1, np.int64)
    for k, l in enumerate(lists):
        rp[k + 1] = rp[k] + len(l)
    flat = [e[key] for l in lists for e in l]
    return rp, np.asarray(flat, np.int32)


def load_pto_graph(path):
    """pto_graph::load (pto_graph.rs:110-118) -> PTOGraphArrays; edge order is the stored (insertion) order"""
    with open(path) as f:
        g = json.load(f)
    nodes = g["nodes"]
    for n in nodes:
        if len(n["state"]) != 2:
            raise ValueError("state is not [f64; 2] (to_pto_node's try_into().unwrap() would panic)")
    xy = np.array([n["state"] for n in nodes], np.float64).reshape(-1, 2)
    nvid = np.array([n["validity_id"] for n in nodes], np.int32)
    rp, col = _csr([n["children"] for n in nodes], "id")
    _, ev = _csr([n["children"] for n in nodes], "validity_id")
    prp, pcol = _csr([n["parents"] for n in nodes], "id")
    _, pev = _csr([n["parents"] for n in nodes], "validity_id")
    if len(col) and (col.min() < 0 or col.max() >= len(nodes)) or len(pcol) and (pcol.min() < 0 or pcol.max() >= len(nodes)):
        raise ValueError("edge id out of range")
    val = np.array([[1 if b else 0 for b in v] for v in g["validities"]], np.uint8)
    return PTOGraphArrays(xy, nvid, rp, col, ev, prp, pcol, pev, val)


def transpose_csr(row_ptr, col, edge_vid, n):
    """parents from children for graphs whose add_edge calls came in (a->b, then b->a) pairs or any order: parents(v) lists the
    sources u of edges u->v in the order the edges u->v appear when rows are walked in node order (== PRM insertion order,
    where parents(k) == children(k) as sequences, prm.rs:99-106)"""
    src = np.repeat(np.arange(n, dtype=np.int32), np.diff(row_ptr))
    order = np.argsort(col, kind="stable")
    prp = np.zeros(n + 1, np.int64)
    np.add.at(prp, col.astype(np.int64) + 1, 1)
    prp = np.cumsum(prp)
    return prp, src[order], np.asarray(edge_vid)[order]


def save_pto_graph(path, g, indent=2):
    """pto_graph::save (pto_graph.rs:105-108); `g` is a PTOGraphArrays"""
    def edges(rp, col, ev, k):
        return [{"id": int(col[e]), "validity_id": int(ev[e])} for e in range(rp[k], rp[k + 1])]
    nodes = [{"state": [float(g.xy[k, 0]), float(g.xy[k, 1])], "validity_id": int(g.node_vid[k]),
              "parents": edges(g.p_row_ptr, g.p_col, g.p_edge_vid, k), "children": edges(g.row_ptr, g.col, g.edge_vid, k)}
             for k in range(g.n_nodes)]
    doc = {"nodes": nodes, "validities": [[bool(b) for b in v] for v in g.validities]}
    with open(path, "w") as f:
        json.dump(doc, f, indent=indent)
