"""PTOGraph <-> JSON in the reference's on-disk format (src/pto_graph.rs:22-118: serde_json of SerializablePTOGraph), and
<-> the CSR arrays the C ABI takes.  A roadmap saved by the Rust planner (`pto_graph::save`) can be loaded here and handed to
porrt_sssp_worlds / porrt_belief_vi; a graph built on the GPU can be written back for `pto_graph::load`.

Format: {"nodes": [{"state": [x, y], "validity_id": n, "parents": [{"id": i, "validity_id": v}, ...], "children": [...]}, ...],
         "validities": [[bool, ...], ...]}          (validities[v][w]: validity v holds in world w, pto_graph.rs:95-101)
f64 values are written with Python's shortest round-trip repr (serde_json/ryu also writes shortest round-trip digits; the two may
differ in exponent spelling, which every JSON reader accepts)."""
import ctypes as C

import numpy as np

from . import _lib


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


def load_pto_graph(path):
    """pto_graph::load (pto_graph.rs:110-118) -> PTOGraphArrays; edge order is the stored (insertion) order.  The reader is the
    library's (csrc/formats.cu: porrt_graph_load_json)."""
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.porrt_graph_load_json(None, str(path).encode(), C.byref(h))
    if rc != 0:
        raise ValueError("porrt_graph_load_json failed [status %d] (malformed file: the reference's serde / try_into unwrap would panic)" % rc)
    try:
        nn, nc, npar = C.c_int64(), C.c_int64(), C.c_int64()
        nv, nw = C.c_int32(), C.c_int32()
        lib.porrt_graph_info(h, C.byref(nn), C.byref(nc), C.byref(npar), C.byref(nv), C.byref(nw))
        V = nn.value
        xy, nvid = np.empty((V, 2)), np.empty(V, np.int32)
        rp, col, ev = np.empty(V + 1, np.int64), np.empty(nc.value, np.int32), np.empty(nc.value, np.int32)
        prp, pcol, pev = np.empty(V + 1, np.int64), np.empty(npar.value, np.int32), np.empty(npar.value, np.int32)
        val = np.empty((nv.value, nw.value), np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.porrt_graph_arrays(h, p(xy), p(nvid), p(rp), p(col), p(ev), p(prp), p(pcol), p(pev), p(val))
    finally:
        lib.porrt_graph_destroy(h)
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
    """pto_graph::save (pto_graph.rs:105-108); `g` is a PTOGraphArrays.  The writer is the library's (porrt_graph_save_json):
    serde_json::to_writer_pretty layout, floats as shortest round-trip digits in ryu's spelling."""
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    val = np.ascontiguousarray(g.validities, np.uint8).reshape(len(g.validities), -1)
    rc = _lib.load().porrt_graph_save_json(None, str(path).encode(), g.n_nodes, p(g.xy), p(g.node_vid), p(g.row_ptr), p(g.col), p(g.edge_vid),
                                           p(g.p_row_ptr), p(g.p_col), p(g.p_edge_vid), p(val), val.shape[0], val.shape[1])
    if rc != 0:
        raise OSError("porrt_graph_save_json failed [status %d]" % rc)
