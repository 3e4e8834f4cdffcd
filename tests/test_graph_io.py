"""PTOGraph JSON (pto_graph.rs:22-118, its test :566-572 only saves and loads): round trip through the reference's on-disk layout,
field names and nesting as serde writes them, and CSR conversion that keeps the stored edge order."""
import json

import numpy as np
import pytest

from oracle import pyoracle as O
from po_rrt_b200 import graph_io, synth


def _oracle_prm_graph():
    occ, zones = synth.shelf_map(120, n_rects=8, n_zones=3, seed=9)
    smap = O.GridMap(occ, zones, [-1.0, -1.0], [1.0, 1.0], O.SHELF, 0.7)
    prm = O.PRM(smap, [-1.0, -1.0], [1.0, 1.0], seed=0)
    prm.init([0.0, 0.0])
    prm.grow_graph(0.15, 3.0, 300)
    xy, nvid, rp, col, ev = prm.graph.export(0)
    _, _, prp, pcol, pev = prm.graph.export(1)
    return graph_io.PTOGraphArrays(xy, nvid, rp, col, ev, prp, pcol, pev, smap.world_validities())


def test_round_trip_and_layout(tmp_path):
    g = _oracle_prm_graph()
    p = tmp_path / "graph.json"
    graph_io.save_pto_graph(str(p), g)
    doc = json.loads(p.read_text())
    assert set(doc) == {"nodes", "validities"}
    assert set(doc["nodes"][0]) == {"state", "validity_id", "parents", "children"}     # SerializablePTONode, pto_graph.rs:22-28
    assert set(doc["nodes"][5]["children"][0]) == {"id", "validity_id"}                # SerializablePTOEdge, :30-34
    assert all(isinstance(b, bool) for b in doc["validities"][0])                      # Vec<Vec<bool>>, :76-80
    h = graph_io.load_pto_graph(str(p))
    for a in ("xy", "node_vid", "row_ptr", "col", "edge_vid", "p_row_ptr", "p_col", "p_edge_vid", "validities"):
        np.testing.assert_array_equal(getattr(h, a), getattr(g, a), err_msg=a)          # f64 states survive bit for bit


def test_parents_are_the_transpose_in_insertion_order():
    g = _oracle_prm_graph()
    prp, pcol, pev = graph_io.transpose_csr(g.row_ptr, g.col, g.edge_vid, g.n_nodes)
    np.testing.assert_array_equal(prp, g.p_row_ptr)
    np.testing.assert_array_equal(np.sort(pcol[prp[7]:prp[8]]), np.sort(g.p_col[g.p_row_ptr[7]:g.p_row_ptr[8]]))
    for k in range(g.n_nodes):                                                          # prm.rs:99-106: same sequences
        np.testing.assert_array_equal(g.p_col[g.p_row_ptr[k]:g.p_row_ptr[k + 1]], g.col[g.row_ptr[k]:g.row_ptr[k + 1]])


def test_rejects_wrong_state_dimension(tmp_path):
    p = tmp_path / "bad.json"
    p.write_text(json.dumps({"nodes": [{"state": [0.0, 1.0, 2.0], "validity_id": 0, "parents": [], "children": []}], "validities": [[True]]}))
    with pytest.raises(ValueError):
        graph_io.load_pto_graph(str(p))


def test_writer_spells_floats_like_ryu_and_pretty_prints_like_serde(tmp_path):
    """serde_json::to_writer_pretty: 2-space indent, one array element per line, floats as ryu's shortest round-trip spelling
    (1.0 not 1, 1e-7 not 1e-07, 0.001 not 1e-3); hand-checked expectations, and every value must survive the trip bit for bit."""
    vals = [1.0, 0.5, -0.0, 1e-7, 0.001, 123456.789, 1e16, 1.5e300, 5e-324, -2.2250738585072014e-308, 0.1 + 0.2, 1e21, 12345678901234567.0]
    xy = np.array(vals + [0.0] * (len(vals) % 2)).reshape(-1, 2)
    V = len(xy)
    g = graph_io.PTOGraphArrays(xy, np.zeros(V, np.int32), np.zeros(V + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32),
                                np.zeros(V + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), np.array([[1, 0]], np.uint8))
    p = tmp_path / "f.json"
    graph_io.save_pto_graph(str(p), g)
    text = p.read_text()
    for want in ("1.0,", "0.5\n", "-0.0,", "1e-7\n", "0.001,", "123456.789\n", "1e16,", "1.5e300\n", "5e-324,", "-2.2250738585072014e-308\n",
                 "0.30000000000000004,", "1e21\n", "1.2345678901234568e16,"):
        assert ("        " + want) in text, want
    assert text.startswith('{\n  "nodes": [\n    {\n      "state": [\n        1.0,\n        0.5\n      ],\n      "validity_id": 0,\n      "parents": [],\n      "children": []\n    },')
    assert text.endswith('  "validities": [\n    [\n      true,\n      false\n    ]\n  ]\n}')
    np.testing.assert_array_equal(np.array(json.loads(text)["nodes"][0]["state"]), xy[0])
    h = graph_io.load_pto_graph(str(p))
    assert h.xy.tobytes() == xy.tobytes()                     # incl. the sign of -0.0 and the denormal


def test_reader_accepts_any_key_order_and_compact_json(tmp_path):
    doc = {"validities": [[True, False], [False, True]],
           "nodes": [{"children": [{"validity_id": 1, "id": 1}], "parents": [], "validity_id": 0, "state": [0.25, -1e-3]},
                     {"state": [1, 2], "validity_id": 1, "parents": [{"id": 0, "validity_id": 1}], "children": []}]}
    p = tmp_path / "c.json"
    p.write_text(json.dumps(doc, separators=(",", ":")))
    g = graph_io.load_pto_graph(str(p))
    np.testing.assert_array_equal(g.xy, [[0.25, -0.001], [1.0, 2.0]])
    np.testing.assert_array_equal(g.row_ptr, [0, 1, 1])
    np.testing.assert_array_equal(g.col, [1])
    np.testing.assert_array_equal(g.edge_vid, [1])
    np.testing.assert_array_equal(g.p_row_ptr, [0, 0, 1])
    np.testing.assert_array_equal(g.validities, [[1, 0], [0, 1]])
    for bad in ('{"nodes": [], "validities": [[1]]}', '{"nodes": [{"state": [0,0], "validity_id": -1, "parents": [], "children": []}], "validities": []}',
                '{"nodes": [{"state": [0,0], "validity_id": 0, "parents": [], "children": [{"id": 5, "validity_id": 0}]}], "validities": []}', '{"nodes": ['):
        p.write_text(bad)
        with pytest.raises(ValueError):
            graph_io.load_pto_graph(str(p))


def test_graph_serialization(tmp_path):  # pto_graph.rs:566-572: save + load of create_minimal_graph (:434-444), 0 -> 1
    g = graph_io.PTOGraphArrays([[0.0, 0.0], [1.0, 0.0]], [0, 0], [0, 1, 1], [1], [0], [0, 0, 1], [0], [0], [[1]])
    p = tmp_path / "test_graph_serialization.json"
    graph_io.save_pto_graph(str(p), g)
    doc = json.loads(p.read_text())
    assert doc["validities"] == [[True]]
    assert [n["state"] for n in doc["nodes"]] == [[0.0, 0.0], [1.0, 0.0]]
    assert doc["nodes"][0]["children"] == [{"id": 1, "validity_id": 0}] and doc["nodes"][0]["parents"] == []
    assert doc["nodes"][1]["parents"] == [{"id": 0, "validity_id": 0}] and doc["nodes"][1]["children"] == []
    h = graph_io.load_pto_graph(str(p))
    for a in ("xy", "node_vid", "row_ptr", "col", "edge_vid", "p_row_ptr", "p_col", "p_edge_vid", "validities"):
        np.testing.assert_array_equal(getattr(h, a), getattr(g, a), err_msg=a)
