//! qmdp_b200.rs -- planner-level swap for `QMdpPolicyExtractor::plan_qmdp` (qmdp_policy_extractor.rs:23-35): child module of
//! `qmdp_policy_extractor` (`#[cfg(feature = "b200")] #[path = "qmdp_b200.rs"] mod b200;`).
//!
//! The reference runs one heap `dijkstra` per world over a `PTOGraphWorldView`; here all worlds are value columns of one roadmap
//! solved on chip by `porrt_sssp_worlds`.  `cost_to_goals[w][v]` is bit-identical (f64) to the reference's.
#![cfg(feature = "b200")]

use super::QMdpPolicyExtractor;
use crate::b200::{words_from_mask, B200Domain};
use crate::b200_ffi::*;
use crate::pto_graph::PTOFuncs;

/// PTOGraph -> the CSR arrays of the C ABI: children adjacency in stored order (pto_graph.rs:204-207)
pub fn export_csr(graph: &crate::pto_graph::PTOGraph<2>) -> (Vec<i64>, Vec<i32>, Vec<i32>, Vec<f64>, Vec<i32>) {
    let n = graph.nodes.len();
    let mut row_ptr = Vec::with_capacity(n + 1);
    let (mut col, mut edge_vid, mut xy, mut node_vid) = (Vec::new(), Vec::new(), Vec::with_capacity(2 * n), Vec::with_capacity(n));
    row_ptr.push(0i64);
    for node in &graph.nodes {
        for e in &node.children {
            col.push(e.id as i32);
            edge_vid.push(e.validity_id as i32);
        }
        row_ptr.push(col.len() as i64);
        xy.extend_from_slice(&node.state);
        node_vid.push(node.validity_id as i32);
    }
    (row_ptr, col, edge_vid, xy, node_vid)
}

impl<'a> QMdpPolicyExtractor<'a, B200Domain<'a>, 2> {
    pub fn plan_qmdp_b200(&mut self) -> Result<(), &'static str> {
        let ctx = self.fns.ctx;
        let n_worlds = *self.n_worlds;
        let (row_ptr, col, _edge_vid, xy, node_vid) = export_csr(self.graph);
        let v = self.graph.nodes.len();
        let mut finals_ptr = vec![0i64; n_worlds + 1];
        let mut finals: Vec<i32> = Vec::new();
        for world in 0..n_worlds {
            let final_nodes = self.conservative_reachability.get_final_nodes_for_world(world);
            if final_nodes.is_empty() {
                return Err(&"We should have final node ids for each world");
            }
            finals.extend(final_nodes.iter().map(|&f| f as i32));
            finals_ptr[world + 1] = finals.len() as i64;
        }
        let mw = self.fns.mask_words();
        let validities: Vec<u64> = self.fns.world_validities().iter().flat_map(|m| words_from_mask(m, mw)).collect();
        let n_validities = (validities.len() / mw) as i32;
        let mut dist = vec![0.0f64; n_worlds * v];
        ctx.check(unsafe {
            porrt_sssp_worlds(ctx.raw(), v as i64, row_ptr.as_ptr(), col.as_ptr(), xy.as_ptr(), node_vid.as_ptr(), validities.as_ptr(),
                              n_validities, mw as i32, n_worlds as i32, finals_ptr.as_ptr(), finals.as_ptr(), dist.as_mut_ptr(),
                              std::ptr::null_mut())
        });
        self.cost_to_goals = dist.chunks(v).map(|c| c.to_vec()).collect();
        Ok(())
    }
}
