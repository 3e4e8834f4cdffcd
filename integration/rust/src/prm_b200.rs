//! prm_b200.rs -- planner-level swap for `PRM` (prm.rs:13-129).  Declared as a CHILD module of `prm` so that it sees the private
//! `continuous_sampler`:
//!
//! ```ignore
//! // at the end of src/prm.rs
//! #[cfg(feature = "b200")]
//! #[path = "prm_b200.rs"]
//! mod b200;
//! ```
//!
//! `PRM::grow_graph` (prm.rs:38-50) adds its samples one by one; nothing in the loop depends on a validity result, so with the
//! feature on the n_iter samples are drawn first and the whole roadmap comes from ONE `porrt_prm_build` call (prefix-restricted
//! radius batch in kd pre-order, one edge batch, CSR in the reference's insertion order -- DESIGN.md 3.4).  `graph.nodes[k]
//! .children / .parents` are filled from the returned CSR and are identical, element for element, to what `add_sample`
//! (prm.rs:52-109) would have produced; the kd-tree is kept in sync for `plan_path`.
#![cfg(feature = "b200")]

use super::PRM;
use crate::b200::B200Domain;
use crate::b200_ffi::*;
use crate::pto_graph::PTOEdge;

impl<'a> PRM<'a, B200Domain<'a>, 2> {
    /// drop-in for `grow_graph(max_step, search_radius, n_iter)` after `init(start)`
    pub fn grow_graph_b200(&mut self, max_step: f64, search_radius: f64, n_iter: usize) {
        let ctx = self.fns.ctx;
        // 1. the sample stream, in the order add_sample would have consumed it
        let first_new = self.graph.nodes.len();
        let mut states: Vec<[f64; 2]> = self.graph.nodes.iter().map(|n| n.state).collect();
        for _ in 0..n_iter {
            states.push(self.continuous_sampler.sample());
        }
        let n = states.len();
        // 2. one build: node k queries the tree as it was before k arrived, radius = heuristic_radius(k + 1, ..) (prm.rs:61-65)
        let mut row_ptr = vec![0i64; n + 1];
        let mut n_edges = 0i64;
        let mut col: Vec<i32> = Vec::new();
        let rc = unsafe {
            porrt_prm_build(ctx.raw(), states.as_ptr() as *const f64, n as i64, max_step, search_radius, row_ptr.as_mut_ptr(),
                            std::ptr::null_mut(), 0, &mut n_edges, std::ptr::null_mut())
        };
        if rc == 4 {
            // PORRT_ERR_CAPACITY: the result is retained on the device, fetch it into a buffer of the right size
            col = vec![0i32; n_edges as usize];
            ctx.check(unsafe { porrt_prm_fetch(ctx.raw(), row_ptr.as_mut_ptr(), col.as_mut_ptr(), n_edges) });
        } else {
            ctx.check(rc);
        }
        // 3. the graph: PRM edges carry validity id 0 (prm.rs:98-106: add_edge(.., .., 0))
        for k in first_new..n {
            self.graph.add_node(states[k], 0);
        }
        for u in 0..n {
            let row = &col[row_ptr[u] as usize..row_ptr[u + 1] as usize];
            self.graph.nodes[u].children = row.iter().map(|&v| PTOEdge { id: v as usize, validity_id: 0 }).collect();
        }
        // parents in insertion order: edge (from, to) pushes `from` onto parents[to] at the time it is added; for a PRM every
        // edge exists in both directions and was added at the same time as its mirror, so parents[u] == children[u]
        for u in 0..n {
            self.graph.nodes[u].parents = self.graph.nodes[u].children.clone();
        }
        // 4. the kd-tree for plan_path's two nearest_neighbor queries
        for k in first_new..n {
            self.kdtree.add(states[k], k);
        }
        self.n_it += n_iter;
    }
}
