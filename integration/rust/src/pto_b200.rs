//! pto_b200.rs -- planner-level swap for `PTO::plan_belief_space` (pto.rs:152-182): child module of `pto`
//! (`#[cfg(feature = "b200")] #[path = "pto_b200.rs"] mod b200;`), so it may fill the private `expected_costs_to_goals`.
//!
//! `build_belief_graph` (pto.rs:185-259) materialises V x B belief nodes and calls `observe` V x B times; with the feature on the
//! belief graph stays IMPLICIT on the device (belief node id = node * B + belief, node types and observation edges from
//! per-visible-zone-set tables): reachable beliefs + one visibility test per NODE + `porrt_belief_vi`, then the
//! policy walk `porrt_extract_policy` (belief_graph.rs:184-267).  Expected costs are bit-identical to conditional_dijkstra's.
//! Roadmap growth (`PTO::grow_graph`, pto.rs:55-139) stays sequential and unchanged; through `B200Domain`'s per-query methods it
//! runs on the same library (batches of one).
#![cfg(feature = "b200")]

use super::PTO;
use crate::b200::{words_from_mask, B200Domain};
use crate::b200_ffi::*;
use crate::common::*;
use crate::pto_graph::PTOFuncs;
use crate::qmdp_policy_extractor::b200::export_csr;

impl<'a> PTO<'a, B200Domain<'a>, 2> {
    pub fn plan_belief_space_b200(&mut self, start_belief_state: &BeliefState) -> Policy<2> {
        assert_belief_state_validity(start_belief_state);
        let ctx = self.fns.ctx;
        let beliefs = self.fns.reachable_belief_states(start_belief_state);
        let b = beliefs.len();
        let n_worlds = self.n_worlds;
        let flat_beliefs: Vec<f64> = beliefs.iter().flatten().copied().collect();
        let (row_ptr, col, edge_vid, xy, node_vid) = export_csr(&self.graph);
        let v = self.graph.nodes.len();
        let states: Vec<[f64; 2]> = self.graph.nodes.iter().map(|n| n.state).collect();
        let visible = self.fns.visible_zones(&states); // observe()'s geometric test, once per node instead of once per (node, belief)
        let mw = self.fns.mask_words();
        let validities: Vec<u64> = self.fns.world_validities().iter().flat_map(|m| words_from_mask(m, mw)).collect();
        let n_validities = (validities.len() / mw) as i32;
        // final nodes with their finality masks (pto.rs:263-271)
        let (mut fin_ids, mut fin_masks): (Vec<i32>, Vec<u64>) = (Vec::new(), Vec::new());
        for (&final_id, validity) in self.conservative_reachability.final_nodes_with_validities() {
            fin_ids.push(final_id as i32);
            fin_masks.extend(words_from_mask(validity, mw));
        }
        // the V x B tables stay in ctx-owned pinned memory (porrt_belief_result): no second 8 * V * B byte copy
        ctx.check(unsafe {
            porrt_belief_vi(ctx.raw(), v as i64, row_ptr.as_ptr(), col.as_ptr(), edge_vid.as_ptr(), xy.as_ptr(), node_vid.as_ptr(),
                            validities.as_ptr(), n_validities, mw as i32, n_worlds as i32, flat_beliefs.as_ptr(), b as i32,
                            visible.as_ptr(), fin_ids.as_ptr(), fin_masks.as_ptr(), fin_ids.len() as i32, std::ptr::null_mut(),
                            std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut())
        });
        let (mut dist_ptr, mut type_ptr): (*const f64, *const u8) = (std::ptr::null(), std::ptr::null());
        ctx.check(unsafe { porrt_belief_result(ctx.raw(), &mut dist_ptr, &mut type_ptr, std::ptr::null_mut(), std::ptr::null_mut()) });
        self.expected_costs_to_goals = unsafe { std::slice::from_raw_parts(dist_ptr, v * b) }.to_vec();
        // policy nodes in the reference's creation order
        let mut cap = 4096i64;
        let (mut n, mut cost) = (0i64, 0.0f64);
        let (mut node, mut belief, mut parent, mut leaf);
        loop {
            node = vec![0i32; cap as usize];
            belief = vec![0i32; cap as usize];
            parent = vec![0i32; cap as usize];
            leaf = vec![0u8; cap as usize];
            let rc = unsafe {
                porrt_extract_policy(ctx.raw(), node.as_mut_ptr(), belief.as_mut_ptr(), parent.as_mut_ptr(), leaf.as_mut_ptr(), cap, &mut n, &mut cost)
            };
            if rc == 4 {
                cap = n;
                continue;
            }
            ctx.check(rc);
            break;
        }
        let mut policy: Policy<2> = Policy { nodes: Vec::new(), leafs: Vec::new(), expected_costs: cost };
        for k in 0..n as usize {
            let original_id = node[k] as usize * b + belief[k] as usize; // the dense belief node id of pto.rs:197-199
            let id = policy.add_node(&states[node[k] as usize], &beliefs[belief[k] as usize], original_id, leaf[k] != 0);
            if parent[k] >= 0 {
                policy.add_edge(parent[k] as usize, id);
            }
        }
        policy
    }
}
