//! refiner_b200.rs -- planner-level swap for `PTOPolicyRefiner::refine_solution` (pto_policy_refiner.rs:85-133): child module of
//! `pto_policy_refiner` (`#[cfg(feature = "b200")] #[path = "refiner_b200.rs"] mod b200;`).
//!
//! Works on a policy that `PTO::plan_belief_space_b200` (pto_b200.rs) just extracted on the same `B200Ctx`: the library still holds
//! that belief-space result, so the policy travels as three index arrays (`original_node_id = node * B + belief`, parents in
//! creation order) and comes back refined.
//!   * `RefinmentStrategy::PartialShortCut(n)` (:135-206): `porrt_refine_policy_shortcut` -- decompose, all pieces' trial loops in
//!     one device launch (one CTA per piece), recompose, expected cost.  The reference draws the trials from `DiscreteSampler::new()`
//!     (seed 0) per piece; so does the library (`sampler_seed = 0`).
//!   * `RefinmentStrategy::Reparent(radius)` (:208-322): `porrt_refine_policy_reparent` -- trees on the host, every candidate
//!     transition of every tree in one device batch, the label-correcting loop on the host in the reference's queue order.
//! The refined states, tree and expected cost equal the reference's bit for bit (tests/test_gpu_parity.py:
//! test_refine_solution_partial_shortcut, test_refine_solution_reparent).  The `Vec<RefinmentTree>` the reference also returns is
//! not rebuilt: callers in the crate (main.rs:221,270,442) only keep it for drawing.
#![cfg(feature = "b200")]

use super::{PTOPolicyRefiner, RefinmentStrategy};
use crate::b200::B200Domain;
use crate::b200_ffi::*;
use crate::common::*;

impl<'a> PTOPolicyRefiner<'a, B200Domain<'a>, 2> {
    pub fn refine_solution_b200(&mut self, strategy: RefinmentStrategy) -> Policy<2> {
        let start_time = std::time::Instant::now();
        let ctx = self.fns.ctx;
        let b = self.belief_graph.reachable_belief_states.len();
        let n_pol = self.policy.nodes.len();
        // the policy as (node, belief, parent) in creation order (Policy::add_node / add_edge, common.rs:42-66)
        let node: Vec<i32> = self.policy.nodes.iter().map(|n| (n.original_node_id / b) as i32).collect();
        let belief: Vec<i32> = self.policy.nodes.iter().map(|n| (n.original_node_id % b) as i32).collect();
        let parent: Vec<i32> = self.policy.nodes.iter().map(|n| n.parent.map_or(-1, |p| p as i32)).collect();
        let mut cap = n_pol.max(64) as i64;
        let (mut n, mut cost) = (0i64, 0.0f64);
        let (mut xy, mut out_node, mut out_belief, mut out_parent, mut out_leaf);
        loop {
            xy = vec![0.0f64; 2 * cap as usize];
            out_node = vec![0i32; cap as usize];
            out_belief = vec![0i32; cap as usize];
            out_parent = vec![0i32; cap as usize];
            out_leaf = vec![0u8; cap as usize];
            let rc = unsafe {
                match strategy {
                    RefinmentStrategy::PartialShortCut(n_iterations) => porrt_refine_policy_shortcut(
                        ctx.raw(), node.as_ptr(), belief.as_ptr(), parent.as_ptr(), n_pol as i64, n_iterations as i32, 0,
                        xy.as_mut_ptr(), out_node.as_mut_ptr(), out_belief.as_mut_ptr(), out_parent.as_mut_ptr(), out_leaf.as_mut_ptr(),
                        cap, &mut n, &mut cost, std::ptr::null_mut()),
                    RefinmentStrategy::Reparent(radius) => porrt_refine_policy_reparent(
                        ctx.raw(), node.as_ptr(), belief.as_ptr(), parent.as_ptr(), n_pol as i64, radius,
                        xy.as_mut_ptr(), out_node.as_mut_ptr(), out_belief.as_mut_ptr(), out_parent.as_mut_ptr(), out_leaf.as_mut_ptr(),
                        cap, &mut n, &mut cost, std::ptr::null_mut(), std::ptr::null_mut()),
                }
            };
            if rc == 4 && n > cap {
                cap = n; // PORRT_ERR_CAPACITY: the recomposed policy is larger than the buffers
                continue;
            }
            ctx.check(rc);
            break;
        }
        // recompose's policy (pto_policy_refiner.rs:324-393): nodes in creation order; a parent of -1 beyond node 0 is a piece the
        // reference leaves unconnected (a one-node piece is a start but never an end there)
        let mut policy: Policy<2> = Policy { nodes: Vec::new(), leafs: Vec::new(), expected_costs: cost };
        for k in 0..n as usize {
            let original_id = out_node[k] as usize * b + out_belief[k] as usize;
            let belief_state = &self.belief_graph.reachable_belief_states[out_belief[k] as usize];
            let id = policy.add_node(&[xy[2 * k], xy[2 * k + 1]], belief_state, original_id, false);
            if out_parent[k] >= 0 {
                policy.add_edge(out_parent[k] as usize, id);
            }
        }
        for k in 0..n as usize {
            if out_leaf[k] != 0 {
                policy.leafs.push(k); // "set remaining leafs" (:384-389): nodes without children, ascending
            }
        }
        self.refinement_s = start_time.elapsed().as_secs_f64();
        policy
    }
}
