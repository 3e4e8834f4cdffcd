//! b200.rs -- safe Rust layer over libporrt_b200 (`crate::b200_ffi`, generated from include/porrt_b200.h), compiled only with
//! `--features b200`.  In `src/lib.rs`:
//!
//! ```ignore
//! #[cfg(feature = "b200")] pub mod b200_ffi;
//! #[cfg(feature = "b200")] pub mod b200;
//! ```
//!
//! What is here:
//!  * `B200Ctx`       -- RAII handle (porrt_ctx_create / porrt_ctx_destroy), status -> panic with the library's message, exactly
//!                       where the reference itself panics (`PORRT_ERR_PANIC`) or on misuse;
//!  * `B200Domain`    -- a `Map` / `MapShelfDomain` whose geometry lives on the GPU: `impl PTOFuncs<2>` and `impl RTTFuncs<2>`
//!                       (per-query methods = batches of one, used by the sequential planners RRT / PTO growth / refiner)
//!                       plus the batched forms the planner-level swaps call (`prm_b200.rs`, `pto_b200.rs`, `qmdp_b200.rs`);
//!  * `B200KdTree`    -- `KdTree<2>`'s interface (`reset`, `add`, `nearest_neighbor[_filtered]`, `nearest_neighbors`) over the
//!                       device-resident vertex set.
//! Not compiled in the image this repository is built in (no cargo / rustc); the same entry points are exercised from C
//! (tests/c_abi_harness.c) and Python (po_rrt_b200/api.py, tests/test_gpu_parity.py).
#![cfg(feature = "b200")]

use crate::b200_ffi::*;
use crate::common::*;
use crate::pto_graph::{PTOFuncs, PTONode};
use crate::rrt::RTTFuncs;
use bitvec::prelude::*;
use std::ffi::CStr;

pub const DOMAIN_DOOR: i32 = 0; // map_io.rs `Map`
pub const DOMAIN_SHELF: i32 = 1; // map_shelves_io.rs `MapShelfDomain`

/// per-element codes of include/porrt_b200.h
const INVALID: i32 = -1;
const PANIC_OOB: i32 = -2;
const PANIC_ZONE_UNWRAP: i32 = -3;
const PANIC_MULTI_ZONE: i32 = -4;

pub struct B200Ctx {
    raw: *mut PorrtCtx,
}

impl B200Ctx {
    pub fn new(device: i32) -> Self {
        let mut raw: *mut PorrtCtx = std::ptr::null_mut();
        let rc = unsafe { porrt_ctx_create(device, &mut raw) };
        assert!(rc == 0 && !raw.is_null(), "porrt_ctx_create failed ({}): no sm_100 device? there is no CPU fallback", rc);
        Self { raw }
    }

    pub fn raw(&self) -> *mut PorrtCtx {
        self.raw
    }

    /// 0 = ok; anything else panics with the library's message (the reference panics in the same situations)
    pub fn check(&self, rc: i32) {
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(porrt_last_error(self.raw)) }.to_string_lossy().into_owned();
            panic!("porrt_b200 error {}: {}", rc, msg);
        }
    }
}

impl Drop for B200Ctx {
    fn drop(&mut self) {
        unsafe { porrt_ctx_destroy(self.raw) };
    }
}

/// validity code of the C ABI -> the reference's Option<usize> (pto_c.rs:17-18: >= 0 id, < 0 invalid), panics where it panics
fn to_validity(code: i32) -> Option<usize> {
    match code {
        c if c >= 0 => Some(c as usize),
        INVALID => None,
        PANIC_OOB => panic!("Image index out of bounds"), // image::get_pixel (map_io.rs:167,226)
        PANIC_ZONE_UNWRAP => panic!("called `Option::unwrap()` on a `None` value"), // map_io.rs:172,231
        PANIC_MULTI_ZONE => panic!("multiple zone traversal not supported"), // map_io.rs:233
        c => panic!("unknown validity code {}", c),
    }
}

/// BitVec<Lsb0, usize> <-> the ABI's u64 words (bit w of word w / 64 = world w)
pub fn mask_from_words(words: &[u64], n_worlds: usize) -> WorldMask {
    (0..n_worlds).map(|w| (words[w / 64] >> (w % 64)) & 1 == 1).collect()
}

pub fn words_from_mask(mask: &WorldMask, mask_words: usize) -> Vec<u64> {
    let mut out = vec![0u64; mask_words];
    for (w, bit) in mask.iter().enumerate() {
        if *bit {
            out[w / 64] |= 1u64 << (w % 64);
        }
    }
    out
}

/// A map whose pixels live in HBM.  Built from the decoded gray images the reference already has in memory
/// (`Map::b200_images()` / `MapShelfDomain::b200_images()`, added next to `open_image` under the feature flag).
pub struct B200Domain<'a> {
    pub ctx: &'a B200Ctx,
    pub kind: i32,
    n_zones: usize,
    n_worlds: usize,
    mask_words: usize,
    world_validities: Vec<WorldMask>,
    zone_positions: Vec<[f64; 2]>,
}

impl<'a> B200Domain<'a> {
    /// `occ` / `zones`: row-major 8-bit gray, row 0 = top (what `image::open(..)` yields as ImageLuma8, map_io.rs:98-105)
    pub fn upload(ctx: &'a B200Ctx, kind: i32, occ: &[u8], zones: Option<&[u8]>, height: u32, width: u32, low: [f64; 2],
                  up: [f64; 2], visibility_distance: f64) -> Self {
        assert_eq!(occ.len(), (height * width) as usize);
        let zp = zones.map_or(std::ptr::null(), |z| z.as_ptr());
        ctx.check(unsafe {
            porrt_map_upload(ctx.raw(), occ.as_ptr(), zp, height as i32, width as i32, low.as_ptr(), up.as_ptr(), kind, visibility_distance)
        });
        let (mut nz, mut nw, mut nv, mut mw) = (0i32, 0i32, 0i32, 0i32);
        ctx.check(unsafe { porrt_map_info(ctx.raw(), &mut nz, &mut nw, &mut nv, &mut mw) });
        let mut words = vec![0u64; (nv * mw) as usize];
        ctx.check(unsafe { porrt_map_world_validities(ctx.raw(), words.as_mut_ptr()) });
        let world_validities = (0..nv as usize).map(|v| mask_from_words(&words[v * mw as usize..(v + 1) * mw as usize], nw as usize)).collect();
        let mut zxy = vec![0.0f64; 2 * nz as usize];
        if nz > 0 {
            ctx.check(unsafe { porrt_map_zone_positions(ctx.raw(), zxy.as_mut_ptr()) });
        }
        let zone_positions = zxy.chunks(2).map(|c| [c[0], c[1]]).collect();
        Self { ctx, kind, n_zones: nz as usize, n_worlds: nw as usize, mask_words: mw as usize, world_validities, zone_positions }
    }

    pub fn n_zones(&self) -> usize {
        self.n_zones
    }
    pub fn mask_words(&self) -> usize {
        self.mask_words
    }
    pub fn zone_positions(&self) -> &[[f64; 2]] {
        &self.zone_positions
    }

    // ---- batched forms (what the planner-level swaps use) -------------------------------------------------------------------
    pub fn state_validity_batch(&self, states: &[[f64; 2]]) -> Vec<i32> {
        let mut out = vec![0i32; states.len()];
        self.ctx.check(unsafe { porrt_state_validity(self.ctx.raw(), states.as_ptr() as *const f64, states.len() as i64, out.as_mut_ptr()) });
        out
    }

    /// edge i runs from `from[i]` to `to[i]` in that direction (Bresenham is direction dependent, map_io.rs:216-241)
    pub fn edge_validity_batch(&self, from: &[[f64; 2]], to: &[[f64; 2]]) -> Vec<i32> {
        assert_eq!(from.len(), to.len());
        let mut out = vec![0i32; from.len()];
        self.ctx.check(unsafe {
            porrt_edge_validity(self.ctx.raw(), from.as_ptr() as *const f64, to.as_ptr() as *const f64, from.len() as i64,
                                out.as_mut_ptr(), std::ptr::null_mut())
        });
        out
    }

    /// transition_validator(&nodes[from_id[i]], &nodes[to_id[i]]) for nodes of the uploaded vertex set: 8 bytes in, 1 byte out per edge
    pub fn edge_validity_by_node_id(&self, from_id: &[i32], to_id: &[i32]) -> Vec<i8> {
        assert_eq!(from_id.len(), to_id.len());
        let mut out = vec![0i8; from_id.len()];
        self.ctx.check(unsafe { porrt_edge_validity_indexed_i8(self.ctx.raw(), from_id.as_ptr(), to_id.as_ptr(), from_id.len() as i64, out.as_mut_ptr()) });
        out
    }

    /// the geometric half of observe(): bit z = zone z is within `visibility_distance` and in line of sight (map_io.rs:285-290)
    pub fn visible_zones(&self, states: &[[f64; 2]]) -> Vec<u64> {
        let mut mask = vec![0u64; states.len()];
        let mut status = vec![0i32; states.len()];
        self.ctx.check(unsafe {
            porrt_visibility(self.ctx.raw(), states.as_ptr() as *const f64, states.len() as i64, mask.as_mut_ptr(), status.as_mut_ptr())
        });
        for s in &status {
            if *s != 0 {
                to_validity(*s); // panics with the reference's message
            }
        }
        mask
    }

    /// get_successor_belief_states (map_io.rs:243-278 / map_shelves_io.rs:205-240): host-side belief algebra, unchanged
    fn successor_belief_states(&self, belief_state: &BeliefState, zone_id: usize) -> Vec<BeliefState> {
        let in_zone_world = |w: usize| -> bool {
            if self.kind == DOMAIN_DOOR { !self.world_validities[zone_id][w] } else { w == zone_id }
        };
        // DOOR: world_validities[zone] = worlds in which the door is OPEN (zone_index_to_world_mask), [closed, open];
        // SHELF: [object there, object not there]
        let first: Vec<f64> = (0..belief_state.len()).map(|w| if in_zone_world(w) { belief_state[w] } else { 0.0 }).collect();
        let second: Vec<f64> = (0..belief_state.len()).map(|w| if in_zone_world(w) { 0.0 } else { belief_state[w] }).collect();
        let mut out = Vec::new();
        for mut b in vec![first, second] {
            let sum = b.iter().fold(0.0, |sum, p| sum + p);
            for p in b.iter_mut() {
                *p /= sum;
            }
            if !b.iter().any(|p| p.is_nan()) {
                out.push(b);
            }
        }
        out
    }
}

impl<'a> PTOFuncs<2> for B200Domain<'a> {
    fn n_worlds(&self) -> usize {
        self.n_worlds
    }

    fn state_validity(&self, state: &[f64; 2]) -> Option<usize> {
        to_validity(self.state_validity_batch(std::slice::from_ref(state))[0])
    }

    fn transition_validator(&self, from: &PTONode<2>, to: &PTONode<2>) -> Option<usize> {
        to_validity(self.edge_validity_batch(std::slice::from_ref(&from.state), std::slice::from_ref(&to.state))[0])
    }

    fn reachable_belief_states(&self, belief_state: &BeliefState) -> Vec<BeliefState> {
        let mut n: i32 = 0;
        let mut cap: i32 = 4096;
        loop {
            let mut out = vec![0.0f64; cap as usize * self.n_worlds];
            let rc = unsafe { porrt_reachable_belief_states(self.ctx.raw(), belief_state.as_ptr(), out.as_mut_ptr(), cap, &mut n) };
            if rc == 4 {
                cap = n; // PORRT_ERR_CAPACITY: n holds the count
                continue;
            }
            self.ctx.check(rc);
            return out[..n as usize * self.n_worlds].chunks(self.n_worlds).map(|c| c.to_vec()).collect();
        }
    }

    fn world_validities(&self) -> Vec<WorldMask> {
        self.world_validities.clone()
    }

    fn observe(&self, state: &[f64; 2], belief_state: &BeliefState) -> Vec<BeliefState> {
        // observe_impl (map_io.rs:281-300): zones ascending, every current belief split in turn; the visibility test is the device's
        let visible = self.visible_zones(std::slice::from_ref(state))[0];
        let mut output_beliefs = vec![belief_state.clone()];
        for zone_id in 0..self.n_zones {
            if (visible >> zone_id) & 1 == 1 {
                let beliefs = output_beliefs.clone();
                output_beliefs.clear();
                for belief in beliefs {
                    output_beliefs.extend(self.successor_belief_states(&belief, zone_id));
                }
            }
        }
        output_beliefs
    }
}

impl<'a> RTTFuncs<2> for B200Domain<'a> {
    // rrt.rs:274-286 (tests' Funcs wrapper): a state is valid iff it is Free, an edge iff it traverses only free space
    fn state_validator(&self, state: &[f64; 2]) -> bool {
        PTOFuncs::state_validity(self, state) == Some(self.world_validities.len() - 1)
    }

    fn transition_validator(&self, from: &[f64; 2], to: &[f64; 2]) -> bool {
        to_validity(self.edge_validity_batch(std::slice::from_ref(from), std::slice::from_ref(to))[0]) == Some(self.world_validities.len() - 1)
    }
}

/// `KdTree<2>` (nearest_neighbor.rs:10-127) over the device-resident vertex set.  The reference returns `&KdNode`; here the
/// callers get `(id, state)` pairs -- every call site only reads `.id` and `.state` (rrt.rs:113-163, prm.rs:72-113, pto.rs:64-126).
pub struct B200KdTree<'a> {
    ctx: &'a B200Ctx,
    states: Vec<[f64; 2]>,
    rank: Option<Vec<i32>>, // kd pre-order rank of the current set (the reference's visit order), computed on demand
}

impl<'a> B200KdTree<'a> {
    pub fn new(ctx: &'a B200Ctx, state: [f64; 2]) -> Self {
        let mut t = Self { ctx, states: Vec::new(), rank: None };
        t.reset(state);
        t
    }

    pub fn reset(&mut self, state: [f64; 2]) {
        self.states = vec![state];
        self.rank = None;
        self.ctx.check(unsafe { porrt_vertices_set(self.ctx.raw(), state.as_ptr(), 1, 0.0) });
    }

    /// ids are assigned in insertion order, like every caller of `KdTree::add(state, id)` does (id == nodes.len() - 1)
    pub fn add(&mut self, state: [f64; 2], id: usize) {
        assert_eq!(id, self.states.len());
        self.states.push(state);
        self.rank = None;
        self.ctx.check(unsafe { porrt_vertices_append(self.ctx.raw(), state.as_ptr(), 1) });
    }

    pub fn nearest_neighbor(&self, state: [f64; 2]) -> (usize, [f64; 2]) {
        self.nearest(state, None, 0)
    }

    /// `validator(id)` of the reference is always "bit `world` of reachability(id)" (pto.rs:74-77): passed as the mask table
    pub fn nearest_neighbor_filtered(&self, state: [f64; 2], reach_words: &[u64], words_per_vertex: usize, world: usize) -> (usize, [f64; 2]) {
        assert_eq!(reach_words.len(), self.states.len() * words_per_vertex);
        self.nearest(state, Some((reach_words, words_per_vertex)), world)
    }

    fn nearest(&self, state: [f64; 2], reach: Option<(&[u64], usize)>, world: usize) -> (usize, [f64; 2]) {
        let (mut id, mut dist, mut ties) = (0i32, 0.0f64, 0i32);
        let w = world as u32;
        let (rp, rw, wp) = match reach {
            Some((r, k)) => (r.as_ptr(), k as i32, &w as *const u32),
            None => (std::ptr::null(), 1, std::ptr::null()),
        };
        self.ctx.check(unsafe { porrt_nearest(self.ctx.raw(), state.as_ptr(), 1, rp, rw, wp, &mut id, &mut dist, &mut ties) });
        if id < 0 {
            return (0, self.states[0]); // nothing passes the filter: the reference returns the root (nearest_neighbor.rs:89)
        }
        // ties > 1: several vertices at exactly the winning distance.  Exact duplicates chain to the right in insertion order, the
        // reference's strict `d < dmin` keeps the first visited = lowest id = what the library returns (argmin over (d2, id)).
        (id as usize, self.states[id as usize])
    }

    /// hits in the reference's visit order (node, left, right = kd pre-order): the device returns the SET (ids ascending), the
    /// order is restored with the pre-order rank of the current tree
    pub fn nearest_neighbors(&mut self, state: [f64; 2], radius: f64) -> Vec<(usize, [f64; 2])> {
        let mut offsets = [0i64; 2];
        let mut total = 0i64;
        let mut cap = 256usize;
        let mut ids: Vec<i32>;
        loop {
            ids = vec![0i32; cap];
            let rc = unsafe {
                porrt_radius_query(self.ctx.raw(), state.as_ptr(), &radius, 1, std::ptr::null(), std::ptr::null(), 1, std::ptr::null(),
                                   offsets.as_mut_ptr(), ids.as_mut_ptr(), cap as i64, &mut total)
            };
            if rc == 4 {
                cap = total as usize;
                continue;
            }
            self.ctx.check(rc);
            break;
        }
        ids.truncate(total as usize);
        if ids.len() > 1 {
            if self.rank.is_none() {
                let mut rank = vec![0i32; self.states.len()];
                // xy == NULL: rank the ctx's own vertex set in place
                self.ctx.check(unsafe { porrt_kd_preorder_rank(self.ctx.raw(), std::ptr::null(), self.states.len() as i64, rank.as_mut_ptr()) });
                self.rank = Some(rank);
            }
            let rank = self.rank.as_ref().unwrap();
            ids.sort_by_key(|&i| rank[i as usize]);
        }
        ids.iter().map(|&i| (i as usize, self.states[i as usize])).collect()
    }
}
