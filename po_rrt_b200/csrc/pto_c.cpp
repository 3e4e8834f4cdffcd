// pto_c.cpp -- the reference's exported C planner API (src/pto_c.rs:63-270) on top of the hot-path ABI (include/porrt_b200.h).
// Builds po_rrt_b200/libpo_rrt_c.so; declarations and the list of deliberate differences: include/po_rrt_c.h.
//
// plan() = the reference's plan_inner! (pto_c.rs:209-224):
//   PTO::grow_graph            pto.rs:55-139       host: one sample at a time, every answer is a callback of the caller
//   PTO::plan_belief_space     pto.rs:152-182      build_belief_graph (:185-259, observer callback per (node, belief)) on the host,
//                                                  conditional_dijkstra + extract_policy on the DEVICE through
//                                                  porrt_conditional_dijkstra_nd / porrt_extract_policy_graph_nd
//   refine_solution(PartialShortCut(n))  pto_policy_refiner.rs:85-206,324-423   host: callbacks again
//   save_paths / save_planning_metrics   pto_c.rs:272-312
// States have a run-time dimension (the reference instantiates N = 2, 3, 7, 9).  No pixel, grid or map is touched here: the
// caller's callbacks own the world model, as in the reference.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/po_rrt_c.h"
#include "../../include/porrt_b200.h"
#include "pcg64.h"

#define PTOC_API extern "C" __attribute__((visibility("default")))

struct CPlanningProblem {
  // input (pto_c.rs:30-49)
  size_t state_dim = 0, n_worlds = 0;
  double* low = nullptr; size_t low_size = 0;
  double* up = nullptr; size_t up_size = 0;
  size_t** world_validities = nullptr; size_t world_validities_size = 0;
  StateValidityCallbackType state_validity_callback = nullptr;
  TransitionValidityCallbackType transition_validity_callback = nullptr;
  CostEvaluatorCallbackType cost_evaluator_callback = nullptr;
  ObserverCallbackType observer_callback = nullptr;
  double* start_belief_state = nullptr; size_t start_belief_state_size = 0;
  double** reachable_belief_states = nullptr; size_t reachable_belief_states_size = 0;
  GoalCallbackType goal_callback = nullptr;
  GoalExampleCallbackType goal_example_callback = nullptr;
  size_t n_iterations_min = 0, n_iterations_max = 0;
  double max_step = 0.0, search_radius = 0.0;
  size_t refine_iterations = 0;
  // output (:51-61)
  std::vector<std::vector<std::vector<double>>> paths;
  std::vector<size_t> paths_lengths;
  double expected_costs = 0.0;
  size_t n_iterations = 0;
  double graph_growth_s = 0.0, belief_space_expansion_s = 0.0, dynamic_programming_s = 0.0, refinement_s = 0.0, total_s = 0.0;
  // additions
  bool seeded = false; uint64_t seed = 0;
  int32_t device = 0; bool free_observer_arrays = true;
  porrt_ctx* ctx = nullptr; int32_t ctx_device = -1;
  bool failed = false; std::string error;
  size_t n_nodes = 0, n_belief_nodes = 0, n_belief_edges = 0, n_sweeps = 0, n_policy_nodes = 0;
};

namespace {

typedef std::chrono::steady_clock Clock;
double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

struct PlanError { std::string msg; };
[[noreturn]] void fail(const std::string& m) { throw PlanError{m}; }

// common.rs:192-213
double norm1(const double* a, const double* b, int n) { double d = 0.0; for (int i = 0; i < n; ++i) d += std::fabs(b[i] - a[i]); return d; }
double norm2(const double* a, const double* b, int n) { double d2 = 0.0; for (int i = 0; i < n; ++i) { const double dx = b[i] - a[i]; d2 += dx * dx; } return std::sqrt(d2); }
// common.rs:352-355, wrapping like a release build
uint64_t belief_hash(const double* bs, size_t n) {
  uint64_t h = 0, p10 = 1;
  for (size_t i = 0; i < n; ++i) {
    const double r = std::round(bs[i] * 1000.0);
    const uint64_t v = (r <= 0.0 || std::isnan(r)) ? 0 : (r >= 18446744073709551615.0 ? UINT64_MAX : (uint64_t)r);
    h += (p10 + 1) * v;
    p10 *= 10;
  }
  return h;
}
inline bool bit(const uint64_t* m, size_t w) { return (m[w >> 6] >> (w & 63)) & 1ull; }

Pcg64 make_rng(const CPlanningProblem* p) {
  if (p->seeded) return Pcg64::seed_from_u64(p->seed);
  std::random_device rd;                       // new_true_random: Pcg64::from_rng(thread_rng()) fills the 32-byte seed
  uint32_t w[8];
  for (uint32_t& x : w) x = rd();
  return Pcg64::from_seed_words(w);
}

// nearest_neighbor.rs:3-126 for states of n doubles; kd node k holds roadmap node k (PTO adds them in the same order, pto.rs:62,124)
struct KdTree {
  int n; const std::vector<double>* xs;
  std::vector<int32_t> left, right;
  void reset() { left.assign(1, -1); right.assign(1, -1); }
  const double* st(int32_t k) const { return xs->data() + (size_t)k * n; }
  void add(int32_t id) {                                         // :29-46
    const double* s = st(id);
    left.push_back(-1); right.push_back(-1);
    int32_t cur = 0;
    for (int axis = 0;; axis = (axis + 1) % n) {
      int32_t& next = s[axis] < st(cur)[axis] ? left[cur] : right[cur];
      if (next >= 0) cur = next; else { next = id; return; }
    }
  }
  template <class F> int32_t nearest_filtered(const double* s, F&& ok) const {   // :52-92
    double dmin = std::numeric_limits<double>::infinity();
    int32_t best = 0;
    struct Frame { int32_t node; int axis; int stage; };
    std::vector<Frame> stack(1, Frame{0, 0, 0});
    while (!stack.empty()) {
      Frame& f = stack.back();
      const double* fs = st(f.node);
      const int next_axis = (f.axis + 1) % n;
      const bool left_first = s[f.axis] < fs[f.axis];
      if (f.stage == 0) {
        const double d = norm2(fs, s, n);
        if (d < dmin && ok(f.node)) { dmin = d; best = f.node; }
      }
      if (f.stage >= 2) { stack.pop_back(); continue; }
      const int stage = f.stage++;
      // stage 0 = the near side, stage 1 = the far side; dmin is read when the branch is taken, as the recursion does
      const bool go_left = (stage == 0) == left_first;
      int32_t child = -1;
      if (go_left) { if (s[f.axis] - dmin < fs[f.axis]) child = left[f.node]; }
      else { if (s[f.axis] + dmin >= fs[f.axis]) child = right[f.node]; }
      if (child >= 0) stack.push_back(Frame{child, next_axis, 0});
    }
    return best;
  }
  void radius(const double* s, double r, std::vector<int32_t>& out) const {     // :94-126, visit order kept
    out.clear();
    struct Frame { int32_t node; int axis; int stage; };
    std::vector<Frame> stack(1, Frame{0, 0, 0});
    while (!stack.empty()) {
      Frame& f = stack.back();
      const double* fs = st(f.node);
      const int next_axis = (f.axis + 1) % n;
      if (f.stage == 0 && norm2(fs, s, n) <= r) out.push_back(f.node);
      if (f.stage >= 2) { stack.pop_back(); continue; }
      const int stage = f.stage++;
      int32_t child = -1;
      if (stage == 0) { if (s[f.axis] - r <= fs[f.axis]) child = left[f.node]; }
      else { if (s[f.axis] + r >= fs[f.axis]) child = right[f.node]; }
      if (child >= 0) stack.push_back(Frame{child, next_axis, 0});
    }
  }
};

void check(int32_t rc, porrt_ctx* ctx, const char* what) {
  if (rc == PORRT_OK) return;
  const char* why = ctx ? porrt_last_error(ctx) : nullptr;
  fail(std::string(what) + ": " + (why && *why ? why : "error " + std::to_string(rc)));
}

struct Policy {                     // common.rs:23-39 over flat arrays, nodes in creation order
  std::vector<double> state;        // dim per node
  std::vector<int32_t> belief, parent, original;
  std::vector<std::vector<int32_t>> children;
  std::vector<int32_t> leafs;
  double expected_costs = 0.0;
  int dim = 0;
  int32_t add_node(const double* s, int32_t b, int32_t orig, bool leaf) {
    const int32_t id = (int32_t)belief.size();
    state.insert(state.end(), s, s + dim);
    belief.push_back(b); parent.push_back(-1); original.push_back(orig); children.emplace_back();
    if (leaf) leafs.push_back(id);
    return id;
  }
  void add_edge(int32_t p, int32_t c) { children[(size_t)p].push_back(c); parent[(size_t)c] = p; }
  size_t size() const { return belief.size(); }
};

void do_plan(CPlanningProblem* P, const double* start, size_t start_size) {
  const auto t_total = Clock::now();
  if (start_size != P->state_dim) fail("assertion failed: start_size == state_dim (pto_c.rs:229)");
  if (P->state_dim == 0 || P->state_dim > PORRT_MAX_STATE_DIM) fail("case not yet handled! (pto_c.rs:238; this library takes 1..16 dimensions)");
  const int N = (int)P->state_dim;
  const size_t nw = P->n_worlds;
  if (nw == 0) fail("n_worlds is 0: set_problem_dimensions first");
  // ---- PTOFuncsAdapter::new (pto_c.rs:326-373)
  if (P->low_size != (size_t)N || !P->low) fail("N:" + std::to_string(N) + ", low size:" + std::to_string(P->low_size) + " (pto_c.rs:328)");
  if (P->up_size != (size_t)N || !P->up) fail("N:" + std::to_string(N) + ", up size:" + std::to_string(P->up_size) + " (pto_c.rs:329)");
  if (!P->state_validity_callback || !P->transition_validity_callback || !P->observer_callback || !P->goal_callback || !P->goal_example_callback)
    fail("called `Option::unwrap()` on a `None` value: a callback is missing");
  if (!P->world_validities || P->world_validities_size == 0) fail("world validities are missing");
  if (!P->start_belief_state || P->start_belief_state_size != nw) fail("start belief state is missing");
  if (!P->reachable_belief_states || P->reachable_belief_states_size == 0) fail("reachable belief states are missing");
  for (int d = 0; d < N; ++d)
    if (!(P->low[d] < P->up[d]) || !std::isfinite(P->low[d]) || !std::isfinite(P->up[d])) fail("Uniform::new called with `low >= high` (rand, gen_range)");
  const int words = (int)((nw + 63) / 64);
  const size_t NV = P->world_validities_size, B = P->reachable_belief_states_size;
  if (B > 0x7fffffffu || NV > 0x7fffffffu || nw > 0x7fffffffu) fail("problem too large");
  std::vector<uint64_t> vmask(NV * words, 0);                     // validity_vec[i] > 0 (:347-349)
  for (size_t k = 0; k < NV; ++k)
    for (size_t w = 0; w < nw; ++w)
      if (P->world_validities[k][w] > 0) vmask[k * words + (w >> 6)] |= 1ull << (w & 63);
  std::vector<double> beliefs(B * nw);
  for (size_t b = 0; b < B; ++b) memcpy(&beliefs[b * nw], P->reachable_belief_states[b], nw * 8);
  auto is_compatible = [&](const double* bs, const uint64_t* mask) {  // common.rs:254-262
    for (size_t w = 0; w < nw; ++w) if (bs[w] > 0.0 && !bit(mask, w)) return false;
    return true;
  };
  std::vector<uint8_t> compat(B * NV);                            // compute_compatibility, common.rs:264-274
  for (size_t b = 0; b < B; ++b) for (size_t v = 0; v < NV; ++v) compat[b * NV + v] = is_compatible(&beliefs[b * nw], &vmask[v * words]);
  auto state_validity = [&](const double* s) -> int64_t {
    const int64_t v = P->state_validity_callback(s, (size_t)N);
    if (v >= (int64_t)NV) fail("index out of bounds: validity id " + std::to_string(v) + " of " + std::to_string(NV));
    return v;
  };
  auto transition_validity = [&](const double* a, const double* b) -> int64_t {
    const int64_t v = P->transition_validity_callback(a, (size_t)N, b, (size_t)N);
    if (v >= (int64_t)NV) fail("index out of bounds: validity id " + std::to_string(v) + " of " + std::to_string(NV));
    return v;
  };

  // ================================================================ PTO::grow_graph (pto.rs:55-139)
  const auto t_grow = Clock::now();
  Pcg64 continuous = make_rng(P), discrete = make_rng(P);         // two independent streams (pto_c.rs:213)
  std::vector<double> xs;                                         // node states
  std::vector<int32_t> node_vid;
  struct Edge { int32_t id, vid; };
  std::vector<std::vector<Edge>> children;                        // PTONode::children in add_edge order
  auto add_node = [&](const double* s, int32_t vid) { xs.insert(xs.end(), s, s + N); node_vid.push_back(vid); children.emplace_back(); return (int32_t)node_vid.size() - 1; };
  const int64_t root_vid = state_validity(start);
  if (root_vid < 0) fail("Start from a valid state! (pto.rs:61)");
  add_node(start, (int32_t)root_vid);
  porrt_reach* reach = nullptr;
  check(porrt_reach_create((int32_t)nw, &vmask[(size_t)root_vid * words], &reach), nullptr, "porrt_reach_create");
  struct ReachGuard { porrt_reach* r; ~ReachGuard() { porrt_reach_destroy(r); } } reach_guard{reach};
  KdTree kd{N, &xs, {}, {}};
  kd.reset();
  std::vector<uint64_t> reach_row((size_t)words), finality((size_t)words);
  std::vector<double> sample((size_t)N);
  std::vector<int32_t> neighbours;
  std::vector<Edge> edges;
  std::vector<uint8_t> goal_validity(nw);
  size_t it = 0;
  auto final_set_complete = [&]() { int32_t c = 0; check(porrt_reach_is_final_set_complete(reach, &c), nullptr, "porrt_reach_is_final_set_complete"); return c != 0; };
  while (it < P->n_iterations_min || (!final_set_complete() && it < P->n_iterations_max)) {
    ++it;
    const size_t world = (size_t)discrete.below(nw);               // sample(): world first, then the state (:141-149)
    if (it % 100 == 0) {
      std::fill(sample.begin(), sample.end(), 0.0);
      P->goal_example_callback(world, sample.data(), (size_t)N);
    } else {
      for (int d = 0; d < N; ++d) sample[(size_t)d] = continuous.range_f64(P->low[d], P->up[d]);
    }
    const int32_t kd_from = kd.nearest_filtered(sample.data(), [&](int32_t id) {
      porrt_reach_masks(reach, id, 1, reach_row.data());
      return bit(reach_row.data(), world);
    });
    {                                                              // steer, common.rs:215-225
      const double* from = &xs[(size_t)kd_from * N];
      const double step = norm1(from, sample.data(), N);
      if (step > P->max_step) {
        const double lambda = P->max_step / step;
        for (int d = 0; d < N; ++d) sample[(size_t)d] = from[d] + (sample[(size_t)d] - from[d]) * lambda;
      }
    }
    const int64_t svid = state_validity(sample.data());
    if (svid < 0) continue;
    const int32_t new_id = add_node(sample.data(), (int32_t)svid);
    check(porrt_reach_add_node(reach, &vmask[(size_t)svid * words]), nullptr, "porrt_reach_add_node");
    double radius = 0.0;
    check(porrt_heuristic_radius((int64_t)node_vid.size(), P->max_step, P->search_radius, N, &radius), nullptr, "porrt_heuristic_radius");
    kd.radius(sample.data(), radius, neighbours);
    if (neighbours.empty()) neighbours.push_back(kd_from);
    edges.clear();
    for (int32_t id : neighbours) {
      const int64_t tv = transition_validity(&xs[(size_t)id * N], &xs[(size_t)new_id * N]);   // neighbour -> new node
      if (tv >= 0) edges.push_back(Edge{id, (int32_t)tv});
    }
    for (const Edge& e : edges) {                                  // neighbours -> new node
      check(porrt_reach_add_edge(reach, e.id, new_id, &vmask[(size_t)e.vid * words]), nullptr, "porrt_reach_add_edge");
      children[(size_t)e.id].push_back(Edge{new_id, e.vid});
    }
    for (const Edge& e : edges) {                                  // new node -> neighbours
      check(porrt_reach_add_edge(reach, new_id, e.id, &vmask[(size_t)e.vid * words]), nullptr, "porrt_reach_add_edge");
      children[(size_t)new_id].push_back(Edge{e.id, e.vid});
    }
    {
      bool flags[4096];
      std::vector<char> big;
      bool* gv = flags;
      if (nw > 4096) { big.assign(nw * sizeof(bool), 0); gv = (bool*)big.data(); }
      for (size_t w = 0; w < nw; ++w) gv[w] = false;
      if (P->goal_callback(&xs[(size_t)new_id * N], (size_t)N, gv, nw)) {
        std::fill(finality.begin(), finality.end(), 0ull);
        for (size_t w = 0; w < nw; ++w) if (gv[w]) finality[w >> 6] |= 1ull << (w & 63);
        check(porrt_reach_add_final_node(reach, new_id, finality.data()), nullptr, "porrt_reach_add_final_node");
      }
    }
    kd.add(new_id);
  }
  P->n_iterations = it;
  P->graph_growth_s = seconds_since(t_grow);
  if (!final_set_complete()) fail("graph not grown up to solution: final nodes are not reached for each world (pto_c.rs:214)");
  const size_t V = node_vid.size();
  P->n_nodes = V;

  // ================================================================ PTO::plan_belief_space (pto.rs:152-182)
  {                                                                // assert_belief_state_validity, common.rs:276-279
    double s = 0.0;
    for (size_t w = 0; w < nw; ++w) s = P->start_belief_state[w] + s;
    if (!(std::fabs(s - 1.0) < 0.000001)) fail("assertion failed: start belief state does not sum to 1 (common.rs:278)");
  }
  const auto t_expand = Clock::now();
  if (V * B > 0x7fffffffull) fail("more than 2^31 belief nodes");
  // ---- build_belief_graph (:185-259): belief node id = node * B + belief; it exists iff the belief is compatible with the node
  std::unordered_map<uint64_t, int32_t> hash_to_id;               // create_belief_states_hash_map, belief_graph.rs:75-87
  std::vector<uint64_t> bhash(B);
  for (size_t b = 0; b < B; ++b) { bhash[b] = belief_hash(&beliefs[b * nw], nw); hash_to_id[bhash[b]] = (int32_t)b; }
  if (hash_to_id.size() != B) fail("collision when hashing the belief states! (belief_graph.rs:84)");
  auto exists = [&](size_t node, size_t b) { return compat[b * NV + (size_t)node_vid[node]] != 0; };
  const size_t VB = V * B;
  std::vector<uint8_t> type(VB, PORRT_NODE_UNKNOWN);
  std::vector<int32_t> obs_parent, obs_child;                     // observation edges in add_edge order (parents ascending)
  for (size_t id = 0; id < V; ++id)
    for (size_t b = 0; b < B; ++b) {
      size_t** ids_pp = nullptr;
      size_t n_succ = 0;
      P->observer_callback(&xs[id * N], (size_t)N, &beliefs[b * nw], nw, &ids_pp, &n_succ);
      size_t* ids = (ids_pp && n_succ) ? *ids_pp : nullptr;
      if (n_succ && !ids) fail("observer callback returned a null id array");
      for (size_t k = 0; k < n_succ; ++k) {
        if (ids[k] >= B) fail("index out of bounds: observer returned belief id " + std::to_string(ids[k]));
        const uint64_t hc = bhash[ids[k]];
        if (bhash[b] == hc) continue;
        const size_t cb = (size_t)hash_to_id[hc];                  // belief_graph.belief_id(&child)
        if (exists(id, b) && exists(id, cb)) {
          type[id * B + b] = PORRT_NODE_OBSERVATION;
          obs_parent.push_back((int32_t)(id * B + b)); obs_child.push_back((int32_t)(id * B + cb));
        }
      }
      if (ids_pp && *ids_pp && P->free_observer_arrays) free(*ids_pp);
    }
  // action edges (:236-255) and the children CSR: a belief node has observation edges or action edges, never both
  std::vector<int64_t> row_ptr(VB + 1, 0);
  for (int32_t p : obs_parent) ++row_ptr[(size_t)p + 1];
  for (size_t id = 0; id < V; ++id)
    for (size_t b = 0; b < B; ++b) {
      if (!exists(id, b) || type[id * B + b] == PORRT_NODE_OBSERVATION) continue;
      int64_t cnt = 0;
      for (const Edge& e : children[id])
        if (exists((size_t)e.id, b) && compat[b * NV + (size_t)e.vid]) ++cnt;
      if (cnt) { type[id * B + b] = PORRT_NODE_ACTION; row_ptr[id * B + b + 1] = cnt; }
    }
  for (size_t k = 0; k < VB; ++k) row_ptr[k + 1] += row_ptr[k];
  const int64_t E = row_ptr[VB];
  if (E > 0x7fffffffll * 4) fail("belief graph too large");
  std::vector<int32_t> col((size_t)std::max<int64_t>(E, 1));
  {
    std::vector<int64_t> fill(row_ptr.begin(), row_ptr.end() - 1);
    for (size_t k = 0; k < obs_parent.size(); ++k) col[(size_t)fill[(size_t)obs_parent[k]]++] = obs_child[k];
    for (size_t id = 0; id < V; ++id)
      for (size_t b = 0; b < B; ++b) {
        if (type[id * B + b] != PORRT_NODE_ACTION) continue;
        int64_t& at = fill[id * B + b];
        for (const Edge& e : children[id])
          if (exists((size_t)e.id, b) && compat[b * NV + (size_t)e.vid]) col[(size_t)at++] = (int32_t)((size_t)e.id * B + b);
      }
  }
  std::vector<double> bxy(VB * N);
  std::vector<int32_t> bid(VB);
  for (size_t id = 0; id < V; ++id)
    for (size_t b = 0; b < B; ++b) { memcpy(&bxy[(id * B + b) * N], &xs[id * N], (size_t)N * 8); bid[id * B + b] = (int32_t)b; }
  P->n_belief_nodes = VB; P->n_belief_edges = (size_t)E;
  P->belief_space_expansion_s = seconds_since(t_expand);
  // ---- compute_expected_costs_to_goals (:261-275) + extract_policy (:277-283) on the device
  const auto t_dp = Clock::now();
  std::vector<int32_t> finals;
  {
    int64_t nf = 0;
    porrt_reach_finals(reach, nullptr, nullptr, 0, &nf);
    std::vector<int64_t> fids((size_t)std::max<int64_t>(nf, 1));
    std::vector<uint64_t> fmasks((size_t)std::max<int64_t>(nf, 1) * words);
    check(porrt_reach_finals(reach, fids.data(), fmasks.data(), nf, &nf), nullptr, "porrt_reach_finals");
    for (int64_t k = 0; k < nf; ++k)
      for (size_t b = 0; b < B; ++b)
        if (exists((size_t)fids[(size_t)k], b) && is_compatible(&beliefs[b * nw], &fmasks[(size_t)k * words]))
          finals.push_back((int32_t)((size_t)fids[(size_t)k] * B + b));
  }
  if (!P->ctx || P->ctx_device != P->device) {
    if (P->ctx) { porrt_ctx_destroy(P->ctx); P->ctx = nullptr; }
    porrt_ctx* c = nullptr;
    const int32_t rc = porrt_ctx_create(P->device, &c);
    if (rc != PORRT_OK) fail("no sm_100 device for the value backups (porrt_ctx_create failed with " + std::to_string(rc) + "): there is no CPU fallback");
    P->ctx = c; P->ctx_device = P->device;
  }
  std::vector<double> dist(VB);
  int32_t sweeps = 0;
  check(porrt_conditional_dijkstra_nd(P->ctx, N, (int64_t)VB, row_ptr.data(), col.data(), bxy.data(), type.data(), bid.data(), beliefs.data(),
                                      (int32_t)B, (int32_t)nw, finals.data(), (int32_t)finals.size(), dist.data(), &sweeps),
        P->ctx, "conditional_dijkstra");
  P->n_sweeps = (size_t)sweeps;
  std::vector<int32_t> pol_bn, pol_parent;
  std::vector<uint8_t> pol_leaf;
  int64_t n_pol = 0;
  double expected = 0.0;
  {
    int64_t cap = 4096;
    for (;;) {
      pol_bn.resize((size_t)cap); pol_parent.resize((size_t)cap); pol_leaf.resize((size_t)cap);
      const int32_t rc = porrt_extract_policy_graph_nd(P->ctx, N, (int64_t)VB, row_ptr.data(), col.data(), bxy.data(), type.data(), bid.data(),
                                                       beliefs.data(), (int32_t)B, (int32_t)nw, dist.data(), pol_bn.data(), pol_parent.data(),
                                                       pol_leaf.data(), cap, &n_pol, &expected);
      if (rc == PORRT_ERR_CAPACITY && n_pol > cap) { cap = n_pol; continue; }
      check(rc, P->ctx, "extract_policy");
      break;
    }
  }
  Policy policy;
  policy.dim = N;
  for (int64_t k = 0; k < n_pol; ++k) {
    const int32_t bn = pol_bn[(size_t)k];
    policy.add_node(&bxy[(size_t)bn * N], bid[(size_t)bn], bn, pol_leaf[(size_t)k] != 0);
    if (pol_parent[(size_t)k] >= 0) policy.add_edge(pol_parent[(size_t)k], (int32_t)k);
  }
  policy.expected_costs = expected;
  P->dynamic_programming_s = seconds_since(t_dp);

  // ================================================================ refine_solution(PartialShortCut(n)) (pto_policy_refiner.rs:85-133)
  const auto t_refine = Clock::now();
  // ---- Policy::decompose (common.rs:85-129)
  std::vector<std::vector<int32_t>> pieces, skeleton;
  {
    std::vector<int32_t> fifo(1, 0);
    int32_t n_pieces = 0;
    for (size_t head = 0; head < fifo.size(); ++head) {
      const int32_t id = fifo[head];
      std::vector<int32_t> ids, successors;
      int32_t cur = id;
      for (;;) {
        if (policy.belief[(size_t)id] != policy.belief[(size_t)cur]) fail("assertion failed: a path piece changes belief state (common.rs:100)");
        ids.push_back(cur);
        const std::vector<int32_t>& ch = policy.children[(size_t)cur];
        if (ch.empty()) break;
        if (ch.size() == 1) { cur = ch[0]; continue; }
        for (int32_t c : ch) { fifo.push_back(c); successors.push_back(++n_pieces); }
        break;
      }
      pieces.push_back(ids);
      skeleton.push_back(successors);
    }
  }
  // ---- build_path_piece (:135-156) + partial_shortcut (:158-206) per piece
  struct Piece { std::vector<double> state; std::vector<int32_t> belief_graph_id; int32_t belief_state_id; };
  std::vector<Piece> trees;
  auto is_transition_valid = [&](const double* from, const double* to, int32_t belief_id) {   // :395-423
    const int64_t fv = state_validity(from);
    const int64_t tv = state_validity(to);
    if (fv < 0 || tv < 0) return false;
    const int64_t v = transition_validity(from, to);
    return v >= 0 && compat[(size_t)belief_id * NV + (size_t)v] != 0;
  };
  for (const std::vector<int32_t>& path : pieces) {
    Piece t;
    for (int32_t pid : path) {
      const int32_t bn = policy.original[(size_t)pid];
      t.state.insert(t.state.end(), &bxy[(size_t)bn * N], &bxy[(size_t)bn * N] + N);
      t.belief_graph_id.push_back(bn);
    }
    t.belief_state_id = bid[(size_t)t.belief_graph_id[0]];
    const size_t len = t.belief_graph_id.size();
    if (len > 2) {
      Pcg64 sampler = Pcg64::seed_from_u64(0);                      // DiscreteSampler::new(), one per piece (:169)
      std::vector<double> shortcut;
      for (size_t trial = 0; trial < P->refine_iterations; ++trial) {
        const size_t joint = (size_t)sampler.below((uint64_t)N);
        const size_t i0 = (size_t)sampler.below(len - 2);
        const size_t i1 = i0 + 2 + (size_t)sampler.below(len - i0 - 2);
        const double* s0 = &t.state[i0 * N];
        const double* s1 = &t.state[i1 * N];
        shortcut.assign(t.state.begin() + (ptrdiff_t)(i0 * N), t.state.begin() + (ptrdiff_t)(i1 * N));
        for (size_t j = i0; j < i1; ++j) {
          const double lambda = (double)(j - i0) / (double)(i1 - i0);
          shortcut[(j - i0) * N + joint] = s0[joint] * (1.0 - lambda) + s1[joint] * lambda;
        }
        bool commit = true;
        for (size_t j = 0; j + 1 < i1 - i0; ++j) commit = commit && is_transition_valid(&shortcut[j * N], &shortcut[(j + 1) * N], t.belief_state_id);
        commit = commit && is_transition_valid(&shortcut[(i1 - i0 - 1) * N], s1, t.belief_state_id);
        if (commit) memcpy(&t.state[i0 * N], shortcut.data(), shortcut.size() * 8);
      }
    }
    trees.push_back(std::move(t));
  }
  // ---- recompose (:324-393)
  Policy refined;
  refined.dim = N;
  std::vector<int32_t> piece_start(trees.size(), -1), piece_end(trees.size(), -1);
  for (size_t i = 0; i < trees.size(); ++i) {
    const size_t len = trees[i].belief_graph_id.size();
    int32_t previous = -1;
    for (size_t j = 0; j < len; ++j) {
      const int32_t bn = trees[i].belief_graph_id[j];
      const int32_t id = refined.add_node(&trees[i].state[j * N], bid[(size_t)bn], bn, false);
      if (j == 0) piece_start[i] = id;
      else { refined.add_edge(previous, id); if (j == len - 1) piece_end[i] = id; }
      previous = id;
    }
  }
  for (size_t i = 0; i < skeleton.size(); ++i)
    for (int32_t nxt : skeleton[i])
      if (piece_end[i] >= 0 && piece_start[(size_t)nxt] >= 0) refined.add_edge(piece_end[i], piece_start[(size_t)nxt]);
  for (size_t k = 0; k < refined.size(); ++k) if (refined.children[k].empty()) refined.leafs.push_back((int32_t)k);
  {                                                                // compute_expected_costs_to_goals, common.rs:131-153 (recursion as a stack)
    struct Frame { int32_t id; double p; size_t next; double acc; double pq; double cost; };
    std::vector<Frame> stack(1, Frame{0, 1.0, 0, 0.0, 0.0, 0.0});
    double result = 0.0;
    while (!stack.empty()) {
      Frame& f = stack.back();
      const std::vector<int32_t>& ch = refined.children[(size_t)f.id];
      if (f.next < ch.size()) {
        const int32_t c = ch[f.next++];
        double q = 0.0;                                            // transition_probability(node, child), common.rs:188-190
        const double* pb = &beliefs[(size_t)refined.belief[(size_t)f.id] * nw];
        const double* cb = &beliefs[(size_t)refined.belief[(size_t)c] * nw];
        for (size_t w = 0; w < nw; ++w) q = q + (cb[w] > 0.0 ? pb[w] : 0.0);
        const double cost = norm2(&refined.state[(size_t)f.id * N], &refined.state[(size_t)c * N], N);
        const double pq = f.p * q;
        stack.push_back(Frame{c, pq, 0, 0.0, pq, cost});
      } else {
        const Frame done = f;
        stack.pop_back();
        if (stack.empty()) result = done.acc;
        else stack.back().acc += done.pq * done.cost + done.acc;   // p * q * cost + recursion
      }
    }
    refined.expected_costs = result;
  }
  P->refinement_s = seconds_since(t_refine);
  P->n_policy_nodes = refined.size();
  // ---- save_planning_metrics / save_paths (pto_c.rs:272-312)
  P->total_s = seconds_since(t_total);
  for (int32_t leaf : refined.leafs) {
    std::vector<std::vector<double>> path;
    for (int32_t cur = leaf; cur >= 0; cur = refined.parent[(size_t)cur])
      path.emplace_back(&refined.state[(size_t)cur * N], &refined.state[(size_t)cur * N] + N);
    std::vector<std::vector<double>> rev(path.rbegin(), path.rend());
    P->paths_lengths.push_back(rev.size());
    P->paths.push_back(std::move(rev));
  }
  P->expected_costs = refined.expected_costs;
}

}  // namespace

PTOC_API CPlanningProblem* new_planning_problem(void) { return new CPlanningProblem(); }
PTOC_API void delete_planning_problem(CPlanningProblem* p) {
  if (!p) return;
  if (p->ctx) porrt_ctx_destroy(p->ctx);
  delete p;
}
PTOC_API void set_problem_dimensions(CPlanningProblem* p, size_t state_dim, size_t n_worlds) { p->state_dim = state_dim; p->n_worlds = n_worlds; }
PTOC_API void set_lower_sampling_bound(CPlanningProblem* p, double* low, size_t n) { p->low = low; p->low_size = n; }
PTOC_API void set_upper_sampling_bound(CPlanningProblem* p, double* up, size_t n) { p->up = up; p->up_size = n; }
PTOC_API void set_world_validities(CPlanningProblem* p, size_t** v, size_t n) { p->world_validities = v; p->world_validities_size = n; }
PTOC_API void set_state_validity_callback(CPlanningProblem* p, StateValidityCallbackType cb) { p->state_validity_callback = cb; }
PTOC_API void set_transition_validity_callback(CPlanningProblem* p, TransitionValidityCallbackType cb) { p->transition_validity_callback = cb; }
PTOC_API void set_cost_evaluator_callback(CPlanningProblem* p, CostEvaluatorCallbackType cb) { p->cost_evaluator_callback = cb; }
PTOC_API void set_observer_callback(CPlanningProblem* p, ObserverCallbackType cb) { p->observer_callback = cb; }
PTOC_API void set_start_belief_state(CPlanningProblem* p, double* start, size_t n, double** reachable, size_t n_reachable) {
  p->start_belief_state = start; p->start_belief_state_size = n;
  p->reachable_belief_states = reachable; p->reachable_belief_states_size = n_reachable;
}
PTOC_API void set_goal_callback(CPlanningProblem* p, GoalCallbackType cb) { p->goal_callback = cb; }
PTOC_API void set_goal_example_callback(CPlanningProblem* p, GoalExampleCallbackType cb) { p->goal_example_callback = cb; }
PTOC_API void set_search_parameters(CPlanningProblem* p, size_t n_min, size_t n_max, double max_step, double search_radius) {
  p->n_iterations_min = n_min; p->n_iterations_max = n_max; p->max_step = max_step; p->search_radius = search_radius;
}
PTOC_API void set_refine_parameters(CPlanningProblem* p, size_t refine_iterations) { p->refine_iterations = refine_iterations; }

PTOC_API void plan(CPlanningProblem* p, double* start, size_t start_size) {
  p->paths.clear(); p->paths_lengths.clear();
  p->expected_costs = 0.0; p->n_iterations = 0;
  p->graph_growth_s = p->belief_space_expansion_s = p->dynamic_programming_s = p->refinement_s = p->total_s = 0.0;
  p->n_nodes = p->n_belief_nodes = p->n_belief_edges = p->n_sweeps = p->n_policy_nodes = 0;
  p->failed = false; p->error.clear();
  try {
    do_plan(p, start, start_size);
  } catch (const PlanError& e) {
    p->failed = true; p->error = e.msg;
    p->paths.clear(); p->paths_lengths.clear();
  } catch (const std::exception& e) {
    p->failed = true; p->error = e.what();
    p->paths.clear(); p->paths_lengths.clear();
  }
}

PTOC_API void get_planning_metrics(CPlanningProblem* p, size_t* n_iterations, double* graph_growth_s, double* belief_space_expansion_s,
                                   double* dynamic_programming_s, double* refinement_s, double* total_s) {
  *n_iterations = p->n_iterations;
  *graph_growth_s = p->graph_growth_s;
  *belief_space_expansion_s = p->belief_space_expansion_s;
  *dynamic_programming_s = p->dynamic_programming_s;
  *refinement_s = p->refinement_s;
  *total_s = p->total_s;
}
PTOC_API void get_paths_info(CPlanningProblem* p, size_t* number_of_paths, size_t** path_lengths, double* expected_cost) {
  *number_of_paths = p->paths.size();
  *path_lengths = p->paths_lengths.data();
  *expected_cost = p->expected_costs;
}
PTOC_API void get_paths_variable(CPlanningProblem* p, size_t path_id, size_t state_id, double** c_state, size_t* state_size) {
  *c_state = p->paths[path_id][state_id].data();
  *state_size = p->paths[path_id][state_id].size();
}

PTOC_API void set_sampler_seed(CPlanningProblem* p, uint64_t seed) { p->seeded = true; p->seed = seed; }
PTOC_API void set_planning_device(CPlanningProblem* p, int32_t device) { p->device = device; }
PTOC_API void set_observer_array_ownership(CPlanningProblem* p, int32_t library_frees) { p->free_observer_arrays = library_frees != 0; }
PTOC_API const char* get_planning_error(CPlanningProblem* p) { return p->failed ? p->error.c_str() : nullptr; }
PTOC_API void get_planning_sizes(CPlanningProblem* p, size_t* n_nodes, size_t* n_belief_nodes, size_t* n_belief_edges, size_t* n_sweeps,
                                 size_t* n_policy_nodes) {
  if (n_nodes) *n_nodes = p->n_nodes;
  if (n_belief_nodes) *n_belief_nodes = p->n_belief_nodes;
  if (n_belief_edges) *n_belief_edges = p->n_belief_edges;
  if (n_sweeps) *n_sweeps = p->n_sweeps;
  if (n_policy_nodes) *n_policy_nodes = p->n_policy_nodes;
}
