/* pto_c_client.c -- a C program written against the reference's exported planner API (src/pto_c.rs:63-270; include/po_rrt_c.h),
 * the way the reference's C++ callers use it: the world model lives HERE (callbacks), the planner is a black box.  Built with
 * gcc -std=c99 and linked against po_rrt_b200/libpo_rrt_c.so by tests/test_pto_c.py::test_c_client.
 *
 * World: states (x, y, z) in [-1, 1]^3 (z is free), a wall at x in [-0.05, 0.05] with two gaps --
 *   a DOOR at y in [-0.15, 0.15]: open in world 0, closed in world 1 (validity id 1 = {1, 0}),
 *   a WINDOW at y in [0.75, 0.95]: always open (the detour).
 * The door's state is observed within 0.35 (L2, in the plane) of the door's centre, from the start side.
 * Start (-0.7, 0, 0), goal: x > 0.6 and |y| < 0.25, in both worlds.  The plan must branch: through the door if it is open,
 * around through the window if not.  Exit code 0 and "pto_c_client: ok" when every check holds. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "po_rrt_c.h"

#define DIM 3
#define N_WORLDS 2

static long n_state_calls = 0, n_transition_calls = 0, n_observe_calls = 0;

static int in_wall(double x) { return x >= -0.05 && x <= 0.05; }
static int in_door(double y) { return y >= -0.15 && y <= 0.15; }
static int in_window(double y) { return y >= 0.75 && y <= 0.95; }

/* >= 0: validity id (0 = both worlds, 1 = world 0 only), -1: invalid */
static int64_t point_validity(double x, double y) {
  if (fabs(x) > 1.0 || fabs(y) > 1.0) return -1;
  if (!in_wall(x)) return 0;
  if (in_door(y)) return 1;
  if (in_window(y)) return 0;
  return -1;
}
static int64_t state_validity(const double* s, size_t n) {
  (void)n;
  ++n_state_calls;
  return point_validity(s[0], s[1]);
}

static int64_t transition_validity(const double* a, size_t na, const double* b, size_t nb) {
  (void)na; (void)nb;
  ++n_transition_calls;
  int64_t worst = 0;
  const double len = fabs(b[0] - a[0]) + fabs(b[1] - a[1]);
  const int steps = (int)(len / 0.004) + 1;
  for (int k = 0; k <= steps; ++k) {
    const double t = (double)k / (double)steps;
    const int64_t v = point_validity(a[0] + (b[0] - a[0]) * t, a[1] + (b[1] - a[1]) * t);
    if (v < 0) return -1;
    if (v > worst) worst = v;
  }
  return worst;
}

/* reachable belief states: 0 = (0.5, 0.5), 1 = (1, 0), 2 = (0, 1).  The planner frees the id array (pto_c.rs:411). */
static size_t* obs_slot = NULL;
static void observe(const double* s, size_t ns, const double* belief, size_t nb, size_t*** out_ids, size_t* out_n) {
  (void)ns; (void)nb;
  ++n_observe_calls;
  const int sees_door = s[0] < -0.05 && sqrt((s[0] + 0.05) * (s[0] + 0.05) + s[1] * s[1]) < 0.35;
  size_t* ids = (size_t*)malloc(2 * sizeof(size_t));
  size_t n = 0;
  if (sees_door && belief[0] > 0.0 && belief[1] > 0.0) { ids[n++] = 1; ids[n++] = 2; }
  else ids[n++] = belief[0] > 0.0 ? (belief[1] > 0.0 ? 0 : 1) : 2;
  obs_slot = ids;
  *out_ids = &obs_slot;
  *out_n = n;
}

static bool goal(const double* s, size_t n, bool* validity, size_t nw) {
  (void)n;
  if (s[0] > 0.6 && fabs(s[1]) < 0.25) { for (size_t w = 0; w < nw; ++w) validity[w] = true; return true; }
  return false;
}
static void goal_example(size_t world, double* s, size_t n) { (void)world; (void)n; s[0] = 0.8; s[1] = 0.0; s[2] = 0.0; }

#define CHECK(c, msg) do { if (!(c)) { printf("pto_c_client: FAILED: %s (line %d)\n", msg, __LINE__); return 1; } } while (0)

int main(void) {
  CPlanningProblem* p = new_planning_problem();
  double low[DIM] = {-1, -1, -1}, up[DIM] = {1, 1, 1}, start[DIM] = {-0.7, 0.0, 0.0};
  size_t v0[N_WORLDS] = {1, 1}, v1[N_WORLDS] = {1, 0};
  size_t* validities[2] = {v0, v1};
  double b0[N_WORLDS] = {0.5, 0.5}, b1[N_WORLDS] = {1, 0}, b2[N_WORLDS] = {0, 1};
  double* reachable[3] = {b0, b1, b2};
  set_problem_dimensions(p, DIM, N_WORLDS);
  set_lower_sampling_bound(p, low, DIM);
  set_upper_sampling_bound(p, up, DIM);
  set_world_validities(p, validities, 2);
  set_state_validity_callback(p, state_validity);
  set_transition_validity_callback(p, transition_validity);
  set_observer_callback(p, observe);
  set_goal_callback(p, goal);
  set_goal_example_callback(p, goal_example);
  set_start_belief_state(p, b0, N_WORLDS, reachable, 3);
  set_search_parameters(p, 4000, 200000, 0.3, 5.0);
  set_refine_parameters(p, 300);
  set_sampler_seed(p, 7);           /* addition: the reference draws its seeds from the OS */
  plan(p, start, DIM);
  if (get_planning_error(p)) { printf("pto_c_client: plan failed: %s\n", get_planning_error(p)); return 2; }

  size_t n_it = 0, n_paths = 0, *lengths = NULL, n_nodes = 0, n_bn = 0, n_be = 0, sweeps = 0, n_pol = 0;
  double t_grow, t_expand, t_dp, t_refine, t_total, cost = 0.0;
  get_planning_metrics(p, &n_it, &t_grow, &t_expand, &t_dp, &t_refine, &t_total);
  get_paths_info(p, &n_paths, &lengths, &cost);
  get_planning_sizes(p, &n_nodes, &n_bn, &n_be, &sweeps, &n_pol);
  printf("iterations %zu, roadmap %zu nodes, belief graph %zu nodes / %zu edges, %zu device sweeps, policy %zu nodes, %zu paths, expected cost %.6f\n",
         n_it, n_nodes, n_bn, n_be, sweeps, n_pol, n_paths, cost);
  printf("callbacks: %ld state, %ld transition, %ld observe; seconds: growth %.3f, belief graph %.3f, backups + policy %.3f, refinement %.3f\n",
         n_state_calls, n_transition_calls, n_observe_calls, t_grow, t_expand, t_dp, t_refine);
  CHECK(n_it >= 4000 && n_bn == 3 * n_nodes && sweeps > 0, "sizes");
  CHECK(n_observe_calls == (long)n_bn, "one observer call per (node, belief)");
  CHECK(n_paths == 2, "the policy branches on the door: one path per outcome");
  int through_door = 0, through_window = 0;
  double shortest = 1e9, longest = 0.0;
  for (size_t k = 0; k < n_paths; ++k) {
    double len = 0.0, *prev = NULL;
    for (size_t j = 0; j < lengths[k]; ++j) {
      double* s = NULL; size_t sz = 0;
      get_paths_variable(p, k, j, &s, &sz);
      CHECK(sz == DIM, "state size");
      if (j == 0) CHECK(memcmp(s, start, sizeof(start)) == 0, "paths start at the start state");
      if (prev) {
        CHECK(transition_validity(prev, DIM, s, DIM) >= 0, "every step is a transition the caller accepts");
        double d2 = 0.0;
        for (int d = 0; d < DIM; ++d) d2 += (s[d] - prev[d]) * (s[d] - prev[d]);
        len += sqrt(d2);
        if (in_wall(s[0]) || (prev[0] < -0.05 && s[0] > 0.05)) {          /* the step that crosses the wall */
          const double yc = prev[1] + (s[1] - prev[1]) * ((0.0 - prev[0]) / (s[0] - prev[0] + 1e-300));
          if (in_door(in_wall(s[0]) ? s[1] : yc)) through_door |= 1 << k;
          if (in_window(in_wall(s[0]) ? s[1] : yc)) through_window |= 1 << k;
        }
      }
      prev = s;
      if (j + 1 == lengths[k]) { bool gv[N_WORLDS]; CHECK(goal(s, DIM, gv, N_WORLDS), "paths end in the goal"); }
    }
    if (len < shortest) shortest = len;
    if (len > longest) longest = len;
  }
  CHECK(through_door != 0 && through_window != 0 && through_door != through_window, "one path through the door, the other through the window");
  CHECK(cost >= shortest - 1e-9 && cost <= longest + 1e-9, "the expected cost is a mean of the paths' lengths");
  CHECK(shortest >= 1.3, "no path is shorter than the straight line start -> goal");
  delete_planning_problem(p);
  printf("pto_c_client: ok\n");
  return 0;
}
