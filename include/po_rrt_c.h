/* po_rrt_c.h -- the reference's exported C planner API (src/pto_c.rs), same names, same argument lists, implemented by
 * po_rrt_b200/libpo_rrt_c.so (source po_rrt_b200/csrc/pto_c.cpp), which is itself a CLIENT of the hot-path ABI in
 * porrt_b200.h -- exactly as pto_c.rs is a client of the crate's planner.  A C/C++ program written against the reference's
 * dylib (Cargo.toml: crate-type = ["rlib", "dylib"]) relinks against this library and keeps its callbacks.
 *
 * What runs where: roadmap growth (PTO::grow_graph, pto.rs:55-139) is sequential and made of the CALLER's callbacks, so it
 * stays on the host; PTO::plan_belief_space (pto.rs:152-182) -- build_belief_graph with the observer callback, then the value
 * backups of conditional_dijkstra over the whole belief graph -- hands the graph to the device
 * (porrt_conditional_dijkstra_nd / porrt_extract_policy_graph_nd); refine_solution(PartialShortCut(n))
 * (pto_policy_refiner.rs:85-206) is again made of callbacks and stays on the host.
 *
 * Differences from the reference, all deliberate:
 *  - the reference takes OWNERSHIP of every array it is handed (Vec::from_raw_parts, pto_c.rs:231,318-361) and frees it with the
 *    system allocator when plan() returns.  This library only BORROWS low / up / world validities / belief states / start; the
 *    id array an observer callback returns is free()d after use, as the reference does (pto_c.rs:411), unless
 *    set_observer_array_ownership(problem, 0) says otherwise.
 *  - where the reference panics (and aborts the process: a panic cannot unwind out of extern "C"), plan() records the message:
 *    get_planning_error() returns it (NULL = success) and the outputs stay empty.
 *  - any state dimension 1..16 is accepted (the reference: 2, 3, 7, 9 -- "case not yet handled!" otherwise, pto_c.rs:238).
 *  - the reference seeds both samplers from the OS (new_true_random, pto_c.rs:213); set_sampler_seed(problem, s) makes a run
 *    reproducible: both streams = Pcg64::seed_from_u64(s) like ContinuousSampler::new / DiscreteSampler::new (sample_space.rs:18,47).
 */
#ifndef PO_RRT_C_H
#define PO_RRT_C_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pto_c.rs:17-24 (usize = size_t, i64 = int64_t) */
typedef int64_t (*StateValidityCallbackType)(const double* state, size_t state_size);
typedef int64_t (*TransitionValidityCallbackType)(const double* from, size_t from_size, const double* to, size_t to_size);
typedef double (*CostEvaluatorCallbackType)(const double* from, size_t from_size, const double* to, size_t to_size);
/* *belief_ids = address of a pointer to a malloc()ed array of *n_successors ids into the reachable belief states (pto_c.rs:403-414) */
typedef void (*ObserverCallbackType)(const double* state, size_t state_size, const double* belief_state, size_t belief_size,
                                     size_t*** belief_ids, size_t* n_successors);
typedef bool (*GoalCallbackType)(const double* state, size_t state_size, bool* world_validity, size_t n_worlds);
typedef void (*GoalExampleCallbackType)(size_t world, double* state, size_t state_size);

typedef struct CPlanningProblem CPlanningProblem; /* opaque (pto_c.rs:29-62) */

CPlanningProblem* new_planning_problem(void);                                                          /* pto_c.rs:64-101 */
void delete_planning_problem(CPlanningProblem* problem);                                               /* :103-104 */
void set_problem_dimensions(CPlanningProblem* problem, size_t state_dim, size_t n_worlds);             /* :106-112 */
void set_lower_sampling_bound(CPlanningProblem* problem, double* low, size_t low_size);                /* :114-121 */
void set_upper_sampling_bound(CPlanningProblem* problem, double* up, size_t up_size);                  /* :123-130 */
/* validities[k][w] > 0 iff validity k holds in world w; validities_size rows of n_worlds entries (:132-137, 339-355) */
void set_world_validities(CPlanningProblem* problem, size_t** validities, size_t validities_size);
void set_state_validity_callback(CPlanningProblem* problem, StateValidityCallbackType callback);       /* :139-144 */
void set_transition_validity_callback(CPlanningProblem* problem, TransitionValidityCallbackType callback); /* :146-151 */
void set_cost_evaluator_callback(CPlanningProblem* problem, CostEvaluatorCallbackType callback);       /* :153-158; stored, never called (as in the reference: PTOFuncsAdapter keeps the default norm2) */
void set_observer_callback(CPlanningProblem* problem, ObserverCallbackType callback);                  /* :161-166 */
void set_start_belief_state(CPlanningProblem* problem, double* start_belief_state, size_t start_belief_state_size,
                            double** reachable_belief_states, size_t reachable_belief_states_size);    /* :168-176 */
void set_goal_callback(CPlanningProblem* problem, GoalCallbackType callback);                          /* :178-183 */
void set_goal_example_callback(CPlanningProblem* problem, GoalExampleCallbackType callback);           /* :185-190 */
void set_search_parameters(CPlanningProblem* problem, size_t n_iterations_min, size_t n_iterations_max, double max_step,
                           double search_radius);                                                      /* :192-200 */
void set_refine_parameters(CPlanningProblem* problem, size_t refine_iterations);                       /* :202-207 */
/* grow_graph -> plan_belief_space -> refine_solution(PartialShortCut(refine_iterations)) -> paths (:209-241) */
void plan(CPlanningProblem* problem, double* start, size_t start_size);
void get_planning_metrics(CPlanningProblem* problem, size_t* n_iterations, double* graph_growth_s, double* belief_space_expansion_s,
                          double* dynamic_programming_s, double* refinement_s, double* total_s);       /* :243-253 */
void get_paths_info(CPlanningProblem* problem, size_t* number_of_paths, size_t** path_lengths, double* expected_cost); /* :255-262 */
void get_paths_variable(CPlanningProblem* problem, size_t path_id, size_t state_id, double** c_state, size_t* state_size); /* :264-270 */

/* ---- additions (not in the reference) */
void set_sampler_seed(CPlanningProblem* problem, uint64_t seed);          /* reproducible sampler streams */
void set_planning_device(CPlanningProblem* problem, int32_t cuda_device); /* default 0 */
void set_observer_array_ownership(CPlanningProblem* problem, int32_t library_frees);
const char* get_planning_error(CPlanningProblem* problem);                /* NULL after a successful plan() */
/* sizes of what the last plan() built: roadmap nodes, belief nodes, belief edges, device sweeps, policy nodes (any may be NULL) */
void get_planning_sizes(CPlanningProblem* problem, size_t* n_nodes, size_t* n_belief_nodes, size_t* n_belief_edges, size_t* n_sweeps,
                        size_t* n_policy_nodes);

#ifdef __cplusplus
}
#endif
#endif
