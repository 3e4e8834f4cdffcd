// refine.cu -- the policy refiner's data-parallel part.
//
//  * porrt_transition_valid : PTOPolicyRefiner::is_transition_valid (reference src/pto_policy_refiner.rs:395-423), batched:
//                             state validity of both ends + transition validity + belief/validity compatibility.
//  * porrt_partial_shortcut : PTOPolicyRefiner::partial_shortcut (:158-206) on one path piece.  The reference runs
//                             n_iterations SEQUENTIAL trials, each a handful of short edge checks -- a per-trial GPU round trip
//                             could never beat that (SURVEY 8(b) granularity caveat).  But the random choices of a trial
//                             (joint, interval) come from a fresh DiscreteSampler::new() and depend only on the piece length,
//                             so the whole trial sequence is known up front.  Trials are therefore evaluated in speculative
//                             WAVES: the next WAVE trials are built on a speculative copy of the path (each trial assumed to
//                             end like most recent ones did), their transitions are checked in one device batch, and the
//                             wave is replayed in order on the real path; a trial's results are used iff the states it read
//                             are bit-identical to the real ones, the first trial that fails this starts the next wave.
//                             The result -- states and commit count -- is exactly the sequential one.
// The sampler is rand_pcg 0.3 Pcg64 (Lcg128Xsl64) seeded by rand_core 0.6 seed_from_u64 and rand 0.8's
// UniformInt<usize>::sample_single, restated from their published algorithms (the crates are not vendored in the reference).
#include <algorithm>
#include <chrono>
#include <cstring>

#include "pcg64.h"
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ device
// valid / status from the three validity codes, in the reference's evaluation order
// compat: rows of n_validities bytes; row[i] (nullable: row 0) selects the belief state of transition i
__global__ void transition_combine_kernel(const int32_t* __restrict__ sv_from, const int32_t* __restrict__ sv_to,
                                          const int32_t* __restrict__ ev, const uint8_t* __restrict__ compat,
                                          const int32_t* __restrict__ row, int32_t n_validities, int64_t n,
                                          uint8_t* __restrict__ out_valid, int32_t* __restrict__ out_status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t f = sv_from[i], t = sv_to[i], e = ev[i];
  int32_t status = 0;
  uint8_t valid = 0;
  if (f < -1) status = f;                 // state_validity(from) panics first (:396)
  else if (t < -1) status = t;            // then state_validity(to) (:397)
  else if (f >= 0 && t >= 0) {            // only then is the transition looked at (:399-414)
    if (e < -1) status = e;
    else if (e >= 0) valid = compat[(row ? (int64_t)row[i] * n_validities : 0) + e] ? 1 : 0;
  }
  out_valid[i] = valid;
  if (out_status) out_status[i] = status;
}

// device buffers in, device buffers out; d_tmp: 3 * n int32
static int32_t transition_valid_dev(porrt_ctx* ctx, const double* d_from, const double* d_to, int64_t n, const uint8_t* d_compat,
                                    const int32_t* d_row, int32_t n_validities, int32_t* d_tmp, uint8_t* d_valid, int32_t* d_status) {
  cudaStream_t st = ctx->stream;
  int32_t rc = porrt_state_validity_dev(ctx, d_from, n, d_tmp);
  if (rc) return rc;
  rc = porrt_state_validity_dev(ctx, d_to, n, d_tmp + n);
  if (rc) return rc;
  rc = map_edge_validity_dev(ctx, d_from, d_to, n, d_tmp + 2 * n, nullptr, st);
  if (rc) return rc;
  transition_combine_kernel<<<div_up(n, 256), 256, 0, st>>>(d_tmp, d_tmp + n, d_tmp + 2 * n, d_compat, d_row, n_validities, n, d_valid, d_status);
  LAUNCH_CHECK(ctx);
  return PORRT_OK;
}

PORRT_API int32_t porrt_transition_valid(porrt_ctx* ctx, const double* from_xy, const double* to_xy, int64_t n,
                                         const uint8_t* compat_row, uint8_t* out_valid, int32_t* out_status) {
  CTX_CHECK(ctx);
  if (!ctx->has_map) return porrt_fail(ctx, PORRT_ERR_NO_MAP, "no map uploaded");
  if (n < 0 || (n > 0 && (!from_xy || !to_xy || !compat_row || !out_valid))) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "porrt_transition_valid: bad arguments");
  if (n == 0) return PORRT_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nv = (size_t)ctx->n_validities;
  CUDA_TRY(ctx, ctx->scratch[3].ensure((size_t)n * (32 + 12 + 4 + 1) + nv + 64));
  char* b = ctx->scratch[3].as<char>();
  double* d_from = (double*)b; b += (size_t)n * 16;
  double* d_to = (double*)b; b += (size_t)n * 16;
  int32_t* d_tmp = (int32_t*)b; b += (size_t)n * 12;
  int32_t* d_status = (int32_t*)b; b += (size_t)n * 4;
  uint8_t* d_valid = (uint8_t*)b; b += (size_t)n;
  uint8_t* d_compat = (uint8_t*)b;
  CUDA_TRY(ctx, cudaMemcpyAsync(d_from, from_xy, (size_t)n * 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_to, to_xy, (size_t)n * 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_compat, compat_row, nv, cudaMemcpyHostToDevice, st));
  int32_t rc = transition_valid_dev(ctx, d_from, d_to, n, d_compat, nullptr, (int32_t)nv, d_tmp, d_valid, d_status);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(out_valid, d_valid, (size_t)n, cudaMemcpyDeviceToHost, st));
  if (out_status) CUDA_TRY(ctx, cudaMemcpyAsync(out_status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  return PORRT_OK;
}

// ------------------------------------------------------------------------------------------------ sampler (host)
namespace {
struct Trial { int joint, a, b; };
}  // namespace

#define SHORTCUT_WAVE 64

namespace {
struct Piece {                       // one path piece of the policy (pto_policy_refiner.rs:127-156 build_path_piece)
  double* st; int L;                 // its states (in the caller's array), number of states
  std::vector<Trial> trials;
  int i0 = 0, K = 0, recent = 0, commits = 0;   // next trial, trials in the current wave, outcome predictor, committed shortcuts
  std::vector<double> spec, snap;    // speculative path; per trial of the wave the states it read (a..b) when it was built
  std::vector<size_t> first, snap_at;
  size_t base = 0;                   // first transition of this piece in the wave's batch
};
inline double interpolate(double x, double y, double lambda) { return x * (1.0 - lambda) + y * lambda; }   // :159-161
}  // namespace

PORRT_API int32_t porrt_partial_shortcut_batch(porrt_ctx* ctx, double* states_xy, const int32_t* piece_ptr, int32_t n_pieces,
                                               const uint8_t* compat_rows, int32_t n_iterations, uint64_t sampler_seed,
                                               int32_t* out_commits, int32_t* out_waves) {
  CTX_CHECK(ctx);
  if (!ctx->has_map) return porrt_fail(ctx, PORRT_ERR_NO_MAP, "no map uploaded");
  if (n_pieces < 0 || n_iterations < 0 || (n_pieces > 0 && (!states_xy || !piece_ptr || !compat_rows)))
    return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "porrt_partial_shortcut_batch: bad arguments");
  if (out_waves) *out_waves = 0;
  for (int p = 0; p < n_pieces; ++p) {
    if (out_commits) out_commits[p] = 0;
    if (piece_ptr[p + 1] < piece_ptr[p]) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "porrt_partial_shortcut_batch: piece_ptr must be non-decreasing");
  }
  if (n_pieces == 0 || n_iterations == 0) return PORRT_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nv = (size_t)ctx->n_validities;

  // every piece draws the same kind of sequence from a fresh sampler (:169 DiscreteSampler::new()): joint, interval start,
  // interval end (:172-174) depend on the piece length only, so all trials are known before anything is checked
  std::vector<Piece> pieces;
  std::vector<int> piece_of;           // index into the caller's arrays
  size_t cap = 0;
  for (int p = 0; p < n_pieces; ++p) {
    const int L = piece_ptr[p + 1] - piece_ptr[p];
    if (L <= 2) continue;              // :163-165: can't shortcut with only 2 states or less
    Piece pc;
    pc.st = states_xy + 2 * (size_t)piece_ptr[p]; pc.L = L;
    pc.trials.resize((size_t)n_iterations);
    Pcg64 rng = Pcg64::seed_from_u64(sampler_seed);
    for (int i = 0; i < n_iterations; ++i) {
      Trial t;
      t.joint = (int)rng.below(2);
      t.a = (int)rng.below((uint64_t)(L - 2));
      t.b = t.a + 2 + (int)rng.below((uint64_t)(L - t.a - 2));
      pc.trials[(size_t)i] = t;
    }
    pc.spec.resize((size_t)2 * L);
    pc.first.resize(SHORTCUT_WAVE + 1); pc.snap_at.resize(SHORTCUT_WAVE + 1);
    cap += (size_t)SHORTCUT_WAVE * (size_t)(L - 1);      // a wave has at most SHORTCUT_WAVE * (L - 1) transitions per piece
    pieces.push_back(std::move(pc));
    piece_of.push_back(p);
  }
  if (pieces.empty()) return PORRT_OK;
  CUDA_TRY(ctx, ctx->pin[2].ensure(cap * (32 + 4 + 4 + 1)));
  CUDA_TRY(ctx, ctx->scratch[3].ensure(cap * (32 + 12 + 4 + 4 + 1) + (size_t)n_pieces * nv + 64));
  double* h_from = ctx->pin[2].as<double>();
  double* h_to = h_from + 2 * cap;
  int32_t* h_row = (int32_t*)(h_to + 2 * cap);
  int32_t* h_status = h_row + cap;
  uint8_t* h_valid = (uint8_t*)(h_status + cap);
  char* b = ctx->scratch[3].as<char>();
  double* d_from = (double*)b; b += cap * 16;
  double* d_to = (double*)b; b += cap * 16;
  int32_t* d_tmp = (int32_t*)b; b += cap * 12;
  int32_t* d_row = (int32_t*)b; b += cap * 4;
  int32_t* d_status = (int32_t*)b; b += cap * 4;
  uint8_t* d_valid = (uint8_t*)b; b += cap;
  uint8_t* d_compat = (uint8_t*)b;
  CUDA_TRY(ctx, cudaMemcpyAsync(d_compat, compat_rows, (size_t)n_pieces * nv, cudaMemcpyHostToDevice, st));

  int waves = 0;
  double t_build = 0, t_dev = 0, t_replay = 0;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (;;) {
    const double tb0 = now();
    // ---- build: the next <= SHORTCUT_WAVE trials of every unfinished piece, on a SPECULATIVE copy of its path: every trial
    // is assumed to end like most recent ones did (commit / reject), so later trials of the wave see the states they will
    // most likely see.  Shortcut states (:182-190) and transitions (:193-197) of all pieces form one device batch.
    size_t m = 0;
    for (size_t q = 0; q < pieces.size(); ++q) {
      Piece& pc = pieces[q];
      pc.K = std::min(SHORTCUT_WAVE, n_iterations - pc.i0);
      if (pc.K <= 0) { pc.K = 0; continue; }
      const bool predict_commit = pc.recent >= 0;
      std::copy(pc.st, pc.st + 2 * pc.L, pc.spec.begin());
      pc.snap.clear();
      pc.base = m;
      double* spec = pc.spec.data();
      for (int w = 0; w < pc.K; ++w) {
        const Trial& t = pc.trials[(size_t)(pc.i0 + w)];
        pc.first[(size_t)w] = m;
        pc.snap_at[(size_t)w] = pc.snap.size();
        pc.snap.insert(pc.snap.end(), spec + 2 * t.a, spec + 2 * (t.b + 1));
        const double sa = spec[2 * t.a + t.joint], sb = spec[2 * t.b + t.joint];
        double prev[2] = {0, 0};
        for (int j = t.a; j < t.b; ++j) {
          const double lambda = (double)(j - t.a) / (double)(t.b - t.a);
          double s2[2] = {spec[2 * j], spec[2 * j + 1]};
          s2[t.joint] = interpolate(sa, sb, lambda);
          if (j > t.a) { h_from[2 * m] = prev[0]; h_from[2 * m + 1] = prev[1]; h_to[2 * m] = s2[0]; h_to[2 * m + 1] = s2[1]; h_row[m] = piece_of[q]; ++m; }
          prev[0] = s2[0]; prev[1] = s2[1];
          if (predict_commit) spec[2 * j + t.joint] = s2[t.joint];   // in place is fine: sa / sb were read before, state a maps to itself
        }
        h_from[2 * m] = prev[0]; h_from[2 * m + 1] = prev[1];                     // last shortcut state -> interval end (:197)
        h_to[2 * m] = spec[2 * t.b]; h_to[2 * m + 1] = spec[2 * t.b + 1];
        h_row[m] = piece_of[q];
        ++m;
      }
      pc.first[(size_t)pc.K] = m;
      pc.snap_at[(size_t)pc.K] = pc.snap.size();
    }
    if (m == 0) break;                 // every piece has run all its trials
    const double tb1 = now();
    t_build += tb1 - tb0;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_from, h_from, m * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_to, h_to, m * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_row, h_row, m * 4, cudaMemcpyHostToDevice, st));
    int32_t rc = transition_valid_dev(ctx, d_from, d_to, (int64_t)m, d_compat, d_row, (int32_t)nv, d_tmp, d_valid, d_status);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(h_valid, d_valid, m, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_status, d_status, m * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    ++waves;
    const double tb2 = now();
    t_dev += tb2 - tb1;
    // ---- replay, piece by piece, in trial order on the REAL path.  A trial's device results are usable iff the states it
    // read while the wave was built are bit-identical to the real ones now; the first trial that fails this test starts
    // the piece's next wave.
    for (size_t q = 0; q < pieces.size(); ++q) {
      Piece& pc = pieces[q];
      int w = 0;
      for (; w < pc.K; ++w) {
        const Trial& t = pc.trials[(size_t)(pc.i0 + w)];
        if (!std::equal(pc.snap.begin() + pc.snap_at[(size_t)w], pc.snap.begin() + pc.snap_at[(size_t)w + 1], pc.st + 2 * t.a,
                        [](double x, double y) { return memcmp(&x, &y, 8) == 0; }))
          break;
        bool should_commit = true;
        for (size_t k = pc.first[(size_t)w]; k < pc.first[(size_t)w + 1]; ++k) {
          if (!h_valid[k]) {
            if (h_status[k] < -1) {                // the reference panics here (the `&&` chain evaluates up to the first false)
              if (out_waves) *out_waves = waves;
              return porrt_fail(ctx, PORRT_ERR_PANIC, "partial_shortcut: a transition check hit a reference panic (code " + std::to_string(h_status[k]) + ")");
            }
            should_commit = false;
            break;
          }
        }
        if (should_commit) {                       // :200-204
          const double sa = pc.st[2 * t.a + t.joint], sb = pc.st[2 * t.b + t.joint];
          for (int j = t.a; j < t.b; ++j) pc.st[2 * j + t.joint] = interpolate(sa, sb, (double)(j - t.a) / (double)(t.b - t.a));
          ++pc.commits;
        }
        pc.recent = std::max(-4, std::min(4, pc.recent + (should_commit ? 1 : -1)));
      }
      pc.i0 += w;                      // w >= 1 whenever K >= 1: the first trial of a wave is built from the real path
      if (out_commits) out_commits[piece_of[q]] = pc.commits;
    }
    t_replay += now() - tb2;
  }
  if (getenv("PORRT_DEBUG")) fprintf(stderr, "[porrt] partial_shortcut: %d waves, build %.2f ms, device %.2f ms, replay %.2f ms\n", waves, t_build, t_dev, t_replay);
  if (out_waves) *out_waves = waves;
  return PORRT_OK;
}

PORRT_API int32_t porrt_partial_shortcut(porrt_ctx* ctx, double* states_xy, int32_t n_states, const uint8_t* compat_row,
                                         int32_t n_iterations, uint64_t sampler_seed, int32_t* out_commits, int32_t* out_waves) {
  if (n_states < 0) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "porrt_partial_shortcut: bad arguments");
  const int32_t ptr[2] = {0, n_states};
  return porrt_partial_shortcut_batch(ctx, states_xy, ptr, 1, compat_row, n_iterations, sampler_seed, out_commits, out_waves);
}

// ================================================================================================ refine_solution(PartialShortCut(n))
// PTOPolicyRefiner::refine_solution with RefinmentStrategy::PartialShortCut (pto_policy_refiner.rs:85-133; what every PTO run of
// the reference's main.rs ends with, e.g. :442 PartialShortCut(1500)) on the policy porrt_extract_policy produced from the last
// porrt_belief_vi of this ctx:
//   Policy::decompose (common.rs:85-129)  ->  build_path_piece (:135-156) + partial_shortcut (:158-206) per piece, all pieces in one
//   batch on the device (porrt_partial_shortcut_batch)  ->  recompose (:324-393) incl. compute_expected_costs_to_goals (common.rs:131-153).
// Everything around the batch is sequential host logic, written here over flat arrays: a policy is (node, belief, parent) per policy
// node in creation order, children of a node = its later nodes in creation order (Policy::add_edge is called right after add_node).
// A piece of ONE node is a start but not an end in recompose (:356-361): its successors are never connected -- kept as is.
PORRT_API int32_t porrt_refine_policy_shortcut(porrt_ctx* ctx, const int32_t* pol_node, const int32_t* pol_belief, const int32_t* pol_parent,
                                               int64_t n_pol, int32_t n_iterations, uint64_t sampler_seed, double* out_xy, int32_t* out_node,
                                               int32_t* out_belief, int32_t* out_parent, uint8_t* out_is_leaf, int64_t cap, int64_t* out_n,
                                               double* out_expected_cost, int64_t* out_commits) {
  CTX_CHECK(ctx);
  auto& R = ctx->bel;
  if (R.V <= 0) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "refine_policy_shortcut: run porrt_belief_vi / porrt_extract_policy first");
  if (n_pol <= 0 || !pol_node || !pol_belief || !pol_parent || n_iterations < 0 || !out_n)
    return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "refine_policy_shortcut: bad arguments");
  const int B = R.B, nw = R.n_worlds, nv = R.n_validities;
  for (int64_t k = 0; k < n_pol; ++k)
    if (pol_node[k] < 0 || pol_node[k] >= R.V || pol_belief[k] < 0 || pol_belief[k] >= B || pol_parent[k] >= k || (k > 0 && pol_parent[k] < 0) || (k == 0 && pol_parent[k] != -1))
      return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "refine_policy_shortcut: not a policy in creation order");
  // children in creation order
  std::vector<int64_t> ch_ptr((size_t)n_pol + 1, 0);
  for (int64_t k = 1; k < n_pol; ++k) ++ch_ptr[(size_t)pol_parent[k] + 1];
  for (int64_t k = 0; k < n_pol; ++k) ch_ptr[(size_t)k + 1] += ch_ptr[(size_t)k];
  std::vector<int64_t> ch((size_t)std::max<int64_t>(n_pol - 1, 0)), fill(ch_ptr.begin(), ch_ptr.end() - 1);
  for (int64_t k = 1; k < n_pol; ++k) ch[(size_t)fill[(size_t)pol_parent[k]]++] = k;
  // ---- decompose: FIFO over piece starts; a piece runs until a leaf or a branching node
  std::vector<int32_t> piece_ptr(1, 0);
  std::vector<int64_t> piece_nodes;                   // policy node ids, piece after piece
  std::vector<std::vector<int32_t>> successors;       // skeleton
  {
    std::vector<int64_t> fifo(1, 0);
    int32_t n_pieces = 0;
    for (size_t head = 0; head < fifo.size(); ++head) {
      int64_t cur = fifo[head];
      std::vector<int32_t> succ;
      for (;;) {
        piece_nodes.push_back(cur);
        const int64_t nc = ch_ptr[(size_t)cur + 1] - ch_ptr[(size_t)cur];
        if (nc == 0) break;
        if (nc == 1) { cur = ch[(size_t)ch_ptr[(size_t)cur]]; continue; }
        for (int64_t e = ch_ptr[(size_t)cur]; e < ch_ptr[(size_t)cur + 1]; ++e) { fifo.push_back(ch[(size_t)e]); succ.push_back(++n_pieces); }
        break;
      }
      piece_ptr.push_back((int32_t)piece_nodes.size());
      successors.push_back(succ);
    }
  }
  const int32_t n_pieces = (int32_t)successors.size();
  const int64_t n_out = (int64_t)piece_nodes.size();   // every policy node ends up in exactly one piece
  *out_n = n_out;
  if (n_out > cap || !out_xy || !out_node || !out_belief || !out_parent || !out_is_leaf)
    return porrt_fail(ctx, PORRT_ERR_CAPACITY, "refine_policy_shortcut: cap too small");
  // ---- build_path_piece: the pieces' states back to back; the piece's belief is the one of its first node
  std::vector<uint8_t> rows((size_t)n_pieces * nv);
  for (int64_t k = 0; k < n_out; ++k) {
    const int64_t pn = piece_nodes[(size_t)k];
    out_xy[2 * k] = R.xy[2 * (size_t)pol_node[pn]]; out_xy[2 * k + 1] = R.xy[2 * (size_t)pol_node[pn] + 1];
    out_node[k] = pol_node[pn]; out_belief[k] = pol_belief[pn];
  }
  for (int32_t p = 0; p < n_pieces; ++p)
    memcpy(&rows[(size_t)p * nv], &R.compat[(size_t)pol_belief[piece_nodes[(size_t)piece_ptr[p]]] * nv], (size_t)nv);
  // ---- partial_shortcut on all pieces (device batch, sequential semantics)
  std::vector<int32_t> commits((size_t)n_pieces, 0);
  int32_t rc = porrt_partial_shortcut_batch(ctx, out_xy, piece_ptr.data(), n_pieces, rows.data(), n_iterations, sampler_seed, commits.data(), nullptr);
  if (rc) return rc;
  if (out_commits) { *out_commits = 0; for (int32_t c : commits) *out_commits += c; }
  // ---- recompose: chains inside the pieces, then end of piece i -> start of its successors (an end exists only for >= 2 nodes)
  for (int32_t p = 0; p < n_pieces; ++p)
    for (int32_t k = piece_ptr[p]; k < piece_ptr[p + 1]; ++k) out_parent[k] = k == piece_ptr[p] ? -1 : k - 1;
  for (int32_t p = 0; p < n_pieces; ++p) {
    if (piece_ptr[p + 1] - piece_ptr[p] < 2) continue;
    for (int32_t q : successors[(size_t)p]) out_parent[piece_ptr[q]] = piece_ptr[p + 1] - 1;
  }
  // children of the recomposed policy in add_edge order: the chain edge, or -- at a piece's end -- the successors in skeleton order
  // (successor pieces come later in the arrays, so increasing index = add_edge order); leaves = nodes without children
  std::vector<int32_t> n_children((size_t)n_out, 0);
  for (int64_t k = 0; k < n_out; ++k) if (out_parent[k] >= 0) ++n_children[(size_t)out_parent[k]];
  for (int64_t k = 0; k < n_out; ++k) out_is_leaf[k] = n_children[(size_t)k] == 0;
  // ---- compute_expected_costs_to_goals: probabilities top down, costs bottom up; children are visited in increasing index
  auto transition_probability = [&](int pb, int cb) {
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s = s + (R.beliefs[(size_t)cb * nw + i] > 0.0 ? R.beliefs[(size_t)pb * nw + i] : 0.0);
    return s;
  };
  std::vector<double> prob((size_t)n_out, 0.0), pq((size_t)n_out, 0.0), edge_term((size_t)n_out, 0.0), below((size_t)n_out, 0.0);
  prob[0] = 1.0;
  for (int64_t k = 1; k < n_out; ++k) {
    const int32_t par = out_parent[k];
    if (par < 0) continue;   // a piece that recompose left unconnected: not under the root
    const double q = transition_probability(out_belief[par], out_belief[k]);
    const double dx = out_xy[2 * k] - out_xy[2 * (size_t)par], dy = out_xy[2 * k + 1] - out_xy[2 * (size_t)par + 1];
    const double cost = std::sqrt(dx * dx + dy * dy);   // norm2(parent, child), common.rs:203-213
    pq[(size_t)k] = prob[(size_t)par] * q;              // p * q
    prob[(size_t)k] = pq[(size_t)k];
    edge_term[(size_t)k] = pq[(size_t)k] * cost;        // (p * q) * cost
  }
  // expected_future_costs(parent) += p * q * cost + expected_future_costs(child), children in order (common.rs:149): a child's own sum
  // is complete before its parent folds it in -- fold the children of each node in increasing index, nodes in decreasing index
  std::vector<int64_t> oc_ptr((size_t)n_out + 1, 0);
  for (int64_t k = 0; k < n_out; ++k) if (out_parent[k] >= 0) ++oc_ptr[(size_t)out_parent[k] + 1];
  for (int64_t k = 0; k < n_out; ++k) oc_ptr[(size_t)k + 1] += oc_ptr[(size_t)k];
  std::vector<int64_t> oc((size_t)oc_ptr[(size_t)n_out]), ofill(oc_ptr.begin(), oc_ptr.end() - 1);
  for (int64_t k = 0; k < n_out; ++k) if (out_parent[k] >= 0) oc[(size_t)ofill[(size_t)out_parent[k]]++] = k;
  for (int64_t k = n_out - 1; k >= 0; --k) {
    double acc = 0.0;
    for (int64_t e = oc_ptr[(size_t)k]; e < oc_ptr[(size_t)k + 1]; ++e) { const int64_t c = oc[(size_t)e]; acc += edge_term[(size_t)c] + below[(size_t)c]; }
    below[(size_t)k] = acc;
  }
  if (out_expected_cost) *out_expected_cost = below[0];
  return PORRT_OK;
}
