// sssp_frontier.cuh -- frontier relaxation over the value table in global memory (sssp_frontier.cu), used by graph.cu for roadmaps
// too large for the on-chip column solver.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct porrt_ctx;

int32_t sssp_frontier_run(porrt_ctx* ctx, const int64_t* d_row, const int32_t* d_col, const double* d_xy, int64_t V, int64_t E,
                          const int32_t* d_node_vid, const uint64_t* d_validities, int32_t mask_words, int32_t wlo, int32_t W,
                          const int32_t* d_fin_node, const int32_t* d_fin_world, int64_t n_fin, double* d_out_wv, int32_t* out_rounds,
                          double* out_offers, cudaStream_t st);
