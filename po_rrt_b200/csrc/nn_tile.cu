// nn_tile.cu -- nearest-neighbour search over TMA-staged vertex tiles (SURVEY.md 8(a) rows B2/B3 + the k-NN metric).
//
// Replaces, batched, KdTree::nearest_neighbors[_filtered] / nearest_neighbor[_filtered] (reference
// src/nearest_neighbor.rs:48-126).  nn.cu's first version ran one thread per query straight on the cell-sorted
// vertex array in global memory: every query re-read the ~9 cells around it through L1/L2 (124 candidates x 16 B at the
// c5 shape) and k-NN kept its candidate list in local memory.  Here
//   * the QUERIES are binned by grid cell as well (count / scan / scatter), so the queries of a run of NT_W cells of
//     one cell row form a contiguous slice;
//   * a CTA takes such a tile, and the vertices any of its queries can reach -- three cell rows x (NT_W + 2) cells,
//     i.e. three CONTIGUOUS ranges of the cell-sorted arrays -- are staged in shared memory by bulk-async copies
//     (cp.async.bulk, mbarrier byte-count completion): each vertex is fetched once per tile instead of once per query;
//   * every query is owned by one THREAD of the tile's CTA, which walks its (<= 3) candidate runs in shared memory; k-NN
//     keeps its k best in shared memory too (column layout, conflict-free).  Warp-per-query variants (ballot-compacted
//     hit lists; a warp-wide sorted list) cost 2-10x the instructions per candidate -- see the notes at the kernels.
// Distances are the reference's f64 arithmetic (common.rs:203-213) and `sqrt(d2) <= r` is evaluated as d2 <= T(r).
// Queries whose radius reaches beyond the neighbouring cells, tiles whose candidates exceed the staging buffer and k-NN
// queries that need a wider ring go to nn.cu's thread-per-query kernels (exact ring search) through an index list.
#include "nn_dev.cuh"

#define NT_W 8               // cells per tile along x
#define NT_THREADS 128
#define NT_CAP 2048          // staged vertices per tile (radius): 32 KiB + 8 KiB ids
#define NT_CAP_KNN 768       // k-NN: leaves room for the per-thread lists

__device__ __forceinline__ uint32_t nt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nt_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(nt_smem_u32(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nt_mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nt_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(nt_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(nt_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void nt_mbar_wait0(uint64_t* bar) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(nt_smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ query binning
// qcell[t] = cell of query t if the tile path can serve it, -1 -> fallback list, -2 -> no result at all (NaN, r < 0)
template <bool KNN>
__global__ void nt_bin_kernel(GridDev g, const double2* __restrict__ q, const double* __restrict__ radius, int64_t m,
                              int32_t* __restrict__ qcell, int32_t* __restrict__ qcount, int32_t* __restrict__ fb_list, int32_t* __restrict__ fb_n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const double2 p = q[t];
  int32_t c = -2;
  if (p.x == p.x && p.y == p.y) {
    const int cx = cell_coord(p.x, g.org_x, g.inv_cell, g.cells_x), cy = cell_coord(p.y, g.org_y, g.inv_cell, g.cells_y);
    bool ok = true, any = true;
    if (!KNN) {
      const double r = radius[t];
      any = radius_threshold(r) >= 0.0;
      const double rr = isinf(r) ? r : __dadd_rn(__dmul_rn(r, 1.000000001), 1e-300);  // the same conservative cell cover as nn.cu
      const int cx0 = cell_coord(p.x - rr, g.org_x, g.inv_cell, g.cells_x), cx1 = cell_coord(p.x + rr, g.org_x, g.inv_cell, g.cells_x);
      const int cy0 = cell_coord(p.y - rr, g.org_y, g.inv_cell, g.cells_y), cy1 = cell_coord(p.y + rr, g.org_y, g.inv_cell, g.cells_y);
      ok = cx0 >= cx - 1 && cx1 <= cx + 1 && cy0 >= cy - 1 && cy1 <= cy + 1;
    }
    if (any) {
      if (ok) { c = cy * g.cells_x + cx; atomicAdd(&qcount[c], 1); }
      else { c = -1; fb_list[atomicAdd(fb_n, 1)] = (int32_t)t; }
    }
  } else if (KNN) { c = -1; fb_list[atomicAdd(fb_n, 1)] = (int32_t)t; }   // nn.cu writes the "nothing found" record
  qcell[t] = c;
}

__global__ void nt_scatter_kernel(const int32_t* __restrict__ qcell, int64_t m, const int64_t* __restrict__ qstart,
                                  int32_t* __restrict__ cursor, int32_t* __restrict__ qorder) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int32_t c = qcell[t];
  if (c >= 0) qorder[qstart[c] + atomicAdd(&cursor[c], 1)] = (int32_t)t;
}

// ------------------------------------------------------------------------------------------------ tile staging
struct Tile {
  int cy, cxa, cxb;          // query cells: row cy, columns cxa..cxb
  int ca;                    // first staged column
  int64_t qa; int nq;        // slice of qorder
  int64_t k0[3]; int len[3]; // staged vertex ranges of rows cy-1, cy, cy+1 (len 0: row outside the grid)
  int xoff[3], ioff[3];      // their offsets in s_xy / s_id (ids keep their 16-byte phase: element j at ioff + j)
  bool staged;
};

__device__ __forceinline__ Tile tile_setup(const GridDev& g, const int64_t* __restrict__ qstart, int tiles_per_row, int cap) {
  Tile T;
  const int tile = blockIdx.x;
  T.cy = tile / tiles_per_row;
  T.cxa = (tile % tiles_per_row) * NT_W;
  T.cxb = min(T.cxa + NT_W - 1, g.cells_x - 1);
  T.qa = qstart[(int64_t)T.cy * g.cells_x + T.cxa];
  T.nq = (int)(qstart[(int64_t)T.cy * g.cells_x + T.cxb + 1] - T.qa);
  T.ca = max(T.cxa - 1, 0);
  const int cb = min(T.cxb + 1, g.cells_x - 1);
  int tot = 0, itot = 0;
  for (int r = 0; r < 3; ++r) {
    const int row = T.cy - 1 + r;
    T.k0[r] = 0; T.len[r] = 0;
    if (T.nq > 0 && row >= 0 && row < g.cells_y) {
      T.k0[r] = g.cell_start[(int64_t)row * g.cells_x + T.ca];
      T.len[r] = (int)min((int64_t)0x3fffffff, g.cell_start[(int64_t)row * g.cells_x + cb + 1] - T.k0[r]);
    }
    T.xoff[r] = tot; tot += T.len[r];
    T.ioff[r] = itot + (int)(T.k0[r] & 3);                       // copy starts at the 16-byte boundary below k0
    itot += (((int)(T.k0[r] & 3) + T.len[r] + 3) & ~3);
  }
  T.staged = tot <= cap;
  return T;
}

// thread 0 issues the bulk copies; everybody waits for the bytes to land
__device__ __forceinline__ void tile_stage(const GridDev& g, const Tile& T, double2* s_xy, int32_t* s_id, uint64_t* bar) {
  if (threadIdx.x == 0) nt_mbar_init(bar);
  __syncthreads();
  if (T.staged) {
    if (threadIdx.x == 0) {
      uint32_t bytes = 0;
      for (int r = 0; r < 3; ++r)
        if (T.len[r] > 0) bytes += (uint32_t)T.len[r] * 16u + (uint32_t)((((int)(T.k0[r] & 3) + T.len[r] + 3) & ~3) * 4);
      nt_mbar_expect(bar, bytes);
      for (int r = 0; r < 3; ++r)
        if (T.len[r] > 0) {
          nt_bulk_g2s(s_xy + T.xoff[r], g.vxy + T.k0[r], (uint32_t)T.len[r] * 16u, bar);
          const int ph = (int)(T.k0[r] & 3);
          nt_bulk_g2s(s_id + T.ioff[r] - ph, g.vid + (T.k0[r] - ph), (uint32_t)(((ph + T.len[r] + 3) & ~3) * 4), bar);
        }
    }
    nt_mbar_wait0(bar);
  }
}

// ------------------------------------------------------------------------------------------------ radius: thread per query
// (A warp-per-query version -- lanes striding the candidates, ballot-compacted coalesced id lists -- was built first and
// measured: 709 + 828 us for the count + fill passes of 1e6 queries, slower than nn.cu's 450 + 450 us.  The per-query set-up
// (exact threshold, cell cover, six cell_start look-ups) was paid by 32 lanes for ONE query and the cross-lane bookkeeping
// cost ~0.8 warp instructions per candidate against ~0.4 for a thread that owns its query.)
template <bool FILL>
__global__ void __launch_bounds__(NT_THREADS) nt_radius_kernel(GridDev g, const double2* __restrict__ q, const double* __restrict__ radius,
                                                               const uint32_t* __restrict__ prefix, const uint64_t* __restrict__ reach,
                                                               const uint32_t* __restrict__ world, const int32_t* __restrict__ qorder,
                                                               const int64_t* __restrict__ qstart, int tiles_per_row,
                                                               int32_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                                               int32_t* __restrict__ out_ids, const uint32_t* __restrict__ prefix_lo) {
  __shared__ __align__(128) double2 s_xy[NT_CAP];
  __shared__ __align__(16) int32_t s_id[NT_CAP + 24];
  __shared__ uint64_t s_bar;
  const Tile T = tile_setup(g, qstart, tiles_per_row, NT_CAP);
  if (T.nq == 0) return;
  tile_stage(g, T, s_xy, s_id, &s_bar);
  for (int qi = threadIdx.x; qi < T.nq; qi += NT_THREADS) {
    const int32_t t = qorder[T.qa + qi];
    const double2 p = q[t];
    const double r = radius[t];
    const double Tr = radius_threshold(r);
    const uint32_t limit = prefix ? prefix[t] : 0xffffffffu;
    const uint32_t lo_limit = prefix_lo ? prefix_lo[t] : 0u;
    const uint32_t wq = (reach && world) ? world[t] : 0u;
    const double rr = isinf(r) ? r : __dadd_rn(__dmul_rn(r, 1.000000001), 1e-300);
    const int cx0 = cell_coord(p.x - rr, g.org_x, g.inv_cell, g.cells_x), cx1 = cell_coord(p.x + rr, g.org_x, g.inv_cell, g.cells_x);
    const int cy0 = cell_coord(p.y - rr, g.org_y, g.inv_cell, g.cells_y), cy1 = cell_coord(p.y + rr, g.org_y, g.inv_cell, g.cells_y);
    int cnt = 0;
    int32_t* out = FILL ? out_ids + offsets[t] : nullptr;
#pragma unroll
    for (int rr3 = 0; rr3 < 3; ++rr3) {
      const int row = T.cy - 1 + rr3;
      if (row < cy0 || row > cy1) continue;
      const int64_t s = g.cell_start[(int64_t)row * g.cells_x + cx0], e = g.cell_start[(int64_t)row * g.cells_x + cx1 + 1];
      const int n_run = (int)(e - s), s_run = (int)(s - T.k0[rr3]);
      const double2* cx = T.staged ? s_xy + T.xoff[rr3] + s_run : g.vxy + s;
      const int32_t* ci = T.staged ? s_id + T.ioff[rr3] + s_run : g.vid + s;
      for (int j = 0; j < n_run; ++j) {
        if (dist2(cx[j], p.x, p.y) <= Tr) {
          const int32_t id = ci[j];
          if ((uint32_t)id < limit && (uint32_t)id >= lo_limit && (!reach || reach_bit(reach, g.reach_words, id, wq))) {
            if (FILL) out[cnt] = id;
            ++cnt;
          }
        }
      }
    }
    if (!FILL) counts[t] = cnt;
  }
}

// ------------------------------------------------------------------------------------------------ 1-NN / k-NN: thread per query
// k = 1: running minimum (+ the number of vertices at exactly the winning d2, nn.cu: knn_kernel<1>).
// k > 1: a sorted list kept by insertion was measured first (1.75 ms for 1e6 queries, k = 16): lanes insert at different
// candidates, so the warp pays the insertion loop at almost every candidate.  Now two uniform passes over the candidates:
//   pass 1 bins d2 into NT_BINS equal-width bins (equal area: candidates are ~uniform in d2) in a per-thread byte
//          histogram and finds the first bin where the cumulative count reaches k;
//   pass 2 collects the candidates of the bins up to that one (k + a few) into the thread's list, which is then
//          insertion-sorted by (d2, id) and cut at k.
// Exact: both passes evaluate the same expression per candidate; a query is finished here only if its k-th best is
// provably closer than anything outside the 3 x 3 cells (ring_lower_bound2, R = 1), otherwise -- or if its list overflows
// (> k + NT_SLACK candidates in the bins, e.g. many exact duplicates) -- it goes to nn.cu's exact ring search.
// dynamic shared memory: [ s_xy NT_CAP_KNN | s_id NT_CAP_KNN + 24 | list cap x NT_THREADS (u16 codes)
//                          | histogram NT_BINS x NT_THREADS (u8) ],  cap = k + NT_SLACK
#define NT_BINS 32
#define NT_SLACK 16
__global__ void __launch_bounds__(NT_THREADS) nt_knn_kernel(GridDev g, const double2* __restrict__ q, int k,
                                                            const uint64_t* __restrict__ reach, const uint32_t* __restrict__ world,
                                                            const int32_t* __restrict__ qorder, const int64_t* __restrict__ qstart,
                                                            int tiles_per_row, int32_t* __restrict__ out_ids, double* __restrict__ out_dist,
                                                            int32_t* __restrict__ out_ties, int32_t* __restrict__ fb_list, int32_t* __restrict__ fb_n) {
  extern __shared__ __align__(128) unsigned char nt_smem[];
  const int cap = k + NT_SLACK;
  double2* s_xy = (double2*)nt_smem;
  int32_t* s_id = (int32_t*)(nt_smem + (size_t)NT_CAP_KNN * 16);
  uint16_t* s_lst = (uint16_t*)(s_id + NT_CAP_KNN + 24);
  uint8_t* s_hist = (uint8_t*)(s_lst + (size_t)cap * NT_THREADS);
  __shared__ uint64_t s_bar;
  const Tile T = tile_setup(g, qstart, tiles_per_row, NT_CAP_KNN);
  if (T.nq == 0) return;
  tile_stage(g, T, s_xy, s_id, &s_bar);
  uint16_t* lst = s_lst + threadIdx.x; // column layout: entry j of this thread at [j * NT_THREADS]
  uint8_t* hist = s_hist + threadIdx.x;
  // bins cover d2 in [0, (1.5 cell)^2): whatever lies beyond cannot be certified by the R = 1 bound anyway (<= 2 cells)
  const double bin_scale = (double)NT_BINS / (2.25 * g.cell * g.cell);
  for (int qi = threadIdx.x; qi < T.nq; qi += NT_THREADS) {
    const int32_t t = qorder[T.qa + qi];
    const double2 p = q[t];
    const uint32_t wq = (reach && world) ? world[t] : 0u;
    const int cx = cell_coord(p.x, g.org_x, g.inv_cell, g.cells_x);
    const int cx0 = max(cx - 1, 0), cx1 = min(cx + 1, g.cells_x - 1);
    // the query's candidate runs (rows cy-1, cy, cy+1)
    const double2* rx[3]; const int32_t* ri[3]; int rn[3];
#pragma unroll
    for (int rr3 = 0; rr3 < 3; ++rr3) {
      const int row = T.cy - 1 + rr3;
      rn[rr3] = 0; rx[rr3] = s_xy; ri[rr3] = s_id;
      if (row >= 0 && row < g.cells_y) {
        const int64_t s = g.cell_start[(int64_t)row * g.cells_x + cx0], e = g.cell_start[(int64_t)row * g.cells_x + cx1 + 1];
        const int s_run = (int)(s - T.k0[rr3]);
        rn[rr3] = (int)(e - s);
        rx[rr3] = T.staged ? s_xy + T.xoff[rr3] + s_run : g.vxy + s;
        ri[rr3] = T.staged ? s_id + T.ioff[rr3] + s_run : g.vid + s;
      }
    }
    bool done = false;
    if (k == 1) {
      int cnt = 0, ties = 0;
      double best = INFINITY;
      int32_t best_id = 0x7fffffff;
#pragma unroll
      for (int rr3 = 0; rr3 < 3; ++rr3)
        for (int j = 0; j < rn[rr3]; ++j) {
          const double d = dist2(rx[rr3][j], p.x, p.y);
          if (d != d || (cnt != 0 && d > best)) continue;
          const int32_t id = ri[rr3][j];
          if (reach && !reach_bit(reach, g.reach_words, id, wq)) continue;
          if (cnt == 0 || d < best) ties = 1; else ++ties;       // here d == best
          if (cnt == 0 || d < best || id < best_id) { best = d; best_id = id; }
          cnt = 1;
        }
      if (cnt == 1 && best < ring_lower_bound2(g, p.x, p.y, cx, T.cy, 1)) {
        out_ids[t] = best_id;
        if (out_dist) out_dist[t] = __dsqrt_rn(best);
        if (out_ties) out_ties[t] = ties;
        done = true;
      }
    } else {
      // pass 1: histogram of d2
#pragma unroll
      for (int bsel = 0; bsel < NT_BINS; ++bsel) hist[bsel * NT_THREADS] = 0;
      int n_far = 0;   // candidates beyond the last bin are only counted
#pragma unroll
      for (int rr3 = 0; rr3 < 3; ++rr3)
        for (int j = 0; j < rn[rr3]; ++j) {
          const double d = dist2(rx[rr3][j], p.x, p.y);
          if (d != d) continue;
          if (reach && !reach_bit(reach, g.reach_words, ri[rr3][j], wq)) continue;
          const double fb = d * bin_scale;
          if (fb < (double)NT_BINS) { uint8_t* h = hist + (int)fb * NT_THREADS; if (*h < 255) ++*h; }
          else ++n_far;
        }
      int cum = 0, B = -1;
      for (int bsel = 0; bsel < NT_BINS; ++bsel) {
        const int h = hist[bsel * NT_THREADS];
        if (h == 255) break;           // saturated counter: let the exact ring search handle it
        cum += h;
        if (cum >= k) { B = bsel; break; }
      }
      if (B >= 0 && cum <= cap && T.staged) {
        // pass 2: collect the candidates of bins 0..B as 16-bit codes (index into the staged tile), unsorted,
        // then insertion sort by (d2, id) with d2 recomputed from the staged tile.  2 bytes per entry instead of 12: the
        // shared-memory footprint of a CTA drops from 56 to 25 KiB and twice as many warps hide the latencies.
        const double lim = (double)(B + 1);
        int cnt = 0;
#pragma unroll
        for (int rr3 = 0; rr3 < 3; ++rr3)
          for (int j = 0; j < rn[rr3]; ++j) {
            const double d = dist2(rx[rr3][j], p.x, p.y);
            if (d != d || !(d * bin_scale < lim)) continue;
            if (reach && !reach_bit(reach, g.reach_words, ri[rr3][j], wq)) continue;
            lst[cnt * NT_THREADS] = (uint16_t)(rx[rr3] - s_xy + j);    // unsorted append: the warp stays together
            ++cnt;
          }
        // code = index into the staged coordinates; the staged ids sit at a per-row offset from it (16-byte phase of the copy)
        const int x1 = T.xoff[1], x2 = T.xoff[2], d0 = T.ioff[0] - T.xoff[0], d1 = T.ioff[1] - T.xoff[1], d2o = T.ioff[2] - T.xoff[2];
        auto cand_xy = [&](uint32_t code) -> double2 { return s_xy[code]; };
        auto cand_id = [&](uint32_t code) -> int32_t { return s_id[(int)code + ((int)code >= x2 ? d2o : ((int)code >= x1 ? d1 : d0))]; };
        for (int i = 1; i < cnt; ++i) {                              // every lane sorts its k + few entries at the same time
          const uint32_t code = lst[i * NT_THREADS];
          const double d = dist2(cand_xy(code), p.x, p.y);
          int pos = i;
          while (pos > 0) {
            const uint32_t pc = lst[(pos - 1) * NT_THREADS];
            const double pd = dist2(cand_xy(pc), p.x, p.y);
            if (!(d < pd || (d == pd && cand_id(code) < cand_id(pc)))) break;
            lst[pos * NT_THREADS] = (uint16_t)pc;
            --pos;
          }
          lst[pos * NT_THREADS] = (uint16_t)code;
        }
        if (cnt >= k) {
          const double dk = dist2(cand_xy(lst[(k - 1) * NT_THREADS]), p.x, p.y);
          if (dk < ring_lower_bound2(g, p.x, p.y, cx, T.cy, 1)) {
            for (int j = 0; j < k; ++j) {
              const uint32_t code = lst[j * NT_THREADS];
              out_ids[(int64_t)t * k + j] = cand_id(code);
              if (out_dist) out_dist[(int64_t)t * k + j] = __dsqrt_rn(dist2(cand_xy(code), p.x, p.y));
            }
            done = true;
          }
        }
      }
      (void)n_far;
    }
    if (!done) fb_list[atomicAdd(fb_n, 1)] = t;
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct NtBins {
  int32_t *qcell, *qorder, *qcount, *cursor, *fb_list, *fb_n;
  int64_t* qstart;
  int tiles_per_row, n_tiles;
};

static int32_t nt_layout(porrt_ctx* ctx, const GridDev& g, int64_t m, NtBins* B) {
  const int64_t n_cells = (int64_t)g.cells_x * g.cells_y;
  CUDA_TRY(ctx, ctx->nn_tmp[0].ensure((size_t)m * 12 + 64));                 // qcell | qorder | fb_list
  CUDA_TRY(ctx, ctx->nn_tmp[1].ensure((size_t)n_cells * 8 + 64));            // qcount | cursor | fb_n
  CUDA_TRY(ctx, ctx->nn_tmp[2].ensure((size_t)(n_cells + 1) * 8));           // qstart
  B->qcell = ctx->nn_tmp[0].as<int32_t>(); B->qorder = B->qcell + m; B->fb_list = B->qorder + m;
  B->qcount = ctx->nn_tmp[1].as<int32_t>(); B->cursor = B->qcount + n_cells; B->fb_n = B->cursor + n_cells;
  B->qstart = ctx->nn_tmp[2].as<int64_t>();
  B->tiles_per_row = (g.cells_x + NT_W - 1) / NT_W;
  B->n_tiles = B->tiles_per_row * g.cells_y;
  return PORRT_OK;
}

template <bool KNN>
static int32_t nt_bin(porrt_ctx* ctx, const GridDev& g, const double* q_dev, const double* radius_dev, int64_t m, NtBins* B) {
  cudaStream_t st = ctx->stream;
  const int64_t n_cells = (int64_t)g.cells_x * g.cells_y;
  int32_t rc0 = nt_layout(ctx, g, m, B);
  if (rc0) return rc0;
  CUDA_TRY(ctx, cudaMemsetAsync(B->qcount, 0, (size_t)n_cells * 8 + 4, st));
  nt_bin_kernel<KNN><<<div_up(m, 256), 256, 0, st>>>(g, (const double2*)q_dev, radius_dev, m, B->qcell, B->qcount, B->fb_list, B->fb_n);
  LAUNCH_CHECK(ctx);
  int32_t rc = scan_exclusive_i64(ctx, B->qcount, n_cells, B->qstart);
  if (rc) return rc;
  nt_scatter_kernel<<<div_up(m, 256), 256, 0, st>>>(B->qcell, m, B->qstart, B->cursor, B->qorder);
  LAUNCH_CHECK(ctx);
  return PORRT_OK;
}

bool nn_tile_usable(const porrt_ctx* ctx, int64_t m) {
  // worth the binning passes only for real batches; the grid must be addressable with 32-bit cell ids and tile counts
  return m >= 2048 && (int64_t)ctx->cells_x * ctx->cells_y < (1ll << 30) && getenv("PORRT_NN_NO_TILES") == nullptr;
}

// counts (zero-initialised by the caller) / fill for the queries the tiles can serve; *fb_list_out / *fb_n_out: the rest
int32_t nn_tile_radius(porrt_ctx* ctx, const GridDev& g, const double* q_dev, const double* radius_dev, int64_t m,
                       const uint32_t* prefix_dev, const uint64_t* reach_dev, const uint32_t* world_dev, bool fill, int32_t* counts_dev,
                       const int64_t* offsets_dev, int32_t* ids_dev, const int32_t** fb_list_out, int32_t* fb_n_out,
                       const uint32_t* prefix_lo_dev) {
  cudaStream_t st = ctx->stream;
  NtBins B;          // the fill pass reuses the bins of the count pass of the same call (same buffers, same layout)
  if (fill) { int32_t rc = nt_layout(ctx, g, m, &B); if (rc) return rc; }
  if (!fill) {
    int32_t rc = nt_bin<false>(ctx, g, q_dev, radius_dev, m, &B);
    if (rc) return rc;
    nt_radius_kernel<false><<<B.n_tiles, NT_THREADS, 0, st>>>(g, (const double2*)q_dev, radius_dev, prefix_dev, reach_dev, world_dev, B.qorder,
                                                              B.qstart, B.tiles_per_row, counts_dev, nullptr, nullptr, prefix_lo_dev);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->nn_fb_n, B.fb_n, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
  } else {
    nt_radius_kernel<true><<<B.n_tiles, NT_THREADS, 0, st>>>(g, (const double2*)q_dev, radius_dev, prefix_dev, reach_dev, world_dev, B.qorder,
                                                             B.qstart, B.tiles_per_row, nullptr, offsets_dev, ids_dev, prefix_lo_dev);
    LAUNCH_CHECK(ctx);
  }
  *fb_list_out = B.fb_list;
  *fb_n_out = ctx->nn_fb_n;
  if (getenv("PORRT_DEBUG")) fprintf(stderr, "[porrt] nn tiles: %lld queries, %d left to the thread-per-query kernels\n", (long long)m, ctx->nn_fb_n);
  return PORRT_OK;
}

int32_t nn_tile_knn(porrt_ctx* ctx, const GridDev& g, const double* q_dev, int64_t m, int k, const uint64_t* reach_dev,
                    const uint32_t* world_dev, int32_t* ids_dev, double* dist_dev, int32_t* ties_dev, const int32_t** fb_list_out,
                    int32_t* fb_n_out) {
  cudaStream_t st = ctx->stream;
  NtBins B;
  int32_t rc = nt_bin<true>(ctx, g, q_dev, nullptr, m, &B);
  if (rc) return rc;
  const size_t smem = (size_t)NT_CAP_KNN * 16 + (size_t)(NT_CAP_KNN + 24) * 4 + (size_t)(k + NT_SLACK) * NT_THREADS * 2 + (size_t)NT_BINS * NT_THREADS;
  static bool attr_set[16] = {};
  if (!attr_set[ctx->device & 15]) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(nt_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set[ctx->device & 15] = true;
  }
  nt_knn_kernel<<<B.n_tiles, NT_THREADS, smem, st>>>(g, (const double2*)q_dev, k, reach_dev, world_dev, B.qorder, B.qstart, B.tiles_per_row,
                                                     ids_dev, dist_dev, ties_dev, B.fb_list, B.fb_n);
  LAUNCH_CHECK(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->nn_fb_n, B.fb_n, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  *fb_list_out = B.fb_list;
  *fb_n_out = ctx->nn_fb_n;
  if (getenv("PORRT_DEBUG")) fprintf(stderr, "[porrt] nn tiles: %lld queries, %d left to the thread-per-query kernels\n", (long long)m, ctx->nn_fb_n);
  return PORRT_OK;
}
