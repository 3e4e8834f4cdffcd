/* c_abi_harness.c -- calls libporrt_b200.so from plain C99 through include/porrt_b200.h, the way a foreign caller (the Rust
 * crate's extern "C" block) does: no Python, no ctypes marshalling in between.
 * Built and run by tests/test_gpu_parity.py::test_c_abi_harness (gcc -std=c99, -m gpu).  Checks, all against values known
 * without any oracle:
 *   1. context + map upload, map info, world validities of a 1-zone door map (map_io.rs:198-214)
 *   2. state / edge validity on a hand-made map whose answers can be read off the picture (incl. direction + world masks)
 *   3. vertices_set + vertices_append, radius / 1-NN / k-NN against brute force in C (sqrt(d2) <= r, ties by index)
 *   4. porrt_sssp_worlds on the reference's golden grid graph (pto_graph.rs:446-486, :640-656)
 *   5. error behaviour: calls before the map / vertex set exist return their status codes, nothing aborts
 * Prints "c_abi_harness: ok" and exits 0, or the first failing check and exits 1. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "porrt_b200.h"

#define CHECK(cond)                                                                      \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      fprintf(stderr, "c_abi_harness: FAILED %s (line %d): %s\n", #cond, __LINE__,       \
              ctx ? porrt_last_error(ctx) : "");                                        \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

enum { N = 64 }; /* map side in pixels; world [-1, 1]^2, 32 px per unit */

static uint64_t lcg(uint64_t* s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return *s >> 11; }
static double urand(uint64_t* s) { return (double)(lcg(s) & 0xfffffffffffffull) / 4503599627370496.0 * 2.0 - 1.0; }

int main(void) {
  porrt_ctx* ctx = NULL;
  /* 5a. nothing exists yet */
  CHECK(porrt_ctx_destroy(NULL) == PORRT_ERR_INVALID_ARG);
  CHECK(porrt_ctx_create(0, &ctx) == PORRT_OK && ctx != NULL);
  {
    double p[2] = {0.0, 0.0};
    int32_t v = 0;
    int64_t offs[2], total = 0;
    int32_t ids[4];
    double r = 0.1;
    CHECK(porrt_state_validity(ctx, p, 1, &v) == PORRT_ERR_NO_MAP);
    CHECK(porrt_radius_query(ctx, p, &r, 1, NULL, NULL, 1, NULL, offs, ids, 4, &total) == PORRT_ERR_NO_VERTICES);
  }

  /* 1. a 64 x 64 door map: free (255) everywhere, a wall (0) in pixel columns 30..33 with a door (gray 128, zone 0) in pixel rows
   *    28..35.  Row 0 is the top of the image: y = 1 - (row + 0.5) / 32. */
  static uint8_t occ[N * N], zone[N * N];
  memset(occ, 255, sizeof(occ));
  memset(zone, 255, sizeof(zone));
  for (int i = 0; i < N; ++i)
    for (int j = 30; j <= 33; ++j) {
      const int door = i >= 28 && i <= 35;
      occ[i * N + j] = door ? 128 : 0;
      if (door) zone[i * N + j] = 0;
    }
  const double low[2] = {-1.0, -1.0}, up[2] = {1.0, 1.0};
  CHECK(porrt_map_upload(ctx, occ, zone, N, N, low, up, PORRT_DOMAIN_DOOR, 0.5) == PORRT_OK);
  int32_t n_zones = 0, n_worlds = 0, n_validities = 0, mask_words = 0;
  CHECK(porrt_map_info(ctx, &n_zones, &n_worlds, &n_validities, &mask_words) == PORRT_OK);
  CHECK(n_zones == 1 && n_worlds == 2 && n_validities == 2 && mask_words == 1);
  uint64_t wv[2];
  CHECK(porrt_map_world_validities(ctx, wv) == PORRT_OK);
  /* validity 0 = "through zone 0": only the world in which door 0 is open (world 1, map_io.rs:198-214); validity 1 = free: both */
  CHECK(wv[0] == 0x2ull && wv[1] == 0x3ull);

  /* 2. states and edges */
  {
    const double st[8] = {-0.5, 0.0, /* free */ 0.0, 0.9, /* wall */ 0.0, 0.0, /* door */ 0.5, -0.5 /* free */};
    int32_t v[4];
    CHECK(porrt_state_validity(ctx, st, 4, v) == PORRT_OK);
    CHECK(v[0] == 1 && v[1] == PORRT_INVALID && v[2] == 0 && v[3] == 1);
    /* left room -> right room through the door, through the wall, inside the left room, and the door edge reversed */
    const double from[8] = {-0.5, 0.0, -0.5, 0.8, -0.9, -0.9, 0.5, 0.0};
    const double to[8] = {0.5, 0.0, 0.5, 0.8, -0.2, 0.9, -0.5, 0.0};
    int32_t ev[4];
    uint64_t em[4];
    int8_t ev8[4];
    CHECK(porrt_edge_validity(ctx, from, to, 4, ev, em) == PORRT_OK);
    CHECK(ev[0] == 0 && ev[1] == PORRT_INVALID && ev[2] == 1 && ev[3] == 0);
    CHECK(em[0] == 0x2ull && em[1] == 0 && em[2] == 0x3ull && em[3] == 0x2ull);
    CHECK(porrt_edge_validity_i8(ctx, from, to, 4, ev8) == PORRT_OK);
    for (int k = 0; k < 4; ++k) CHECK(ev8[k] == (int8_t)ev[k]);
  }

  /* 3. vertex set grown by append; queries against brute force */
  {
    enum { V = 2000, Q = 200, K = 4 };
    static double pts[2 * V], q[2 * Q], rad[Q];
    uint64_t seed = 12345;
    for (int i = 0; i < 2 * V; ++i) pts[i] = urand(&seed);
    for (int i = 0; i < 2 * Q; ++i) q[i] = urand(&seed);
    for (int i = 0; i < Q; ++i) rad[i] = 0.1;
    CHECK(porrt_vertices_set(ctx, pts, 500, 0.1) == PORRT_OK);
    CHECK(porrt_vertices_append(ctx, pts + 2 * 500, 1) == PORRT_OK);
    CHECK(porrt_vertices_append(ctx, pts + 2 * 501, V - 501) == PORRT_OK);
    int64_t nv = 0;
    CHECK(porrt_vertices_count(ctx, &nv) == PORRT_OK && nv == V);
    static int64_t offs[Q + 1];
    static int32_t ids[Q * 256], nn[Q], ties[Q], knn[Q * K];
    static double nnd[Q], knnd[Q * K];
    int64_t total = 0;
    CHECK(porrt_radius_query(ctx, q, rad, Q, NULL, NULL, 1, NULL, offs, ids, Q * 256, &total) == PORRT_OK);
    CHECK(porrt_nearest(ctx, q, Q, NULL, 1, NULL, nn, nnd, ties) == PORRT_OK);
    CHECK(porrt_knn(ctx, q, Q, K, knn, knnd) == PORRT_OK);
    int64_t at = 0;
    for (int i = 0; i < Q; ++i) {
      CHECK(offs[i] == at);
      int best = -1;
      double bd = INFINITY;
      for (int j = 0; j < V; ++j) {
        const double dx = pts[2 * j] - q[2 * i], dy = pts[2 * j + 1] - q[2 * i + 1];
        const double d = sqrt(dx * dx + dy * dy); /* common.rs:203-213 */
        if (d <= rad[i]) { CHECK(at < total && ids[at] == j); ++at; }
        if (d < bd) { bd = d; best = j; }
      }
      CHECK(nn[i] == best && nnd[i] == bd && ties[i] == 1);
      CHECK(knn[i * K] == best && knnd[i * K] == bd);
      for (int k = 1; k < K; ++k) CHECK(knnd[i * K + k] >= knnd[i * K + k - 1]);
    }
    CHECK(at == total && offs[Q] == total);
    /* too small an output buffer: the status says so and the needed size comes back */
    int64_t need = 0;
    CHECK(porrt_radius_query(ctx, q, rad, Q, NULL, NULL, 1, NULL, offs, ids, 1, &need) == PORRT_ERR_CAPACITY && need == total);
  }

  /* 4. dijkstra golden vectors (pto_graph.rs:640-656) on the 3 x 3 grid graph (:446-486), children adjacency as CSR */
  {
    static const double xy[18] = {0, 0, 1, 0, 2, 0, 0, 1, 1, 1, 2, 1, 0, 2, 1, 2, 2, 2};
    static const int pairs[12][2] = {{0, 1}, {1, 2}, {0, 3}, {1, 4}, {2, 5}, {3, 4}, {4, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 7}, {7, 8}};
    int deg[9] = {0};
    int64_t row_ptr[10] = {0};
    int32_t col[24];
    for (int e = 0; e < 12; ++e) { ++deg[pairs[e][0]]; ++deg[pairs[e][1]]; }
    for (int u = 0; u < 9; ++u) row_ptr[u + 1] = row_ptr[u] + deg[u];
    int fill[9] = {0};
    for (int e = 0; e < 12; ++e) { /* add_bi_edge(a, b): children of a gets b, children of b gets a, in call order */
      const int a = pairs[e][0], b = pairs[e][1];
      col[row_ptr[a] + fill[a]++] = b;
      col[row_ptr[b] + fill[b]++] = a;
    }
    double dist[9];
    int32_t sweeps = 0;
    {
      const int64_t fptr[2] = {0, 1};
      const int32_t fin[1] = {8};
      static const double want[9] = {4, 3, 2, 3, 2, 1, 2, 1, 0};
      CHECK(porrt_sssp_worlds(ctx, 9, row_ptr, col, xy, NULL, NULL, 0, 0, 0, fptr, fin, dist, &sweeps) == PORRT_OK);
      for (int u = 0; u < 9; ++u) CHECK(dist[u] == want[u]);
    }
    {
      const int64_t fptr[2] = {0, 2};
      const int32_t fin[2] = {7, 5};
      static const double want[9] = {3, 2, 1, 2, 1, 0, 1, 0, 1};
      CHECK(porrt_sssp_worlds(ctx, 9, row_ptr, col, xy, NULL, NULL, 0, 0, 0, fptr, fin, dist, &sweeps) == PORRT_OK);
      for (int u = 0; u < 9; ++u) CHECK(dist[u] == want[u]);
    }
    {
      const int64_t fptr[2] = {0, 0};
      CHECK(porrt_sssp_worlds(ctx, 9, row_ptr, col, xy, NULL, NULL, 0, 0, 0, fptr, NULL, dist, &sweeps) == PORRT_OK);
      for (int u = 0; u < 9; ++u) CHECK(isinf(dist[u]) && dist[u] > 0);
    }
    /* 5b. a malformed graph is refused, not dereferenced */
    col[3] = 99;
    {
      const int64_t fptr[2] = {0, 1};
      const int32_t fin[1] = {8};
      CHECK(porrt_sssp_worlds(ctx, 9, row_ptr, col, xy, NULL, NULL, 0, 0, 0, fptr, fin, dist, &sweeps) == PORRT_ERR_INVALID_ARG);
      CHECK(strlen(porrt_last_error(ctx)) > 0);
    }
  }
  CHECK(porrt_ctx_launch_count(ctx) > 0);
  CHECK(porrt_ctx_destroy(ctx) == PORRT_OK);
  printf("c_abi_harness: ok\n");
  return 0;
}
