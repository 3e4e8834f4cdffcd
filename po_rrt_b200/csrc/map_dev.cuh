// map_dev.cuh -- device helpers shared by the occupancy-grid kernels (map.cu, edge3.cu): pixel mapping (A1), the
// tiled fused grid address, the Bresenham closed form (A5) and the exact sequential walk (A3/A3').
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t tile_addr(int i, int j, int tiles_x) {
  return ((uint32_t)((i >> 3) * tiles_x + (j >> 4)) << 7) | ((uint32_t)(i & 4) << 4) | ((uint32_t)(j & 8) << 2) |
         ((uint32_t)(i & 3) << 3) | (uint32_t)(j & 7);
}

// Map::to_pixel_coordinates (map_io.rs:176-181): one sub, one mul, one sub, saturating cast -- in that order, no FMA
// (the library is compiled with -fmad=false; __dmul_rn/__dsub_rn make the intent explicit).
__device__ __forceinline__ void to_pixel(const MapDev& m, double x, double y, uint32_t& i, uint32_t& j) {
  i = __double2uint_rz(__dsub_rn(m.hm1, __dmul_rn(__dsub_rn(y, m.low1), m.ppm)));  // cvt.rzi.u32.f64 saturates, NaN -> 0 (= Rust `as u32`)
  j = __double2uint_rz(__dmul_rn(__dsub_rn(x, m.low0), m.ppm));
}

// walk results (internal): >= 0 zone id | R_* below | PORRT_PANIC_* codes
#define R_BLOCKED (-1)  // DOOR: Obstacle, SHELF: HighObstacle
#define R_FREE (-5)
#define R_LOW (-6)      // SHELF: LowObstacle

struct EdgeSetup {      // 32 bytes, one per lane, broadcast through shared memory
  int32_t ai, aj;       // start pixel
  int32_t dxo, dyo;     // octant-space deltas (dxo >= dyo >= 0)
  uint32_t m_lo, m_hi;  // floor(2^64 / dxo) + 1 : exact floor(k*dyo/dxo) by one 64-bit mulhi
  int32_t steps;        // packed unit steps: U (major) and V (minor) as (di,dj) in {-1,0,1}, 2 bits each, biased by 1
  int32_t flags;        // bit0: start out of bounds, bit1: end out of bounds
};

// line_drawing::Octant::new + to/from (crate line_drawing 0.8, octant.rs) folded into the two unit steps of the
// closed form  pixel_k = a + k*U + floor(k*dyo/dxo)*V   (checked against the iterator in tests).
__device__ __forceinline__ EdgeSetup make_setup(const MapDev& m, double ax, double ay, double bx, double by) {
  uint32_t ai, aj, bi, bj;
  to_pixel(m, ax, ay, ai, aj);
  to_pixel(m, bx, by, bi, bj);
  EdgeSetup s;
  s.ai = (int32_t)ai; s.aj = (int32_t)aj;
  s.flags = ((ai >= (uint32_t)m.H || aj >= (uint32_t)m.W) ? 1 : 0) | ((bi >= (uint32_t)m.H || bj >= (uint32_t)m.W) ? 2 : 0);
  int32_t dx = (int32_t)bi - (int32_t)ai, dy = (int32_t)bj - (int32_t)aj;
  int oct = 0;
  if (dy < 0) { dx = -dx; dy = -dy; oct += 4; }
  if (dx < 0) { int32_t t = dx; dx = dy; dy = -t; oct += 2; }
  if (dx < dy) { int32_t t = dx; dx = dy; dy = t; oct += 1; }
  s.dxo = dx; s.dyo = dy;
  // from_octant(1,0) and from_octant(0,1) per octant, as (di,dj):
  //  o0 U(1,0) V(0,1) | o1 U(0,1) V(1,0) | o2 U(0,1) V(-1,0) | o3 U(-1,0) V(0,1)
  //  o4 U(-1,0) V(0,-1) | o5 U(0,-1) V(-1,0) | o6 U(0,-1) V(1,0) | o7 U(1,0) V(0,-1)
  int ui, uj, vi, vj;
  switch (oct) {
    case 0: ui = 1; uj = 0; vi = 0; vj = 1; break;
    case 1: ui = 0; uj = 1; vi = 1; vj = 0; break;
    case 2: ui = 0; uj = 1; vi = -1; vj = 0; break;
    case 3: ui = -1; uj = 0; vi = 0; vj = 1; break;
    case 4: ui = -1; uj = 0; vi = 0; vj = -1; break;
    case 5: ui = 0; uj = -1; vi = -1; vj = 0; break;
    case 6: ui = 0; uj = -1; vi = 1; vj = 0; break;
    default: ui = 1; uj = 0; vi = 0; vj = -1; break;
  }
  s.steps = (ui + 1) | ((uj + 1) << 2) | ((vi + 1) << 4) | ((vj + 1) << 6);
  uint64_t M = dx > 1 ? (0xFFFFFFFFFFFFFFFFull / (uint64_t)dx) + 1ull : 0ull;
  s.m_lo = (uint32_t)M; s.m_hi = (uint32_t)(M >> 32);
  return s;
}

struct Walker {  // per-edge state shared by all lanes of the warp
  int32_t ai, aj, dxo, dyo, ui, uj, vi, vj;
  uint64_t M;
  __device__ __forceinline__ void load(const EdgeSetup& s) {
    ai = s.ai; aj = s.aj; dxo = s.dxo; dyo = s.dyo;
    ui = (s.steps & 3) - 1; uj = ((s.steps >> 2) & 3) - 1; vi = ((s.steps >> 4) & 3) - 1; vj = ((s.steps >> 6) & 3) - 1;
    M = ((uint64_t)s.m_hi << 32) | s.m_lo;
  }
  __device__ __forceinline__ void pixel(int32_t k, int32_t& i, int32_t& j) const {
    // floor(k*dyo/dxo); dxo == 1 -> k*dyo, dxo == 0 -> only k == 0 exists
    int32_t mnr = dxo > 1 ? (int32_t)__umul64hi((uint64_t)((uint32_t)k * (uint32_t)dyo), M) : k * dyo;
    i = ai + k * ui + mnr * vi;
    j = aj + k * uj + mnr * vj;
  }
};

// Exact sequential semantics, including the order of the reference's panics; one lane, rare.
template <int KIND>
__device__ __noinline__ int32_t walk_sequential(const MapDev& m, const Walker& w) {
  int32_t prev = -1;
  uint32_t lowest = 255;
  for (int32_t k = 0; k <= w.dxo; ++k) {
    int32_t i, j;
    w.pixel(k, i, j);
    if ((uint32_t)i >= (uint32_t)m.H || (uint32_t)j >= (uint32_t)m.W) return PORRT_PANIC_OOB;
    uint32_t c = __ldg(m.grid + tile_addr(i, j, m.tiles_x));
    if (KIND == PORRT_DOMAIN_SHELF) {
      lowest = min(lowest, c);
      if (lowest == 0) return R_BLOCKED;   // map_shelves_io.rs:199
    } else {
      if (c == 255) continue;
      if (c == 0) return R_BLOCKED;        // map_io.rs:229
      if (c == 254) return PORRT_PANIC_ZONE_UNWRAP;
      int32_t z = (int32_t)c - 1;
      if (prev >= 0 && prev != z) return PORRT_PANIC_MULTI_ZONE;  // map_io.rs:233
      prev = z;
    }
  }
  if (KIND == PORRT_DOMAIN_SHELF) return lowest == 255 ? R_FREE : (lowest >= 127 ? R_LOW : R_BLOCKED);
  return prev >= 0 ? prev : R_FREE;
}

// Warp-cooperative walk of one edge; every lane returns the same result.
template <int KIND>
__device__ __forceinline__ int32_t walk_warp(const MapDev& m, const EdgeSetup& s, int lane) {
  if (s.flags & 1) return PORRT_PANIC_OOB;  // the first pixel read already panics
  Walker w;
  w.load(s);
  bool slow = (s.flags & 2) != 0;           // end pixel outside: order of events matters -> sequential
  int32_t result = R_FREE;
  if (!slow) {
    const int32_t n_px = w.dxo + 1;
    int32_t zone_seen = -1;
    uint32_t lowest = 255;
    bool done = false;
    for (int32_t k0 = 0; k0 < n_px && !done; k0 += 128) {
      uint32_t code[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int32_t k = k0 + c * 32 + lane;
        code[c] = 255;
        if (k < n_px) {
          int32_t i, j;
          w.pixel(k, i, j);
          code[c] = __ldg(m.grid + tile_addr(i, j, m.tiles_x));
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (done) break;
        if (KIND == PORRT_DOMAIN_SHELF) {
          lowest = min(lowest, __reduce_min_sync(0xffffffffu, code[c]));
          if (lowest == 0) done = true;
        } else {
          uint32_t ob = __ballot_sync(0xffffffffu, code[c] == 0);
          uint32_t before = ob ? ((1u << (__ffs(ob) - 1)) - 1u) : 0xffffffffu;  // lanes ahead of the first obstacle
          bool is_gray = code[c] != 0 && code[c] != 255 && ((before >> lane) & 1u);
          uint32_t gray = __ballot_sync(0xffffffffu, is_gray);
          if (gray) {
            uint32_t zmin = __reduce_min_sync(0xffffffffu, is_gray ? code[c] : 255u);
            uint32_t zmax = __reduce_max_sync(0xffffffffu, is_gray ? code[c] : 0u);
            if (zmin != zmax || zmax == 254u || (zone_seen >= 0 && zone_seen != (int32_t)zmin - 1)) { slow = true; done = true; }
            zone_seen = (int32_t)zmin - 1;
          }
          if (ob && !slow) { result = R_BLOCKED; done = true; }
        }
      }
    }
    if (!slow) {
      if (KIND == PORRT_DOMAIN_SHELF) result = lowest == 255 ? R_FREE : (lowest >= 127 ? R_LOW : R_BLOCKED);
      else if (result != R_BLOCKED) result = zone_seen >= 0 ? zone_seen : R_FREE;
    }
  }
  if (slow) {
    int32_t r = 0;
    if (lane == 0) r = walk_sequential<KIND>(m, w);
    result = __shfl_sync(0xffffffffu, r, 0);
  }
  return result;
}

__device__ __forceinline__ int32_t walk_to_validity(const MapDev& m, int32_t r) {
  if (r >= 0) return r;                    // Zone(z) -> Some(z)
  if (r == R_FREE) return m.free_vid;      // Free -> Some(world_validities.len() - 1)
  if (r == R_LOW) return PORRT_INVALID;
  return r;                                // blocked (-1) or panic code
}


// is_state_valid + state_validity (map_io.rs:165-174,487-493 / map_shelves_io.rs:158-163,464-469) of one state
__device__ __forceinline__ int32_t state_validity_of(const MapDev& m, double x, double y) {
  uint32_t i, j;
  to_pixel(m, x, y, i, j);
  if (i >= (uint32_t)m.H || j >= (uint32_t)m.W) return PORRT_PANIC_OOB;
  const uint32_t c = __ldg(m.grid + tile_addr((int)i, (int)j, m.tiles_x));
  if (m.kind == PORRT_DOMAIN_SHELF) return c == 255 ? m.free_vid : PORRT_INVALID;
  return c == 255 ? m.free_vid : (c == 0 ? PORRT_INVALID : (c == 254 ? PORRT_PANIC_ZONE_UNWRAP : (int32_t)c - 1));
}

// One edge's outputs: the validity id as int32 or as a signed byte (ids are < 128: n_validities <= 65; the negative codes
// are the same), and optionally its per-world bitvec = world_validities[id] (map_io.rs:548-550), all zero when invalid.
__device__ __forceinline__ void store_edge_result(const MapDev& m, int64_t e, int32_t vid, int32_t* __restrict__ out_vid,
                                                  int8_t* __restrict__ out_vid8, uint64_t* __restrict__ out_mask,
                                                  const uint64_t* __restrict__ validities) {
  if (out_vid8) out_vid8[e] = (int8_t)vid;
  else out_vid[e] = vid;
  if (out_mask) {
    if (m.mask_words == 1) out_mask[e] = vid >= 0 ? validities[vid] : 0ull;
    else
      for (int wd = 0; wd < m.mask_words; ++wd)
        out_mask[e * m.mask_words + wd] = vid >= 0 ? validities[(int64_t)vid * m.mask_words + wd] : 0ull;
  }
}
