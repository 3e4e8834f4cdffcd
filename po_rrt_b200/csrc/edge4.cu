// edge4.cu -- batched edge validity, fourth design: same data structures as edge3.cu (class plane staged in shared
// memory, 256-bit block bitmaps, one-IMAD exact minor offset), different work distribution.
//
// ncu on edge3 (profiles/r1_v3_edge_ncu_summary.txt): the kernel is bound by integer issue (ALU pipe 74 %, FMA-heavy
// pipe 49 %, "math pipe throttle" the top stall), and a quarter of its 44 warp instructions per edge were spent on
// FLATTENING (owner search per round, exclusive scans for the queue, 25 % of the strip slots empty because items are
// groups of 8 strips).  Here every lane walks ITS OWN edge strip by strip -- no owner search, no empty slots -- and the
// load balance comes from sorting: a warp takes 256 edges (8 per lane), counting-sorts them by strip count in shared
// memory (64 bins, ~0.3 warp instructions per edge) and then runs 8 groups of 32 edges of similar length in lock step
// (simulated efficiency 0.87 on the c5 workload; 0.44 without the sort).  Undecided strips are still queued and
// resolved by full warps from the block bitmaps; the queue now collects the strips of 256 edges, so its rounds are
// full.  Exactness rules are those of edge3.cu.
#include "edge_common.cuh"

#define E4_NB 256            // edges per warp batch
#define E4_PER_LANE (E4_NB / 32)
#define E4_QB 384            // bitmap-queue entries per warp
#define E4_QG 64             // byte-queue entries per warp
#define E4_BINS 64
#define E4_MAX_WARPS 26

struct WarpMem4 {
  uint32_t recS[E4_NB];      // fixed-point slope S
  uint32_t recN[E4_NB];      // mirrored start minor coordinate n0m | dxo << 16
  uint32_t recC[E4_NB];      // start major coordinate c0 | dirs << 16 | eflags << 20   (dirs: bit0 major axis is i, bit1 major
                             // step -1, bit2 minor step -1; eflags: bit0 start outside the map, bit1 end outside)
  uint32_t hist[E4_BINS];
  uint32_t qb[E4_QB];        // slot | strip << 8
  uint32_t qg[E4_QG];        // slot | strip << 8
  uint16_t z[E4_NB];         // zone codes seen on gray pixels: min | max << 8 (0x00ff: none)
  uint8_t perm[E4_NB];       // slots sorted by strip count, longest first
  uint8_t obst[E4_NB];       // != 0: a blocking pixel is on the line
};

struct EdgeView {
  uint32_t S;
  int n0m, dxo, c0, dirs, eflags;
  int sm_, sn_;              // block-index step per strip / per mirrored minor block (signed)
  int idx0, lo_raw0, n_strips;
};

__device__ __forceinline__ EdgeView load_edge(const WarpMem4& wm, int slot, int cw, int ch, int guard) {
  EdgeView v;
  const uint32_t N = wm.recN[slot], C = wm.recC[slot];
  v.S = wm.recS[slot];
  v.n0m = (int)(N & 0xffffu); v.dxo = (int)(N >> 16);
  v.c0 = (int)(C & 0xffffu); v.dirs = (int)((C >> 16) & 7u); v.eflags = (int)((C >> 20) & 3u);
  const bool major_i = (v.dirs & 1) != 0;
  const int sgn = (v.dirs & 2) ? -1 : 1;
  const int unit_major = major_i ? cw : 1, unit_minor = major_i ? 1 : cw;
  const int b0 = v.c0 >> E3_LOG_BS, b1 = (v.c0 + sgn * v.dxo) >> E3_LOG_BS;
  v.sm_ = sgn * unit_major;
  v.idx0 = guard + b0 * unit_major;
  v.sn_ = unit_minor;
  if (v.dirs & 4) { v.idx0 += ((major_i ? cw : ch) - 1) * unit_minor; v.sn_ = -unit_minor; }
  v.lo_raw0 = (v.dirs & 2) ? (v.c0 & (E3_BS - 1)) - (E3_BS - 1) : -(v.c0 & (E3_BS - 1));
  v.n_strips = (b1 > b0 ? b1 - b0 : b0 - b1) + 1;
  return v;
}

__device__ __forceinline__ void z_update(uint16_t* z, int slot, uint32_t zmin, uint32_t zmax) {   // rare
  uint32_t* w = (uint32_t*)z + (slot >> 1);
  const int sh = (slot & 1) * 16;
  uint32_t old = *w, assumed;
  do {
    assumed = old;
    const uint32_t cur = (assumed >> sh) & 0xffffu;
    const uint32_t nv = min(cur & 0xffu, zmin) | (max(cur >> 8, zmax) << 8);
    old = atomicCAS(w, assumed, (assumed & ~(0xffffu << sh)) | (nv << sh));
  } while (old != assumed);
}

// One persistent CTA per SM; dynamic shared memory: [ class plane | mbarrier (16) | WarpMem4 x warps ]
template <int KIND, bool INDEXED>
__global__ void __launch_bounds__(E4_MAX_WARPS * 32, 1)
edge_validity_v4_kernel(MapDev m, const double2* __restrict__ from, const double2* __restrict__ to, int64_t n,
                        int32_t* __restrict__ out_vid, uint64_t* __restrict__ out_mask,
                        const uint64_t* __restrict__ validities, const int32_t* __restrict__ from_idx,
                        const int32_t* __restrict__ to_idx, unsigned long long* __restrict__ ticket) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* s_mbar = (uint64_t*)(smem + m.plane_bytes);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpMem4& wm = *(WarpMem4*)(smem + m.plane_bytes + 16 + (size_t)wib * sizeof(WarpMem4));
  const uint32_t lt_mask = (1u << lane) - 1u;

  // ---- stage the class plane: bulk-async copies global -> shared, completion counted in bytes on an mbarrier
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_mbar)), "r"((uint32_t)m.plane_bytes) : "memory");
    for (int off = 0; off < m.plane_bytes; off += 32768) {
      const int sz = min(32768, m.plane_bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + off)),
                   "l"((const unsigned char*)m.plane + off), "r"((uint32_t)sz), "r"(smem_u32(s_mbar))
                   : "memory");
    }
  }
  bool plane_ready = false;

  const int cw = m.plane_cw, ch = m.plane_ch, guard = m.plane_guard;
  int qb_n = 0, qg_n = 0;   // queue fill, warp-uniform

  // ---- bitmap pass: drop the strips of edges already blocked, then one lane per strip (see edge3.cu for the arithmetic)
  auto drain_bits = [&]() {
    __syncwarp();
    int live_n = 0;
    for (int q0 = 0; q0 < qb_n; q0 += 32) {
      const int q = q0 + lane;
      uint32_t ent = 0;
      bool live = false;
      if (q < qb_n) { ent = wm.qb[q]; live = wm.obst[ent & 0xffu] == 0; }
      const unsigned lv = __ballot_sync(0xffffffffu, live);
      __syncwarp();
      if (live) wm.qb[live_n + __popc(lv & lt_mask)] = ent;
      live_n += __popc(lv);
      __syncwarp();
    }
    for (int q0 = 0; q0 < live_n; q0 += 32) {
      const int q = q0 + lane;
      if (q < live_n) {
        const uint32_t ent = wm.qb[q];
        const int slot = ent & 0xffu, ts = (int)(ent >> 8);
        if (wm.obst[slot] == 0) {
          const EdgeView v = load_edge(wm, slot, cw, ch, guard);
          const int lo_raw = v.lo_raw0 + ts * E3_BS;
          const int k_lo = max(0, lo_raw), k_hi = min(v.dxo, lo_raw + E3_BS - 1);
          const bool neg_major = (v.dirs & 2) != 0;
          const int bn_a = minor_m((uint32_t)k_lo, v.S, v.n0m) >> E3_LOG_BS;
          const int bn_b = minor_m((uint32_t)k_hi, v.S, v.n0m) >> E3_LOG_BS;
          const int blk_a = v.idx0 - guard + ts * v.sm_ + bn_a * v.sn_;
          const uint32_t* tiles = m.bits + (size_t)(((v.dirs & 1) << 1) | ((v.dirs >> 2) & 1)) * (size_t)m.bits_var_words;
          uint32_t A[8], B[8];
#pragma unroll
          for (int w = 0; w < 8; ++w) B[w] = 0;
          ld256(A, tiles + (size_t)blk_a * 8);
          if (bn_b != bn_a) ld256(B, tiles + (size_t)(blk_a + v.sn_) * 8);
          const int k0 = neg_major ? lo_raw + E3_BS - 1 : lo_raw;          // k at major position t = 0
          const int c_hi = v.n0m - (bn_a << E3_LOG_BS);
          const int t_lo = neg_major ? lo_raw + E3_BS - 1 - k_hi : k_lo - lo_raw;
          const int t_hi = neg_major ? lo_raw + E3_BS - 1 - k_lo : k_hi - lo_raw;
          const uint32_t Vm = ((2u << t_hi) - 1u) & ~((1u << t_lo) - 1u);  // positions that are pixels of this edge
          uint32_t acc = 0;
          uint64_t Y = ((uint64_t)(uint32_t)c_hi << 32) + (uint64_t)((int64_t)k0 * (int64_t)(uint64_t)v.S) + E3_BIAS;
          const uint64_t D = (neg_major ? (uint64_t)0 - (uint64_t)v.S : (uint64_t)v.S) - (1ull << 32);
#pragma unroll
          for (int t = 0; t < E3_BS; ++t) {
            const uint32_t V = (t & 1) ? __byte_perm(A[t >> 1], B[t >> 1], 0x7632) : __byte_perm(A[t >> 1], B[t >> 1], 0x5410);
            const uint32_t x = Vm & (1u << t);
            acc |= __funnelshift_l(x, x, (uint32_t)(Y >> 32)) & V;     // bit t rotated to bit u_t; hi32(Y_t) = u_t - t
            Y += D;
          }
          if (acc) wm.obst[slot] = 1;
        }
      }
      __syncwarp();
    }
    qb_n = 0;
  };
  // ---- byte pass: strips through blocks with gray pixels, per pixel on the fused byte grid, 16 lanes per strip
  auto drain_gray = [&]() {
    __syncwarp();
    for (int q0 = 0; q0 < qg_n; q0 += 2) {
      const int q = q0 + (lane >> 4);
      uint32_t code = 255;
      int slot = 0;
      if (q < qg_n) {
        const uint32_t ent = wm.qg[q];
        slot = ent & 0xffu;
        const int ts = (int)(ent >> 8);
        const EdgeView v = load_edge(wm, slot, cw, ch, guard);
        const int lo_raw = v.lo_raw0 + ts * E3_BS;
        const int k = max(0, lo_raw) + (lane & 15);
        if (k <= min(v.dxo, lo_raw + E3_BS - 1)) {
          const int mk = minor_m((uint32_t)k, v.S, 0);
          const int n0 = (v.dirs & 4) ? ((v.dirs & 1) ? cw : ch) * E3_BS - 1 - v.n0m : v.n0m;   // un-mirror
          const int major = v.c0 + ((v.dirs & 2) ? -k : k), minor = n0 + ((v.dirs & 4) ? -mk : mk);
          const int i = (v.dirs & 1) ? major : minor, j = (v.dirs & 1) ? minor : major;
          code = __ldg(m.grid + tile_addr(i, j, m.tiles_x));
        }
      }
      const bool blocking = KIND == PORRT_DOMAIN_SHELF ? code != 255 : code == 0;
      const bool gray = KIND != PORRT_DOMAIN_SHELF && code != 0 && code != 255;
      const unsigned half_mask = 0xffffu << (lane & 16);
      const unsigned bl = __ballot_sync(0xffffffffu, blocking) & half_mask;
      const unsigned gr_all = __ballot_sync(0xffffffffu, gray);
      if (gr_all) {   // warp-uniform
        uint32_t zmin = gray ? code : 255u, zmax = gray ? code : 0u;
#pragma unroll
        for (int o = 8; o; o >>= 1) {   // within each half warp
          zmin = min(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
          zmax = max(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
        }
        if ((gr_all & half_mask) && (lane & 15) == 0) z_update(wm.z, slot, zmin, zmax);
      }
      if (bl && (lane & 15) == 0) wm.obst[slot] = 1;
    }
    __syncwarp();
    qg_n = 0;
  };

  // batches of E4_NB edges are handed out by a global ticket counter: batch times vary (edge lengths, obstacles), a
  // static split leaves SMs idle at the end (ncu: 24 % of the SM cycles inactive on 4 Mi edges)
  while (true) {
    unsigned long long tk = 0;
    if (lane == 0) tk = atomicAdd(ticket, 1ull);
    const int64_t base = (int64_t)__shfl_sync(0xffffffffu, tk, 0) * E4_NB;
    if (base >= n) break;
    if (!plane_ready) { mbar_wait0(s_mbar); plane_ready = true; }
    if (lane < E4_BINS / 2) { wm.hist[2 * lane] = 0; wm.hist[2 * lane + 1] = 0; }
    __syncwarp();
    // ---- setup: lane handles edges base + j * 32 + lane
    uint32_t keyrank[E4_PER_LANE];   // key | rank within its bin << 8
#pragma unroll 2
    for (int j = 0; j < E4_PER_LANE; ++j) {
      const int slot = j * 32 + lane;
      const int64_t eidx = base + slot;
      uint32_t S = 0, N = 0, C = 0, pre_blocked = 0;
      int key = 0;
      if (eidx < n) {
        const double2 a = INDEXED ? from[from_idx[eidx]] : from[eidx];
        const double2 b = INDEXED ? to[to_idx[eidx]] : to[eidx];
        uint32_t ai, aj, bi, bj;
        to_pixel(m, a.x, a.y, ai, aj);
        to_pixel(m, b.x, b.y, bi, bj);
        const int eflags = ((ai >= (uint32_t)m.H || aj >= (uint32_t)m.W) ? 1 : 0) | ((bi >= (uint32_t)m.H || bj >= (uint32_t)m.W) ? 2 : 0);
        C = (uint32_t)eflags << 20;
        if (!eflags) {
          // line_drawing's octant (Octant::new) reduces to: major axis = the longer delta (ties give the same
          // pixels), unit steps = the signs of the two deltas
          const int di = (int)bi - (int)ai, dj = (int)bj - (int)aj;
          const int adi = abs(di), adj = abs(dj);
          const bool major_i = adi >= adj;
          const int dxo = major_i ? adi : adj, dyo = major_i ? adj : adi;
          const int d_major = major_i ? di : dj, d_minor = major_i ? dj : di;
          const int c0 = major_i ? (int)ai : (int)aj, n0 = major_i ? (int)aj : (int)ai;
          const int dirs = (major_i ? 1 : 0) | (d_major < 0 ? 2 : 0) | (d_minor < 0 ? 4 : 0);
          // the start pixel's block blocks entirely: Obstacle at k = 0, nothing to walk
          const int sb = guard + ((int)ai >> E3_LOG_BS) * cw + ((int)aj >> E3_LOG_BS);
          pre_blocked = plane_class<false>(smem, (uint32_t)sb) == K_BLOCKED ? 1u : 0u;
          if (dxo > 0) S = slope_fixed_point(dyo, dxo);
          const int n0m = (dirs & 4) ? (major_i ? cw : ch) * E3_BS - 1 - n0 : n0;
          N = (uint32_t)n0m | ((uint32_t)dxo << 16);
          C |= (uint32_t)c0 | ((uint32_t)dirs << 16);
          if (!pre_blocked) {
            const int sgn = (dirs & 2) ? -1 : 1;
            const int b0 = c0 >> E3_LOG_BS, b1 = (c0 + sgn * dxo) >> E3_LOG_BS;
            key = min((b1 > b0 ? b1 - b0 : b0 - b1) + 1, E4_BINS - 1);
          }
        }
      }
      wm.recS[slot] = S; wm.recN[slot] = N; wm.recC[slot] = C;
      wm.obst[slot] = (uint8_t)pre_blocked;
      wm.z[slot] = 0x00ffu;
      keyrank[j] = (uint32_t)key | (atomicAdd(&wm.hist[key], 1u) << 8);
    }
    __syncwarp();
    // ---- counting sort by strip count, longest first: bin starts by a warp scan over (bin 63, 62, ..., 0)
    {
      const uint32_t h_hi = wm.hist[E4_BINS - 1 - 2 * lane], h_lo = wm.hist[E4_BINS - 2 - 2 * lane];
      uint32_t incl = h_hi + h_lo;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      __syncwarp();
      wm.hist[E4_BINS - 1 - 2 * lane] = incl - h_hi - h_lo;   // start of the longer bin
      wm.hist[E4_BINS - 2 - 2 * lane] = incl - h_lo;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < E4_PER_LANE; ++j) wm.perm[wm.hist[keyrank[j] & 0xffu] + (keyrank[j] >> 8)] = (uint8_t)(j * 32 + lane);
    }
    __syncwarp();

    // ---- pass 1: groups of 32 edges of similar length, every lane walks its own edge, two strips per iteration.
    // The strips that need the bitmaps / the byte grid are collected in per-lane bit masks (bit = strip index mod 32)
    // and moved to the warp's queues once per group: nothing but class lookups in the loop.
    for (int r = 0; r < E4_PER_LANE; ++r) {
      const int slot = wm.perm[r * 32 + lane];
      const EdgeView v = load_edge(wm, slot, cw, ch, guard);
      int left = (v.eflags || wm.obst[slot] || base + slot >= n) ? 0 : v.n_strips;   // strips still to look at
      if (!__any_sync(0xffffffffu, left > 0)) break;              // sorted: the remaining groups are empty too
      int lo_raw = v.lo_raw0, idx_m = v.idx0, ts = 0;
      int step_k = E3_BS, step_i = v.sm_;
      uint32_t wmask = 0, gmask = 0;
      bool blocked = false;
      // flush: masks -> queues (strip index = ts_base + bit)
      auto flush = [&](int ts_base) {
        if (blocked) { wm.obst[slot] = 1; wmask = 0; }             // the bitmaps cannot change the outcome any more
        const int cnt = __popc(wmask);
        int sc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, sc, o);
          if (lane >= o) sc += t;
        }
        const int tot = __shfl_sync(0xffffffffu, sc, 31);
        if (tot) {
          if (qb_n + tot > E4_QB) drain_bits();
          if (tot <= E4_QB) {
            int pos = qb_n + sc - cnt;
            while (wmask) {
              const int bit = __ffs(wmask) - 1;
              wmask &= wmask - 1;
              wm.qb[pos++] = (uint32_t)slot | ((uint32_t)(ts_base + bit) << 8);
            }
            qb_n += tot;
          } else {                                                   // pathological (> 12 mixed strips per edge on average)
            while (true) {
              const unsigned bal = __ballot_sync(0xffffffffu, wmask != 0);
              if (!bal) break;
              if (wmask) {
                const int bit = __ffs(wmask) - 1;
                wmask &= wmask - 1;
                wm.qb[qb_n + __popc(bal & lt_mask)] = (uint32_t)slot | ((uint32_t)(ts_base + bit) << 8);
              }
              qb_n += __popc(bal);
              if (qb_n > E4_QB - 32) drain_bits();
            }
          }
        }
        while (true) {                                               // rare: strips through gray pixels
          const unsigned bal = __ballot_sync(0xffffffffu, gmask != 0);
          if (!bal) break;
          if (gmask) {
            const int bit = __ffs(gmask) - 1;
            gmask &= gmask - 1;
            wm.qg[qg_n + __popc(bal & lt_mask)] = (uint32_t)slot | ((uint32_t)(ts_base + bit) << 8);
          }
          qg_n += __popc(bal);
          if (qg_n > E4_QG - 32) drain_gray();
        }
      };
      while (true) {
        uint32_t oh[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int k_lo = max(0, lo_raw), k_hi = min(v.dxo, lo_raw + E3_BS - 1);
          const uint32_t ia = (uint32_t)idx_m + (uint32_t)((minor_m((uint32_t)k_lo, v.S, v.n0m) >> E3_LOG_BS) * v.sn_);
          const uint32_t ib = (uint32_t)idx_m + (uint32_t)((minor_m((uint32_t)k_hi, v.S, v.n0m) >> E3_LOG_BS) * v.sn_);
          const uint32_t ca = plane_class<false>(smem, ia), cb = plane_class<false>(smem, ib);
          oh[g] = g < left ? ((1u << ca) | (1u << cb)) : 0u;        // one-hot: bit1 mixed, bit2 blocked, bit3 special
          lo_raw += step_k; idx_m += step_i;
        }
        blocked = blocked || ((oh[0] | oh[1]) & 4u) != 0;
        const uint32_t sh = (uint32_t)ts & 31u;
        wmask |= (((oh[0] & 14u) == 2u ? 1u : 0u) | ((oh[1] & 14u) == 2u ? 2u : 0u)) << sh;
        gmask |= (((oh[0] >> 3) & 1u) | ((oh[1] >> 2) & 2u)) << sh;
        ts += 2;
        left = blocked ? 0 : left - 2;
        // a lane that is done stays where it is (its reads are masked): its indices must not run away while the rest
        // of the group is still walking
        if (left <= 0) { step_k = 0; step_i = 0; }
        const bool more = __any_sync(0xffffffffu, left > 0);
        if ((ts & 31) == 0 && more) flush(ts - 32);                  // edges longer than 32 strips
        if (!more) break;
      }
      flush((ts - 1) & ~31);
    }
    // ---- pass 2: whatever is still queued
    if (qg_n) drain_gray();
    if (qb_n) drain_bits();
    __syncwarp();

    // ---- results, coalesced: lane writes edges base + j * 32 + lane
#pragma unroll 1
    for (int j = 0; j < E4_PER_LANE; ++j) {
      const int slot = j * 32 + lane;
      const int64_t eidx = base + slot;
      if (eidx < n) {
        const int eflags = (int)((wm.recC[slot] >> 20) & 3u);
        int32_t r;
        bool slow = false;
        if (eflags & 1) r = PORRT_PANIC_OOB;               // the first pixel read already panics
        else if (eflags & 2) slow = true;                  // end pixel outside: order of events matters
        else {
          const bool is_blocked = wm.obst[slot] != 0;
          if (KIND == PORRT_DOMAIN_SHELF) r = is_blocked ? R_BLOCKED : R_FREE;   // Low and High obstacle both invalidate the edge
          else {
            const uint32_t zz = wm.z[slot], zmin = zz & 0xffu, zmax = zz >> 8;
            if (zmax != 0 && (zmin != zmax || zmax == 254)) slow = true;         // order of events decides: re-walk
            else r = is_blocked ? R_BLOCKED : (zmax ? (int32_t)zmin - 1 : R_FREE);
          }
        }
        if (slow) {   // rare: exact sequential semantics from the original end points
          const double2 a = INDEXED ? from[from_idx[eidx]] : from[eidx];
          const double2 b = INDEXED ? to[to_idx[eidx]] : to[eidx];
          const EdgeSetup s = make_setup(m, a.x, a.y, b.x, b.y);
          Walker wk;
          wk.load(s);
          r = walk_sequential<KIND>(m, wk);
          if (KIND == PORRT_DOMAIN_SHELF && r == R_LOW) r = R_BLOCKED;
        }
        const int32_t vid = walk_to_validity(m, r);
        out_vid[eidx] = vid;
        if (out_mask) {
          if (m.mask_words == 1) out_mask[eidx] = vid >= 0 ? validities[vid] : 0ull;
          else
            for (int wd = 0; wd < m.mask_words; ++wd)
              out_mask[eidx * m.mask_words + wd] = vid >= 0 ? validities[(int64_t)vid * m.mask_words + wd] : 0ull;
        }
      }
    }
    __syncwarp();
  }
  if (!plane_ready) mbar_wait0(s_mbar);  // a warp without work must not leave while the copy is in flight
}

static const int E4_SMEM_LIMIT = 227 * 1024;

static int edge4_max_warps(const porrt_ctx* ctx) {
  const int64_t room = (int64_t)E4_SMEM_LIMIT - ctx->map.plane_bytes - 16;
  const int64_t w = room / (int64_t)sizeof(WarpMem4);
  return (int)(w > E4_MAX_WARPS ? E4_MAX_WARPS : w);
}

// the class plane has to fit in shared memory next to at least 8 warps' state; coordinates must fit the packed records
bool edge4_usable(const porrt_ctx* ctx) {
  return ctx->map.plane != nullptr && edge4_max_warps(ctx) >= 8 && ctx->map.H <= 32768 && ctx->map.W <= 32768 &&
         ctx->map.plane_guard + (int64_t)ctx->map.plane_cw * ctx->map.plane_ch < (1 << 30);
}

template <int KIND, bool INDEXED>
static int32_t edge4_launch_t(porrt_ctx* ctx, const double2* from, const double2* to, int64_t n, int32_t* out_vid, uint64_t* out_mask,
                              const int32_t* from_idx, const int32_t* to_idx, cudaStream_t st) {
  auto kern = edge_validity_v4_kernel<KIND, INDEXED>;
  static bool attr_set[16] = {};
  if (!attr_set[ctx->device & 15]) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, E4_SMEM_LIMIT));
    attr_set[ctx->device & 15] = true;
  }
  const int max_warps = edge4_max_warps(ctx);
  const int64_t warps_needed = (n + E4_NB - 1) / E4_NB;
  // one CTA per SM; small batches are spread over the SMs with fewer warps each
  int warps = (int)((warps_needed + ctx->sm_count - 1) / ctx->sm_count);
  warps = warps < 4 ? 4 : (warps > max_warps ? max_warps : warps);
  int64_t ctas = (warps_needed + warps - 1) / warps;
  if (ctas > ctx->sm_count) ctas = ctx->sm_count;
  const size_t smem = (size_t)ctx->map.plane_bytes + 16 + (size_t)warps * sizeof(WarpMem4);
  CUDA_TRY(ctx, ctx->d_ticket.ensure(8));
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_ticket.p, 0, 8, st));
  kern<<<(int)ctas, warps * 32, smem, st>>>(ctx->map, from, to, n, out_vid, out_mask, ctx->d_validities.as<uint64_t>(), from_idx, to_idx,
                                            ctx->d_ticket.as<unsigned long long>());
  LAUNCH_CHECK(ctx);
  return PORRT_OK;
}

int32_t edge4_launch(porrt_ctx* ctx, const double* from_dev, const double* to_dev, int64_t n, int32_t* out_vid_dev,
                     uint64_t* out_mask_dev, const int32_t* from_idx_dev, const int32_t* to_idx_dev, cudaStream_t st) {
  const double2* f = (const double2*)from_dev;
  const double2* t = (const double2*)to_dev;
  const bool shelf = ctx->map.kind == PORRT_DOMAIN_SHELF;
  if (from_idx_dev)
    return shelf ? edge4_launch_t<PORRT_DOMAIN_SHELF, true>(ctx, f, t, n, out_vid_dev, out_mask_dev, from_idx_dev, to_idx_dev, st)
                 : edge4_launch_t<PORRT_DOMAIN_DOOR, true>(ctx, f, t, n, out_vid_dev, out_mask_dev, from_idx_dev, to_idx_dev, st);
  return shelf ? edge4_launch_t<PORRT_DOMAIN_SHELF, false>(ctx, f, t, n, out_vid_dev, out_mask_dev, nullptr, nullptr, st)
               : edge4_launch_t<PORRT_DOMAIN_DOOR, false>(ctx, f, t, n, out_vid_dev, out_mask_dev, nullptr, nullptr, st);
}
