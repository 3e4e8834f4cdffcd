// pcg64.h -- the reference's random streams on the host: rand_pcg 0.3 `Pcg64` (Lcg128Xsl64) seeded by
// rand_core 0.6 `SeedableRng::seed_from_u64`, with rand 0.8's `gen_range` for f64 and usize (sample_space.rs:18,33,47,58).
// Sequential by nature: the streams decide WHAT is computed (samples, worlds, refiner trials), never how; they stay on the host.
#pragma once
#include <stdint.h>
#include <string.h>

struct Pcg64 {
  typedef unsigned __int128 u128;
  u128 state, inc;
  static u128 mul() { return ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull; }
  // seed_from_u64: a PCG32 stream (XSH-RR) fills the 32-byte seed word by word; from_seed reads state | stream, stream forced odd
  static Pcg64 seed_from_u64(uint64_t s) {
    const uint64_t M = 6364136223846793005ull, I = 11634580027462260723ull;
    uint32_t w[8];
    for (int c = 0; c < 8; ++c) {
      s = s * M + I;
      const uint32_t x = (uint32_t)(((s >> 18) ^ s) >> 27), rot = (uint32_t)(s >> 59);
      w[c] = (x >> rot) | (x << ((32 - rot) & 31));
    }
    return from_seed_words(w);
  }
  // SeedableRng::from_seed for Lcg128Xsl64: 32 seed bytes, little endian: state (16 bytes) | stream (16 bytes, forced odd)
  static Pcg64 from_seed_words(const uint32_t w[8]) {
    Pcg64 p;
    p.state = (u128)((uint64_t)w[0] | ((uint64_t)w[1] << 32)) | ((u128)((uint64_t)w[2] | ((uint64_t)w[3] << 32)) << 64);
    p.inc = ((u128)((uint64_t)w[4] | ((uint64_t)w[5] << 32)) | ((u128)((uint64_t)w[6] | ((uint64_t)w[7] << 32)) << 64)) | 1;
    p.state = p.state + p.inc;               // Lcg128Xsl64::from_state_incr
    p.state = p.state * mul() + p.inc;
    return p;
  }
  uint64_t next_u64() {
    state = state * mul() + inc;
    const uint32_t rot = (uint32_t)(state >> 122);
    const uint64_t x = (uint64_t)(state >> 64) ^ (uint64_t)state;
    return (x >> rot) | (x << ((64 - rot) & 63));
  }
  // gen_range(0..range) for usize: UniformInt::sample_single_inclusive, widening multiply with a rejection zone
  uint64_t below(uint64_t range) {
    if (range == 0) return next_u64();
    const uint64_t zone = (range << __builtin_clzll(range)) - 1;
    for (;;) {
      const u128 m = (u128)next_u64() * range;
      if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
    }
  }
  // gen_range(low..high) for f64: UniformFloat::sample_single -- 52 random mantissa bits -> [1, 2) -> [0, 1) * scale + low,
  // retried with `scale` one ulp smaller in the rounding corner res == high
  double range_f64(double low, double high) {
    double scale = high - low;
    for (;;) {
      const uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ull;
      double v12;
      memcpy(&v12, &bits, 8);
      const double res = (v12 - 1.0) * scale + low;
      if (res < high) return res;
      uint64_t sb;
      memcpy(&sb, &scale, 8);
      sb -= 1;
      memcpy(&scale, &sb, 8);
    }
  }
};
