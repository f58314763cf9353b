// Counter-based RNG (Philox4x32-10) and the keyed bijection used for the epoch shuffle.
// The CPU restatement is oracle/philox.py; the two must stay bit-identical.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NCF_HD __host__ __device__ __forceinline__
#else
#define NCF_HD inline
#endif

struct Philox4 {
  uint32_t x, y, z, w;
};

NCF_HD uint32_t ncf_mulhi32(uint32_t a, uint32_t b) {
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
}

NCF_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                             uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = ncf_mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = ncf_mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o = {c0, c1, c2, c3};
  return o;
}

NCF_HD uint32_t ncf_fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu;
  h ^= h >> 13; h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

// Keyed bijection of [0, S): balanced 6-round Feistel network over 2^(2*half_bits) >= S with
// cycle walking.  Round keys are 6 Philox words of (seed, epoch).
struct ShufflePerm {
  uint64_t S;
  uint32_t half_bits;
  uint32_t rk[6];
};

NCF_HD ShufflePerm make_shuffle_perm(uint64_t S, uint64_t seed, uint64_t epoch) {
  ShufflePerm p;
  p.S = S;
  uint32_t bits = 2;
  while (bits < 64 && (1ull << bits) < S) ++bits;
  p.half_bits = (bits + 1) / 2;
  Philox4 a = philox4x32_10(0u, 0u, 0x53485546u /*'SHUF'*/, (uint32_t)epoch, (uint32_t)seed,
                            (uint32_t)(seed >> 32));
  Philox4 b = philox4x32_10(1u, 0u, 0x53485546u, (uint32_t)epoch, (uint32_t)seed,
                            (uint32_t)(seed >> 32));
  p.rk[0] = a.x; p.rk[1] = a.y; p.rk[2] = a.z; p.rk[3] = a.w; p.rk[4] = b.x; p.rk[5] = b.y;
  return p;
}

NCF_HD uint64_t shuffle_perm_apply(const ShufflePerm& p, uint64_t q) {
  const uint32_t mask = (p.half_bits >= 32) ? 0xffffffffu : ((1u << p.half_bits) - 1u);
  uint64_t x = q;
  do {
    uint32_t L = (uint32_t)(x >> p.half_bits) & mask;
    uint32_t R = (uint32_t)x & mask;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      uint32_t t = L ^ (ncf_fmix32(R + p.rk[r]) & mask);
      L = R;
      R = t;
    }
    x = ((uint64_t)L << p.half_bits) | (uint64_t)R;
  } while (x >= p.S);
  return x;
}
