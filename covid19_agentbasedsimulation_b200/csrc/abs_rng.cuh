// Counter-based draws for the agent step: Philox4x32-10 keyed (seed, agent id, step).
//
// The reference consumes one sequential MT19937 stream (np.random.randn/rand/random at
// covid_abs/abs.py:174-175,188,206,219,150), which a data-parallel step cannot do; every
// draw here is a pure function of (seed, id, step, stream) so that the result does not
// depend on how agents are laid out in memory or spread over GPUs.
//
// The uniform -> normal transform is Box-Muller written only with correctly rounded
// binary32 add/mul/div/sqrt (no FMA contraction, no libdevice transcendentals), so it is
// reproducible bit for bit by an independent CPU statement of the same arithmetic.
#pragma once
#include <stdint.h>

namespace cabs {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// Per-agent draws need two words at a time, so they use the two-word variant (Philox2x32-10: half the multiplies
// of 4x32 per call); the counter is (agent id, step ^ seed_hi) and the key is seed_lo mixed with the stream:
constexpr uint32_t kStreamMove = 0;     // r0, r1: the two move normals            (abs.py:174-175)
constexpr uint32_t kStreamEconomy = 1;  // r0: income uniform, r1: severity test   (abs.py:188, 206)
constexpr uint32_t kStreamDeath = 2;    // r0: death test                          (abs.py:219)
// Pair draws stay on Philox4x32-10: ctr = (lower id, step, 2, higher id), key = seed; w0: contagion test (abs.py:150)
constexpr uint32_t kStreamContact = 2;
constexpr uint32_t kPhilox2M = 0xD256D193u;
constexpr uint32_t kStreamMix = 0x85EBCA6Bu;   // odd, unrelated to the round-key increment

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi = __umulhi(kPhilox2M, c0), lo = kPhilox2M * c0;
    c0 = hi ^ k ^ c1;
    c1 = lo;
    k += kPhiloxW0;
  }
  return make_uint2(c0, c1);
}
// the two words of (agent, step, stream) under seed = (k0, k1)
__device__ __forceinline__ uint2 agent_draw(uint32_t id, uint32_t step, uint32_t stream, uint32_t k0, uint32_t k1) {
  return philox2x32_10(id, step ^ k1, k0 ^ (stream * kStreamMix));
}

// 32 random bits -> double in (0,1): (w + 0.5) * 2^-32 (exact in binary64).
__device__ __forceinline__ double u01(uint32_t w) { return __dmul_rn(__dadd_rn((double)w, 0.5), 0x1p-32); }

// Two standard normals from two words.  Every operation is one IEEE-rounded binary32 op.
__device__ __forceinline__ void normal_pair(uint32_t w0, uint32_t w1, float &z0, float &z1) {
  // radius: u1 = (2k+1) 2^-24, k = 23 high bits;  ln u1 = e ln2 + 2 atanh-series((m-1)/(m+1))
  const float u1 = __fmul_rn((float)(int)(2u * (w0 >> 9) + 1u), 0x1p-24f);
  const uint32_t bits = __float_as_uint(u1);
  int e = (int)(bits >> 23) - 127;
  float m = __uint_as_float((bits & 0x7FFFFFu) | 0x3F800000u);
  if (m > 1.41421354f) {
    m = __fmul_rn(m, 0.5f);
    e += 1;
  }
  const float s = __fdiv_rn(__fadd_rn(m, -1.0f), __fadd_rn(m, 1.0f));
  const float s2 = __fmul_rn(s, s);
  float p = __fmul_rn(s2, (float)(1.0 / 7.0));
  p = __fadd_rn(p, (float)(1.0 / 5.0));
  p = __fmul_rn(p, s2);
  p = __fadd_rn(p, (float)(1.0 / 3.0));
  p = __fmul_rn(p, s2);
  p = __fadd_rn(p, 1.0f);
  p = __fmul_rn(p, s);
  const float lnm = __fmul_rn(p, 2.0f);
  const float L = __fadd_rn(__fmul_rn((float)e, 0.693147182f), lnm);
  const float t = fmaxf(__fmul_rn(L, -2.0f), 0.0f);
  const float r = __fsqrt_rn(t);
  // angle: quadrant from the top 2 bits, phi in (0, pi/2) from the next 23
  const uint32_t quad = w1 >> 30;
  const float tt = __fmul_rn((float)(int)(2u * ((w1 >> 7) & 0x7FFFFFu) + 1u), 0x1p-24f);
  const float phi = __fmul_rn(tt, 1.57079637f);
  const float x2 = __fmul_rn(phi, phi);
  float ps = __fmul_rn(x2, (float)(-1.0 / 39916800.0));
  ps = __fmul_rn(__fadd_rn(ps, (float)(1.0 / 362880.0)), x2);
  ps = __fmul_rn(__fadd_rn(ps, (float)(-1.0 / 5040.0)), x2);
  ps = __fmul_rn(__fadd_rn(ps, (float)(1.0 / 120.0)), x2);
  ps = __fmul_rn(__fadd_rn(ps, (float)(-1.0 / 6.0)), x2);
  const float sn = __fmul_rn(__fadd_rn(ps, 1.0f), phi);
  float pc = __fmul_rn(x2, (float)(1.0 / 479001600.0));
  pc = __fmul_rn(__fadd_rn(pc, (float)(-1.0 / 3628800.0)), x2);
  pc = __fmul_rn(__fadd_rn(pc, (float)(1.0 / 40320.0)), x2);
  pc = __fmul_rn(__fadd_rn(pc, (float)(-1.0 / 720.0)), x2);
  pc = __fmul_rn(__fadd_rn(pc, (float)(1.0 / 24.0)), x2);
  pc = __fmul_rn(__fadd_rn(pc, -0.5f), x2);
  const float cs = __fadd_rn(pc, 1.0f);
  const float co = (quad == 0) ? cs : (quad == 1) ? -sn : (quad == 2) ? -cs : sn;
  const float so = (quad == 0) ? sn : (quad == 1) ? cs : (quad == 2) ? -sn : -cs;
  z0 = __fmul_rn(r, co);
  z1 = __fmul_rn(r, so);
}

// largest |z| normal_pair can return: sqrt(-2 ln 2^-24) = 5.7673..., times |cos|,|sin| <= 1 (+ rounding)
constexpr float kMaxAbsNormal = 5.78f;

}  // namespace cabs
