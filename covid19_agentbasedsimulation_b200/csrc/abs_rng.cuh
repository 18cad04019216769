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

// ------------------------------------------------------------------------------------------
// The move of abs.py:174-175 only needs  ix = int(z0 * amp), iy = int(z1 * amp)  (truncation toward zero), not the
// normals themselves.  `normal_pair` above DEFINES z0, z1 (about 110 instructions of correctly rounded binary32);
// `displacement_fast` evaluates the same Box-Muller map on the special-function unit (lg2 / rsqrt / sin / cos
// .approx, about 35 instructions) and proves, per draw, that truncation gives the same integers: the approximate
// product v differs from the defining one by at most `e` (bound below), so whenever v is farther than e from every
// non-zero integer both truncate alike.  Draws that fail the test (a few in 10^5) take the defining path.  The result
// is therefore bit-identical to normal_pair + truncation for every input; tests/test_gpu_normals.py measures the
// actual errors exhaustively on the device (all 2^23 radius inputs, all 2^25 angle inputs) against the constants.
//   |r_f - r| <= kFastA / r_f + kFastB * r_f     (lg2.approx: 2^-22 absolute on (0.5, 2), 2^-22 relative elsewhere;
//                                                 rsqrt.approx 2^-22; the defining series / roundings ~2^-23)
//   |c_f - c|, |s_f - s| <= kFastC                (sin/cos.approx: 2^-20.9 absolute on [-pi, pi]; angle rounding 3e-7)
// each constant carries a factor >= 2 over the measured maximum.
constexpr float kFastA = 8e-7f;
constexpr float kFastB = 1.0e-6f;
constexpr float kFastC = 2.5e-6f;
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// radius and (cos, sin) of the fast evaluation, separately (also what the self-test compares with normal_pair's)
__device__ __forceinline__ void fast_radius(uint32_t w0, float &r, float &rs) {
  const float u1 = __fmul_rn((float)(int)(2u * (w0 >> 9) + 1u), 0x1p-24f);
  const float t = __fmul_rn(mufu_lg2(u1), -1.38629436f);   // -2 ln u1
  rs = mufu_rsqrt(t);
  r = __fmul_rn(t, rs);
}
__device__ __forceinline__ void fast_angle(uint32_t w1, float &co, float &so) {
  // the 25 angle bits (quadrant : fraction) as one signed odd integer: angle - pi in (-pi, pi)
  const int kk = (int)(w1 >> 7) - (1 << 24);
  const float ang = __fmul_rn((float)(2 * kk + 1), 0x1p-24f * 1.57079637f);
  co = -mufu_cos(ang);
  so = -mufu_sin(ang);
}
// true: (ix, iy) are certain to equal int(z0 * amp), int(z1 * amp) of normal_pair; false: call the defining path
__device__ __forceinline__ bool displacement_fast(uint32_t w0, uint32_t w1, float amp, int &ix, int &iy) {
  float r, rs, co, so;
  fast_radius(w0, r, rs);
  fast_angle(w1, co, so);
  const float ra = __fmul_rn(r, amp);
  const float v0 = __fmul_rn(ra, co), v1 = __fmul_rn(ra, so);
  ix = __float2int_rz(v0);
  iy = __float2int_rz(v1);
  // |v_f - v| <= |amp| (r (kFastC + kFastB + 2^-21) + kFastA / r)
  const float e = __fmul_rn(fabsf(amp), __fmaf_rn(r, kFastC + kFastB + 0x1p-21f, __fmul_rn(rs, kFastA)));
  const float n0 = rintf(v0), n1 = rintf(v1);
  const bool safe0 = fabsf(__fadd_rn(v0, -n0)) > e || n0 == 0.0f;   // NaN (t <= 0 after rounding) compares false
  const bool safe1 = fabsf(__fadd_rn(v1, -n1)) > e || n1 == 0.0f;
  return safe0 && safe1 && e < 0.5f;   // n == 0 is only conclusive while |v_f| + e < 1
}
// int(z0 * amp), int(z1 * amp): the defining evaluation
__device__ __forceinline__ void displacement_exact(uint32_t w0, uint32_t w1, float amp, int &ix, int &iy) {
  float z0, z1;
  normal_pair(w0, w1, z0, z1);
  ix = __float2int_rz(__fmul_rn(z0, amp));  // int() truncates toward zero (abs.py:174-175)
  iy = __float2int_rz(__fmul_rn(z1, amp));
}

// ------------------------------------------------------------------------------------------
// Self-test of displacement_fast (cabs_selftest_normals): exhaustive over the inputs of each factor.
//   out[0] max over the 2^23 radius inputs of |r_f - r| / (kFastA / r_f + kFastB r_f)      (must stay < 1)
//   out[1] max over the 2^25 angle inputs of max(|c_f - c|, |s_f - s|) / kFastC            (must stay < 1)
//   out[2] radius inputs whose fast evaluation is not finite (they always take the defining path)
//   out[3] draws of the sampled (w0, w1, amp) sweep that the fast path accepted
//   out[4] accepted draws whose integers differ from the defining path                       (must be 0)
//   out[5] draws of the sweep
static __global__ void k_selftest_normals(unsigned long long *out, unsigned sweep_log2) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  float worst_r = 0.f, worst_c = 0.f;
  unsigned bad_r = 0;
  for (unsigned k = tid; k < (1u << 23); k += stride) {
    float z0, z1, r, rs;
    normal_pair(k << 9, 0x00000040u, z0, z1);   // angle bits 0 -> phi = 2^-24 pi/2: cos rounds to 1, z0 = r exactly
    fast_radius(k << 9, r, rs);
    if (!(r == r) || !(rs == rs) || rs > 3.0e38f) { ++bad_r; continue; }
    worst_r = fmaxf(worst_r, fabsf(r - z0) / (kFastA * rs + kFastB * r));
  }
  float rr, unused;
  normal_pair(0x80000000u, 0x00000040u, rr, unused);   // u1 ~ 0.5, angle ~ 0: rr = sqrt(2 ln 2)
  for (unsigned k = tid; k < (1u << 25); k += stride) {
    float z0, z1, co, so;
    normal_pair(0x80000000u, k << 7, z0, z1);          // the same radius for every angle k
    fast_angle(k << 7, co, so);
    // z = fl(r c): compare c_f with z / r in binary64 (the rounding of the product is part of the allowance)
    const double c = (double)z0 / (double)rr, s = (double)z1 / (double)rr;
    worst_c = fmaxf(worst_c, (float)(fmax(fabs((double)co - c), fabs((double)so - s)) / (double)kFastC));
  }
  unsigned long long accepted = 0, wrong = 0, total = 0;
  const float amps[8] = {0.3f, 1.5f, 5.0f, 10.0f, 77.0f, 1000.0f, 5000.0f, -5.0f};
  for (unsigned long long k = tid; k < (1ull << sweep_log2); k += stride) {
    const uint2 w = philox2x32_10((uint32_t)k, (uint32_t)(k >> 32) ^ 0x5EEDu, 0xC0FFEEu);
    const float amp = amps[k & 7u];
    int fx, fy, ex, ey;
    const bool ok = displacement_fast(w.x, w.y, amp, fx, fy);
    displacement_exact(w.x, w.y, amp, ex, ey);
    ++total;
    if (ok) {
      ++accepted;
      if (fx != ex || fy != ey) ++wrong;
    }
  }
  atomicMax(reinterpret_cast<unsigned *>(&out[0]), __float_as_uint(worst_r));
  atomicMax(reinterpret_cast<unsigned *>(&out[1]), __float_as_uint(worst_c));
  atomicAdd(&out[2], (unsigned long long)bad_r);
  atomicAdd(&out[3], accepted);
  atomicAdd(&out[4], wrong);
  atomicAdd(&out[5], total);
}

}  // namespace cabs
