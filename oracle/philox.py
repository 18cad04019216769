"""Philox4x32-10 and the deterministic draw transforms (TEST INFRASTRUCTURE ONLY).

Independent numpy statement of the counter-based generator the CUDA path uses
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123
`philox4x32_R(10, ...)`).  The reference itself draws from numpy's legacy
MT19937 stream (abs.py:98-99,112-113,150,174-175,188,206,219), which cannot be
consumed in parallel, so stochastic parity with the reference is statistical;
bit parity is between this file and the CUDA kernels.

Counter / key layout (shared contract with csrc/abs_rng.cuh):
    (k0, k1) = (seed & 0xffffffff, seed >> 32)
    per-agent draws: Philox2x32-10, ctr = (agent id, step ^ k1), key = k0 ^ (stream * 0x85EBCA6B)
      stream 0: r0, r1 -> the two move normals
      stream 1: r0 -> income uniform, r1 -> severity-test uniform
      stream 2: r0 -> death-test uniform
    pair draws: Philox4x32-10, ctr = (lower id, step, 2, higher id), key = (k0, k1): w0 -> contagion test

Every floating-point operation below is a single IEEE-754 correctly rounded
add / mul / div / sqrt (no fused multiply-add), so the same sequence of
`__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn` on the GPU gives the same bits.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_MOVE = 0
STREAM_ECONOMY = 1
STREAM_DEATH = 2
STREAM_CONTACT = 2
M2 = np.uint64(0xD256D193)
STREAM_MIX = 0x85EBCA6B


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs broadcast; returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def philox2x32_10(c0, c1, k):
    """Vectorised Philox2x32-10 (Random123 `philox2x32_R(10, ...)`).  Returns two uint32 arrays."""
    c0, c1 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1)]
    k = int(k) & 0xFFFFFFFF
    for _ in range(10):
        prod = M2 * c0
        hi, lo = prod >> np.uint64(32), prod & MASK
        c0, c1 = (hi ^ np.uint64(k) ^ c1), lo
        k = (k + W0) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32)


def split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def u01(w):
    """32 random bits -> double in (0, 1): (w + 0.5) * 2**-32, exact in binary64."""
    return (np.asarray(w, dtype=np.float64) + 0.5) * (2.0 ** -32)


_F = np.float32
_SQRT2 = _F(1.41421354)
_LN2 = _F(0.693147182)
_HALF_PI = _F(1.57079637)
_LOG_C = [_F(1.0 / 7.0), _F(1.0 / 5.0), _F(1.0 / 3.0)]
_SIN_C = [_F(-1.0 / 39916800.0), _F(1.0 / 362880.0), _F(-1.0 / 5040.0), _F(1.0 / 120.0), _F(-1.0 / 6.0)]
_COS_C = [_F(1.0 / 479001600.0), _F(-1.0 / 3628800.0), _F(1.0 / 40320.0), _F(-1.0 / 720.0), _F(1.0 / 24.0),
          _F(-0.5)]


def normal_pair(w0, w1):
    """Two standard normals (float32) from two uint32 words: Box-Muller in exactly rounded binary32.

    radius: u1 = (2*(w0>>9)+1) * 2**-24 in (0,1);  r = sqrt(-2 ln u1) with
            ln u1 = e*ln2 + 2*(s + s^3/3 + s^5/5 + s^7/7), s=(m-1)/(m+1), m in (sqrt.5, sqrt2]
    angle : quadrant = w1>>30, phi = (pi/2) * (2*((w1>>7)&0x7fffff)+1) * 2**-24, Taylor sin/cos on (0, pi/2)
    """
    w0 = np.asarray(w0, dtype=np.uint32)
    w1 = np.asarray(w1, dtype=np.uint32)
    k1 = (w0 >> np.uint32(9)).astype(np.int64)
    u1 = (2 * k1 + 1).astype(np.float32) * _F(2.0 ** -24)
    bits = u1.view(np.uint32)
    e = (bits >> np.uint32(23)).astype(np.int32) - 127
    m = ((bits & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)).view(np.float32)
    big = m > _SQRT2
    m = np.where(big, m * _F(0.5), m).astype(np.float32)
    e = np.where(big, e + 1, e)
    s = (m - _F(1.0)) / (m + _F(1.0))
    s2 = s * s
    p = s2 * _LOG_C[0]
    p = p + _LOG_C[1]
    p = p * s2
    p = p + _LOG_C[2]
    p = p * s2
    p = p + _F(1.0)
    p = p * s
    lnm = p * _F(2.0)
    L = e.astype(np.float32) * _LN2 + lnm
    t = L * _F(-2.0)
    t = np.maximum(t, _F(0.0))
    r = np.sqrt(t).astype(np.float32)

    quad = (w1 >> np.uint32(30)).astype(np.int64)
    f = ((w1 >> np.uint32(7)) & np.uint32(0x7FFFFF)).astype(np.int64)
    tt = (2 * f + 1).astype(np.float32) * _F(2.0 ** -24)
    phi = tt * _HALF_PI
    x2 = phi * phi
    ps = x2 * _SIN_C[0]
    for c in _SIN_C[1:]:
        ps = ps + c
        ps = ps * x2
    ps = ps + _F(1.0)
    sn = ps * phi
    pc = x2 * _COS_C[0]
    for c in _COS_C[1:]:
        pc = pc + c
        pc = pc * x2
    cs = pc + _F(1.0)
    # rotate by quadrant: (cos, sin)(phi + quad*pi/2)
    c_out = np.select([quad == 0, quad == 1, quad == 2], [cs, -sn, -cs], sn).astype(np.float32)
    s_out = np.select([quad == 0, quad == 1, quad == 2], [sn, cs, -sn], -cs).astype(np.float32)
    z0 = (r * c_out).astype(np.float32)
    z1 = (r * s_out).astype(np.float32)
    assert z0.dtype == np.float32 and z1.dtype == np.float32
    return z0, z1


def agent_draws(seed, ids, step, stream=STREAM_MOVE):
    """The two words of (agent, step, stream)."""
    k0, k1 = split_seed(seed)
    return philox2x32_10(ids, (int(step) ^ k1) & 0xFFFFFFFF, k0 ^ ((stream * STREAM_MIX) & 0xFFFFFFFF))


def pair_draws(seed, lo_ids, hi_ids, step):
    k0, k1 = split_seed(seed)
    return philox4x32_10(lo_ids, step, STREAM_CONTACT, hi_ids, k0, k1)
