"""Known-answer tests of the oracle's Philox4x32-10 / Philox2x32-10 (Random123 kat_vectors) and of the draw
transforms."""
import numpy as np

from oracle import philox


def test_random123_kat_vectors():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == out


def test_random123_kat_vectors_2x32():
    kat = [((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
           ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
           ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))]
    for ctr, key, out in kat:
        got = philox.philox2x32_10(*ctr, key)
        assert tuple(int(v) for v in got) == out


def test_agent_draw_layout_and_vectorisation():
    """ctr = (agent id, step ^ seed_hi), key = seed_lo ^ (stream * 0x85EBCA6B); streams are unrelated words."""
    ids = np.arange(100, dtype=np.uint64)
    seed = 0x123456789ABCDEF
    for stream in (0, 1, 2):
        w = philox.agent_draws(seed, ids, 7, stream)
        for k in (0, 13, 99):
            s = philox.philox2x32_10(k, 7 ^ 0x1234567, 0x89ABCDEF ^ ((stream * 0x85EBCA6B) & 0xFFFFFFFF))
            assert [int(a[k]) for a in w] == [int(v) for v in s]
    a, b = philox.agent_draws(seed, ids, 7, 0), philox.agent_draws(seed, ids, 7, 1)
    assert not np.array_equal(a[0], b[0]) and not np.array_equal(a[0], philox.agent_draws(seed, ids, 8, 0)[0])


def test_u01_open_interval_and_exact():
    assert philox.u01(np.uint32(0)) == 2.0 ** -33
    assert philox.u01(np.uint32(0xFFFFFFFF)) == 1.0 - 2.0 ** -33


def test_normal_pair_is_float32_standard_normal():
    rng = np.random.default_rng(0)
    w = rng.integers(0, 2 ** 32, size=(2, 1_000_000), dtype=np.uint64).astype(np.uint32)
    z0, z1 = philox.normal_pair(w[0], w[1])
    assert z0.dtype == np.float32 and z1.dtype == np.float32
    for z in (z0, z1):
        assert abs(float(z.mean())) < 4e-3 and abs(float(z.std()) - 1.0) < 4e-3
    assert abs(float(np.corrcoef(z0, z1)[0, 1])) < 4e-3
    assert float(np.abs(z0).max()) <= 5.78 and float(np.abs(z1).max()) <= 5.78
    # against float64 Box-Muller of the same uniforms
    u1 = (2 * (w[0] >> 9).astype(np.float64) + 1) * 2.0 ** -24
    ang = ((w[1] >> 30).astype(np.float64) + (2 * ((w[1] >> 7) & 0x7FFFFF).astype(np.float64) + 1) * 2.0 ** -24) * np.pi / 2
    r = np.sqrt(-2 * np.log(u1))
    assert np.abs(z0 - r * np.cos(ang)).max() < 5e-6 and np.abs(z1 - r * np.sin(ang)).max() < 5e-6


def test_extreme_words():
    for w0 in (0, 0xFFFFFFFF, 1 << 9, 0x80000000):
        for w1 in (0, 0xFFFFFFFF, 0x40000000, 0x80000000, 0xC0000000):
            z0, z1 = philox.normal_pair(np.uint32(w0), np.uint32(w1))
            assert np.isfinite(z0) and np.isfinite(z1)
