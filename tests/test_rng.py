"""RNG known-answer tests (integer-exact): SURVEY.md appendix C = reference rnd.h:5-39 and the seed formula kernels.cu:542."""
import ctypes as C

import numpy as np

KAT = [  # pixelId, wang_hash, seed, xs[0..2], rnd[0..2]
    (0, 0xc0a9496a, 0xbedeee8b, (0xbc89bf0c, 0xb6347ab8, 0x8a67e609), (0.538071394, 0.20499754, 0.405853808)),
    (1, 0x27922c9d, 0x6d52c7ad, (0x8b187d7e, 0x7b257f25, 0x0e633545), (0.0956648588, 0.146471322, 0.387531579)),
    (12345, 0x0ddeec13, 0x7a886803, (0x6e6bb3c7, 0xb7f45fce, 0xfce981c8), (0.420711935, 0.954586864, 0.912136555)),
    (959999, 0x98f933a6, 0x76b7c387, (0x3cb564e4, 0x4651acf0, 0xb9441517), (0.708570719, 0.319045067, 0.265946805)),
]


def test_wang_hash_seed_xorshift_kat(oracle):
    L = oracle.lib()
    for pid, wh, seed, xs, rs in KAT:
        assert L.oracleWangHash(pid) == wh
        assert L.oraclePathSeed(pid) == seed
        st = C.c_uint(seed)
        assert tuple(L.oracleXorShift(C.byref(st)) for _ in range(3)) == xs
        st = C.c_uint(seed)
        got = [L.oracleRnd(C.byref(st)) for _ in range(3)]
        assert np.allclose(got, rs, rtol=0, atol=5e-9)
        # rnd = low 24 bits / 2^24 exactly (rnd.h:17)
        assert got == [np.float32((x & 0xFFFFFF) / 16777216.0) for x in xs]


def test_rnd_range_and_period_fragment(oracle):
    L = oracle.lib()
    st = C.c_uint(L.oraclePathSeed(7))
    seen = set()
    for _ in range(20000):
        r = L.oracleRnd(C.byref(st))
        assert 0.0 <= r < 1.0
        seen.add(st.value)
    assert len(seen) == 20000 and 0 not in seen  # xorshift32 never reaches 0 from an odd seed


def test_unit_samplers_draw_order(oracle):
    """x is the FIRST draw, y the second, z the third (the nvcc device order), rejection loops re-draw."""
    L = oracle.lib()
    for pid in (0, 1, 12345):
        seed = L.oraclePathSeed(pid)
        st = C.c_uint(seed)
        out = (C.c_float * 3)()
        L.oracleUnitSphere(C.byref(st), out)
        st2 = C.c_uint(seed)
        while True:
            a, b, c = (L.oracleRnd(C.byref(st2)) for _ in range(3))
            p = np.float32(2) * np.array([a, b, c], np.float32) - np.float32(1)
            if np.float32(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) < 1.0:
                break
        assert list(out) == list(p) and st.value == st2.value
        st = C.c_uint(seed)
        L.oracleUnitDisk(C.byref(st), out)
        st2 = C.c_uint(seed)
        while True:
            a, b = (L.oracleRnd(C.byref(st2)) for _ in range(2))
            p = np.float32(2) * np.array([a, b], np.float32) - np.float32(1)
            if np.float32(p[0] * p[0] + p[1] * p[1]) < 1.0:
                break
        assert list(out)[:2] == list(p) and out[2] == 0.0 and st.value == st2.value


def test_get_ray_draws_even_without_lens(oracle, crt):
    """camera.h:8-12: random_in_unit_disk consumes >= 2 numbers even when lens_radius == 0; the ray direction is unit."""
    L = oracle.lib()
    cam = crt.staircase_camera(1200, 800)
    assert cam.lens_radius == 0.0
    st = C.c_uint(L.oraclePathSeed(3))
    before = st.value
    o, d = (C.c_float * 3)(), (C.c_float * 3)()
    L.oracleGetRay(C.byref(cam), 0.5, 0.5, C.byref(st), o, d)
    assert st.value != before
    assert list(o) == list(cam.origin.e)
    assert abs(np.linalg.norm(np.array(list(d))) - 1.0) < 1e-6
    assert d[2] < -0.99  # looking down -z (staircase_scene.h:63-64)
