"""CPU checks of the renderer's own acceleration structure (csrc/wide_bvh.cpp), run in the GPU-less container:
the build is structurally sound (every triangle once, every quantised child box contains what lies below it, padded), and a
scalar walk that mirrors the CUDA traversal statement by statement (oracle/wide_walk.cpp: group stack, octant order, paired
plane decode, near-tie margin, reference-leaf certificate) returns, for every ray it does not flag, exactly what the CPU
restatement of the reference's hitBvh (kernels.cu:154-224) returns on the caller's tree -- closest hit and any-hit alike."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLT_MAX = np.float32(3.4028234663852886e38)


@pytest.fixture(scope="module")
def walk(crt, oracle):
    oracle.lib()
    W = C.CDLL(os.path.join(ROOT, "oracle", "build", "libwide_walk.so"))
    W.wideWalkBatch.argtypes = [C.POINTER(crt.KernelScene), C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_ulonglong), C.c_int]
    W.wideCheckStructure.argtypes = [C.POINTER(crt.KernelScene), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    return W


def make_rays(crt, scene, n, seed):
    rng = np.random.default_rng(seed)
    bmin, bmax = scene.bounds()
    cam = crt.staircase_camera(1200, 800)
    o = np.zeros((n, 4), np.float32)
    d = np.zeros((n, 4), np.float32)
    h = n // 2
    uv = rng.random((h, 2), dtype=np.float32)
    llc, hor, ver, org = (np.array(list(v.e), np.float32) for v in (cam.lower_left_corner, cam.horizontal, cam.vertical, cam.origin))
    o[:h, :3] = org
    d[:h, :3] = llc + uv[:, :1] * hor + uv[:, 1:] * ver - org
    o[h:, :3] = (bmin + rng.random((n - h, 3)) * (bmax - bmin)).astype(np.float32)
    v = rng.normal(size=(n - h, 3)).astype(np.float32)
    d[h:, :3] = v / np.linalg.norm(v, axis=1, keepdims=True)
    # axis-parallel directions (zero components) and origins on mesh vertices: the cases padding and clamping exist for
    k = min(512, n - h)
    axes = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice(np.array([-1.0, 1.0], np.float32), (k, 1))
    d[h:h + k, :3] = axes
    tris = scene.triangles()
    real = tris[np.isfinite(tris[:, 0])]
    o[h + k:h + 2 * k, :3] = real[rng.integers(0, len(real), min(k, n - h - k)), 0:3]
    o[:, 3] = 0.01
    d[:, 3] = FLT_MAX
    return o, d


@pytest.mark.parametrize("threads", [1, 4])
def test_structure(crt, walk, small_scene, medium_scene, threads):
    for scene in (small_scene, medium_scene):
        ms, nn = C.c_double(), C.c_ulonglong()
        assert walk.wideCheckStructure(C.byref(scene.ks), threads, C.byref(ms), C.byref(nn)) == 0
        assert 0 < nn.value < scene.num_real_triangles


@pytest.mark.parametrize("any_hit", [0, 1])
def test_walk_equals_reference_walk(crt, oracle, walk, medium_scene, any_hit):
    n = 60000
    o, d = make_rays(crt, medium_scene, n, 7 + any_hit)
    if any_hit:
        d[:, 3] = (np.random.default_rng(3).random(n) * 600 + 1).astype(np.float32)
    ref, _, cnt_ref = oracle.intersect_batch(medium_scene, o, d, any_hit=bool(any_hit), count=True)
    hit = np.zeros((n, 4), np.float32)
    flag = np.zeros(n, np.uint8)
    cnt = (C.c_ulonglong * 8)()
    assert walk.wideWalkBatch(C.byref(medium_scene.ks), o.ctypes.data, d.ctypes.data, n, any_hit, hit.ctypes.data, flag.ctypes.data, cnt, 2) == 0
    same = (hit.view(np.uint32) == ref.view(np.uint32)).all(axis=1)
    assert not (~same & (flag == 0)).any(), "an unflagged ray differs from the reference's walk"
    assert flag.mean() < 2e-3, "the certificate should fail for a handful of rays only"
    assert cnt[3] <= cnt[5], "the stack never holds more entries than the tree has levels"
    # the point of the exercise: far fewer node visits and triangle tests than the caller's median-split heap
    assert cnt[0] < 0.4 * cnt_ref["nodeVisits"] and cnt[1] < 0.6 * cnt_ref["triTests"]


def test_tiny_and_degenerate_scenes(crt, oracle, walk):
    rng = np.random.default_rng(5)
    for ntri in (1, 2, 7):
        t = np.zeros((ntri, 16), np.float32)
        t[:, :9] = rng.random((ntri, 9), dtype=np.float32) * 10
        scene = crt.Scene.from_triangles(t, 5, 16)
        n = 4000
        o = np.zeros((n, 4), np.float32)
        d = np.zeros((n, 4), np.float32)
        o[:, :3] = rng.random((n, 3), dtype=np.float32) * 30 - 10
        target = rng.random((n, 3), dtype=np.float32) * 10
        d[:, :3] = target - o[:, :3]
        o[:, 3] = 0.01
        d[:, 3] = FLT_MAX
        ref, _ = oracle.intersect_batch(scene, o, d)
        hit = np.zeros((n, 4), np.float32)
        flag = np.zeros(n, np.uint8)
        assert walk.wideWalkBatch(C.byref(scene.ks), o.ctypes.data, d.ctypes.data, n, 0, hit.ctypes.data, flag.ctypes.data, None, 1) == 0
        same = (hit.view(np.uint32) == ref.view(np.uint32)).all(axis=1)
        assert not (~same & (flag == 0)).any()
        assert (ref[:, 0] < 1e30).any()
        scene.close()
    # coincident triangles (exact ties): whatever the walk finds, it must flag the rays that hit them
    t = np.zeros((4, 16), np.float32)
    t[:, :9] = np.array([0, 0, 0, 10, 0, 0, 0, 10, 0], np.float32)
    scene = crt.Scene.from_triangles(t, 5, 16)
    n = 500
    o = np.zeros((n, 4), np.float32)
    d = np.zeros((n, 4), np.float32)
    o[:, :3] = np.c_[rng.random((n, 2)) * 4 + 0.5, np.full(n, 5.0)].astype(np.float32)
    d[:, :3] = (0, 0, -1)
    o[:, 3] = 0.01
    d[:, 3] = FLT_MAX
    ref, _ = oracle.intersect_batch(scene, o, d)
    hit = np.zeros((n, 4), np.float32)
    flag = np.zeros(n, np.uint8)
    assert walk.wideWalkBatch(C.byref(scene.ks), o.ctypes.data, d.ctypes.data, n, 0, hit.ctypes.data, flag.ctypes.data, None, 1) == 0
    assert (ref[:, 0] < 1e30).all() and flag.all()
    scene.close()
