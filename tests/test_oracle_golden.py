"""Pins the CPU restatement (oracle/cpu_oracle.cpp) to golden vectors produced by the REFERENCE's own CUDA build
(tests/golden/README.md).  The oracle is fp-contract-off host code, the reference fuses multiply-adds on the GPU:
floats agree to rounding (stated per test), ids agree except where a ray grazes an edge."""
import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(G, "golden.json")))
FLT_MAX = np.float32(3.4028234663852886e38)


@pytest.fixture(scope="module")
def golden_scene(small_scene):
    assert f"{small_scene.hash():016x}" == META["scene_hash"], "procedural scene changed: regenerate tests/golden (make_golden.py)"
    return small_scene


def _frame_check(img, gold, frac_1e3, psnr_min):
    d = np.abs(img.astype(np.float64) - gold).max(axis=2)
    mse = ((img.astype(np.float64) - gold) ** 2).mean()
    psnr = 10 * np.log10(gold.max() ** 2 / mse) if mse > 0 else np.inf
    assert (d <= 1e-3).mean() >= frac_1e3, f"only {(d <= 1e-3).mean():.4f} of pixels within 1e-3"
    assert psnr >= psnr_min, f"PSNR {psnr:.1f} dB"


@pytest.mark.parametrize("name,depth", [("staircase_96x64_8spp.ref", 64), ("staircase_96x64_8spp_d3.ref", 3)])
def test_oracle_frame_vs_reference_kernel(oracle, crt, golden_scene, name, depth):
    """Same seeds, same scene: >= 99.5 % of pixels within 1e-3 absolute and PSNR >= 55 dB against libref.so's frame."""
    gold = crt.read_ref(os.path.join(G, name), META["nx"], META["ny"])
    img, _ = oracle.render(golden_scene, META["nx"], META["ny"], META["ns"], depth)
    assert np.isfinite(img).all()
    _frame_check(img, gold, 0.995, 55.0)


def test_oracle_spheres_vs_reference_derived_kernel(oracle, crt):
    gold = crt.read_ref(os.path.join(G, "rtiow_96x64_8spp.ref"), META["nx"], META["ny"])
    img, _ = oracle.render_spheres(crt.rtiow_scene(1), META["nx"], META["ny"], META["ns"], 50)
    _frame_check(img, gold, 0.99, 40.0)  # silhouettes of 488 small spheres: a grazing sample flips more easily


def test_oracle_ray_batch_vs_reference_hitmesh(oracle, golden_scene):
    """hitMesh() closest hit: ids equal on >= 99.9 % of rays (edge grazes may flip), t within 1e-5 relative where ids agree."""
    z = np.load(os.path.join(G, "rays_8192.npz"))
    hit, mesh = oracle.intersect_batch(golden_scene, z["ray_o"], z["ray_d"])
    ids, gids = hit[:, 3].view(np.uint32), z["hit"][:, 3].view(np.uint32)
    same = ids == gids
    assert same.mean() >= 0.999
    assert np.array_equal(mesh[same], z["mesh"][same])
    h = same & (gids != 0xFFFFFFFF)
    assert h.sum() > 4000
    rel = np.abs(hit[h, 0] - z["hit"][h, 0]) / np.abs(z["hit"][h, 0])
    assert rel.max() <= 1e-5
    assert (hit[same & ~h, 0] == FLT_MAX).all()
    assert np.abs(hit[h, 1:3] - z["hit"][h, 1:3]).max() <= 1e-3


def test_oracle_any_hit_vs_reference(oracle, golden_scene):
    z = np.load(os.path.join(G, "rays_8192.npz"))
    rd = z["ray_d"].copy()
    rd[:, 3] = z["shadow_tmax"]
    hit, _ = oracle.intersect_batch(golden_scene, z["ray_o"], rd, any_hit=True)
    occluded = hit[:, 0] == 0.0
    assert (occluded == z["occluded"]).mean() >= 0.999
    assert 0.05 < occluded.mean() < 0.95  # both outcomes are exercised


def test_oracle_traversal_equals_brute_force(oracle, crt):
    """The bit-stack walk finds what testing every triangle finds (same triangleHit arithmetic, so ids are exact)."""
    rng = np.random.default_rng(11)
    n_t = 300
    t = np.zeros((n_t, 16), np.float32)
    base = rng.uniform(-5, 5, (n_t, 1, 3)).astype(np.float32)
    t[:, :9] = (base + rng.uniform(-1, 1, (n_t, 3, 3)).astype(np.float32)).reshape(n_t, 9)
    t.view(np.uint8).reshape(-1, 64)[:, 60] = rng.integers(0, 20, n_t)
    bvh = crt.Scene.from_triangles(t, 5, 4)
    flat = crt.Scene.from_triangles(t, n_t, 4)  # one leaf holding everything = brute force
    assert flat.num_nodes == 2
    n = 4000
    ro = np.zeros((n, 4), np.float32)
    rd = np.zeros((n, 4), np.float32)
    ro[:, :3] = rng.uniform(-8, 8, (n, 3))
    v = rng.normal(size=(n, 3))
    rd[:, :3] = v / np.linalg.norm(v, axis=1, keepdims=True)
    ro[:, 3] = 0.01
    rd[:, 3] = FLT_MAX
    h1, m1 = oracle.intersect_batch(bvh, ro, rd)
    h2, m2 = oracle.intersect_batch(flat, ro, rd)
    assert np.array_equal(h1[:, 0], h2[:, 0]) and np.array_equal(m1, m2)
    # ids index different slot layouts: compare the triangles they name
    a, b = h1[:, 3].view(np.uint32), h2[:, 3].view(np.uint32)
    hitmask = a != 0xFFFFFFFF
    assert hitmask.sum() > 500 and np.array_equal(hitmask, b != 0xFFFFFFFF)
    assert np.array_equal(bvh.triangles()[a[hitmask]][:, :9], flat.triangles()[b[hitmask]][:, :9])
    bvh.close()
    flat.close()


def test_oracle_edge_cases(oracle, crt):
    import ctypes as C
    L = oracle.lib()
    f3 = C.c_float * 3
    # slab test: direction with zero components (inf/NaN slabs, intersections.h:28-35), box t_min fixed at 0.001
    assert L.oracleBoxDist(f3(-1, -1, -1), f3(1, 1, 1), f3(0, 0, -5), f3(0, 0, 1), FLT_MAX) == np.float32(4.0)
    assert L.oracleBoxDist(f3(-1, -1, -1), f3(1, 1, 1), f3(2, 0, -5), f3(0, 0, 1), FLT_MAX) == FLT_MAX
    assert L.oracleBoxDist(f3(-1, -1, -1), f3(1, 1, 1), f3(0, 0, 0), f3(0, 0, 1), FLT_MAX) == np.float32(0.001)  # origin inside
    assert L.oracleBoxDist(f3(-1, -1, -1), f3(1, 1, 1), f3(0, 0, -5), f3(0, 0, 1), 3.0) == FLT_MAX             # clipped by t_max
    # triangle: parallel ray, hit outside (t_min, t_max), edge conditions (intersections.h:62-81)
    tri = crt.Triangle()
    for k, v in enumerate([(0, 0, 0), (1, 0, 0), (0, 1, 0)]):
        tri.v[k].e[:] = v
    u, v = C.c_float(), C.c_float()
    assert L.oracleTriangleHit(C.byref(tri), f3(0.25, 0.25, 1), f3(0, 0, -1), 0.01, FLT_MAX, C.byref(u), C.byref(v)) == 1.0
    assert (u.value, v.value) == (0.25, 0.25)
    assert L.oracleTriangleHit(C.byref(tri), f3(0.25, 0.25, 1), f3(1, 0, 0), 0.01, FLT_MAX, C.byref(u), C.byref(v)) == FLT_MAX
    assert L.oracleTriangleHit(C.byref(tri), f3(0.25, 0.25, 1), f3(0, 0, -1), 0.01, 1.0, C.byref(u), C.byref(v)) == FLT_MAX   # t == t_max
    assert L.oracleTriangleHit(C.byref(tri), f3(0.25, 0.25, 1), f3(0, 0, -1), 1.0, FLT_MAX, C.byref(u), C.byref(v)) == FLT_MAX  # t == t_min
    assert L.oracleTriangleHit(C.byref(tri), f3(0.75, 0.75, 1), f3(0, 0, -1), 0.01, FLT_MAX, C.byref(u), C.byref(v)) == FLT_MAX  # u+v > 1
    # sphere: both roots (intersections.h:93-101): from inside the far root is returned
    sp = crt.Sphere()
    sp.center.e[:] = (0, 0, 0)
    sp.radius = 1.0
    assert L.oracleSphereHit(C.byref(sp), f3(0, 0, -3), f3(0, 0, 1), 0.01, FLT_MAX) == 2.0
    assert L.oracleSphereHit(C.byref(sp), f3(0, 0, 0), f3(0, 0, 1), 0.01, FLT_MAX) == 1.0
    assert L.oracleSphereHit(C.byref(sp), f3(0, 2, -3), f3(0, 0, 1), 0.01, FLT_MAX) == FLT_MAX


def test_oracle_depth_zero_and_counts(oracle, small_scene):
    img, cnt = oracle.render(small_scene, 16, 12, 2, 0, count=True)
    assert (img == 0).all() and cnt["primary"] == 0
    img, cnt = oracle.render(small_scene, 16, 12, 2, 5, count=True)
    assert cnt["primary"] == 16 * 12 * 2 and cnt["secondary"] <= 4 * cnt["primary"] and cnt["shadow"] <= cnt["primary"] + cnt["secondary"]


def test_cpu_oracle_bsdf_presets_vs_golden_reference_outputs(oracle):
    """The CPU restatement of the reference's whole BSDF library (material.h behind the presets of scene_materials.h:22-93)
    against outputs of the reference's own device functions (tests/golden/bsdf_presets.npz, minted on a B200 through
    oracle/ref_shim.cu). Integer results (RNG state after the call = number of draws, specular/refracted flags) and the
    subsurface free-flight distance decisions must agree on every item where the decision is not a float tie; directions and
    throughputs to 2e-5 (the GPU fuses multiply-adds and uses its own expf/logf/powf, the oracle does neither)."""
    z = np.load(os.path.join(G, "bsdf_presets.npz"))
    items = z["items"]
    for preset in range(10):
        ref = z["ref_%d" % preset]
        got = oracle.scatter_batch(preset, items)
        same = (got[:, 8].view(np.uint32) == ref[:, 8].view(np.uint32)) & (got[:, 7].view(np.int32) == ref[:, 7].view(np.int32))
        assert same.mean() >= 0.99, (preset, same.mean())   # a Fresnel / free-flight comparison can flip on the last bit
        ok = same
        assert np.allclose(got[ok, 0:3], ref[ok, 0:3], rtol=2e-5, atol=2e-5), preset
        assert np.allclose(got[ok, 3], ref[ok, 3], rtol=2e-5, atol=1e-6), preset
        assert np.allclose(got[ok, 4:7], ref[ok, 4:7], rtol=2e-5, atol=1e-6), preset


def test_sah_tree_gives_the_same_frame_with_fewer_visits(oracle, crt, golden_scene):
    """The SAH-built tree (same layout, same triangles) is a pure acceleration change: the CPU oracle renders the same frame
    bit for bit and the same hit records, visiting fewer nodes and testing fewer triangles."""
    sah = crt.Scene.staircase(0.1, 32, 5, sah=True)
    a, ca = oracle.render(golden_scene, META["nx"], META["ny"], 4, 16, count=True)
    b, cb = oracle.render(sah, META["nx"], META["ny"], 4, 16, count=True)
    assert np.array_equal(a, b)
    assert (ca["primary"], ca["secondary"], ca["shadow"]) == (cb["primary"], cb["secondary"], cb["shadow"])
    assert cb["nodeVisits"] < 0.9 * ca["nodeVisits"] and cb["triTests"] < 0.8 * ca["triTests"]
    sah.close()
