"""Parity of the CUDA path (through the C ABI of include/kernels.h) with the golden vectors, the CPU oracle and --
where oracle/_ref travelled to the box -- the reference's own CUDA kernel run side by side.

Stated bars (BASELINE.json north_star / SURVEY.md 8d):
  * ray batches: hit triangle ids and mesh ids bit-exact; t within 1e-5 relative (observed: t, u, v bit-exact);
  * frames with the reference's seeding: per-pixel max-abs <= 1e-3 on >= 99.9 % of pixels and PSNR >= 50 dB
    against the reference kernel (observed: bit-exact frames);
  * against the CPU oracle (no FMA): >= 99.5 % of pixels within 1e-3, ids equal on >= 99.9 % of rays."""
import json
import os
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(G, "golden.json")))
FLT_MAX = np.float32(3.4028234663852886e38)


def gpu_intersect(crt, ro, rd):
    n = ro.shape[0]
    hit = np.zeros((n, 4), np.float32)
    mesh = np.zeros(n, np.int32)
    L = crt.device_lib()
    ptrs = [L.rendererDeviceAlloc(16 * n) for _ in range(3)] + [L.rendererDeviceAlloc(4 * n)]
    L.rendererCopyToDevice(ptrs[0], np.ascontiguousarray(ro).ctypes.data, 16 * n)
    L.rendererCopyToDevice(ptrs[1], np.ascontiguousarray(rd).ctypes.data, 16 * n)
    ms = L.intersectBatchDevice(ptrs[0], ptrs[1], n, ptrs[2], ptrs[3])
    L.rendererCopyToHost(hit.ctypes.data, ptrs[2], 16 * n)
    L.rendererCopyToHost(mesh.ctypes.data, ptrs[3], 4 * n)
    for p in ptrs:
        L.rendererDeviceFree(p)
    return hit, mesh, ms


def frame_stats(a, b):
    d = np.abs(a.astype(np.float64) - b).max(axis=2)
    mse = ((a.astype(np.float64) - b) ** 2).mean()
    return (d <= 1e-3).mean(), (10 * np.log10(b.max() ** 2 / mse) if mse > 0 else np.inf), (d == 0).mean()


@pytest.mark.parametrize("name,depth", [("staircase_96x64_8spp.ref", 64), ("staircase_96x64_8spp_d3.ref", 3)])
def test_frame_vs_golden_reference_frame(crt, small_scene, name, depth):
    assert f"{small_scene.hash():016x}" == META["scene_hash"]
    gold = crt.read_ref(os.path.join(G, name), META["nx"], META["ny"])
    with crt.Frame(small_scene, META["nx"], META["ny"], depth) as fr:
        img = fr.run(META["ns"])
        again = fr.run(META["ns"])  # runRenderer may be called repeatedly between init and cleanup
    within, psnr, exact = frame_stats(img, gold)
    assert within >= 0.999 and psnr >= 50.0, (within, psnr, exact)
    assert exact == 1.0, f"expected a bit-exact frame, {exact:.6f} of pixels are"
    assert np.array_equal(img, again)


def test_spheres_vs_golden_reference_derived_frame(crt):
    gold = crt.read_ref(os.path.join(G, "rtiow_96x64_8spp.ref"), META["nx"], META["ny"])
    with crt.Frame(crt.rtiow_scene(1), META["nx"], META["ny"], 50) as fr:
        img = fr.run(META["ns"])
    within, psnr, exact = frame_stats(img, gold)
    # The golden frame comes from oracle/ref_spheres.cu compiled on its own: nvcc/ptxas fuse a*b+c*d there as that
    # context suggests, our kernels pin the forms of the reference's mesh kernel (csrc/vecmath.cuh). Same samples, same
    # paths; a percent of the pixels differs in the last bits.
    assert within >= 0.999 and psnr >= 50.0 and exact >= 0.98, (within, psnr, exact)


def test_frames_survive_cache_release_and_size_changes(crt, small_scene, medium_scene):
    """cleanupRenderer keeps the device arena / streams / pinned buffers for the next frame; a bigger frame, a smaller
    frame, another scene and an explicit rendererReleaseCaches() in between must not change a bit."""
    L = crt.device_lib()

    def render(scene, nx, ny, ns):
        with crt.Frame(scene, nx, ny, 16) as fr:
            return fr.run(ns)

    a = render(small_scene, 96, 64, 4)
    b = render(medium_scene, 320, 200, 4)   # grows the arena
    assert np.array_equal(render(small_scene, 96, 64, 4), a)   # shrinks again: same bits
    L.rendererReleaseCaches()
    assert np.array_equal(render(medium_scene, 320, 200, 4), b)
    with crt.Frame(small_scene, 96, 64, 16) as fr:
        L.rendererReleaseCaches()           # no-op while a frame is live
        assert np.array_equal(fr.run(4), a)
    L.rendererReleaseCaches()


def test_progressive_rendering_and_checkpoint_are_exact(crt, medium_scene, tmp_path):
    """runRenderer(a) + continueRenderer(b) == runRenderer(a + b), bit for bit -- also through a checkpoint file and a fresh
    initRenderer (a pixel's samples are one RNG stream, so the split is invisible)."""
    nx, ny, depth = 320, 200, 64
    L = crt.device_lib()
    with crt.Frame(medium_scene, nx, ny, depth) as fr:
        whole = fr.run(12)
        part = fr.run(5)
        assert L.getRendererSamplesDone() == 5
        both = fr.continue_run(7)
        assert L.getRendererSamplesDone() == 12
        assert not np.array_equal(part, whole) and np.array_equal(both, whole)
        fr.run(5)
        ck = str(tmp_path / "frame.ckpt").encode()
        assert L.saveRendererCheckpoint(ck) == 0
    with crt.Frame(medium_scene, nx, ny, depth) as fr:      # a fresh renderer
        assert L.continueRenderer(7, 8, 8) != 0             # nothing to continue yet
        assert L.loadRendererCheckpoint(ck) == 0
        assert np.array_equal(fr.continue_run(7), whole)
    with crt.Frame(medium_scene, nx + 8, ny, depth):
        assert L.loadRendererCheckpoint(ck) != 0            # another frame size


def _scatter(crt, preset, items):
    out = np.zeros_like(items)
    assert crt.device_lib().scatterBatch(preset, items.shape[0], items.ctypes.data, out.ctypes.data) == 0
    return out


def _compare_scatter(ours, ref):
    """rng state after (= number of draws), flags and t must be identical; directions / throughputs to 2e-6. Observed: seven
    presets bit for bit; floor_coat, floor_checker and model_coat differ by <= 2e-7 in wi on 8-24 % of the items -- the
    reference's compiler fuses length()'s a*b + c*d differently where diffuse_bsdf is inlined into those callers, ours is
    pinned to the form of the render kernel (csrc/vecmath.cuh)."""
    assert np.array_equal(ours[:, 8].view(np.uint32), ref[:, 8].view(np.uint32))
    assert np.array_equal(ours[:, 7].view(np.int32), ref[:, 7].view(np.int32))
    assert np.array_equal(ours[:, 3], ref[:, 3])
    assert np.allclose(ours[:, 0:3], ref[:, 0:3], rtol=2e-6, atol=2e-6) and np.allclose(ours[:, 4:7], ref[:, 4:7], rtol=2e-6, atol=1e-7)
    return float((ours.view(np.uint32) == ref.view(np.uint32)).all(axis=1).mean())


def test_bsdf_library_vs_golden_reference_outputs(crt):
    """csrc/bsdf.cuh's restatement of the reference's whole BSDF library (material.h) behind its scene presets
    (scene_materials.h:22-93) against outputs of the reference's own device functions (tests/golden/bsdf_presets.npz,
    minted by tools/make_bsdf_golden.py through oracle/ref_shim.cu)."""
    z = np.load(os.path.join(G, "bsdf_presets.npz"))
    items = np.ascontiguousarray(z["items"])
    for preset in range(10):
        exact = _compare_scatter(_scatter(crt, preset, items), z["ref_%d" % preset])
        assert exact >= (0.7 if preset in (0, 2, 3) else 1.0), (preset, exact)
    assert crt.device_lib().scatterBatch(10, 1, items.ctypes.data, items.ctypes.data) != 0  # unknown preset


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "ref_shim_driver")), reason="oracle/_ref not built")
def test_bsdf_library_side_by_side(crt, oracle):
    from bsdf_inputs import make_items
    items = make_items(50000, seed=11)
    tmp = tempfile.mkdtemp()
    for preset in range(10):
        exact = _compare_scatter(_scatter(crt, preset, items), oracle.ref_scatter_batch(preset, items, tmp))
        assert exact >= (0.7 if preset in (0, 2, 3) else 1.0), (preset, exact)


def test_sphere_bvh_equals_the_brute_force_loop(crt, monkeypatch):
    """The sphere BVH returns, ray by ray, what the loop over all spheres returns (closest root, ties to the lowest index):
    whole frames are bit-identical, ray counts equal."""
    nx, ny, ns = 300, 200, 8
    monkeypatch.setenv("CRT_SPHERES_BRUTE", "1")
    with crt.Frame(crt.rtiow_scene(1), nx, ny, 50) as fr:
        brute = fr.run(ns)
        rays_brute = crt.stats().raysExtend
    monkeypatch.delenv("CRT_SPHERES_BRUTE")
    with crt.Frame(crt.rtiow_scene(1), nx, ny, 50) as fr:
        bvh = fr.run(ns)
        rays_bvh = crt.stats().raysExtend
    assert rays_bvh == rays_brute and rays_bvh > nx * ny * ns
    assert np.array_equal(bvh, brute)


def test_sphere_persistent_kernel_equals_the_wavefront(crt, monkeypatch):
    """The persistent kernel (a lane owns a pixel from its first camera ray to its last sample, path in registers) and the
    wavefront kernels (extend / shade launches over queues) run the same device functions in the same per-pixel order: frames
    are bit-identical, ray counts equal -- with the sphere BVH and, for the wavefront, the brute-force loop beneath."""
    nx, ny, ns = 300, 200, 8
    monkeypatch.setenv("CRT_SPHERES_WAVEFRONT", "1")
    with crt.Frame(crt.rtiow_scene(1), nx, ny, 50) as fr:
        wave = fr.run(ns)
        rays_wave = crt.stats().raysExtend
    monkeypatch.delenv("CRT_SPHERES_WAVEFRONT")
    with crt.Frame(crt.rtiow_scene(1), nx, ny, 50) as fr:
        mega = fr.run(ns)
        st = crt.stats()
        again = fr.run(ns)
    assert st.raysExtend == rays_wave and st.kernelLaunches <= 2
    assert np.array_equal(mega, wave) and np.array_equal(mega, again)


def _sphere_set(crt, rng, n, kind):
    sph = (crt.Sphere * 1024)()
    mats = (crt.Material * 1024)()
    for i in range(n):
        if kind == "ties" and i >= n // 2:       # every sphere twice: equal roots, the lower index must win
            src = sph[i - n // 2]
            c, r = (src.center.e[0], src.center.e[1], src.center.e[2]), src.radius
        elif i < 3:                              # spheres the BVH keeps in its always-list (radius > 100), overlapping
            c, r = (float(rng.uniform(-30, 30)), -150.0 - 10.0 * i, float(rng.uniform(-30, 30))), 150.0 + 10.0 * i
        else:
            big = rng.random() < 0.1
            c = tuple(float(x) for x in rng.uniform(-6, 6, 3))
            r = float(rng.uniform(1.0, 2.5) if big else rng.uniform(0.05, 0.6))   # overlapping, nested, tiny
        for a in range(3):
            sph[i].center.e[a] = c[a]
        sph[i].radius = r
        mats[i].type = int(rng.integers(0, 3))   # DIFFUSE / METAL / GLASS
        for a in range(3):
            mats[i].color.e[a] = float(rng.uniform(0.2, 1.0))
        mats[i].param = float(rng.uniform(0.0, 0.4)) if mats[i].type == 1 else 1.5
        mats[i].texId = -1
    return sph, mats, n


@pytest.mark.parametrize("kind,n", [("random", 300), ("ties", 200), ("single", 1), ("full", 1024)])
def test_sphere_paths_agree_on_adversarial_sets(crt, monkeypatch, kind, n):
    """Sphere sets that are not the README scene -- overlapping and nested spheres, three overlapping giants in the always-list,
    exact duplicates (ties by index), one sphere, the full constant table: the persistent kernel over the SAH-split, octant-threaded
    BVH and the wavefront over the brute-force loop give identical frames and ray counts."""
    rng = np.random.default_rng(len(kind) + n)
    scene = _sphere_set(crt, rng, n, kind)
    nx, ny, ns = 160, 120, 6
    monkeypatch.setenv("CRT_SPHERES_BRUTE", "1")
    with crt.Frame(scene, nx, ny, 50) as fr:
        brute = fr.run(ns)
        rays_brute = crt.stats().raysExtend
    monkeypatch.delenv("CRT_SPHERES_BRUTE")
    with crt.Frame(scene, nx, ny, 50) as fr:
        mega = fr.run(ns)
        rays_mega = crt.stats().raysExtend
    assert rays_mega == rays_brute and rays_mega >= nx * ny * ns
    assert np.isfinite(mega).all() and np.array_equal(mega, brute)


def test_ray_batch_vs_golden_hitmesh(crt, small_scene):
    z = np.load(os.path.join(G, "rays_8192.npz"))
    with crt.Frame(small_scene, 8, 8, 1):
        hit, mesh, _ = gpu_intersect(crt, z["ray_o"], z["ray_d"])
        # host-pointer entry point gives the same answers
        n = 1000
        t = np.zeros(n, np.float32)
        tid = np.zeros(n, np.int32)
        mid = np.zeros(n, np.int32)
        o3 = np.ascontiguousarray(z["ray_o"][:n, :3])
        d3 = np.ascontiguousarray(z["ray_d"][:n, :3])
        crt.device_lib().intersectBatch(o3.ctypes.data, d3.ctypes.data, n, 0.01, float(FLT_MAX), t.ctypes.data, tid.ctypes.data, mid.ctypes.data)
    assert np.array_equal(hit[:, 3].view(np.uint32), z["hit"][:, 3].view(np.uint32)), "triangle ids differ from the reference"
    assert np.array_equal(mesh, z["mesh"])
    h = z["mesh"] >= 0
    rel = np.abs(hit[h, 0] - z["hit"][h, 0]) / np.abs(z["hit"][h, 0])
    assert rel.max() <= 1e-5
    assert np.array_equal(hit.view(np.uint32), z["hit"].view(np.uint32)), "t/u/v are expected to be bit-exact as well"
    assert np.array_equal(t, hit[:n, 0]) and np.array_equal(tid.view(np.uint32), hit[:n, 3].view(np.uint32)) and np.array_equal(mid, mesh[:n])


def test_frame_vs_cpu_oracle(crt, oracle, medium_scene):
    nx, ny, ns, depth = 120, 80, 6, 64
    ref, cnt = oracle.render(medium_scene, nx, ny, ns, depth, count=True)
    with crt.Frame(medium_scene, nx, ny, depth) as fr:
        img = fr.run(ns)
        st = crt.stats()
    within, psnr, _ = frame_stats(img, ref)
    assert within >= 0.995 and psnr >= 50.0, (within, psnr)
    # the same paths are traced: ray counts agree to the handful of paths where rounding flips a branch
    assert st.samples == nx * ny * ns
    assert abs(int(st.raysExtend) - (cnt["primary"] + cnt["secondary"])) <= 2e-3 * st.raysExtend
    assert abs(int(st.raysShadow) - cnt["shadow"]) <= 2e-3 * st.raysShadow
    assert st.kernelLaunches > 0


def test_ray_batch_vs_cpu_oracle_and_counters(crt, oracle, medium_scene):
    n = 1 << 16
    L = crt.device_lib()
    crt.set_traversal(crt.TRAVERSAL_EXACT)  # the counters compared below are those of the reference's walk over the caller's tree
    with crt.Frame(medium_scene, 64, 64, 1):
        crt.set_traversal(-1)
        dO, dD, dH, dM = L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(4 * n)
        L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, float(FLT_MAX))
        L.setRendererCounting(1)
        L.intersectBatchDevice(dO, dD, n, dH, dM)
        import ctypes as C
        nv, tt = C.c_ulonglong(), C.c_ulonglong()
        L.getRendererTraversalCounts(C.byref(nv), C.byref(tt))
        L.setRendererCounting(0)
        ro, rd, hit, mesh = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros(n, np.int32)
        for a, p in ((ro, dO), (rd, dD), (hit, dH)):
            L.rendererCopyToHost(a.ctypes.data, p, 16 * n)
        L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
        for p in (dO, dD, dH, dM):
            L.rendererDeviceFree(p)
    assert np.allclose(np.linalg.norm(rd[:, :3], axis=1), 1.0, atol=1e-5)
    chit, cmesh, cnt = oracle.intersect_batch(medium_scene, ro, rd, count=True)
    same = hit[:, 3].view(np.uint32) == chit[:, 3].view(np.uint32)
    assert same.mean() >= 0.999 and np.array_equal(mesh[same], cmesh[same])
    h = same & (mesh >= 0)
    assert (np.abs(hit[h, 0] - chit[h, 0]) / np.abs(chit[h, 0])).max() <= 1e-5
    # visit counters (flop side of the roofline) agree with the CPU walk to the same tolerance
    assert abs(nv.value - cnt["nodeVisits"]) <= 2e-3 * cnt["nodeVisits"] and abs(tt.value - cnt["triTests"]) <= 2e-3 * cnt["triTests"]


def test_edge_cases(crt, oracle):
    """Empty scene, single triangle, ragged leaf, zero samples, depth 0, 1x1 frame, non-multiple-of-32 sizes."""
    empty = crt.Scene.from_triangles(np.zeros((0, 16), np.float32), 5, 4)
    with crt.Frame(empty, 5, 3, 8) as fr:
        img = fr.run(4)
    assert np.array_equal(img, np.full((3, 5, 3), 0.5, np.float32))  # every path misses: constant grey sky (kernels.cu:424)
    empty.close()

    t = np.zeros((1, 16), np.float32)
    t[0, :9] = [-300, 0, 300, 300, 0, 300, 0, 400, 300]
    t.view(np.uint8).reshape(-1, 64)[0, 60] = 4  # ChairSeat: a dark diffuse albedo, so hit pixels differ from the sky
    one = crt.Scene.from_triangles(t, 5, 4)
    ref, _ = oracle.render(one, 33, 17, 3, 4)
    with crt.Frame(one, 33, 17, 4) as fr:
        img = fr.run(3)
    assert np.abs(img - ref).max() <= 1e-3 and (img != 0.5).any()
    with crt.Frame(one, 33, 17, 0) as fr:  # maxDepth 0: the bounce loop never runs (kernels.cu:402)
        assert (fr.run(3) == 0).all()
    with crt.Frame(one, 1, 1, 4) as fr:
        assert fr.run(1).shape == (1, 1, 3)
    one.close()


def test_sample_streams_and_slots_are_statistically_consistent(crt, medium_scene):
    """Multi-GPU shards use other RNG streams (kernels.h renderer_options.sampleStream) and the throughput mode uses
    several slots per pixel: different noise, same expectation."""
    nx, ny, ns = 96, 64, 32
    with crt.Frame(medium_scene, nx, ny, 16) as fr:
        base = fr.run(ns)
    crt.set_options(sample_stream=1)
    with crt.Frame(medium_scene, nx, ny, 16) as fr:
        other = fr.run(ns)
    crt.set_options(slots_per_pixel=4)
    with crt.Frame(medium_scene, nx, ny, 16) as fr:
        multi = fr.run(ns)
        again = fr.run(ns)
    crt.set_options()
    assert not np.array_equal(base, other) and not np.array_equal(base, multi)
    for img in (other, multi):
        assert abs(float(img.mean()) - float(base.mean())) <= 0.05 * float(base.mean())
    # slot 0 of the 4-slot run is the reference stream; the frame is reproducible up to atomic-add order
    assert np.abs(multi - again).max() <= 1e-5


def test_full_size_properties(crt):
    """BASELINE size (1200x800) at reduced spp: finite, deterministic, linear in the sample count split, sky pixels exact."""
    scene = crt.Scene.staircase(1.0, 256, 5)
    nx, ny = 1200, 800
    with crt.Frame(scene, nx, ny, 64) as fr:
        a = fr.run(2)
        st = crt.stats()
        b = fr.run(2)
    assert a.shape == (ny, nx, 3) and np.isfinite(a).all() and (a >= 0).all()
    assert np.array_equal(a, b)
    assert st.samples == nx * ny * 2 and st.raysExtend >= st.samples
    # two 1-spp frames on streams 0 and 1 average to a frame with the same mean as a 2-spp frame (linearity of the estimator)
    crt.set_options(sample_stream=1)
    with crt.Frame(scene, nx, ny, 64) as fr:
        s1 = fr.run(1)
    crt.set_options()
    with crt.Frame(scene, nx, ny, 64) as fr:
        s0 = fr.run(1)
    assert abs(float((0.5 * (s0 + s1)).mean()) - float(a.mean())) <= 0.03 * float(a.mean())
    scene.close()


def test_frame_does_not_depend_on_who_runs_a_slot(crt, medium_scene, monkeypatch):
    """The wavefront alone, the chaser alone (every slot handed over before the first iteration) and mixes with slots handed
    over in every state (fresh / parked extend ray, pending / parked shadow ray, deferred shade, last FINAL shadow ray) give
    the same bits: the schedule decides only WHEN a slot's next bounce is computed, never what it computes."""
    nx, ny, ns, depth = 400, 300, 16, 64

    def render(env):
        for k in ("CRT_CHASER", "CRT_CHASE_MOVE_ALL", "CRT_CHASE_CAPACITY", "CRT_CHASE_LAG_PCT", "CRT_CHASE_TARGET", "CRT_CHASE_EXCLUSIVE_PCT"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        crt.set_options(mega_batch=2)  # hand-overs every 2 iterations
        with crt.Frame(medium_scene, nx, ny, depth) as fr:
            img = fr.run(ns)
            st = crt.stats()
        crt.set_options()
        return img, st.raysExtend + st.raysShadow

    wave, rays = render({"CRT_CHASER": "0"})
    for env in ({"CRT_CHASE_MOVE_ALL": "100000000", "CRT_CHASE_CAPACITY": "100000000"},          # chaser alone
                {"CRT_CHASE_MOVE_ALL": "0", "CRT_CHASE_LAG_PCT": "90", "CRT_CHASE_EXCLUSIVE_PCT": "50"},  # steady trickle of hand-overs
                {"CRT_CHASE_MOVE_ALL": "60000", "CRT_CHASE_LAG_PCT": "60"},                          # big final hand-over mid-flight
                {}):                                                                                 # defaults
        img, r = render(env)
        assert r == rays, (env, r, rays)
        assert np.array_equal(img, wave), env


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "ref_driver")), reason="oracle/_ref not built")
def test_side_by_side_with_reference_kernel(crt, oracle, medium_scene):
    """The reference's own render kernel (libref.so, separate process) and ours on the same box, same seeds, a config
    that is not in the golden set."""
    nx, ny, ns, depth = 200, 150, 12, 64
    tmp = tempfile.mkdtemp()
    ref, _ = oracle.ref_render(0.25, 64, 5, nx, ny, ns, depth, os.path.join(tmp, "r.ref"))
    with crt.Frame(medium_scene, nx, ny, depth) as fr:
        img = fr.run(ns)
    within, psnr, exact = frame_stats(img, ref)
    assert within >= 0.999 and psnr >= 50.0 and exact == 1.0, (within, psnr, exact)
    # the drop-in proof: ref_driver.cpp compiled against the REFERENCE's headers, linked to OUR library
    drop, _ = oracle.ref_render(0.25, 64, 5, nx, ny, ns, depth, os.path.join(tmp, "d.ref"), driver="dropin_driver")
    assert np.array_equal(drop, img)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "ref_driver")), reason="oracle/_ref not built")
def test_side_by_side_on_a_sah_built_tree(crt, oracle, tmp_path):
    """The opt-in SAH builder writes the same BVH_00.04 layout: the reference kernel reads the file like any scene of its own,
    and on the same tree both kernels visit in the same order -- identical frames, as with the median tree."""
    nx, ny, ns, depth = 200, 150, 12, 64
    scene = crt.Scene.staircase(0.25, 64, 5, sah=True)
    path = str(tmp_path / "sah.bvh")
    assert scene.save_bvh(path) == 0
    ref, _ = oracle.ref_render(path, 64, 5, nx, ny, ns, depth, str(tmp_path / "r.ref"))
    with crt.Frame(scene, nx, ny, depth) as fr:
        img = fr.run(ns)
    within, psnr, exact = frame_stats(img, ref)
    assert within >= 0.999 and psnr >= 50.0 and exact == 1.0, (within, psnr, exact)
    scene.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "ref_shim_driver")), reason="oracle/_ref not built")
def test_large_ray_batch_side_by_side(crt, oracle, medium_scene):
    n = 1 << 20
    L = crt.device_lib()
    with crt.Frame(medium_scene, 64, 64, 1):
        dO, dD, dH, dM = L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(4 * n)
        L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, float(FLT_MAX))
        L.intersectBatchDevice(dO, dD, n, dH, dM)
        ro, rd, hit, mesh = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros(n, np.int32)
        for a, p in ((ro, dO), (rd, dD), (hit, dH)):
            L.rendererCopyToHost(a.ctypes.data, p, 16 * n)
        L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
        for p in (dO, dD, dH, dM):
            L.rendererDeviceFree(p)
    rhit, rmesh, _ = oracle.ref_intersect_batch(0.25, 64, 5, ro, rd, False, tempfile.mkdtemp())
    assert np.array_equal(hit[:, 3].view(np.uint32), rhit[:, 3].view(np.uint32)) and np.array_equal(mesh, rmesh)
    h = rmesh >= 0
    assert (np.abs(hit[h, 0] - rhit[h, 0]) / np.abs(rhit[h, 0])).max() <= 1e-5


# ------------------------------------------------------------------------------------------------------------ round 2 --
HAVE_REF = os.path.exists(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "ref_driver"))
HAVE_SHIM = os.path.exists(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "ref_shim_driver"))


def test_light_sampler_trig_equals_libm_on_every_argument(crt):
    """sinCosSmall (the math library's fast path restated so that no large-argument reduction with its local-memory table is
    compiled into the kernels) returns the bits of sinf / cosf for EVERY float in [0, 2*pi]."""
    assert crt.device_lib().rendererTrigSelfTest() == 0


def test_any_hit_ray_batch_vs_golden_occluded(crt, small_scene):
    """The shadow-ray query on the GPU (hitMesh(.., isShadow = true), kernels.cu:207,500) against the reference's own answers
    for the golden batch, through both traversals."""
    z = np.load(os.path.join(G, "rays_8192.npz"))
    rd = z["ray_d"].copy()
    rd[:, 3] = z["shadow_tmax"]
    n = rd.shape[0]
    L = crt.device_lib()
    for mode in (crt.TRAVERSAL_WIDE, crt.TRAVERSAL_EXACT):
        crt.set_traversal(mode)
        with crt.Frame(small_scene, 8, 8, 1):
            crt.set_traversal(-1)
            ptrs = [L.rendererDeviceAlloc(16 * n) for _ in range(3)] + [L.rendererDeviceAlloc(4 * n)]
            L.rendererCopyToDevice(ptrs[0], np.ascontiguousarray(z["ray_o"]).ctypes.data, 16 * n)
            L.rendererCopyToDevice(ptrs[1], np.ascontiguousarray(rd).ctypes.data, 16 * n)
            L.intersectBatchDeviceEx(ptrs[0], ptrs[1], n, ptrs[2], ptrs[3], 1)
            hit = np.zeros((n, 4), np.float32)
            L.rendererCopyToHost(hit.ctypes.data, ptrs[2], 16 * n)
            for p in ptrs:
                L.rendererDeviceFree(p)
            assert crt.wide_info().active == (1 if mode == crt.TRAVERSAL_WIDE else 0)
        occluded = hit[:, 0] == 0.0
        assert np.array_equal(occluded, z["occluded"]), mode
        assert np.all((hit[:, 0] == 0.0) | (hit[:, 0] == FLT_MAX))
        assert 0.05 < occluded.mean() < 0.95


def test_wide_tree_equals_exact_walk(crt, medium_scene):
    """The renderer's own wide tree (+ certificate + exact re-trace) against the order-exact walk of the caller's tree: ray
    batches (closest hit and any-hit) and a frame, bit for bit; the certificate hands only a handful of rays to the exact kernel."""
    n = 1 << 20
    L = crt.device_lib()
    res = {}
    for mode in (crt.TRAVERSAL_EXACT, crt.TRAVERSAL_WIDE):
        crt.set_traversal(mode)
        with crt.Frame(medium_scene, 240, 160, 64) as fr:
            crt.set_traversal(-1)
            dO, dD, dH, dM = (L.rendererDeviceAlloc(16 * n) for _ in range(4))
            out = []
            for any_hit, tmax in ((0, float(FLT_MAX)), (1, 250.0)):
                L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, tmax)
                L.intersectBatchDeviceEx(dO, dD, n, dH, dM, any_hit)
                hit, mesh = np.zeros((n, 4), np.float32), np.zeros(n, np.int32)
                L.rendererCopyToHost(hit.ctypes.data, dH, 16 * n)
                L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
                out += [hit, mesh]
                if mode == crt.TRAVERSAL_WIDE:
                    assert crt.wide_info().active == 1 and crt.wide_info().lastBatchRedo < 1e-3 * n
            for p in (dO, dD, dH, dM):
                L.rendererDeviceFree(p)
            img = fr.run(8)
            st = crt.stats()
            out += [img, st.raysExtend + st.raysShadow, crt.wide_info().lastFrameRedo]
        res[mode] = out
    e, w = res[crt.TRAVERSAL_EXACT], res[crt.TRAVERSAL_WIDE]
    for k in range(4):
        assert np.array_equal(e[k].view(np.uint32), w[k].view(np.uint32)), k
    assert np.array_equal(e[4], w[4]) and e[5] == w[5]
    assert e[6] == 0 and w[6] < 1e-3 * w[5]


def _soup(rng, n, kind):
    """n triangles in the 64-byte layout ((n, 16) float32: v[3], texCoords[3], meshID) of a kind that stresses the builder / walk."""
    t = np.zeros((n, 16), np.float32)
    if kind == "uniform":          # small triangles everywhere
        c = rng.uniform(-50, 50, (n, 1, 3))
        v = c + rng.uniform(-2, 2, (n, 3, 3))
    elif kind == "mixed":          # a few huge triangles over many tiny ones, axis-aligned quads, needles
        c = rng.uniform(-50, 50, (n, 1, 3))
        size = np.where(rng.random((n, 1, 1)) < 0.02, 80.0, 0.5)
        v = c + rng.uniform(-1, 1, (n, 3, 3)) * size
        flat = rng.random(n) < 0.3
        v[flat, :, 1] = np.round(v[flat, :1, 1])           # axis-aligned: zero extent in y
        needle = rng.random(n) < 0.05
        v[needle, 2] = v[needle, 1] + 1e-4                  # nearly degenerate
    else:                          # "ties": every triangle twice (exact ties), some degenerate
        half = n // 2
        c = rng.uniform(-20, 20, (half, 1, 3))
        v = c + rng.uniform(-3, 3, (half, 3, 3))
        v = np.concatenate([v, v], axis=0)
        v[:8, 2] = v[:8, 1]                                 # zero area
        n = v.shape[0]
        t = np.zeros((n, 16), np.float32)
    t[:, :9] = v.reshape(n, 9).astype(np.float32)
    t[:, 9:15] = rng.random((n, 6)).astype(np.float32)
    t[:, 15] = (np.arange(n) % 20).astype(np.int32).view(np.float32)
    return t


@pytest.mark.parametrize("kind", ["uniform", "mixed", "ties"])
def test_wide_tree_equals_exact_walk_on_triangle_soups(crt, kind, tmp_path):
    """Geometry that is not the staircase: random soups with huge-over-tiny triangles, axis-aligned and nearly degenerate ones, and
    exact duplicates (every hit a tie). Closest-hit and any-hit batches -- random rays, axis-parallel rays, -0.0 components, origins
    on vertices -- through the renderer's own tree + certificate must equal the order-exact walk of the caller's tree bit for bit."""
    rng = np.random.default_rng({"uniform": 1, "mixed": 2, "ties": 3}[kind])
    tris = _soup(rng, 6000, kind)
    scene = crt.Scene.from_triangles(tris, 5, 16)
    n = 1 << 17
    ro = np.zeros((n, 4), np.float32)
    rd = np.zeros((n, 4), np.float32)
    ro[:, :3] = rng.uniform(-60, 60, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 8
    d[:k] = 0.0
    d[np.arange(k), rng.integers(0, 3, k)] = rng.choice([-1.0, 1.0], k)          # axis-parallel
    d[k:2 * k, 0] = -0.0                                                          # a signed zero component
    verts = tris[:, :9].reshape(-1, 3)
    ro[2 * k:3 * k, :3] = verts[rng.integers(0, verts.shape[0], k)]               # origins exactly on vertices
    target = verts[rng.integers(0, verts.shape[0], k)]
    aim = target - ro[3 * k:4 * k, :3]
    d[3 * k:4 * k] = aim / np.maximum(np.linalg.norm(aim, axis=1, keepdims=True), 1e-9)  # aimed exactly at vertices (edge hits)
    rd[:, :3] = d.astype(np.float32)
    ro[:, 3] = 0.001
    L = crt.device_lib()
    res = {}
    for mode in (crt.TRAVERSAL_EXACT, crt.TRAVERSAL_WIDE):
        crt.set_traversal(mode)
        with crt.Frame(scene, 64, 64, 8) as fr:
            crt.set_traversal(-1)
            out = []
            dO, dD, dH, dM = (L.rendererDeviceAlloc(16 * n) for _ in range(4))
            for any_hit, tmax in ((0, float(FLT_MAX)), (1, 40.0)):
                rd[:, 3] = tmax
                L.rendererCopyToDevice(dO, ro.ctypes.data, 16 * n)
                L.rendererCopyToDevice(dD, rd.ctypes.data, 16 * n)
                L.intersectBatchDeviceEx(dO, dD, n, dH, dM, any_hit)
                hit, mesh = np.zeros((n, 4), np.float32), np.zeros(n, np.int32)
                L.rendererCopyToHost(hit.ctypes.data, dH, 16 * n)
                L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
                out += [hit, mesh]
            for p in (dO, dD, dH, dM):
                L.rendererDeviceFree(p)
            if mode == crt.TRAVERSAL_WIDE:
                assert crt.wide_info().active == 1
            out.append(fr.run(4))
        res[mode] = out
    e, w = res[crt.TRAVERSAL_EXACT], res[crt.TRAVERSAL_WIDE]
    if HAVE_SHIM:  # and the reference's own hitMesh on the same file and rays: ids and mesh ids bit-exact, any-hit verdicts equal
        sys.path.insert(0, os.path.join(os.path.dirname(G), "..", "oracle"))
        import oracle as orc
        path = str(tmp_path / "soup.bvh")
        assert scene.save_bvh(path) == 0
        rd[:, 3] = FLT_MAX
        rhit, rmesh, _ = orc.ref_intersect_batch(path, 16, 5, ro, rd, False, str(tmp_path))
        assert np.array_equal(w[0][:, 3].view(np.uint32), rhit[:, 3].view(np.uint32)) and np.array_equal(w[1], rmesh)
        hit_ref = rmesh >= 0
        assert np.array_equal(w[0][hit_ref, 0].view(np.uint32), rhit[hit_ref, 0].view(np.uint32))  # t to the bit
        rd[:, 3] = 40.0
        ohit, _, _ = orc.ref_intersect_batch(path, 16, 5, ro, rd, True, str(tmp_path))
        assert np.array_equal(w[2][:, 0] == 0.0, ohit[:, 0] == 0.0)  # hitBvh returns 0.0f for an occluded shadow ray (kernels.cu:207)
    scene.close()
    assert (e[1] >= 0).mean() > 0.05  # the batch does hit things
    for i in range(4):
        assert np.array_equal(e[i].view(np.uint32), w[i].view(np.uint32)), i
    assert np.array_equal(e[4], w[4])


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_benchmark_scene_side_by_side_with_reference_kernel(crt, oracle):
    """BASELINE config 3's own scene and frame size (detail 1.0, 1024^2 textures, 1200x800, depth 64) at 4 spp against the
    reference's kernel on the same box. Stated bar: >= 99.9 % of pixels within 1e-3 and PSNR >= 50 dB; observed: identical
    bits. Ray counts: the reference cannot count its rays (its STATS build does not compile on Linux), bench.py's reference arm
    takes them from profiles/raycounts.json, which rests on the frames being identical -- checked here at 4 spp, and the
    100-spp count in the file must be what 25 of these frames extrapolate to within the run-to-run spread of path lengths."""
    nx, ny, ns, depth = 1200, 800, 4, 64
    tmp = tempfile.mkdtemp()
    ref, _ = oracle.ref_render(1.0, 1024, 5, nx, ny, ns, depth, os.path.join(tmp, "r.ref"))
    scene = crt.Scene.staircase(1.0, 1024, 5)
    with crt.Frame(scene, nx, ny, depth) as fr:
        img = fr.run(ns)
        st = crt.stats()
        info = crt.wide_info()
    scene.close()
    within, psnr, exact = frame_stats(img, ref)
    assert info.active == 1
    assert within >= 0.999 and psnr >= 50.0, (within, psnr, exact)
    assert exact == 1.0, f"expected a bit-exact frame, {exact:.6f} of pixels are"
    rays4 = st.raysExtend + st.raysShadow
    counts = json.load(open(os.path.join(os.path.dirname(G), "..", "profiles", "raycounts.json")))
    rays100 = counts["staircase:1.000:1024:1200x800x100:d64"]
    assert abs(rays100 / 25.0 - rays4) <= 0.01 * rays4, (rays100, rays4)


@pytest.mark.skipif(not HAVE_SHIM, reason="oracle/_ref not built")
def test_config5_rays_on_the_benchmark_bvh_side_by_side(crt, oracle):
    """4 Mi rays of BASELINE config 5 (half camera rays, half incoherent) against the config-3 BVH (detail 1.0): triangle and
    mesh ids bit-exact against the reference's hitMesh (oracle/ref_shim.cu), t within 1e-5 relative (observed: t, u, v exact)."""
    n = 1 << 22
    L = crt.device_lib()
    scene = crt.Scene.staircase(1.0, 32, 5)
    with crt.Frame(scene, 64, 64, 1):
        dO, dD, dH, dM = (L.rendererDeviceAlloc(16 * n) for _ in range(4))
        L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, float(FLT_MAX))
        L.intersectBatchDevice(dO, dD, n, dH, dM)
        ro, rd, hit, mesh = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros(n, np.int32)
        for a, p in ((ro, dO), (rd, dD), (hit, dH)):
            L.rendererCopyToHost(a.ctypes.data, p, 16 * n)
        L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
        for p in (dO, dD, dH, dM):
            L.rendererDeviceFree(p)
        assert crt.wide_info().active == 1
    scene.close()
    rhit, rmesh, _ = oracle.ref_intersect_batch(1.0, 32, 5, ro, rd, False, tempfile.mkdtemp())
    assert np.array_equal(hit[:, 3].view(np.uint32), rhit[:, 3].view(np.uint32)) and np.array_equal(mesh, rmesh)
    h = rmesh >= 0
    assert h.mean() > 0.9
    assert (np.abs(hit[h, 0] - rhit[h, 0]) / np.abs(rhit[h, 0])).max() <= 1e-5
    assert np.array_equal(hit.view(np.uint32), rhit.view(np.uint32)), "t/u/v are expected to be bit-exact as well"


def test_device_built_wide_tree_is_sound(crt, oracle, medium_scene, small_scene):
    """The tree the device build (csrc/wide_build.cuh) produces, downloaded and checked on the host: every real triangle exactly
    once, every quantised child box contains (padded) everything below it, reported depth = real depth. Also for a few tiny
    scenes (1, 2, 7 triangles; coincident centroids)."""
    import ctypes as C
    oracle.lib()
    W = C.CDLL(os.path.join(os.path.dirname(G), "..", "oracle", "build", "libwide_walk.so"))
    W.wideCheckGiven.argtypes = [C.POINTER(crt.KernelScene), C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.POINTER(C.c_int)]
    rng = np.random.default_rng(11)
    tiny = []
    for ntri in (1, 2, 7):
        t = np.zeros((ntri, 16), np.float32)
        t[:, :9] = rng.random((ntri, 9), dtype=np.float32) * 10
        tiny.append(crt.Scene.from_triangles(t, 5, 16))
    same = np.zeros((9, 16), np.float32)
    same[:, :9] = np.array([0, 0, 0, 10, 0, 0, 0, 10, 0], np.float32)  # nine coincident triangles: no plane separates their centroids
    tiny.append(crt.Scene.from_triangles(same, 5, 16))
    for scene in [small_scene, medium_scene] + tiny:
        with crt.Frame(scene, 8, 8, 1):
            info = crt.wide_info()
            assert info.active == 1 and info.buildThreads == 0, "the device build is the default"
            nodes = np.zeros((info.numNodes, 96), np.uint8)
            orig = np.zeros(info.numTriangles, np.uint32)
            assert crt.device_lib().getRendererWideTree(nodes.ctypes.data, info.numNodes, orig.ctypes.data, info.numTriangles) == info.numNodes
        assert info.numTriangles == scene.num_real_triangles
        depth = C.c_int()
        assert W.wideCheckGiven(C.byref(scene.ks), nodes.ctypes.data, info.numNodes, orig.ctypes.data, info.numTriangles, C.byref(depth)) == 0
        assert depth.value == info.depth
    for s in tiny:
        s.close()


def _gpu_count():
    try:
        import subprocess
        return len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout.strip().splitlines())
    except Exception:
        return 0


def test_changing_the_accumulator_between_runs_gets_a_fresh_graph(crt, medium_scene):
    """The captured CUDA graph carries the accumulator pointer by value (round-1 advisor finding): a caller that hands in its own
    device buffer AFTER a first runRenderer of the same initRenderer must still get the frame in that buffer, not in the old one."""
    L = crt.device_lib()
    nx, ny, ns = 96, 64, 4
    crt.set_options(defer_finalize=1)
    with crt.Frame(medium_scene, nx, ny, 16) as fr:
        L.runRenderer(ns, 8, 8)
        L.finalizeFrame(ns)
        first = fr.frame(copy=True)
        own = L.rendererDeviceAlloc(16 * nx * ny)
        L.setRendererAccumDevice(own)
        L.runRenderer(ns, 8, 8)
        sums = np.zeros((ny * nx, 4), np.float32)
        L.rendererCopyToHost(sums.ctypes.data, own, sums.nbytes)
        L.finalizeFrame(ns)
        second = fr.frame(copy=True)
        L.rendererDeviceFree(own)
    crt.set_options()
    assert np.array_equal(first, second)
    assert np.array_equal((sums[:, :3] / np.float32(ns)).reshape(ny, nx, 3), first)


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs in the box")
def test_multi_gpu_inside_the_library(crt, medium_scene):
    """setRendererGpus(2): one host thread per device inside libcrt_b200.so, sample streams 0 and 1, one ncclReduce, fb on device 0.
    The frame is the average of the two single-GPU frames of streams 0 and 1 (same sums, one float addition apart), the ray
    count their sum, and it agrees statistically with a one-GPU frame of the same total sample count."""
    nx, ny, ns, depth = 240, 160, 16, 64
    L = crt.device_lib()
    L.setRendererGpus(2)
    with crt.Frame(medium_scene, nx, ny, depth) as fr:
        assert L.getRendererGpus() == 2
        both = fr.run(ns)
        st = crt.stats()
        again = fr.run(ns)
    L.setRendererGpus(1)
    halves, rays = [], 0
    for g in (0, 1):
        crt.set_options(sample_stream=g)
        with crt.Frame(medium_scene, nx, ny, depth) as fr:
            halves.append(fr.run(ns // 2).astype(np.float64))
            s = crt.stats()
            rays += s.raysExtend + s.raysShadow
    crt.set_options()
    assert L.getRendererGpus() == 1
    assert np.array_equal(both, again)
    assert st.samples == nx * ny * ns and st.raysExtend + st.raysShadow == rays
    assert np.abs(both - 0.5 * (halves[0] + halves[1])).max() <= 1e-5 * max(1.0, float(both.max()))
    with crt.Frame(medium_scene, nx, ny, depth) as fr:
        single = fr.run(ns)
    assert abs(float(both.mean()) - float(single.mean())) <= 0.03 * float(single.mean())
