"""Host side (libcrt_host.so): ABI layouts, the BVH_00.04 builder/container, camera, frame files, scenes. CPU only."""
import ctypes as C
import os

import numpy as np
import pytest


def test_abi_layouts_match_reference_headers(crt):
    """Sizes/offsets of helper_structs.h / vec3.h as measured on the reference headers."""
    assert C.sizeof(crt.Vec3) == 12 and C.alignment(crt.Vec3) == 4
    assert C.sizeof(crt.Triangle) == 64 and crt.Triangle.texCoords.offset == 36 and crt.Triangle.meshID.offset == 60
    assert C.sizeof(crt.BvhNode) == 24
    assert C.sizeof(crt.Mesh) == 56 and (crt.Mesh.tris.offset, crt.Mesh.numTris.offset, crt.Mesh.bvh.offset,
                                          crt.Mesh.numBvhNodes.offset, crt.Mesh.bounds.offset) == (0, 8, 16, 24, 28)
    assert C.sizeof(crt.Material) == 24 and (crt.Material.color.offset, crt.Material.param.offset, crt.Material.texId.offset) == (4, 16, 20)
    assert C.sizeof(crt.STexture) == 16 and C.sizeof(crt.Sphere) == 16 and C.sizeof(crt.Plane) == 24
    assert C.sizeof(crt.Camera) == 88 and crt.Camera.lens_radius.offset == 84
    ks = crt.KernelScene
    assert C.sizeof(ks) == 64 and (ks.m.offset, ks.floor.offset, ks.materials.offset, ks.numMaterials.offset, ks.textures.offset,
                                    ks.numTextures.offset, ks.numPrimitivesPerLeaf.offset) == (0, 8, 32, 40, 48, 56, 60)


def test_c_abi_library_exports_every_declared_symbol(crt):
    """include/kernels.h: every prototype resolves in build/libcrt_b200.so (no compute call: there is no GPU here)."""
    import re
    root = crt.ROOT
    header = open(os.path.join(root, "include", "kernels.h")).read()
    body = header[header.index('extern "C" {'):]
    names = set(re.findall(r"^\s*(?:[\w\s\*]+?)\b(\w+)\(", body, flags=re.M)) - {"defined"}
    assert {"initRenderer", "runRenderer", "cleanupRenderer"} <= names
    lib = crt.device_lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in kernels.h but not exported"
    assert set(crt.DEVICE_SYMBOLS) == names


def _check_tree(nodes, tris, ppl):
    n_nodes = nodes.shape[0]
    assert n_nodes & (n_nodes - 1) == 0
    first_leaf = n_nodes // 2
    assert tris.shape[0] == first_leaf * ppl
    real = ~np.isinf(tris[:, 0])
    v = tris[:, :9].reshape(-1, 3, 3)
    # leaves bound their triangles exactly; unused slots come after the used ones
    for leaf in range(first_leaf):
        sl = slice(leaf * ppl, (leaf + 1) * ppl)
        r = real[sl]
        assert not (np.diff(r.astype(int)) > 0).any(), "a real triangle follows a sentinel inside a leaf"
        box = nodes[first_leaf + leaf]
        if r.any():
            pts = v[sl][r].reshape(-1, 3)
            assert np.array_equal(box[:3], pts.min(0)) and np.array_equal(box[3:], pts.max(0))
        else:
            assert (box[:3] > box[3:]).all()  # empty box: every slab test fails
    # parents are the union of their children
    for i in range(first_leaf - 1, 0, -1):
        l, r = nodes[2 * i], nodes[2 * i + 1]
        assert np.array_equal(nodes[i][:3], np.minimum(l[:3], r[:3])) and np.array_equal(nodes[i][3:], np.maximum(l[3:], r[3:]))


def test_bvh_builder_invariants(small_scene):
    """Implicit complete tree of kernels.cu:154-224/:614 and the BVH_00.04 slot layout (staircase_scene.h:75-101)."""
    s = small_scene
    tris, nodes = s.triangles().copy(), s.nodes().copy()
    ppl = s.ks.numPrimitivesPerLeaf
    assert ppl == 5 and int((~np.isinf(tris[:, 0])).sum()) == s.num_real_triangles
    _check_tree(nodes, tris, ppl)
    bmin, bmax = s.bounds()
    assert np.array_equal(bmin, nodes[1][:3]) and np.array_equal(bmax, nodes[1][3:])
    assert s.ks.numMaterials == 20 and s.ks.numTextures == 9
    mesh_ids = tris.view(np.uint8).reshape(-1, 64)[:, 60]
    assert mesh_ids[~np.isinf(tris[:, 0])].max() < 20


def test_sah_builder_same_layout_same_triangles(crt, small_scene):
    """BUILD_SAH: the same complete-tree BVH_00.04 layout, every triangle stored exactly once, all the tree
    invariants of the traversal -- only the split positions differ."""
    sah = crt.Scene.staircase(0.1, 32, 5, sah=True)
    assert sah.num_nodes == small_scene.num_nodes and sah.num_slots == small_scene.num_slots
    assert sah.num_real_triangles == small_scene.num_real_triangles and sah.hash() != small_scene.hash()
    tris, nodes = sah.triangles().copy(), sah.nodes().copy()
    _check_tree(nodes, tris, 5)
    real = lambda t: sorted(map(tuple, t[~np.isinf(t[:, 0])].view(np.uint32).tolist()))
    assert real(tris) == real(small_scene.triangles().copy())
    assert np.array_equal(sah.bounds()[0], small_scene.bounds()[0]) and np.array_equal(sah.bounds()[1], small_scene.bounds()[1])
    sah.close()


@pytest.mark.parametrize("n,ppl", [(0, 5), (1, 5), (5, 5), (6, 5), (37, 1), (100, 3)])
def test_bvh_builder_ragged_inputs(crt, n, ppl):
    rng = np.random.default_rng(n * 31 + ppl)
    t = np.zeros((n, 16), np.float32)
    base = rng.uniform(-10, 10, (n, 1, 3)).astype(np.float32)
    t[:, :9] = (base + rng.uniform(-1, 1, (n, 3, 3)).astype(np.float32)).reshape(n, 9)
    s = crt.Scene.from_triangles(t, ppl, 4)
    assert s.num_real_triangles == n
    assert s.num_slots >= n and s.num_slots == (s.num_nodes // 2) * ppl
    _check_tree(s.nodes().copy(), s.triangles().copy(), ppl)
    # every input triangle is stored exactly once
    stored = s.triangles()[~np.isinf(s.triangles()[:, 0])][:, :9]
    assert sorted(map(tuple, stored.tolist())) == sorted(map(tuple, t[:, :9].tolist()))
    s.close()


def test_bvh_file_round_trip(crt, small_scene, tmp_path):
    path = str(tmp_path / "scene.bvh")
    assert small_scene.save_bvh(path) == 0
    raw = open(path, "rb").read()
    assert raw[:10] == b"BVH_00.04\x00"
    n_tris = int(np.frombuffer(raw[10:14], np.int32)[0])
    assert n_tris == small_scene.num_slots
    assert len(raw) == 10 + 4 + 64 * n_tris + 4 + 24 * small_scene.num_nodes + 24 + 4
    s2 = crt.Scene.from_bvh_file(path, 32)
    assert s2.hash() == small_scene.hash()
    assert np.array_equal(s2.triangles(), small_scene.triangles()) and np.array_equal(s2.nodes(), small_scene.nodes())
    s2.close()
    # truncated file and bad magic are refused
    open(path, "wb").write(raw[:-30])
    with pytest.raises(RuntimeError):
        crt.Scene.from_bvh_file(path, 32)
    open(path, "wb").write(b"BVH_00.03\x00" + raw[10:])
    with pytest.raises(RuntimeError):
        crt.Scene.from_bvh_file(path, 32)


def test_scene_is_deterministic(crt, small_scene):
    again = crt.Scene.staircase(0.1, 32, 5)
    assert again.hash() == small_scene.hash()
    again.close()


def test_staircase_camera_matches_reference_ctor(crt):
    """setup_camera (staircase_scene.h:62-73) through the camera ctor (helper_structs.h:194-206), recomputed in float32."""
    nx, ny = 1200, 800
    cam = crt.staircase_camera(nx, ny)
    f = np.float32
    frm = np.array([5.555139, 173.679901, 494.515045], f)
    at = np.array([5.555139, 173.679901, 493.515045], f)
    w = frm - at
    w = w / f(np.sqrt(f(w @ w)))
    assert np.allclose(list(cam.w.e), w) and np.allclose(list(cam.u.e), [1, 0, 0]) and np.allclose(list(cam.v.e), [0, 1, 0])
    assert cam.lens_radius == 0.0 and list(cam.origin.e) == list(frm)
    half_h = np.tan(f(42.0) * f(np.pi) / f(180) / f(2))
    assert np.allclose(list(cam.vertical.e), [0, 2 * half_h, 0], rtol=1e-6)
    assert np.allclose(list(cam.horizontal.e), [2 * half_h * nx / ny, 0, 0], rtol=1e-6)
    llc = np.array(list(cam.lower_left_corner.e))
    assert np.allclose(llc + np.array(list(cam.horizontal.e)) / 2 + np.array(list(cam.vertical.e)) / 2, at, atol=1e-4)


def test_rtiow_scene_shape(crt):
    """README.md:3-6 lineage: 1 ground + 22x22 small + 3 big spheres from the host LCG (main.cpp:17-20), seed 1."""
    sph, mats, n = crt.rtiow_scene(1)
    assert n == 488
    assert sph[0].radius == 1000.0 and [sph[i].radius for i in (485, 486, 487)] == [1.0, 1.0, 1.0]
    types = np.array([mats[i].type for i in range(1, 485)])
    frac = [(types == k).mean() for k in (0, 1, 2)]
    assert 0.7 < frac[0] < 0.9 and 0.08 < frac[1] < 0.22 and 0.01 < frac[2] < 0.1
    sph2, mats2, n2 = crt.rtiow_scene(1)
    assert bytes(sph) == bytes(sph2) and bytes(mats) == bytes(mats2)
    # LCG known answer: state = 214013*1 + 2531011 -> ((state >> 16) & 0x7FFF) / 32767
    st = (214013 * 1 + 2531011) & 0xFFFFFFFF
    first = np.float32(((st >> 16) & 0x7FFF)) / np.float32(32767)
    st = (214013 * st + 2531011) & 0xFFFFFFFF
    second = np.float32(((st >> 16) & 0x7FFF)) / np.float32(32767)
    assert sph[1].center.e[0] == np.float32(-11) + second and first < 1.0


def test_frame_files_and_srgb(crt, tmp_path):
    """REF_00.01 (main.cpp:25-60) round trip; PPM/sRGB (staircase_scene.h:22-43); RMSE (main.cpp:108-128)."""
    rng = np.random.default_rng(5)
    img = rng.random((7, 11, 3), dtype=np.float32)
    p = str(tmp_path / "f.ref")
    assert crt.write_ref(p, img) == 0
    raw = open(p, "rb").read()
    assert raw[:10] == b"REF_00.01\x00" and np.frombuffer(raw[10:18], np.int32).tolist() == [11, 7] and len(raw) == 18 + 7 * 11 * 12
    assert np.array_equal(crt.read_ref(p, 11, 7), img)
    with pytest.raises(IOError):
        crt.read_ref(p, 7, 11)  # wrong size is refused (main.cpp:52-55)
    H = crt.host_lib()
    for x, want in [(-1.0, 0), (0.0, 0), (1.0, 255), (4.0, 255), (0.5, int((1.055 * 0.5 ** 0.416666667 - 0.055) * 255.9))]:
        assert H.crtLinearToSRGB(x) == want
    ppm = str(tmp_path / "f.ppm")
    assert crt.write_ppm(ppm, img) == 0
    toks = open(ppm).read().split()
    assert toks[:4] == ["P3", "11", "7", "255"] and len(toks) == 4 + 7 * 11 * 3
    assert int(toks[4]) == H.crtLinearToSRGB(float(img[6, 0, 0]))  # rows are written top-down
    other = img + np.float32(0.5)
    assert abs(H.crtRmse(img.ctypes.data, other.ctypes.data, 11, 7) - 0.5) < 1e-6
