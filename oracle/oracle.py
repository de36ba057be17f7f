"""TEST INFRASTRUCTURE: ctypes loader for the CPU restatement (oracle/build/liboracle_cpu.so) and helpers that run
the reference's own CUDA build (oracle/_ref/*) in a subprocess.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402  (struct layouts + the host library; not the CUDA library)

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "build", "liboracle_cpu.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "cpu"], stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.oracleRender.argtypes = [C.POINTER(crt.KernelScene), C.POINTER(crt.Camera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.oracleRenderSpheres.argtypes = [C.POINTER(crt.Sphere), C.POINTER(crt.Material), C.c_int, C.POINTER(crt.Camera), C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.oracleIntersectBatch.argtypes = [C.POINTER(crt.KernelScene), C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p,
                                           C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.oracleWangHash.restype = C.c_uint
        L.oracleWangHash.argtypes = [C.c_uint]
        L.oraclePathSeed.restype = C.c_uint
        L.oraclePathSeed.argtypes = [C.c_uint]
        L.oracleXorShift.restype = C.c_uint
        L.oracleXorShift.argtypes = [C.POINTER(C.c_uint)]
        L.oracleRnd.restype = C.c_float
        L.oracleRnd.argtypes = [C.POINTER(C.c_uint)]
        L.oracleUnitSphere.argtypes = [C.POINTER(C.c_uint), C.c_float * 3]
        L.oracleUnitDisk.argtypes = [C.POINTER(C.c_uint), C.c_float * 3]
        L.oracleGetRay.argtypes = [C.POINTER(crt.Camera), C.c_float, C.c_float, C.POINTER(C.c_uint), C.c_float * 3, C.c_float * 3]
        L.oracleTriangleHit.restype = C.c_float
        L.oracleTriangleHit.argtypes = [C.POINTER(crt.Triangle), C.c_float * 3, C.c_float * 3, C.c_float, C.c_float,
                                        C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracleSphereHit.restype = C.c_float
        L.oracleSphereHit.argtypes = [C.POINTER(crt.Sphere), C.c_float * 3, C.c_float * 3, C.c_float, C.c_float]
        L.oracleBoxDist.restype = C.c_float
        L.oracleBoxDist.argtypes = [C.c_float * 3, C.c_float * 3, C.c_float * 3, C.c_float * 3, C.c_float]
        L.oracleNumThreads.restype = C.c_int
        L.oracleScatterBatch.argtypes = [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def render(scene, nx, ny, ns, max_depth, cam=None, stream=0, rows=None, count=False, row_stride=1):
    """CPU restatement of render() (kernels.cu:535). Returns (frame (ny,nx,3), counters dict or None)."""
    cam = cam or crt.staircase_camera(nx, ny)
    fb = np.zeros((ny, nx, 3), dtype=np.float32)
    r0, r1 = rows if rows else (0, ny)
    cnt = (C.c_ulonglong * 5)() if count else None
    lib().oracleRender(C.byref(scene.ks), C.byref(cam), nx, ny, ns, max_depth, stream, r0, r1, row_stride, fb.ctypes.data, cnt)
    return fb, (dict(primary=cnt[0], secondary=cnt[1], shadow=cnt[2], nodeVisits=cnt[3], triTests=cnt[4]) if count else None)


def render_spheres(spheres, nx, ny, ns, max_depth, cam=None, stream=0, rows=None, count=False, row_stride=1):
    sph, mats, n = spheres
    cam = cam or crt.rtiow_camera(nx, ny)
    fb = np.zeros((ny, nx, 3), dtype=np.float32)
    r0, r1 = rows if rows else (0, ny)
    cnt = (C.c_ulonglong * 5)() if count else None
    lib().oracleRenderSpheres(sph, mats, n, C.byref(cam), nx, ny, ns, max_depth, stream, r0, r1, row_stride, fb.ctypes.data, cnt)
    return fb, (dict(primary=cnt[0], secondary=cnt[1]) if count else None)


def intersect_batch(scene, ray_o, ray_d, any_hit=False, count=False):
    """ray_o/ray_d: (n,4) float32 {o,tMin} {d,tMax}. Returns hit (n,4) float32 view {t,u,v,triId-bits}, meshId (n,) int32."""
    ray_o = np.ascontiguousarray(ray_o, dtype=np.float32)
    ray_d = np.ascontiguousarray(ray_d, dtype=np.float32)
    n = ray_o.shape[0]
    hit = np.zeros((n, 4), dtype=np.float32)
    mesh = np.zeros(n, dtype=np.int32)
    cnt = (C.c_ulonglong * 5)() if count else None
    lib().oracleIntersectBatch(C.byref(scene.ks), ray_o.ctypes.data, ray_d.ctypes.data, n, int(any_hit), hit.ctypes.data,
                               mesh.ctypes.data, cnt)
    if count:
        return hit, mesh, dict(nodeVisits=cnt[3], triTests=cnt[4])
    return hit, mesh


# ---------------------------------------------------------------- the reference's own CUDA build (GPU box only) --
def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("ref_driver", "ref_shim_driver", "libref.so", "libref_shim.so"))


def _run(cmd, timeout=1800):
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(f"{' '.join(map(str, cmd))} failed ({out.returncode}):\n{out.stdout}\n{out.stderr}")
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else {}


def ref_render(scene_spec, tex_size, ppl, nx, ny, ns, max_depth, out_path, warmup=0, steps=1, driver="ref_driver"):
    """Runs the UNMODIFIED reference kernel (libref.so) in its own process; returns (frame, info)."""
    info = _run([os.path.join(REF_DIR, driver), "render", str(scene_spec), str(tex_size), str(ppl), str(nx), str(ny), str(ns),
                 str(max_depth), str(warmup), str(steps), out_path])
    return (crt.read_ref(out_path, nx, ny) if out_path != "-" else None), info


def ref_spheres(seed, nx, ny, ns, max_depth, out_path, warmup=0, steps=1):
    info = _run([os.path.join(REF_DIR, "ref_shim_driver"), "spheres", str(seed), str(nx), str(ny), str(ns), str(max_depth),
                 str(warmup), str(steps), out_path])
    return (crt.read_ref(out_path, nx, ny) if out_path != "-" else None), info


def ref_intersect_batch(scene_spec, tex_size, ppl, ray_o, ray_d, any_hit, workdir):
    """hitMesh() of the reference (through oracle/ref_shim.cu) on a ray batch."""
    n = ray_o.shape[0]
    rays = os.path.join(workdir, "rays.bin")
    hits = os.path.join(workdir, "hits.bin")
    with open(rays, "wb") as f:
        f.write(np.int64(n).tobytes())
        f.write(np.ascontiguousarray(ray_o, dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(ray_d, dtype=np.float32).tobytes())
    info = _run([os.path.join(REF_DIR, "ref_shim_driver"), "batch", str(scene_spec), str(tex_size), str(ppl), rays, str(int(any_hit)), hits])
    raw = np.fromfile(hits, dtype=np.uint8)
    hit = raw[8:8 + 16 * n].view(np.float32).reshape(n, 4)
    mesh = raw[8 + 16 * n:8 + 20 * n].view(np.int32)
    return hit, mesh, info


def ref_scatter_batch(preset, items, workdir):
    """The reference's own BSDF preset `preset` (scene_materials.h, through oracle/ref_shim.cu) on `items` (n, 12) float32."""
    n = items.shape[0]
    fin, fout = os.path.join(workdir, "scatter_in.bin"), os.path.join(workdir, "scatter_out.bin")
    with open(fin, "wb") as f:
        f.write(np.int64(n).tobytes())
        f.write(np.ascontiguousarray(items, dtype=np.float32).tobytes())
    _run([os.path.join(REF_DIR, "ref_shim_driver"), "scatter", str(preset), fin, fout])
    raw = np.fromfile(fout, dtype=np.uint8)
    return raw[8:8 + 48 * n].view(np.float32).reshape(n, 12).copy()


def scatter_batch(preset, items):
    """CPU restatement of the BSDF presets (cpu_oracle.cpp preset_scatter) on `items` (n, 12) float32."""
    items = np.ascontiguousarray(items, dtype=np.float32)
    out = np.zeros_like(items)
    assert lib().oracleScatterBatch(preset, items.shape[0], items.ctypes.data, out.ctypes.data) == 0
    return out
