"""ctypes bindings of the two product libraries (no compute in Python, no fallback).

  build/libcrt_host.so   host side: scenes, BVH_00.04, camera, frame files   (host/host_api.h)
  build/libcrt_b200.so   device side: the C ABI of include/kernels.h         (csrc/renderer.cu)

The three reference entry points (reference kernels.h:6-8) are bound with the reference's own
signatures -- kernel_scene and camera BY VALUE -- exactly as a C++ caller sees them.  If the CUDA
library is missing or cannot be loaded this module raises: there is deliberately no CPU path.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BUILD = os.path.join(ROOT, "build")


class Vec3(C.Structure):
    _fields_ = [("e", C.c_float * 3)]


class BBox(C.Structure):
    _fields_ = [("min", Vec3), ("max", Vec3)]


class Triangle(C.Structure):
    _fields_ = [("v", Vec3 * 3), ("texCoords", C.c_float * 6), ("meshID", C.c_ubyte)]


class BvhNode(C.Structure):
    _fields_ = [("a", Vec3), ("b", Vec3)]


class Mesh(C.Structure):
    _fields_ = [("tris", C.POINTER(Triangle)), ("numTris", C.c_uint32), ("bvh", C.POINTER(BvhNode)),
                ("numBvhNodes", C.c_int), ("bounds", BBox)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int), ("color", Vec3), ("param", C.c_float), ("texId", C.c_int)]


class STexture(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("width", C.c_int), ("height", C.c_int)]


class Plane(C.Structure):
    _fields_ = [("norm", Vec3), ("point", Vec3)]


class Sphere(C.Structure):
    _fields_ = [("center", Vec3), ("radius", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("origin", Vec3), ("lower_left_corner", Vec3), ("horizontal", Vec3), ("vertical", Vec3),
                ("u", Vec3), ("v", Vec3), ("w", Vec3), ("lens_radius", C.c_float)]


class KernelScene(C.Structure):
    _fields_ = [("m", C.POINTER(Mesh)), ("floor", Plane), ("materials", C.POINTER(Material)), ("numMaterials", C.c_int),
                ("textures", C.POINTER(STexture)), ("numTextures", C.c_int), ("numPrimitivesPerLeaf", C.c_int)]


class RendererOptions(C.Structure):
    _fields_ = [("device", C.c_int), ("sampleStream", C.c_uint), ("deferFinalize", C.c_int),
                ("resetDeviceOnCleanup", C.c_int), ("megaBatch", C.c_int), ("reserved", C.c_int * 3)]


class RendererWideInfo(C.Structure):
    _fields_ = [("active", C.c_int), ("traversal", C.c_int), ("numNodes", C.c_uint), ("numTriangles", C.c_uint), ("depth", C.c_int),
                ("buildThreads", C.c_int), ("buildMs", C.c_float), ("sahCost", C.c_float), ("lastBatchRedo", C.c_ulonglong),
                ("lastFrameRedo", C.c_ulonglong)]


class RendererStats(C.Structure):
    _fields_ = [("raysExtend", C.c_ulonglong), ("raysShadow", C.c_ulonglong), ("samples", C.c_ulonglong),
                ("kernelLaunches", C.c_ulonglong), ("iterations", C.c_ulonglong), ("resumes", C.c_ulonglong),
                ("deferred", C.c_ulonglong), ("msTotal", C.c_float), ("msTrace", C.c_float), ("msShade", C.c_float),
                ("msOther", C.c_float), ("profiled", C.c_int)]


assert C.sizeof(Vec3) == 12 and C.sizeof(Triangle) == 64 and C.sizeof(BvhNode) == 24 and C.sizeof(Mesh) == 56
assert C.sizeof(Material) == 24 and C.sizeof(STexture) == 16 and C.sizeof(Camera) == 88 and C.sizeof(KernelScene) == 64

_host = None
_dev = None


def host_lib():
    global _host
    if _host is None:
        path = os.path.join(BUILD, "libcrt_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make` (or __graft_entry__.build()) first")
        L = C.CDLL(path)
        L.crtSceneCreateStaircase.restype = C.c_void_p
        L.crtSceneCreateStaircase.argtypes = [C.c_float, C.c_int, C.c_int]
        L.crtSceneCreateStaircaseEx.restype = C.c_void_p
        L.crtSceneCreateStaircaseEx.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int]
        L.crtSceneLoadBVH.restype = C.c_void_p
        L.crtSceneLoadBVH.argtypes = [C.c_char_p, C.c_int]
        L.crtSceneFromTriangles.restype = C.c_void_p
        L.crtSceneFromTriangles.argtypes = [C.POINTER(Triangle), C.c_int, C.c_int, C.c_int]
        L.crtSceneDestroy.argtypes = [C.c_void_p]
        L.crtSceneSaveBVH.argtypes = [C.c_void_p, C.c_char_p]
        L.crtSceneLoadTexturePNG.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.crtSceneLoadTextureDir.argtypes = [C.c_void_p, C.c_char_p]
        L.crtSceneTextureInfo.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_float))]
        L.crtDecodePNG.argtypes = [C.c_char_p, C.c_ulonglong, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_ulonglong]
        L.crtSceneKernelScene.restype = C.POINTER(KernelScene)
        L.crtSceneKernelScene.argtypes = [C.c_void_p]
        L.crtSceneNumRealTriangles.argtypes = [C.c_void_p]
        L.crtSceneHash.restype = C.c_ulonglong
        L.crtSceneHash.argtypes = [C.c_void_p]
        L.crtMakeCamera.argtypes = [C.c_float * 3, C.c_float * 3, C.c_float * 3, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.POINTER(Camera)]
        L.crtStaircaseCamera.argtypes = [C.c_int, C.c_int, C.POINTER(Camera)]
        L.crtRtiowScene.argtypes = [C.c_uint, C.POINTER(Sphere), C.POINTER(Material), C.c_int]
        L.crtRtiowCamera.argtypes = [C.c_int, C.c_int, C.POINTER(Camera)]
        L.crtWritePPM.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.crtWriteRef.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.crtReadRef.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.crtLinearToSRGB.restype = C.c_uint
        L.crtLinearToSRGB.argtypes = [C.c_float]
        L.crtRmse.restype = C.c_double
        L.crtRmse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _host = L
    return _host


DEVICE_SYMBOLS = [
    "initRenderer", "runRenderer", "cleanupRenderer", "setRendererOptions", "initRendererSpheres", "intersectBatch",
    "intersectBatchDevice", "generateRayBatchDevice", "rendererDeviceAlloc", "rendererDeviceFree", "rendererCopyToHost",
    "rendererCopyToDevice", "getRendererStats", "setRendererProfiling", "getRendererAccumDevice", "setRendererAccumDevice",
    "finalizeFrame", "setRendererCounting", "getRendererTraversalCounts", "rendererDebugRead", "rendererReleaseCaches", "getRendererChaserCounts", "continueRenderer", "getRendererSamplesDone",
    "saveRendererCheckpoint", "loadRendererCheckpoint", "scatterBatch", "intersectBatchDeviceEx", "setRendererTraversal",
    "getRendererWideInfo", "rendererTrigSelfTest", "getRendererWideTree", "setRendererGpus", "getRendererGpus",
]

TRAVERSAL_WIDE, TRAVERSAL_EXACT, TRAVERSAL_WIDE_UNCERTIFIED = 0, 1, 2


def device_lib():
    """The CUDA library. Loading needs libcudart's dependencies only; calling anything needs a GPU."""
    global _dev
    if _dev is None:
        path = os.environ.get("CRT_B200_LIB") or os.path.join(BUILD, "libcrt_b200.so")  # override: A/B runs of two builds
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: the CUDA extension was not built; there is no CPU fallback")
        L = C.CDLL(path)
        L.initRenderer.restype = None
        L.initRenderer.argtypes = [KernelScene, Camera, C.POINTER(C.POINTER(Vec3)), C.c_int, C.c_int, C.c_int]
        L.runRenderer.restype = None
        L.runRenderer.argtypes = [C.c_int, C.c_int, C.c_int]
        L.cleanupRenderer.restype = None
        L.setRendererOptions.argtypes = [C.POINTER(RendererOptions)]
        L.initRendererSpheres.restype = None
        L.initRendererSpheres.argtypes = [C.POINTER(Sphere), C.POINTER(Material), C.c_int, Camera, C.POINTER(C.POINTER(Vec3)),
                                          C.c_int, C.c_int, C.c_int]
        L.intersectBatch.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.intersectBatchDevice.restype = C.c_float
        L.intersectBatchDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        L.intersectBatchDeviceEx.restype = C.c_float
        L.intersectBatchDeviceEx.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int]
        L.setRendererTraversal.argtypes = [C.c_int]
        L.rendererTrigSelfTest.restype = C.c_longlong
        L.setRendererGpus.argtypes = [C.c_int]
        L.getRendererWideTree.restype = C.c_uint
        L.getRendererWideTree.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
        L.getRendererWideInfo.argtypes = [C.POINTER(RendererWideInfo)]
        L.generateRayBatchDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_float, C.c_float]
        L.rendererDeviceAlloc.restype = C.c_void_p
        L.rendererDeviceAlloc.argtypes = [C.c_size_t]
        L.rendererDeviceFree.argtypes = [C.c_void_p]
        L.rendererCopyToHost.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.rendererCopyToDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.getRendererStats.argtypes = [C.POINTER(RendererStats)]
        L.setRendererProfiling.argtypes = [C.c_int]
        L.getRendererAccumDevice.restype = C.c_void_p
        L.setRendererAccumDevice.argtypes = [C.c_void_p]
        L.finalizeFrame.argtypes = [C.c_int]
        L.rendererDebugRead.restype = C.c_size_t
        L.rendererDebugRead.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
        L.setRendererCounting.argtypes = [C.c_int]
        L.getRendererTraversalCounts.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        L.getRendererChaserCounts.argtypes = [C.POINTER(C.c_ulonglong)] * 4
        L.scatterBatch.argtypes = [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
        L.continueRenderer.argtypes = [C.c_int, C.c_int, C.c_int]
        L.saveRendererCheckpoint.argtypes = [C.c_char_p]
        L.loadRendererCheckpoint.argtypes = [C.c_char_p]
        _dev = L
    return _dev


class Scene:
    """Owns a crt_scene (host memory: triangles, BVH nodes, materials, textures)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("scene creation failed")
        self.handle = handle
        self.ks = host_lib().crtSceneKernelScene(handle).contents

    @classmethod
    def staircase(cls, detail=1.0, tex_size=1024, prims_per_leaf=5, sah=False):
        if sah:  # same mesh, BVH split by the surface-area heuristic inside the same complete-tree layout
            return cls(host_lib().crtSceneCreateStaircaseEx(detail, tex_size, prims_per_leaf, 1))
        return cls(host_lib().crtSceneCreateStaircase(detail, tex_size, prims_per_leaf))

    @classmethod
    def from_bvh_file(cls, path, tex_size=1024):
        return cls(host_lib().crtSceneLoadBVH(os.fsencode(path), tex_size))

    @classmethod
    def from_triangles(cls, tris, prims_per_leaf=5, tex_size=16):
        """tris: numpy array (n, 16) float32 in the 64-byte triangle layout, or a ctypes Triangle array."""
        if isinstance(tris, np.ndarray):
            arr = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 16)
            ptr = arr.ctypes.data_as(C.POINTER(Triangle))
            n = arr.shape[0]
        else:
            ptr, n = tris, len(tris)
        return cls(host_lib().crtSceneFromTriangles(ptr, n, prims_per_leaf, tex_size))

    def save_bvh(self, path):
        return host_lib().crtSceneSaveBVH(self.handle, os.fsencode(path))

    def load_texture_png(self, index, path):
        """loadTexture (staircase_scene.h:103-118): texture `index` from a PNG file; 0 on success."""
        return host_lib().crtSceneLoadTexturePNG(self.handle, index, os.fsencode(path))

    def load_texture_dir(self, path):
        """The nine file names of load_scene (staircase_scene.h:125-133) from a directory; returns how many loaded."""
        return host_lib().crtSceneLoadTextureDir(self.handle, os.fsencode(path))

    def texture(self, index):
        """Texture `index` as the floats the device library uploads: (height, width, 3) float32."""
        w, h, d = C.c_int(), C.c_int(), C.POINTER(C.c_float)()
        if host_lib().crtSceneTextureInfo(self.handle, index, C.byref(w), C.byref(h), C.byref(d)) != 0:
            raise IndexError(index)
        return np.ctypeslib.as_array(d, shape=(h.value, w.value, 3)).copy()

    @property
    def num_slots(self):
        return int(self.ks.m.contents.numTris)

    @property
    def num_real_triangles(self):
        return host_lib().crtSceneNumRealTriangles(self.handle)

    @property
    def num_nodes(self):
        return int(self.ks.m.contents.numBvhNodes)

    def triangles(self):
        m = self.ks.m.contents
        return np.ctypeslib.as_array(C.cast(m.tris, C.POINTER(C.c_float)), shape=(m.numTris, 16))

    def nodes(self):
        m = self.ks.m.contents
        return np.ctypeslib.as_array(C.cast(m.bvh, C.POINTER(C.c_float)), shape=(m.numBvhNodes, 6))

    def bounds(self):
        b = self.ks.m.contents.bounds
        return np.array(list(b.min.e)), np.array(list(b.max.e))

    def hash(self):
        return host_lib().crtSceneHash(self.handle)

    def close(self):
        if self.handle:
            host_lib().crtSceneDestroy(self.handle)
            self.handle = None


def staircase_camera(nx, ny):
    cam = Camera()
    host_lib().crtStaircaseCamera(nx, ny, C.byref(cam))
    return cam


def make_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus):
    cam = Camera()
    f3 = C.c_float * 3
    host_lib().crtMakeCamera(f3(*lookfrom), f3(*lookat), f3(*vup), vfov, aspect, aperture, focus, C.byref(cam))
    return cam


def rtiow_scene(seed=1):
    sph = (Sphere * 1024)()
    mats = (Material * 1024)()
    n = host_lib().crtRtiowScene(seed, sph, mats, 1024)
    return sph, mats, n


def rtiow_camera(nx, ny):
    cam = Camera()
    host_lib().crtRtiowCamera(nx, ny, C.byref(cam))
    return cam


def set_options(device=-1, sample_stream=0, defer_finalize=0, reset_on_cleanup=0, mega_batch=0, slots_per_pixel=0, trace_budget=0,
                trace_min_active=0):
    o = RendererOptions(device, sample_stream, defer_finalize, reset_on_cleanup, mega_batch,
                        (C.c_int * 3)(slots_per_pixel, trace_budget, trace_min_active))
    device_lib().setRendererOptions(C.byref(o))


def set_traversal(mode):
    """TRAVERSAL_WIDE (default) / TRAVERSAL_EXACT / TRAVERSAL_WIDE_UNCERTIFIED for the next initRenderer; -1 = default."""
    device_lib().setRendererTraversal(mode)


def wide_info():
    w = RendererWideInfo()
    device_lib().getRendererWideInfo(C.byref(w))
    return w


def stats():
    s = RendererStats()
    device_lib().getRendererStats(C.byref(s))
    return s


class Frame:
    """initRenderer ... cleanupRenderer around one scene; run() returns the frame as (ny, nx, 3) float32 (row 0 = bottom)."""

    def __init__(self, scene, nx, ny, max_depth, cam=None):
        self.nx, self.ny = nx, ny
        self.fb = C.POINTER(Vec3)()
        L = device_lib()
        if isinstance(scene, Scene):
            L.initRenderer(scene.ks, cam or staircase_camera(nx, ny), C.byref(self.fb), nx, ny, max_depth)
        else:
            sph, mats, n = scene
            L.initRendererSpheres(sph, mats, n, cam or rtiow_camera(nx, ny), C.byref(self.fb), nx, ny, max_depth)
        self.open = True

    def run(self, ns, copy=True):
        device_lib().runRenderer(ns, 8, 8)
        return self.frame(copy)

    def continue_run(self, ns_more, copy=True):
        if device_lib().continueRenderer(ns_more, 8, 8) != 0:
            raise RuntimeError("continueRenderer: nothing to continue")
        return self.frame(copy)

    def frame(self, copy=True):
        a = np.ctypeslib.as_array(C.cast(self.fb, C.POINTER(C.c_float)), shape=(self.ny, self.nx, 3))
        return a.copy() if copy else a

    def close(self):
        if self.open:
            device_lib().cleanupRenderer()
            self.open = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def write_ppm(path, img):
    img = np.ascontiguousarray(img, dtype=np.float32)
    return host_lib().crtWritePPM(os.fsencode(path), img.shape[1], img.shape[0], img.ctypes.data)


def write_ref(path, img):
    img = np.ascontiguousarray(img, dtype=np.float32)
    return host_lib().crtWriteRef(os.fsencode(path), img.shape[1], img.shape[0], img.ctypes.data)


def decode_png(data, flip=False):
    """PNG bytes -> (height, width, 3) uint8 through the host library's decoder; None when it rejects the stream."""
    w, h = C.c_int(), C.c_int()
    if host_lib().crtDecodePNG(data, len(data), 1 if flip else 0, C.byref(w), C.byref(h), None, 0) != 0:
        return None
    out = np.zeros((h.value, w.value, 3), np.uint8)
    rc = host_lib().crtDecodePNG(data, len(data), 1 if flip else 0, C.byref(w), C.byref(h), out.ctypes.data, out.size)
    return out if rc == 0 else None


def read_ref(path, nx, ny):
    img = np.zeros((ny, nx, 3), dtype=np.float32)
    rc = host_lib().crtReadRef(os.fsencode(path), nx, ny, img.ctypes.data)
    if rc != 0:
        raise IOError(f"cannot read REF_00.01 frame {path} (rc={rc})")
    return img
