#!/usr/bin/env python
"""GPU-box diagnostic: times the ray-batch kernel on the coherent (camera) half and the incoherent half of the config-5
batch separately, and on a shuffled copy of the camera half -- how much does ray order matter to the traversal kernel?"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
scene = crt.Scene.staircase(1.0, 64, 5)
with crt.Frame(scene, 64, 64, 8) as fr:
    L = crt.device_lib()
    dO, dD = L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n)
    dH, dM = L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(4 * n)
    L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, 3.4028234663852886e38)
    h = n // 2
    res = {}
    for name, off in (("camera_half", 0), ("incoherent_half", h)):
        ms = [L.intersectBatchDevice(C.c_void_p(dO + 16 * off), C.c_void_p(dD + 16 * off), h, dH, dM) for _ in range(4)]
        res[name] = dict(ms=min(ms), grays=h / min(ms) / 1e6)
    # shuffled camera half
    ro, rd = np.zeros((h, 4), np.float32), np.zeros((h, 4), np.float32)
    L.rendererCopyToHost(ro.ctypes.data, dO, 16 * h)
    L.rendererCopyToHost(rd.ctypes.data, dD, 16 * h)
    perm = np.random.default_rng(1).permutation(h)
    ro, rd = np.ascontiguousarray(ro[perm]), np.ascontiguousarray(rd[perm])
    L.rendererCopyToDevice(dO, ro.ctypes.data, 16 * h)
    L.rendererCopyToDevice(dD, rd.ctypes.data, 16 * h)
    ms = [L.intersectBatchDevice(dO, dD, h, dH, dM) for _ in range(4)]
    res["camera_half_shuffled"] = dict(ms=min(ms), grays=h / min(ms) / 1e6)
    print(json.dumps(res))
