#!/usr/bin/env python
"""GPU-box diagnostic: how much does the ORDER of real secondary rays matter to the traversal kernel?
Renders 1 spp, reads back every slot's last extend ray (a mix of bounce depths), and times the ray-batch kernel on those
rays in slot order, shuffled, and sorted by direction octant / origin Morton code."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

FLT_MAX = 3.4028234663852886e38


def morton3(q):
    def spread(v):
        v = v.astype(np.uint64) & 0x3FF
        v = (v | (v << 16)) & 0x30000FF
        v = (v | (v << 8)) & 0x300F00F
        v = (v | (v << 4)) & 0x30C30C3
        v = (v | (v << 2)) & 0x9249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


scene = crt.Scene.staircase(1.0, 64, 5)
nx, ny = 1200, 800
n = nx * ny
with crt.Frame(scene, nx, ny, 64) as fr:
    L = crt.device_lib()
    fr.run(2, copy=False)
    ro, rd = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    L.rendererDebugRead(b"rayO", ro.ctypes.data, ro.nbytes)
    L.rendererDebugRead(b"rayD", rd.ctypes.data, rd.nbytes)
    ro[:, 3] = 0.01
    rd[:, 3] = FLT_MAX
    ok = np.isfinite(ro[:, :3]).all(axis=1) & np.isfinite(rd[:, :3]).all(axis=1) & (np.abs(rd[:, :3]).sum(axis=1) > 0)
    ro, rd = ro[ok], rd[ok]
    n = len(ro)
    reps = 4  # replicate to 4x for a longer launch
    dO, dD = L.rendererDeviceAlloc(16 * n * reps), L.rendererDeviceAlloc(16 * n * reps)
    dH, dM = L.rendererDeviceAlloc(16 * n * reps), L.rendererDeviceAlloc(4 * n * reps)
    d = rd[:, :3] / np.linalg.norm(rd[:, :3], axis=1, keepdims=True)
    octant = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2)
    lo, hi = ro[:, :3].min(axis=0), ro[:, :3].max(axis=0)
    q = ((ro[:, :3] - lo) / np.maximum(hi - lo, 1e-9) * 1023).astype(np.uint32)
    mort = morton3(q)
    qd = ((d * 0.5 + 0.5) * 7.999).astype(np.uint64)  # 8x8x8 direction cells
    dircell = qd[:, 0] | (qd[:, 1] << 3) | (qd[:, 2] << 6)
    orders = {
        "slot_order": np.arange(n),
        "shuffled": np.random.default_rng(1).permutation(n),
        "by_octant": np.argsort(octant, kind="stable"),
        "by_octant_then_morton": np.lexsort((mort, octant)),
        "by_morton": np.argsort(mort, kind="stable"),
        "by_morton_coarse_then_dircell": np.lexsort((dircell, mort >> np.uint64(15))),
        "by_dircell_then_morton": np.lexsort((mort, dircell)),
    }
    res = {"rays": int(n)}
    for name, perm in orders.items():
        a = np.ascontiguousarray(np.tile(ro[perm], (reps, 1)))
        b = np.ascontiguousarray(np.tile(rd[perm], (reps, 1)))
        # tiling repeats the same rays back to back: interleave the copies so that equal rays are far apart
        L.rendererCopyToDevice(dO, a.ctypes.data, a.nbytes)
        L.rendererCopyToDevice(dD, b.ctypes.data, b.nbytes)
        ms = [L.intersectBatchDevice(dO, dD, n * reps, dH, dM) for _ in range(4)]
        res[name] = dict(ms=min(ms), grays=n * reps / min(ms) / 1e6)
    print(json.dumps(res, indent=1))
