#!/usr/bin/env python
"""Ray-batch A/B of the two traversals on the GPU: the renderer's own wide tree (+ certificate + exact re-trace) against
the order-exact walk of the caller's tree. Prints one JSON line: equality of the answers, time of each, visits per ray,
rays handed to the exact kernel, host build time.

    python tools/wide_check.py [detail] [log2 rays] [anyhit]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

FLT_MAX = 3.4028234663852886e38


def run(scene, mode, n, any_hit, count, rays=None):
    L = crt.device_lib()
    crt.set_traversal(mode)
    L.setRendererCounting(1 if count else 0)
    out = {}
    with crt.Frame(scene, 64, 64, 1):
        dO, dD, dH, dM = (L.rendererDeviceAlloc(16 * n) for _ in range(4))
        L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, 300.0 if any_hit else FLT_MAX)
        ms = [L.intersectBatchDeviceEx(dO, dD, n, dH, dM, any_hit) for _ in range(4 if not count else 1)]
        hit = np.zeros((n, 4), np.float32)
        mesh = np.zeros(n, np.int32)
        L.rendererCopyToHost(hit.ctypes.data, dH, 16 * n)
        L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
        for p in (dO, dD, dH, dM):
            L.rendererDeviceFree(p)
        w = crt.wide_info()
        nv, tt = crt.C.c_ulonglong(), crt.C.c_ulonglong()
        L.getRendererTraversalCounts(crt.C.byref(nv), crt.C.byref(tt))
        out = dict(ms=min(ms), mrays=n / min(ms) / 1e3, active=w.active, nodes=w.numNodes, depth=w.depth, build_ms=w.buildMs,
                   threads=w.buildThreads, redo=int(w.lastBatchRedo), visits_per_ray=nv.value / n if count else None,
                   tests_per_ray=tt.value / n if count else None)
    L.setRendererCounting(0)
    crt.set_traversal(-1)
    return hit, mesh, out


def main():
    detail = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
    n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
    any_hit = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    scene = crt.Scene.staircase(detail, 32, 5)
    he, me, oe = run(scene, crt.TRAVERSAL_EXACT, n, any_hit, False)
    hw, mw, ow = run(scene, crt.TRAVERSAL_WIDE, n, any_hit, False)
    hu, mu, ou = run(scene, crt.TRAVERSAL_WIDE_UNCERTIFIED, n, any_hit, False)
    _, _, ce = run(scene, crt.TRAVERSAL_EXACT, n, any_hit, True)
    _, _, cw = run(scene, crt.TRAVERSAL_WIDE, n, any_hit, True)
    same = (he.view(np.uint32) == hw.view(np.uint32)).all(axis=1) & (me == mw)
    same_u = (he.view(np.uint32) == hu.view(np.uint32)).all(axis=1) & (me == mu)
    print(json.dumps(dict(detail=detail, rays=n, any_hit=any_hit, slots=scene.num_slots, equal=float(same.mean()), differing=int((~same).sum()),
                          uncertified_differing=int((~same_u).sum()), hit_fraction=float((me >= 0).mean()) if not any_hit else float((he[:, 0] == 0).mean()),
                          exact=oe, wide=ow, uncertified=ou, exact_counts=dict(visits=ce["visits_per_ray"], tests=ce["tests_per_ray"]),
                          wide_counts=dict(visits=cw["visits_per_ray"], tests=cw["tests_per_ray"]))))
    bad = np.nonzero(~same)[0][:5]
    for b in bad:
        print("DIFF", int(b), he[b].tolist(), hw[b].tolist(), he.view(np.uint32)[b, 3], hw.view(np.uint32)[b, 3], file=sys.stderr)


if __name__ == "__main__":
    main()
