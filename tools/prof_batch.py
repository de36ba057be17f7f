#!/usr/bin/env python
"""One scene, one ray batch, a few launches of the ray-batch kernel: the command line for ncu captures of the traversal kernels.
    python tools/prof_batch.py [detail] [log2 rays] [traversal: 0 wide, 1 exact, 2 uncertified] [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

detail = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L = crt.device_lib()
scene = crt.Scene.staircase(detail, 32, 5)
crt.set_traversal(mode)
with crt.Frame(scene, 64, 64, 1):
    dO, dD, dH, dM = (L.rendererDeviceAlloc(16 * n) for _ in range(4))
    L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, 3.4028234663852886e38)
    ms = [L.intersectBatchDevice(dO, dD, n, dH, dM) for _ in range(launches)]
    print("ms", ms, "Mrays/s", n / min(ms) / 1e3, "redo", crt.wide_info().lastBatchRedo)
