#!/usr/bin/env python
"""GPU-box diagnostic: host wall time of initRenderer / runRenderer / frame read / cleanupRenderer (the e2e step of bench.py)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 100
scene = crt.Scene.staircase(1.0, 1024, 5)
L = crt.device_lib()
for it in range(4):
    t0 = time.perf_counter()
    fr = crt.Frame(scene, 1200, 800, 64)
    t1 = time.perf_counter()
    L.runRenderer(ns, 8, 8)
    t2 = time.perf_counter()
    img = fr.frame(copy=True)
    t3 = time.perf_counter()
    fr.close()
    t4 = time.perf_counter()
    print(json.dumps(dict(iter=it, init_ms=(t1 - t0) * 1e3, run_ms=(t2 - t1) * 1e3, device_ms=crt.stats().msTotal, read_ms=(t3 - t2) * 1e3,
                          cleanup_ms=(t4 - t3) * 1e3, total_ms=(t4 - t0) * 1e3)))
