#!/usr/bin/env python
"""GPU-box diagnostic: one RTIOW frame (BASELINE config 2 at reduced spp) through initRendererSpheres / runRenderer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 16
if len(sys.argv) > 3:  # 'defer': keep the sums on the device (no finalize into the pinned frame buffer)
    crt.set_options(defer_finalize=1)
with crt.Frame(crt.rtiow_scene(1), 1200, 800, 50) as fr:
    fr.run(ns, copy=False)
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
        fr.run(ns, copy=False)
        st = crt.stats()
        print("ms", st.msTotal, "Mrays/s", (st.raysExtend + st.raysShadow) / (st.msTotal * 1e3), "iterations", st.iterations, "launches", st.kernelLaunches)
