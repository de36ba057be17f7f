#!/usr/bin/env python
"""GPU-box tool: counts the rays (calls of hit()) of the benchmark frame at the sample counts bench.py's reference arm uses
(spp x world size) and writes them into profiles/raycounts.json. Our frame is bit-identical to the reference kernel's under
the same seeding (tools/parity_report.py), so the reference traces exactly these rays."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

path = os.path.join(ROOT, "profiles", "raycounts.json")
out = os.path.join(ROOT, "gpurun_out", "raycounts.json")
counts = json.load(open(path))
scene = crt.Scene.staircase(1.0, 1024, 5)
for ns in [int(a) for a in sys.argv[1:]] or [100, 200, 400, 800]:
    with crt.Frame(scene, 1200, 800, 64) as fr:
        fr.run(ns, copy=False)
        st = crt.stats()
    counts["staircase:1.000:1024:1200x800x%d:d64" % ns] = int(st.raysExtend + st.raysShadow)
    print(ns, st.raysExtend + st.raysShadow, st.msTotal)
json.dump(counts, open(out, "w"), indent=1)
