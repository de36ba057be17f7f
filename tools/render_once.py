#!/usr/bin/env python
"""GPU-box diagnostic: renders the benchmark frame with libcrt_b200.so only (no reference run) and prints device time,
ray counts and iteration counts. Environment knobs of the library (CRT_DUMP_LANES, CRT_CHASER, ...) apply."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--detail", type=float, default=1.0)
    ap.add_argument("--tex", type=int, default=1024)
    ap.add_argument("--nx", type=int, default=1200)
    ap.add_argument("--ny", type=int, default=800)
    ap.add_argument("--ns", type=int, default=100)
    ap.add_argument("--depth", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--budget", type=int, default=0)
    ap.add_argument("--min-active", type=int, default=0)
    ap.add_argument("--sah", action="store_true", help="BVH split by the surface-area heuristic (same layout, same hits)")
    ap.add_argument("--count", action="store_true", help="one extra frame with the traversal counters on: node visits / triangle tests per ray")
    ap.add_argument("--profile", action="store_true", help="per-kernel-family CUDA events (serialised iterations)")
    args = ap.parse_args()
    crt.set_options(slots_per_pixel=args.slots, trace_budget=args.budget, trace_min_active=args.min_active)
    scene = crt.Scene.staircase(args.detail, args.tex, 5, sah=args.sah)
    with crt.Frame(scene, args.nx, args.ny, args.depth) as fr:
        fr.run(args.ns, copy=False)
        if args.profile:
            crt.device_lib().setRendererProfiling(1)
        for _ in range(args.steps):
            fr.run(args.ns, copy=False)
            st = crt.stats()
            rays = st.raysExtend + st.raysShadow
            extra = {}
            if args.count and _ == args.steps - 1:
                L = crt.device_lib()
                L.setRendererCounting(1)
                fr.run(args.ns, copy=False)
                L.setRendererCounting(0)
                nv, tt = crt.C.c_ulonglong(), crt.C.c_ulonglong()
                L.getRendererTraversalCounts(crt.C.byref(nv), crt.C.byref(tt))
                w = crt.wide_info()
                extra = dict(visits_per_ray=nv.value / rays, tests_per_ray=tt.value / rays, wide_nodes=w.numNodes, depth=w.depth, redo=int(w.lastFrameRedo), build_ms=w.buildMs)
            print(json.dumps(dict(extra, ms=st.msTotal, mrays=rays / (st.msTotal * 1e3), rays=rays, iterations=st.iterations, launches=st.kernelLaunches,
                                  resumes=st.resumes, deferred=st.deferred, ms_trace=st.msTrace, ms_shade=st.msShade)))


if __name__ == "__main__":
    main()
