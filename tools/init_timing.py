import sys, os, time
sys.path.insert(0, "cuda-raytracing-optimized_b200/python")
import crt_b200 as crt
scene = crt.Scene.staircase(1.0, 1024, 5)
for k in range(3):
    t0 = time.time()
    with crt.Frame(scene, 1200, 800, 64) as fr:
        t1 = time.time()
        fr.run(1, copy=False)
        w = crt.wide_info()
    print(f"frame {k}: init {1e3*(t1-t0):.1f} ms, levels/depth {w.depth}, build {w.buildMs:.2f} ms", file=sys.stderr)
