import sys
sys.path.insert(0, "cuda-raytracing-optimized_b200/python")
import crt_b200 as crt
scene = crt.Scene.staircase(1.0, 1024, 5)
for mb in (16, 8, 24, 32, 16):
    crt.set_options(mega_batch=mb)
    with crt.Frame(scene, 1200, 800, 64) as fr:
        fr.run(100, copy=False)
        ms = []
        for _ in range(2):
            fr.run(100, copy=False); ms.append(crt.stats().msTotal)
    print("megaBatch", mb, ["%.1f" % m for m in ms])
