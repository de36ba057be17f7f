"""TEST/DESIGN INFRASTRUCTURE: dumps the per-ray traversal step log of a small CPU-oracle render (see sched_sim.cpp).
usage: OMP_NUM_THREADS=1 python oracle/steplog_dump.py out.bin [nx ny ns detail]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

out = sys.argv[1]
nx, ny, ns = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (240, 160, 2)
detail = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
L = C.CDLL(os.path.join(ROOT, "oracle", "build", "liboracle_steplog.so"))
L.oracleRender.argtypes = [C.POINTER(crt.KernelScene), C.POINTER(crt.Camera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint,
                           C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong)]
L.oracleStepLogFetch.restype = C.c_ulonglong
L.oracleStepLogFetch.argtypes = [C.c_void_p, C.c_ulonglong]
scene = crt.Scene.staircase(detail, 64, 5)
cam = crt.staircase_camera(nx, ny)
fb = np.zeros((ny, nx, 3), dtype=np.float32)
L.oracleStepLogStart()
L.oracleRender(C.byref(scene.ks), C.byref(cam), nx, ny, ns, 64, 0, 0, ny, 1, fb.ctypes.data, None)
n = L.oracleStepLogFetch(None, 0)
buf = np.zeros(n, dtype=np.uint8)
L.oracleStepLogFetch(buf.ctypes.data, n)
buf.tofile(out)
print(n, "bytes")
