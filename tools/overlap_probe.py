#!/usr/bin/env python
"""GPU-box diagnostic: do two wavefronts on ONE device overlap usefully (one's shade launches under the other's trace launches)?
K host threads, each with its own renderer context (the library's contexts are per host thread), render spp/K samples of the same
frame concurrently on RNG streams 0..K-1; the wall time from a common barrier to the last thread's return is compared with one
context rendering all spp. (Sample-split, so only a timing probe: a pixel-split with the reference's streams would cost the same.)"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
MODE = sys.argv[2] if len(sys.argv) > 2 else "samples"  # "samples": split the samples; "rows": split the frame into K bands of rows


def band_camera(k, K):
    """The camera whose ny/K rows are band k of the full frame: same rays as those rows (v = (py + r)/(ny/K) over vertical/K)."""
    cam = crt.staircase_camera(1200, 800)
    for a in range(3):
        cam.lower_left_corner.e[a] += cam.vertical.e[a] * k / K
        cam.vertical.e[a] /= K
    return cam
scene = crt.Scene.staircase(1.0, 1024, 5)


def run(k_threads, reps=3):
    barrier = threading.Barrier(k_threads + 1)
    warm = threading.Lock()  # initRenderer synchronises the device, which is not allowed while another thread captures a graph
    ready = threading.Barrier(k_threads)
    times = []

    def worker(i):
        rows = MODE == "rows"
        crt.set_options(sample_stream=0 if rows else i)
        n = spp if rows else spp // k_threads
        warm.acquire()
        fr = crt.Frame(scene, 1200, 800 // k_threads if rows else 800, 64, cam=band_camera(i, k_threads) if rows else None)
        fr.run(n, copy=False)  # warm-up (graph capture, allocations)
        warm.release()
        ready.wait()
        with fr:
            for _ in range(reps):
                barrier.wait()
                fr.run(n, copy=False)
                barrier.wait()

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(k_threads)]
    for t in threads:
        t.start()
    for _ in range(reps):
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        times.append((time.perf_counter() - t0) * 1e3)
    for t in threads:
        t.join()
    return times


for k in (1, 2, 3):
    print(json.dumps(dict(mode=MODE, contexts=k, wall_ms=run(k))))
