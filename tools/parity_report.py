#!/usr/bin/env python
"""GPU-box diagnostic (not a test, not the bench): renders the same frames with the reference's CUDA kernel
(oracle/_ref/ref_driver, own process) and with libcrt_b200.so, intersects the same ray batch with both, and prints
how far apart they are and how long each took.  Writes gpurun_out/parity_report.json."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

crt = oracle.crt


def image_stats(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    mse = float((d ** 2).mean())
    peak = max(float(b.max()), 1e-9)
    return dict(max_abs=float(d.max()), mean_abs=float(d.mean()), exact_pixels=float((d.max(axis=2) == 0).mean()),
                within_1e3=float((d.max(axis=2) <= 1e-3).mean()), rmse=mse ** 0.5,
                psnr=float(10 * np.log10(peak * peak / mse)) if mse > 0 else float("inf"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--detail", type=float, default=0.25)
    ap.add_argument("--tex", type=int, default=128)
    ap.add_argument("--nx", type=int, default=300)
    ap.add_argument("--ny", type=int, default=200)
    ap.add_argument("--ns", type=int, default=16)
    ap.add_argument("--depth", type=int, default=64)
    ap.add_argument("--rays", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--spheres", action="store_true")
    args = ap.parse_args()
    report = {"args": vars(args)}
    tmp = tempfile.mkdtemp()

    # ---- image parity, mesh scene
    ref_img, ref_info = oracle.ref_render(args.detail, args.tex, 5, args.nx, args.ny, args.ns, args.depth, os.path.join(tmp, "ref.ref"),
                                          warmup=1, steps=args.steps)
    scene = crt.Scene.staircase(args.detail, args.tex, 5)
    with crt.Frame(scene, args.nx, args.ny, args.depth) as fr:
        fr.run(args.ns)
        ms = []
        for _ in range(args.steps):
            img = fr.run(args.ns)
            ms.append(crt.stats().msTotal)
        st = crt.stats()
        rays = st.raysExtend + st.raysShadow
        report["mesh"] = dict(ref_ms=ref_info["ms"], ours_ms=ms, rays=rays, extend=st.raysExtend, shadow=st.raysShadow,
                              iterations=st.iterations, launches=st.kernelLaunches, ours_mrays=rays / (min(ms) * 1e3),
                              ref_mrays=rays / (min(ref_info["ms"]) * 1e3), **image_stats(img, ref_img))
        print("MESH", json.dumps(report["mesh"]))
        if args.cpu:
            t0 = time.time()
            cpu_img, cnt = oracle.render(scene, args.nx, args.ny, args.ns, args.depth, count=True)
            report["mesh_cpu"] = dict(seconds=time.time() - t0, counters=cnt, vs_ref=image_stats(cpu_img, ref_img))
            print("CPU ", json.dumps(report["mesh_cpu"]))

        # ---- ray batch parity
        n = args.rays
        L = crt.device_lib()
        dO, dD = L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n)
        dH, dM = L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(4 * n)
        L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, 3.4028234663852886e38)
        kms = [L.intersectBatchDevice(dO, dD, n, dH, dM) for _ in range(3)]
        ro, rd = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
        hit, mesh = np.zeros((n, 4), np.float32), np.zeros(n, np.int32)
        L.rendererCopyToHost(ro.ctypes.data, dO, 16 * n)
        L.rendererCopyToHost(rd.ctypes.data, dD, 16 * n)
        L.rendererCopyToHost(hit.ctypes.data, dH, 16 * n)
        L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
        for p in (dO, dD, dH, dM):
            L.rendererDeviceFree(p)
    rhit, rmesh, rinfo = oracle.ref_intersect_batch(args.detail, args.tex, 5, ro, rd, False, tmp)
    ids, rids = hit[:, 3].view(np.uint32), rhit[:, 3].view(np.uint32)
    hitmask = rids != 0xFFFFFFFF
    rel = np.abs(hit[hitmask, 0] - rhit[hitmask, 0]) / np.maximum(np.abs(rhit[hitmask, 0]), 1e-30)
    report["batch"] = dict(n=n, hits=int(hitmask.sum()), id_mismatch=int((ids != rids).sum()), mesh_mismatch=int((mesh != rmesh).sum()),
                           t_bits_mismatch=int((hit[:, 0].view(np.uint32) != rhit[:, 0].view(np.uint32)).sum()),
                           uv_bits_mismatch=int((hit[:, 1:3].view(np.uint32) != rhit[:, 1:3].view(np.uint32)).any(axis=1).sum()),
                           t_rel_max=float(rel.max()) if rel.size else 0.0, ours_kernel_ms=kms, ref_kernel_ms=rinfo["kernel_ms"],
                           ours_mrays=n / (min(kms) * 1e3), ref_mrays=n / (rinfo["kernel_ms"] * 1e3))
    print("BATCH", json.dumps(report["batch"]))

    if args.spheres:
        sref, sinfo = oracle.ref_spheres(1, args.nx, args.ny, args.ns, 50, os.path.join(tmp, "sph.ref"), warmup=1, steps=args.steps)
        with crt.Frame(crt.rtiow_scene(1), args.nx, args.ny, 50) as fr:
            fr.run(args.ns)
            ms = []
            for _ in range(args.steps):
                img = fr.run(args.ns)
                ms.append(crt.stats().msTotal)
            st = crt.stats()
        report["spheres"] = dict(ref_ms=sinfo["ms"], ours_ms=ms, rays=st.raysExtend, iterations=st.iterations, **image_stats(img, sref))
        print("SPHERES", json.dumps(report["spheres"]))

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
