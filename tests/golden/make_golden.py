#!/usr/bin/env python
"""Mints the golden vectors in this directory from the REFERENCE's own CUDA code (oracle/_ref, built from
/root/reference by oracle/Makefile).  Needs a GPU: run on the B200 box as

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

and copy gpurun_out/golden/* here.  The reference ships no fixtures of its own (its *.ref frames are git-ignored,
reference .gitignore:11), so these files are what pins the CPU oracle and the CUDA path:
  staircase_96x64_8spp.ref     REF_00.01 frame from libref.so (the unmodified reference kernel), depth 64
  staircase_96x64_8spp_d3.ref  same, max depth 3 (no Russian roulette reached)
  rays_8192.npz                ray batch + the reference's hitMesh() results (closest hit and any-hit) via ref_shim.cu
  rtiow_96x64_8spp.ref         reference-DERIVED sphere megakernel (oracle/ref_spheres.cu), depth 50
  golden.json                  scene hash and parameters
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

crt = oracle.crt
DETAIL, TEX, PPL, NX, NY, NS = 0.1, 32, 5, 96, 64, 8


def make_rays(scene, n, seed=1234):
    rng = np.random.default_rng(seed)
    bmin, bmax = scene.bounds()
    cam = crt.staircase_camera(NX, NY)
    o = np.zeros((n, 4), np.float32)
    d = np.zeros((n, 4), np.float32)
    h = n // 2
    uv = rng.random((h, 2), dtype=np.float32)
    llc, hor, ver, org = (np.array(list(v.e), np.float32) for v in (cam.lower_left_corner, cam.horizontal, cam.vertical, cam.origin))
    o[:h, :3] = org
    d[:h, :3] = llc + uv[:, :1] * hor + uv[:, 1:] * ver - org
    o[h:, :3] = (bmin + rng.random((n - h, 3)) * (bmax - bmin)).astype(np.float32)
    v = rng.normal(size=(n - h, 3)).astype(np.float32)
    d[h:, :3] = v / np.linalg.norm(v, axis=1, keepdims=True)
    o[:, 3] = 0.01
    d[:, 3] = np.float32(3.4028234663852886e38)
    return o, d


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    tmp = tempfile.mkdtemp()
    scene = crt.Scene.staircase(DETAIL, TEX, PPL)
    meta = dict(detail=DETAIL, tex=TEX, ppl=PPL, nx=NX, ny=NY, ns=NS, scene_hash=f"{scene.hash():016x}",
                slots=scene.num_slots, nodes=scene.num_nodes, real_triangles=scene.num_real_triangles)
    oracle.ref_render(DETAIL, TEX, PPL, NX, NY, NS, 64, os.path.join(out, "staircase_96x64_8spp.ref"))
    oracle.ref_render(DETAIL, TEX, PPL, NX, NY, NS, 3, os.path.join(out, "staircase_96x64_8spp_d3.ref"))
    oracle.ref_spheres(1, NX, NY, NS, 50, os.path.join(out, "rtiow_96x64_8spp.ref"))
    ro, rd = make_rays(scene, 8192)
    hit, mesh, _ = oracle.ref_intersect_batch(DETAIL, TEX, PPL, ro, rd, False, tmp)
    # any-hit with a finite tMax (shadow rays stop at the light, kernels.cu:500)
    rd_sh = rd.copy()
    rd_sh[:, 3] = 400.0
    occ, _, _ = oracle.ref_intersect_batch(DETAIL, TEX, PPL, ro, rd_sh, True, tmp)
    np.savez_compressed(os.path.join(out, "rays_8192.npz"), ray_o=ro, ray_d=rd, hit=hit, mesh=mesh, shadow_tmax=np.float32(400.0),
                        occluded=(occ[:, 0] == 0.0))
    with open(os.path.join(out, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
