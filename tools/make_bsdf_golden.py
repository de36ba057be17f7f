#!/usr/bin/env python
"""GPU-box tool: mints tests/golden/bsdf_presets.npz -- inputs and the REFERENCE's outputs (oracle/_ref/ref_shim_driver scatter,
i.e. the device functions of the reference's material.h / scene_materials.h) for the ten BSDF presets. Written to gpurun_out/."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from bsdf_inputs import make_items  # noqa: E402

n = 512
items = make_items(n, seed=7)
tmp = tempfile.mkdtemp()
out = {"items": items}
for preset in range(10):
    out["ref_%d" % preset] = oracle.ref_scatter_batch(preset, items, tmp)
path = os.path.join(ROOT, "gpurun_out", "bsdf_presets.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()})
