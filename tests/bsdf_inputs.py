"""Surface points for the BSDF probe (scatterBatch, include/kernels.h): 12 floats per item
{normal.xyz, t, p.xyz, inside, wo.xyz, rng state bits}. Shared by the golden-vector script and the tests."""
import numpy as np


def make_items(n, seed):
    g = np.random.default_rng(seed)
    nrm = g.normal(size=(n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    wo = g.normal(size=(n, 3)).astype(np.float32)
    wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    flip = (wo * nrm).sum(axis=1) > 0          # the normal always faces the ray (kernels.cu:354-355)
    wo[flip] = -wo[flip]
    wo[: n // 8] *= g.uniform(0.5, 2.0, size=(n // 8, 1)).astype(np.float32)  # some un-normalised directions (subsurface passes wo on)
    graze = slice(n // 8, n // 4)                # grazing incidence: total internal reflection branch
    t = nrm[graze] * 0.05 + np.cross(nrm[graze], g.normal(size=nrm[graze].shape)).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    wo[graze] = np.where(((t * nrm[graze]).sum(axis=1) > 0)[:, None], -t, t).astype(np.float32)
    items = np.zeros((n, 12), np.float32)
    items[:, 0:3] = nrm
    items[:, 3] = g.uniform(0.011, 60.0, size=n)
    items[:, 4:7] = g.uniform(-120.0, 120.0, size=(n, 3))
    items[:, 7] = (g.uniform(size=n) < 0.5).astype(np.float32)
    items[:, 8:11] = wo
    items[:, 11] = (g.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32) | np.uint32(1)).view(np.float32)
    return items
