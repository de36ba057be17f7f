#!/usr/bin/env python
"""GPU-box diagnostic: the same frame rendered by the wavefront alone, by the chaser alone (every slot handed over at the first
partition) and by the mix; bitwise comparison."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
import crt_b200 as crt  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nx, ny = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1200, 800)
scene = crt.Scene.staircase(1.0, 1024, 5)


def render(env):
    for k in ("CRT_CHASER", "CRT_CHASE_MOVE_ALL", "CRT_CHASE_CAPACITY"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with crt.Frame(scene, nx, ny, 64) as fr:
        img = fr.run(ns)
        st = crt.stats()
    return img, st.raysExtend + st.raysShadow, st.msTotal


a, ra, ta = render({"CRT_CHASER": "0"})
b, rb, tb = render({"CRT_CHASE_MOVE_ALL": "100000000", "CRT_CHASE_CAPACITY": "100000000"})
c1, rc1, tc1 = render({})
c2, rc2, tc2 = render({})


def diff(x, y):
    d = np.any(x != y, axis=2)
    idx = np.argwhere(d)
    return dict(pixels=int(d.sum()), first=[(int(j), int(i), x[j, i].tolist(), y[j, i].tolist()) for j, i in idx[:5]])


print(json.dumps(dict(rays=dict(wavefront=ra, chaser=rb, mix1=rc1, mix2=rc2), ms=dict(wavefront=ta, chaser=tb, mix1=tc1, mix2=tc2),
                      wavefront_vs_chaser=diff(a, b), wavefront_vs_mix1=diff(a, c1), mix1_vs_mix2=diff(c1, c2)), indent=1))
