"""Phase timeline of one lone C2 cull from %globaltimer stamps (LB200_CULL_TRACE=1): per kernel and phase boundary, when the first / median /
last block passed it, relative to the first stamp of the classify kernel."""
import os, sys
os.environ["LB200_CULL_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
ctx = lb.Context(0)
scene = scenes.c2_scene(10_000_000)
cs = lb.CullingSystem(ctx); cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
A = scenes.c2_frustum_args()
views = {"c2_default": lb.frustum_perspective(**A), "nothing": lb.frustum_perspective(**dict(A, position=(1e6, 0.0, 1e6), far=100.0))}
names = [["start", "A1 done", "A2 done", "B done", "grid dep", "all rounds done"], []]
for name, f in views.items():
    cs.time_lone_cull(f, 5)
    for rep in range(2):
        t = cs.time_lone_cull(f, 1)[0] * 1e3
        out = np.zeros((2, 2048, 8), np.uint64)
        cs._err(cs.L.lb200_culling_read_trace(cs.h, out.ctypes.data_as(C.c_void_p)))
        grids = [(cs.ctx and (10_000_000 // 180)), 0]
        k0 = out[0]; nb0 = int((k0[:, 0] > 0).sum()); k1 = out[1]; nb1 = int((k1[:, 0] > 0).sum())
        t0 = int(k0[:nb0, 0].min())
        print(f"TRACE {name} rep {rep}: lone {t:.2f} us; classify blocks {nb0}, second kernel blocks {nb1}")
        for k, (arr, nb) in enumerate(((k0, nb0), (k1, nb1))):
            for p, label in enumerate(names[k]):
                v = (arr[:nb, p].astype(np.int64) - t0) / 1e3
                print(f"TRACE   k{k} {label:16s} first {v.min():7.2f}  median {np.median(v):7.2f}  last {v.max():7.2f} us")
cs.close(); ctx.close()
