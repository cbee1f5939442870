"""Device latency of one lone C2 cull and of the fixed costs inside it (nothing / one empty kernel between the events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
ctx = lb.Context(0)
scene = scenes.c2_scene(10_000_000)
cs = lb.CullingSystem(ctx); cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
A = scenes.c2_frustum_args()
views = {"c2_default": lb.frustum_perspective(**A), "nothing": lb.frustum_perspective(**dict(A, position=(1e6, 0.0, 1e6), far=100.0))}
for name, f in views.items():
    for mode, what in ((0, "cull"), (1, "empty interval"), (2, "empty kernel")):
        cs.time_lone_cull(f, 10, mode=mode)
        t = np.sort(cs.time_lone_cull(f, 100, mode=mode)) * 1e3
        print(f"LONE {name:11s} {what:15s} mean {t.mean():6.2f} us  median {t[len(t)//2]:6.2f}  min {t[0]:6.2f}  p90 {t[int(len(t)*0.9)]:6.2f}")
cs.close(); ctx.close()
