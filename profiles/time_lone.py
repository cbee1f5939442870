"""Device latency of one lone C2 cull (LB200_LONE_MODE / LB200_DEBUG_SKIP_WORK select what sits between the events; see culling.cu)."""
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
    cs.time_lone_cull(f, 10)
    t = np.sort(cs.time_lone_cull(f, 60)) * 1e3
    print(f"LONE mode={os.environ.get('LB200_LONE_MODE', '0')} skip_work={os.environ.get('LB200_DEBUG_SKIP_WORK', '0')} {name:11s} median {t[len(t)//2]:6.2f} us  min {t[0]:6.2f}  p90 {t[int(len(t)*0.9)]:6.2f}")
cs.close(); ctx.close()
