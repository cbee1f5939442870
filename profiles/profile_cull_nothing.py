import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
ctx = lb.Context(0)
scene = scenes.c2_scene(10_000_000)
cs = lb.CullingSystem(ctx); cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
A = scenes.c2_frustum_args()
f = lb.frustum_perspective(**dict(A, position=(1e6, 0.0, 1e6), far=100.0)) if (len(sys.argv) < 2 or sys.argv[1] == "nothing") else lb.frustum_perspective(**A)
if len(sys.argv) > 2 and sys.argv[2] == "lanes":
    cs.cull_device_n(f, 12)  # the submission form the bench times: half-occupancy grids on the internal lanes
else:
    for _ in range(12): cs.cull_device(f, want_counts=False)
ctx.synchronize(); cs.close(); ctx.close()
