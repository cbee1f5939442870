"""Profiling driver (run under ncu on the GPU box): the C2 cull (10 M entities, 8 scene replicas), single culls back to back on one stream."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lumixengine_b200 as lb
from lumixengine_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx = lb.Context(0)
scene = scenes.c2_scene(10_000_000)
cs = lb.CullingSystem(ctx)
cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
cs.flush()
f = lb.frustum_perspective(**scenes.c2_frustum_args())
for _ in range(n):
    cs.cull_device(f, want_counts=False)
ctx.synchronize()
cs.close()
ctx.close()
