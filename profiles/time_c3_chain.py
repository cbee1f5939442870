"""Config 3 as a chain on the device: root transforms -> propagate -> sphere refresh -> re-binning -> cull; device time of each link."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
ctx = lb.Context(0)
parents, locals_, roots = scenes.hierarchy_forest(1_000_000, 8, 7, seed=3)
n = len(parents)
h = lb.Hierarchy(ctx, parents); h.setLocalTransforms(locals_); h.setRootTransforms(roots); h.propagate()
bounding = np.full(n, 1.0, np.float32)
pos0, rad0 = h.getSpheres(bounding)
cs = lb.CullingSystem(ctx); cs.add(np.arange(n, dtype=np.int32), np.zeros(n, np.uint8), pos0, rad0); cs.flush()
root_ids = np.nonzero(parents < 0)[0].astype(np.uint32)
sets = [roots[root_ids].copy(), roots[root_ids].copy()]; sets[1]["pos"] += np.array([37.0, 4.0, -29.0])
f = lb.frustum_perspective(**scenes.c2_frustum_args())
ev = [ctx.event() for _ in range(6)]
acc = np.zeros(5); changers = 0
for k in range(1, 14):
    ctx.synchronize(); ctx.record(ev[0])
    h.setSubset(root_ids, sets[k & 1], globals_=True); ctx.record(ev[1])
    h.propagate(); ctx.record(ev[2])
    d_pos, d_rad = h.refreshSpheres(bounding if k == 1 else None); ctx.record(ev[3])
    changers = cs.set_many_device(d_pos, d_rad, n); ctx.record(ev[4])
    cs.cull_device(f, want_counts=False); ctx.record(ev[5])
    ctx.synchronize()
    if k > 3: acc += np.array([ctx.elapsed_ms(ev[i], ev[i + 1]) for i in range(5)])
acc /= 10
a = acc * 1e3
print("C3CHAIN us: roots upload %.1f  propagate %.1f  sphere refresh %.1f  re-binning %.1f (%d changers)  cull %.1f  total %.1f" % (a[0], a[1], a[2], a[3], changers, a[4], a.sum()))
cs.close(); h.close(); ctx.close()
