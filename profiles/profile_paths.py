"""Profiling driver (run under ncu on the GPU box): exercises the secondary paths a few times each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lumixengine_b200 as lb
from lumixengine_b200 import scenes

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = lb.Context(0)
if which in ("all", "propagate"):
    parents, locals_, roots = scenes.hierarchy_forest(1_000_000, 8, 7, seed=3)
    h = lb.Hierarchy(ctx, parents)
    h.setLocalTransforms(locals_)
    h.setRootTransforms(roots)
    for _ in range(3):
        h.propagate()
    ctx.synchronize()
    h.close()
if which in ("all", "anim"):
    sk = scenes.skeleton(64)
    clips = [scenes.clip(sk, frames=60, seed=s) for s in (1, 2, 3, 4)]
    mesh = scenes.mesh(sk, 5000)
    n = 100_000 if which == "anim" else 20_000
    anim = lb.AnimationSystem(ctx, sk, clips, mesh, max_instances=n)
    ci, tt = scenes.instance_times(n, clips)
    anim.setInstances(ci, tt)
    for _ in range(3):
        anim.update(1.0 / 60.0, lb.PALETTE_DUAL_QUAT)
    anim.update(0.0, lb.PALETTE_MATRIX)
    for _ in range(2):
        anim.skin()
    ctx.synchronize()
    anim.close()
ctx.close()
