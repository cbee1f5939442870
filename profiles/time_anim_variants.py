import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
ctx = lb.Context(0)
sk = scenes.skeleton(64); clips = [scenes.clip(sk, frames=60, seed=s) for s in (1, 2, 3, 4)]; mesh = scenes.mesh(sk, 5000)
n = 100_000
anim = lb.AnimationSystem(ctx, sk, clips, mesh, max_instances=n)
ci, tt = scenes.instance_times(n, clips); anim.setInstances(ci, tt)
def t(fn, k):
    for _ in range(3): fn()
    e0, e1 = ctx.event(), ctx.event(); ctx.synchronize(); ctx.record(e0)
    for _ in range(k): fn()
    ctx.record(e1); return ctx.elapsed_ms(e0, e1) / k
p = t(lambda: anim.update(1 / 60, lb.PALETTE_DUAL_QUAT), 50)
anim.update(0.0, lb.PALETTE_MATRIX)
s = t(anim.skin, 5)
print(f"ANIMVAR pose_lanes={os.environ.get('LB200_POSE_LANES')} skin_group={os.environ.get('LB200_SKIN_GROUP')} pose {p*1e3:.1f} us skin {s:.3f} ms")
anim.close(); ctx.close()
