"""Per-cull device time for frustums that exercise different mixes of page classes (10M-entity C2 scene, 8 replicas)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
ctx = lb.Context(0)
scene = scenes.c2_scene(10_000_000)
cs = lb.CullingSystem(ctx); cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
A = scenes.c2_frustum_args()
cases = {
    "c2_default": lb.frustum_perspective(**A),
    "nothing": lb.frustum_perspective(**dict(A, position=(1e6, 0.0, 1e6), far=100.0)),
    "all_visible": lb.frustum_ortho((0.0, 0.0, 20000.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 20000.0, 20000.0, 0.0, 40000.0),
    "narrow_long": lb.frustum_perspective(**dict(A, fov=0.2, far=12000.0, position=(-6000.0, 0.0, 6000.0), direction=(1.0, 0.0, -1.0))),
    "wide_far": lb.frustum_perspective(**dict(A, far=9000.0, position=(0.0, 0.0, 6000.0))),
}
for name, f in cases.items():
    for _ in range(30): cs.cull_device(f, want_counts=False)
    e0, e1 = ctx.event(), ctx.event(); ctx.synchronize(); ctx.record(e0)
    n = 300
    for _ in range(n): cs.cull_device(f, want_counts=False)
    ctx.record(e1); ms_py = ctx.elapsed_ms(e0, e1) / n
    ctx.synchronize(); ctx.record(e0); cs.cull_device_n(f, n); ctx.record(e1); ms = ctx.elapsed_ms(e0, e1) / n
    lone = cs.time_lone_cull(f, 40)  # one cull on an idle device, launches pre-queued behind a delay kernel
    lone_us = float(sorted(lone)[len(lone) // 2]) * 1e3
    _, res = cs.cull_device(f, want_counts=True)
    b = cs.last_algorithmic_bytes()
    print(f"CULLVAR {name:12s} {ms*1e3:7.2f} us (python loop {ms_py*1e3:6.2f})  visible {res.total:9d} tested_pages {res.pages_tested:6d} inside {res.pages_inside:6d} ent_tested {res.entities_tested:9d} bytes {b/1e6:7.1f} MB  {b/ms/1e6:7.0f} GB/s  lone {lone_us:6.2f} us = {b/lone_us/1e3:6.0f} GB/s")
cs.close(); ctx.close()
