"""Round-2 first step: parity + timing of the instruction-lean cull kernel (csrc/cull_kernel_lean.cuh, LB200_CULL_LEAN=1) against the
oracle and against the default kernel.  Run on a GPU box:   LB200_CULL_LEAN=1 python profiles/check_lean_kernel.py
(the knob is read once per process, so the default kernel's numbers come from a second run without it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
from oracle import pyoracle as po

ctx = lb.Context(0)
tag = "LEAN" if os.environ.get("LB200_CULL_LEAN", "0") not in ("", "0") else "DEFAULT"
bad = 0
for seed, n, box, kw in ((1, 100_000, (2000.0, 200.0, 2000.0), {}), (11, 300_000, (3000.0, 300.0, 3000.0), dict(big_fraction=0.01, type_probs=(0.6, 0.2, 0.1, 0.1))),
                         (5, 50_000, (400.0, 100.0, 400.0), dict(big_fraction=0.2))):
    scene = scenes.cull_scene(n, box, seed=seed, **kw)
    cs = lb.CullingSystem(ctx); cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    oc = po.OracleCulling(); oc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    a = scenes.c1_frustum_args()
    views = [lb.frustum_perspective(**a), lb.frustum_perspective(**dict(a, far=2500.0, position=(123.4, -20.0, 987.0), direction=(0.3, -0.1, -0.9))),
             lb.frustum_perspective(**dict(a, position=(1e6, 0.0, -2e6), far=100.0)),
             lb.frustum_ortho((0.0, 0.0, 4000.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 3000.0, 3000.0, 0.0, 8000.0)]
    for vi, f in enumerate(views):
        for type_filter in (0xFF, 1):
            res = cs.cull(f, type_filter) if type_filter != 0xFF else cs.cull(f)
            oi, ot, st = oc.cull(lb.culling.frustum_bytes(f), type_filter if type_filter != 0xFF else -1)
            same = res.total == len(oi) and np.array_equal(np.sort(res.ids.astype(np.int64) * 256 + res.types()), np.sort(oi.astype(np.int64) * 256 + ot))
            stats_ok = all(res.stats[k] == st[k] for k in ("pages_tested", "pages_inside", "pages_outside", "entities_tested"))
            cs.cull_device_n(f, 5, type_filter)
            _, last = cs.last_result()
            bits = int(np.unpackbits(cs.read_bitmask().view(np.uint8)).sum())
            ok = same and stats_ok and last.total == len(oi) and bits == len(oi)
            bad += not ok
            print(f"CHECK {tag} scene {seed} view {vi} filter {type_filter:#x}: visible {res.total:7d} {'ok' if ok else 'MISMATCH'}", flush=True)
    cs.close()
# special radii (tests/golden/cull_kat.npz: special_*): the lean kernel follows the reference for NaN radii, the default kernel is known not to
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cull_kat.npz"))
rad = g["special_radius_bits"].view(np.float32)
cs = lb.CullingSystem(ctx); cs.add(np.arange(len(rad), dtype=np.int32), np.zeros(len(rad), np.uint8), g["special_pos"], rad)
f0 = culling_frustum = lb.culling.frustum_from_bytes(g["frusta"][0]) if hasattr(lb.culling, "frustum_from_bytes") else lb.frustum_perspective(**scenes.c1_frustum_args())
res = cs.cull(f0)
same = np.array_equal(np.sort(res.ids), g["special_visible"])
print(f"CHECK {tag} special radii (NaN / inf / -0.0 / negative): {'ok' if same else 'differs from the reference (expected for the default kernel)'}")
cs.close()
print(f"CHECK {tag} mismatches: {bad}")
# timing on the 10 M scene (the five views of time_cull_variants.py)
scene = scenes.c2_scene(10_000_000)
cs = lb.CullingSystem(ctx); cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
A = scenes.c2_frustum_args()
cases = {"c2_default": lb.frustum_perspective(**A), "nothing": lb.frustum_perspective(**dict(A, position=(1e6, 0.0, 1e6), far=100.0)),
         "all_visible": lb.frustum_ortho((0.0, 0.0, 20000.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 20000.0, 20000.0, 0.0, 40000.0),
         "narrow_long": lb.frustum_perspective(**dict(A, fov=0.2, far=12000.0, position=(-6000.0, 0.0, 6000.0), direction=(1.0, 0.0, -1.0))),
         "wide_far": lb.frustum_perspective(**dict(A, far=9000.0, position=(0.0, 0.0, 6000.0)))}
for name, f in cases.items():
    cs.cull_device_n(f, 60); ctx.synchronize()
    e0, e1 = ctx.event(), ctx.event(); ctx.record(e0); cs.cull_device_n(f, 300); ctx.record(e1)
    print(f"TIME {tag} {name:12s} {ctx.elapsed_ms(e0, e1) / 300 * 1e3:7.2f} us per cull", flush=True)
cs.close(); ctx.close()
