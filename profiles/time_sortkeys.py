"""Device time of the sort-key stage (createSortKeys + radixSort) behind a C2 cull, with and without the sort."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lumixengine_b200 as lb
from lumixengine_b200 import scenes, sortkeys as skm
ctx = lb.Context(0)
N = 10_000_000
scene = scenes.c2_scene(N)
cs = lb.CullingSystem(ctx)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
sk = scenes.sortkey_setup(N, scene["types"], scene["pos"], seed=40)
S = lb.SortKeys(ctx, N, sk["max_sort_key"] + 1, max_keys=1 << 22, max_instances=1 << 22)
S.setModels(sk["models"], sk["meshes"]); S.setInstances(sk["model_of"], sk["lod"], sk["flags"], sk["pose_frame"], sk["decal_sort_key"], sk["decal_layer"]); S.setTransforms(sk["transforms"])
fa = scenes.c2_frustum_args()
f = lb.frustum_perspective(**fa)
view = skm.make_view(fa["position"], fa["position"], 1.0 / 60.0, 1.0, 7, False, sk["max_sort_key"], sk["layer_to_bucket"], sk["depth_sorted_buckets"])
cs.cull_device(f, want_counts=False)
for sort in (False, True):
    for _ in range(5): S.createSortKeys(cs, view, sort=sort, want_counts=False)
    e0, e1 = ctx.event(), ctx.event(); ctx.synchronize(); ctx.record(e0)
    for _ in range(20): S.createSortKeys(cs, view, sort=sort, want_counts=False)
    ctx.record(e1); ms = ctx.elapsed_ms(e0, e1) / 20
    r = S.createSortKeys(cs, view, sort=sort)
    print(f"SORTKEYS sort={sort}: {ms*1e3:8.1f} us per view  keys {r.n_keys} instances {r.n_instances} pose {r.n_pose}")
S.close(); cs.close(); ctx.close()
