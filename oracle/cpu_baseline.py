"""TEST / BENCH INFRASTRUCTURE — times the reference's CPU path of the hot path on the host cores.

Run as a subprocess (`python -m oracle.cpu_baseline ...`) by bench.py's cpu_baseline leg and by `bench.py --impl reference`;
prints one JSON object and leaves with os._exit (the reference's static teardown is broken on Linux).

  cull      : the reference's own CullingSystemImpl::cull on its own job_system (oracle/_ref, kind "reference"),
              W = min(host cores, 64) workers (studio's cap, studio_app.cpp:583); falls back to the serial C restatement
              (kind "port") if oracle/_ref is absent.
  propagate : serial DFS restatement of World::transformEntity (the reference is serial) — kind "port".
  pose/skin : restatement of updateAnimable / evaluateSkin, serial — kind "port".
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bench_cull(n, steps, warmup, workers, scene_name):
    from lumixengine_b200 import scenes  # numpy scene generators only (inputs, not the product's compute path)
    from oracle import pyoracle as po
    po.build()
    scene = scenes.c2_scene(n) if scene_name == "c2" else scenes.c1_scene(n)
    fa = scenes.c2_frustum_args() if scene_name == "c2" else scenes.c1_frustum_args()
    f = po.frustum_perspective(fa["position"], fa["direction"], fa["up"], fa["fov"], fa["ratio"], fa["near"], fa["far"])
    out = {}
    if po.ref_available():
        cores = _cores()
        w = max(1, min(cores, 64)) if workers <= 0 else workers
        rc = po.RefCulling(workers=w)
        t0 = time.time()
        rc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
        build_s = time.time() - t0
        ids, tys, _ = rc.cull(f, cap=n)  # one cull with the ids kept: the digest bench.py compares with the GPU's
        ids = ids.astype(np.uint64)
        out["digest"] = [[int((tys == t).sum()), int(ids[tys == t].sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(ids[tys == t])) if (tys == t).any() else 0] for t in range(4)]
        if warmup:
            rc.cull(f, cap=0, iters=warmup)
        times = []
        count = 0
        for _ in range(steps):
            _, _, info = rc.cull(f, cap=0, iters=1)
            times.append(info["best_s"])
            count = info["count"]
        out.update(kind="reference", cores=rc.workers, impl="CullingSystemImpl::cull on jobs:: (oracle/_ref overlay build)")
    else:
        oc = po.OracleCulling()
        t0 = time.time()
        oc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
        build_s = time.time() - t0
        times = []
        count = 0
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _, _, st = oc.cull(f, want_ids=False)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            count = st["visible"]
        out.update(kind="port", cores=1, impl="serial C restatement (oracle/oracle_cull.c)")
    times = np.array(times)
    out.update(n=n, visible=int(count), build_s=build_s, total_s=float(times.sum()), median_s=float(np.median(times)), best_s=float(times.min()),
               steps=steps, value=float(n / np.median(times) / 1e6), unit="M entities culled/s",
               sample=f"{steps} x cull() of the full {n}-entity {scene_name.upper()} scene, result freed each call")
    return out


def bench_sortkeys(n, steps, seed_offset=40):
    """The stage behind the cull on the host: PipelineImpl::createSortKeys on the visible list of the C2 cull (C restatement, one worker:
    pipeline.cpp cannot be compiled here — kind "port") and PipelineImpl::radixSort (the reference's own function cut out of pipeline.cpp at
    build time when oracle/_ref exists, on its job system; else the restatement).  Same inputs as bench.py's e2e step."""
    from lumixengine_b200 import scenes
    from oracle import pyoracle as po
    po.build()
    scene = scenes.c2_scene(n)
    fa = scenes.c2_frustum_args()
    f = po.frustum_perspective(fa["position"], fa["direction"], fa["up"], fa["fov"], fa["ratio"], fa["near"], fa["far"])
    if po.ref_available():
        rc = po.RefCulling(workers=1)
        rc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
        ids, tys, _ = rc.cull(f, cap=n)
    else:
        oc = po.OracleCulling()
        oc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
        ids, tys, _ = oc.cull(f)
    sk = scenes.sortkey_setup(n, scene["types"], scene["pos"], seed=seed_offset)
    view = np.zeros(1, po.SK_VIEW_DTYPE)
    from lumixengine_b200 import sortkeys as skm  # host-side view packing only (numpy)
    keys_s, sort_s, nk, ni = [], [], 0, 0
    use_ref_sort = po.ref_available() and hasattr(po.ref(), "ref_radix_sort")
    for k in range(steps):
        view = skm.make_view(fa["position"], fa["position"], 1.0 / 60.0, 1.0, 100 + k, False, sk["max_sort_key"], sk["layer_to_bucket"], sk["depth_sorted_buckets"])
        t0 = time.perf_counter()
        out = po.create_sort_keys(ids, tys, sk["transforms"], sk["model_of"], sk["lod"], sk["flags"], sk["pose_frame"], sk["decal_sort_key"], sk["decal_layer"],
                                  sk["models"], sk["meshes"], view, sort=False)
        t1 = time.perf_counter()
        if use_ref_sort:
            po.ref_radix_sort(out["keys"], out["values"], workers=2)
        else:
            po.radix_sort(out["keys"], out["values"])
        t2 = time.perf_counter()
        keys_s.append(t1 - t0)
        sort_s.append(t2 - t1)
        nk, ni = len(out["keys"]), len(out["group_renderables"])
    return dict(kind="port", cores=1, n=n, visible=int(len(ids)), n_keys=nk, n_instances=ni, create_keys_median_s=float(np.median(keys_s)), sort_median_s=float(np.median(sort_s)),
                median_s=float(np.median(np.array(keys_s) + np.array(sort_s))), radix_sort="reference (pipeline.cpp:4100-4144 cut out at build time)" if use_ref_sort else "port",
                sample=f"{steps} x (createSortKeys over the {len(ids)} visible renderables of the C2 cull, one worker, + radixSort of the keys); includes the Python-side output allocation")


def bench_propagate(n, steps):
    from lumixengine_b200 import scenes
    from oracle import pyoracle as po
    po.build()
    parents, locals_, roots = scenes.hierarchy_forest(n, 8, 7, seed=3)
    lb = np.ascontiguousarray(locals_).view(np.uint8).reshape(len(parents), 56)
    gb = np.ascontiguousarray(roots).view(np.uint8).reshape(len(parents), 56)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        po.propagate(parents, lb, gb)
        times.append(time.perf_counter() - t0)
    m = float(np.median(times))
    return dict(kind="port", cores=1, n=len(parents), median_s=m, value=len(parents) / m / 1e6, unit="M nodes/s",
                sample=f"{steps} x serial DFS over {len(parents)} nodes (depth 8), includes the 56 MB globals copy")


def bench_c3chain(n, steps):
    """BASELINE configs[2] on the host: serial DFS propagate (port; the reference is serial), RenderModule::onModelInstanceMoved's sphere refresh,
    the reference's own CullingSystemImpl::set for every node and its cull (oracle/_ref when present, else the C restatement)."""
    from lumixengine_b200 import scenes
    from oracle import pyoracle as po
    po.build()
    parents, locals_, roots = scenes.hierarchy_forest(n, 8, 7, seed=3)
    lb_ = np.ascontiguousarray(locals_).view(np.uint8).reshape(len(parents), 56)
    root_sets = [roots, roots.copy()]
    root_sets[1]["pos"] += np.array([37.0, 4.0, -29.0])
    fa = scenes.c2_frustum_args()
    f = po.frustum_perspective(fa["position"], fa["direction"], fa["up"], fa["fov"], fa["ratio"], fa["near"], fa["far"])
    use_ref = po.ref_available()
    cs = po.RefCulling(workers=1) if use_ref else po.OracleCulling()
    ents = np.arange(len(parents), dtype=np.int32)

    def globals_of(k):
        gb = np.ascontiguousarray(root_sets[k & 1]).view(np.uint8).reshape(len(parents), 56)
        return po.propagate(parents, lb_, gb)
    g = globals_of(0).view(scenes.TRANSFORM_DTYPE).reshape(-1)
    cs.add(ents, np.zeros(len(parents), np.uint8), np.ascontiguousarray(g["pos"]), np.max(g["scale"], axis=1).astype(np.float32))
    times, parts = [], []
    for k in range(1, steps + 1):
        t0 = time.perf_counter()
        g = globals_of(k).view(scenes.TRANSFORM_DTYPE).reshape(-1)
        t1 = time.perf_counter()
        pos, rad = np.ascontiguousarray(g["pos"]), np.max(g["scale"], axis=1).astype(np.float32)
        t2 = time.perf_counter()
        cs.set(ents, pos, rad)
        t3 = time.perf_counter()
        if use_ref:
            cs.cull(f, cap=0, iters=1)
        else:
            cs.cull(f, want_ids=False)
        t4 = time.perf_counter()
        times.append(t4 - t0)
        parts.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3])
    m = float(np.median(times))
    return dict(kind="reference" if use_ref else "port", cores=1, n=len(parents), median_s=m, value=len(parents) / m / 1e6, unit="M nodes/s",
                parts_ms=dict(zip(("propagate", "sphere_refresh", "culling_set", "cull"), [float(x) for x in np.median(np.array(parts), axis=0)])),
                sample=f"{steps} x (serial DFS over {len(parents)} nodes + sphere refresh + CullingSystem::set x {len(parents)} + cull), one host core")


def bench_anim(n_inst, n_verts, skin_instances):
    from lumixengine_b200 import scenes
    from oracle import pyoracle as po
    po.build()
    sk = scenes.skeleton(64)
    clips = [scenes.clip(sk, frames=60, seed=s) for s in (1, 2, 3, 4)]
    mesh = scenes.mesh(sk, n_verts)
    ci, tt = scenes.instance_times(n_inst, clips)
    t0 = time.perf_counter()
    out = po.animate_instances(sk, clips, ci, tt, want=("dq", "mtx"))
    pose_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(skin_instances):
        po.skin_vertices(out["mtx"][i], mesh.positions, mesh.weights, mesh.indices)
    skin_s = time.perf_counter() - t0
    return dict(kind="port", cores=1,
                pose=dict(value=n_inst * 64 / pose_s / 1e6, unit="M bone-instances/s", sample=f"{n_inst} instances x 64 bones, DQ + matrix palettes, serial"),
                skin=dict(value=skin_instances * n_verts / skin_s / 1e6, unit="M skinned verts/s", sample=f"{skin_instances} instances x {n_verts} verts, serial"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cull")
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--scene", default="c2")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--workers", type=int, default=0)
    a = ap.parse_args()
    if a.workload == "cull":
        res = bench_cull(a.n, a.steps, a.warmup, a.workers, a.scene)
    elif a.workload == "sortkeys":
        res = bench_sortkeys(a.n, a.steps)
    elif a.workload == "propagate":
        res = bench_propagate(a.n, a.steps)
    elif a.workload == "c3chain":
        res = bench_c3chain(a.n, a.steps)
    elif a.workload == "anim":
        res = bench_anim(a.n, 5000, 200)
    else:
        raise SystemExit("unknown workload")
    sys.stdout.write("CPU_BASELINE_JSON " + json.dumps(res) + "\n")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
