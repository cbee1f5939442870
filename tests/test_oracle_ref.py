"""Direct cross-checks of the restatement against the reference's own compiled code (oracle/_ref/libref_lumix.so).
Skipped where that library is absent; tests/test_oracle_golden.py carries the same pins as committed vectors."""
import numpy as np
import pytest

from lumixengine_b200 import scenes


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_lumix.so not built (needs /root/reference)")
    return oracle


def test_struct_sizes(ref):
    L = ref.ref()
    sizes = {"Sphere": 16, "Frustum": 224, "ShiftedFrustum": 256, "ShiftedFrustum.origin": 224, "Transform": 56, "CullResult": 4096,
             "CullResult.entities": 16, "LocalRigidTransform": 28, "DualQuat": 32, "Matrix": 64}
    for k, v in sizes.items():
        assert L.ref_sizeof(k.encode()) == v, k


@pytest.mark.parametrize("workers", [1, 4])
def test_cull_sets_equal_reference_job_system(ref, workers):
    """The reference's CullingSystemImpl on its own fiber job system vs the restatement: identical sorted ids and types."""
    scene = scenes.cull_scene(400_000, (3000.0, 300.0, 3000.0), seed=123, big_fraction=0.005, type_probs=(0.5, 0.3, 0.2))
    rc = ref.RefCulling(workers=workers)  # the job system is process-wide: the first test's worker count wins
    oc = ref.OracleCulling()
    rc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    oc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    for a in (scenes.c1_frustum_args(), dict(scenes.c1_frustum_args(), position=(900.0, -50.0, 400.0), direction=(-0.7, 0.1, -0.7), far=3500.0)):
        f = ref.frustum_perspective(a["position"], a["direction"], a["up"], a["fov"], a["ratio"], a["near"], a["far"])
        fr = ref.ref_frustum_perspective(a["position"], a["direction"], a["up"], a["fov"], a["ratio"], a["near"], a["far"])
        assert np.array_equal(f[:248], fr[:248])
        ids, tys, st = oc.cull(f)
        rids, rtys, info = rc.cull(fr, cap=len(scene["entities"]), iters=2)
        assert info["count"] == len(ids) > 1000
        o, ro = np.argsort(ids), np.argsort(rids)
        assert np.array_equal(ids[o], rids[ro]) and np.array_equal(tys[o], rtys[ro])
        assert info["pages"] == st["pages_total"] - st["pages_filtered"]  # one result page per processed cell page (:337)
