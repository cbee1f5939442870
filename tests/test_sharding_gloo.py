"""world_size-2 gloo test of the multi-GPU host logic: index-range sharding + the padded all-gather layout + merge.
Each rank culls its shard with the CPU oracle (the GPU kernel itself is covered by -m gpu tests); the merged result must equal
the unsharded cull."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from lumixengine_b200 import scenes, sharding
    from oracle import pyoracle as po
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = scenes.cull_scene(40_001, (2000.0, 200.0, 2000.0), seed=17, type_probs=(0.6, 0.3, 0.1))
    mine = sharding.shard_scene(scene, rank, world)
    oc = po.OracleCulling()
    oc.add(mine["entities"], mine["types"], mine["pos"], mine["radius"])
    a = scenes.c1_frustum_args()
    f = po.frustum_perspective(a["position"], a["direction"], a["up"], a["fov"], a["ratio"], a["near"], a["far"])
    ids, tys, _ = oc.cull(f)
    # pack like lb200_culling_allgather: ids grouped by type, counts[256], slab padded to the max over ranks
    order = np.argsort(tys, kind="stable")
    packed = ids[order].astype(np.int64)
    counts = np.bincount(tys, minlength=256).astype(np.int64)
    n = torch.tensor([len(packed)])
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    slab = torch.zeros(int(n.item()) + 8, dtype=torch.int64)
    slab[:len(packed)] = torch.from_numpy(packed)
    slabs = [torch.zeros_like(slab) for _ in range(world)]
    cnts = [torch.zeros(256, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(slabs, slab)
    dist.all_gather(cnts, torch.from_numpy(counts))
    merged = sharding.merge_gathered([s.numpy() for s in slabs], np.stack([c.numpy() for c in cnts]))
    if rank == 0:
        full = po.OracleCulling()
        full.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
        fi, ft, _ = full.cull(f)
        ok = True
        for t in range(3):
            ok &= np.array_equal(np.sort(merged.get(t, np.zeros(0, np.int64))), np.sort(fi[ft == t].astype(np.int64)))
        ok &= sum(len(v) for v in merged.values()) == len(fi) and len(fi) > 1000
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather_merge():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)


def test_index_range_partition():
    sys.path.insert(0, ROOT)
    from lumixengine_b200 import sharding
    for n, w in ((10, 3), (50_000_000, 8), (7, 8), (0, 2)):
        r = [sharding.index_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [e - b for b, e in r]
        assert max(sizes) - min(sizes) <= 1
