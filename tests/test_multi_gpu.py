"""Multi-GPU path on real devices (needs >= 2 GPUs: `gpurun --gpus 2`): each rank culls its index-range shard on its own GPU,
lb200_culling_allgather exchanges the compacted visible lists over NCCL, and the merged result equals the oracle's unsharded cull."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


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

    import lumixengine_b200 as lb
    from lumixengine_b200 import scenes, sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # rendezvous only; the data path is our own NCCL communicator
    ctx = lb.Context(rank)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid.copy_(torch.from_numpy(ctx.comm_unique_id()))
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, uid.numpy())
    scene = scenes.cull_scene(300_001, (3000.0, 300.0, 3000.0), seed=23, big_fraction=0.002, type_probs=(0.6, 0.3, 0.1))
    mine = sharding.shard_scene(scene, rank, world)
    cs = lb.CullingSystem(ctx)
    cs.add(mine["entities"], mine["types"], mine["pos"], mine["radius"])
    f = lb.frustum_perspective(**dict(scenes.c1_frustum_args(), far=2500.0))
    _, res = cs.cull_device(f, want_counts=True)
    n = torch.tensor([int(res.total)])
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    slab = int(n.item()) + 64
    dev_ptr, counts = cs.allgather(slab, world)
    slabs, counts2 = cs.read_gathered(dev_ptr, slab, world, stride=256 + slab)
    assert np.array_equal(counts, counts2)
    merged = sharding.merge_gathered(slabs, counts)
    # the asynchronous per-frame form gives the same buffer: first over NCCL, then over the NVLink peer path (several epochs)
    def check_async():
        dev2 = cs.cull_gather(f, slab)
        slabs_b, counts_b = cs.read_gathered(dev2, slab, world)
        assert np.array_equal(counts_b, counts)
        for r in range(world):
            for t in range(3):
                o = int(counts[r, :t].sum())
                assert np.array_equal(np.sort(slabs_b[r][o:o + counts[r, t]]), np.sort(slabs[r][o:o + counts[r, t]]))
    check_async()
    ctx.comm_enable_p2p(slab)
    for _ in range(5):
        check_async()
    for _ in range(50):  # back-to-back epochs without host synchronisation in between
        cs.cull_gather(f, slab)
    check_async()
    ok = True
    if rank == 0:
        from oracle import pyoracle as po
        oc = po.OracleCulling()
        oc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
        oi, ot, _ = oc.cull(lb.culling.frustum_bytes(f))
        for t in range(3):
            ok &= np.array_equal(np.sort(merged.get(t, np.zeros(0, np.uint32))).astype(np.int64), np.sort(oi[ot == t]).astype(np.int64))
        ok &= len(oi) > 5000
    # every rank sees the same gathered data
    digest = torch.tensor([sum(int(np.sort(v).astype(np.uint64).sum()) % (1 << 40) for v in merged.values())])
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    ok &= all(int(g.item()) == int(digest.item()) for g in gathered)
    q.put((rank, bool(ok)))
    dist.barrier()
    cs.close()
    ctx.close()
    dist.destroy_process_group()


def test_two_gpu_cull_allgather():
    import lumixengine_b200 as lb
    if lb.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok in results), results
    assert all(p.exitcode == 0 for p in procs)
