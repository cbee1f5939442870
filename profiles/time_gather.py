"""2-rank timing of the per-frame exchange variants (run with torchrun --nproc-per-node 2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = lb.Context(rank)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: uid.copy_(torch.from_numpy(ctx.comm_unique_id()))
dist.broadcast(uid, 0)
ctx.comm_init(world, rank, uid.cpu().numpy())
scene = scenes.c2_scene(10_000_000, seed=2 + rank)
cs = lb.CullingSystem(ctx); cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
f = lb.frustum_perspective(**scenes.c2_frustum_args())
t = torch.tensor([cs.cull(f).total], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); slab = int(t.item()) + 1024  # same slab size on every rank
def timeit(fn, n=200):
    for _ in range(20): fn()
    ctx.synchronize(); dist.barrier()
    e0, e1 = ctx.event(), ctx.event(); ctx.record(e0)
    for _ in range(n): fn()
    ctx.record(e1); ms = ctx.elapsed_ms(e0, e1) / n
    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
r = {}
r["cull_only"] = timeit(lambda: cs.cull_device(f, want_counts=False))
r["nccl"] = timeit(lambda: cs.cull_gather(f, slab))
ctx.comm_enable_p2p(slab)
r["p2p"] = timeit(lambda: cs.cull_gather(f, slab))
if rank == 0: print("GATHER_TIMES_US", {k: round(v * 1e3, 1) for k, v in r.items()}, "push_grid", os.environ.get("LB200_PUSH_GRID"))
cs.close(); ctx.close(); dist.destroy_process_group()
