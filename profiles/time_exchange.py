"""Where an exchange step's time goes (run under torchrun with 2+ ranks): device time vs host issue time of the cull alone, before and
after the communicator / peer mappings exist, and of the exchange step issued from Python and from C."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import lumixengine_b200 as lb
from lumixengine_b200 import scenes
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
ctx = lb.Context(local)
scene = scenes.c2_scene(10_000_000, seed=2 + rank)
cs = lb.CullingSystem(ctx); cs.set_replicas(8)
cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"]); cs.flush()
f = lb.frustum_perspective(**scenes.c2_frustum_args())
N = 300

def timed(tag, fn):
    fn(); ctx.synchronize()
    e0, e1 = ctx.event(), ctx.event()
    ctx.synchronize(); ctx.record(e0); t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1)
    if rank == 0:
        print(f"XCHG {tag:34s} device {ms / N * 1e3:7.2f} us/step   host issue {(t1 - t0) / N * 1e6:7.2f} us/step", flush=True)

timed("cull_device_n, no torch/nccl yet", lambda: cs.cull_device_n(f, N))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
timed("cull_device_n, torch nccl up", lambda: cs.cull_device_n(f, N))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.from_numpy(ctx.comm_unique_id()))
dist.broadcast(uid, 0)
ctx.comm_init(world, rank, uid.cpu().numpy())
timed("cull_device_n, own comm up", lambda: cs.cull_device_n(f, N))
t = torch.tensor([cs.exchange_slab_words()], dtype=torch.int64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ctx.comm_enable_p2p(int(t.item()) - 256)
timed("cull_device_n, peers mapped", lambda: cs.cull_device_n(f, N))
dist.barrier()
timed("cull_exchange x N from python", lambda: [cs.cull_exchange(f) for _ in range(N)])
dist.barrier()
timed("cull_exchange_n from C (lanes)", lambda: cs.cull_exchange_n(f, N))
dist.barrier()
os.environ["X"] = "1"
cs.close(); ctx.close(); dist.destroy_process_group()
