"""Host-side logic of the multi-GPU path: entities shard by index range (one process per GPU, each rank
owns whole cell pages of its own CullingSystem); the only exchange is the all-gather of the compacted visible lists.

Visibility of an entity depends only on its own sphere and its cell's origin (culling_system.cpp:342-363), never on its
neighbours, so any partition of the entities gives the same union of visible sets as the unsharded cull.
"""
import numpy as np


def index_range(n, rank, world):
    """[begin, end) of rank's share of n entities: contiguous, sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_scene(scene, rank, world):
    """Rank's slice of a scene dict (entities keep their global ids)."""
    b, e = index_range(len(scene["entities"]), rank, world)
    return {k: v[b:e] for k, v in scene.items()}


def slab_layout(counts):
    """counts[r, t] (what lb200_culling_allgather returns) -> per rank: offsets of each type inside that rank's slab."""
    counts = np.asarray(counts, np.int64)
    offs = np.zeros_like(counts)
    offs[:, 1:] = np.cumsum(counts, axis=1)[:, :-1]
    return offs


def merge_gathered(slabs, counts):
    """slabs[r] = rank r's id slab (its visible ids packed type after type), counts[r, t] -> {type: ids of all ranks}."""
    counts = np.asarray(counts, np.int64)
    offs = slab_layout(counts)
    out = {}
    for t in np.nonzero(counts.sum(axis=0))[0]:
        parts = [np.asarray(slabs[r])[offs[r, t]:offs[r, t] + counts[r, t]] for r in range(counts.shape[0]) if counts[r, t]]
        out[int(t)] = np.concatenate(parts)
    return out
