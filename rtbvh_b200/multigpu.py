"""Multi-GPU plumbing: the tree is replicated, rays are sharded by contiguous index range, hit
records are gathered.  One process per GPU over torch.distributed (NCCL on the box, gloo in the CPU tests).
Nothing here touches the kernels: it is the host-side partition / broadcast / gather logic."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Ray index range [g*n/G, (g+1)*n/G) of rank g."""
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_counts(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def broadcast_arrays(arrays: dict | None, src: int = 0, device: str = "cpu") -> dict:
    """Broadcasts a dict of numpy arrays (the tree: nodes, indices, ...) from `src` to every rank.
    Structured dtypes travel as raw bytes; metadata goes through broadcast_object_list."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        meta[0] = {k: (v.dtype.descr if v.dtype.names else v.dtype.str, v.shape) for k, v in arrays.items()}
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for k, (descr, shape) in meta[0].items():
        dt = np.dtype([tuple(d) if isinstance(d, list) else d for d in descr]) if isinstance(descr, list) else np.dtype(descr)
        nbytes = int(np.prod(shape)) * dt.itemsize
        if rank == src:
            buf = torch.from_numpy(np.ascontiguousarray(arrays[k]).view(np.uint8).reshape(-1).copy()).to(device)
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if nbytes:
            dist.broadcast(buf, src=src)
        out[k] = arrays[k] if rank == src else buf.cpu().numpy().view(dt).reshape(shape)
    return out


def all_gather_ragged(local, counts: list[int]):
    """all_gather of per-rank tensors with different leading sizes (contiguous shards): pads to the largest
    shard, gathers, and returns the concatenation in rank order == global ray order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    m = max(counts)
    pad = local
    if local.shape[0] != m:
        pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    if all(c == m for c in counts):
        return out
    return torch.cat([out[r * m: r * m + counts[r]] for r in range(world)])
