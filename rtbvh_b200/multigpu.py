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


class FusedGather:
    """Gather of the hit records fused into the traversal kernel (rtbvh_gpu_*_device_scatter): every rank owns
    `buffers` gather buffers of `world * rays_per_rank` records that its peers map with cudaIpc; the kernel writes
    each record into slot [rank * rays_per_rank + i] of EVERY rank's buffer (P2P stores over NVLink / NVSwitch), and a
    one-block device barrier (rtbvh_gpu_peer_barrier) closes the step — no NCCL call on the data path.  Handles travel
    once, out of band, through torch.distributed (any backend).

    The barrier is OFF the critical path: step k's barrier runs on a side stream behind an event recorded after step k's
    kernel, and the kernel of step k only waits for the barrier of step k - 2 (long finished by then), so neither the
    barrier's launch nor the wait for the slowest rank sits between two traversal kernels (8 GPUs: 4.41 -> see
    profiles/r5n_gather_ab.txt for what the in-line barrier cost).

    Buffer discipline: step k uses buffer k % buffers (default 4).  `wait(k, stream)` makes `stream` wait until buffer
    k % buffers holds the records of ALL ranks; a consumer of step k must be enqueued (behind `wait(k, ...)`) before the
    call of step k + 2 on the same stream: the barrier of step k + 1 then certifies that every rank has consumed step k,
    and step k + 4 — the next writer of that buffer — waits for the barrier of step k + 2."""

    def __init__(self, rays_per_rank: int, record_bytes: int, buffers: int = 4):
        import torch
        import torch.distributed as dist
        from . import api
        self.api, self.torch = api, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.rays_per_rank, self.record_bytes = rays_per_rank, record_bytes
        self.own = [api.PeerBuffer(self.world * rays_per_rank * record_bytes) for _ in range(buffers)]
        self.flags = api.PeerBuffer(64)
        mine = [b.handle_bytes() for b in self.own] + [self.flags.handle_bytes()]
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        self.opened = []
        self.dests = []       # dests[b][r] = pointer to rank r's buffer b as seen from this rank
        for b in range(buffers + 1):
            row = []
            for r in range(self.world):
                if r == self.rank:
                    row.append((self.own[b] if b < buffers else self.flags).ptr.value)
                else:
                    p = api.PeerBuffer.open(everyone[r][b])
                    self.opened.append(p)
                    row.append(p)
            self.dests.append(row)
        self.flag_dests = self.dests.pop()
        self.step = 0
        self.side = torch.cuda.Stream()
        self.done = {}        # step -> event recorded on the side stream behind that step's barrier
        dist.barrier()

    def buffer_ptr(self, k: int) -> int:
        return self.own[k % len(self.own)].ptr.value

    def _main(self, stream: int):
        return self.torch.cuda.ExternalStream(stream) if stream else self.torch.cuda.default_stream()

    def wait(self, k: int, stream: int = 0):
        """Stream-orders `stream` behind the barrier of step k (its gather buffer is complete on this rank afterwards)."""
        ev = self.done.get(k)
        if ev is not None:
            self._main(stream).wait_event(ev)

    def intersect(self, scene, d_rays, n: int, k: int, d_hits=None, tree=None, stream: int = 0, any_hit: bool = False):
        """Traces this rank's shard for step k, scattering into every rank's buffer k % buffers; the step barrier follows on
        the side stream.  Steps must be issued with consecutive k."""
        tree = self.api.TREE_MBVH if tree is None else tree
        main = self._main(stream)
        self.wait(k - 2, stream)  # every rank is past step k - 2, hence has consumed what step k overwrites (see above)
        fn = scene.occluded_device_scatter if any_hit else scene.intersect_device_scatter
        fn(d_rays, n, self.dests[k % len(self.own)], self.rank * self.rays_per_rank, d_hits, tree, stream)
        after = self.torch.cuda.Event()
        after.record(main)
        self.side.wait_event(after)
        self.step += 1
        self.api.peer_barrier(self.flag_dests, self.rank, self.step, self.side.cuda_stream)
        ev = self.torch.cuda.Event()
        ev.record(self.side)
        self.done[k] = ev
        self.done.pop(k - 8, None)

    def close(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        dist.barrier()
        for p in self.opened:
            self.api.PeerBuffer.close(p)
        self.opened = []
        dist.barrier()
        for b in self.own:
            b.free()
        self.flags.free()
