"""Multi-GPU plumbing of the ray stage: contiguous ray-id ranges per rank and one integer all-reduce.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU tests).  Rays
are independent -- the RNG is a pure function of the global ray id and the frame
(Raytrace2D.compute:51) -- so rank r of W traces the thread ids [r*N/W, (r+1)*N/W) of one dispatch into
its private Q23.40 histogram, and a single all-reduce(sum, int64) of bins x bands words produces the
complete histogram on every rank.  Integer addition commutes, so the result is bit-identical for any W.
Batched listeners / streams shard by contiguous batch range with no collective at all.

Two interchangeable data paths for that all-reduce: `PeerExchange` -- the library's own kernel over
CUDA-IPC-mapped peer memory (rar_exchange_*; the process group only carries the 80-byte handles at set-up
and the tear-down barrier) -- and `allreduce_histogram`, ncclAllReduce on a zero-copy view of the slot.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` items for `rank` of `world`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def interleaved_chunks(total: int, rank: int, world: int, chunk_log2: int):
    """The [begin, end) id ranges rar_trace_interleaved(rank, world, chunk_log2) traces of a dispatch of `total` thread
    ids: chunks rank, rank + world, ... of 2^chunk_log2 ids (the last chunk of the dispatch may be short)."""
    if world <= 0 or not (0 <= rank < world) or total < 0 or chunk_log2 < 0:
        raise ValueError("bad shard arguments")
    chunk = 1 << chunk_log2
    return [(c * chunk, min((c + 1) * chunk, total)) for c in range(rank, -(-total // chunk), world)]


def dispatched_threads(ray_count: int, exact: bool = False) -> int:
    """Threads the reference's Trace dispatch runs: ceil(rayCount/64)*64 unless `exact`
    (Raytrace2D.compute:49-52, Helpers/ComputeHelper.cs:27-31)."""
    return ray_count if exact else (ray_count + 63) // 64 * 64


def allreduce_histogram(hist_tensor, group=None) -> None:
    """In-place sum of the int64 histogram over all ranks (a no-op outside a process group)."""
    import torch
    import torch.distributed as dist
    if hist_tensor.dtype != torch.int64:
        raise TypeError("the impulse-response histogram is int64 fixed point")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist_tensor, op=dist.ReduceOp.SUM, group=group)


class DeviceHistogram:
    """Zero-copy torch view of a context's IR slot (rar_ir_device_ptr) for the collective.

    The library treats a slot whose address has been handed out as externally writable from then on (its cached
    spectra are never reused) and refuses to reallocate it; `check()` re-queries the address for callers that want
    to assert the view is still the slot."""

    def __init__(self, ctx, slot: int, device):
        import torch
        self._ctx, self._slot = ctx, slot
        ptr, n = ctx.ir_device_ptr(slot)
        self._ptr, self._n = ptr, n
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        self.tensor = torch.as_tensor(self, device=device)

    def check(self) -> None:
        ptr, n = self._ctx.ir_device_ptr(self._slot)
        if (ptr, n) != (self._ptr, self._n):
            raise RuntimeError("the slot was reconfigured after its device address was taken: re-create the view")


def gather_handles(handle: bytes, group=None):
    """All ranks' exchange handles in rank order (the only use of the process group on the exchange path)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, bytes(handle), group=group)
    if any(not isinstance(h, (bytes, bytearray)) or len(h) != len(handle) for h in out):
        raise RuntimeError("a rank delivered a malformed exchange handle")
    return [bytes(h) for h in out]


class PeerExchange:
    """The all-reduce of the ray-range sharding as one kernel per rank over NVLink peer memory.

    Collective construction: every rank of `group` must create it with the same capacity.  `allreduce(slot)`
    only enqueues on the context's stream; `close()` is collective too (it fences the peers before unmapping).
    """

    def __init__(self, ctx, capacity_words: int, group=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group (one process per GPU)")
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        # Set-up is all-or-nothing across the group: every rank reports whether it could create and map the
        # regions, and either all ranks end up connected or all raise (so a caller can fall back consistently).
        err = None
        try:
            handle = ctx.exchange_create(capacity_words)
        except Exception as ex:  # noqa: BLE001 - reported to the peers below
            handle, err = b"\0" * 80, f"create: {ex}"
        handles = gather_handles(handle, group)
        if err is None:
            try:
                ctx.exchange_connect(self.rank, self.world, handles)
            except Exception as ex:  # noqa: BLE001
                err = f"connect: {ex}"
        status = [None] * self.world
        dist.all_gather_object(status, err, group=group)   # also the barrier: nobody signals an unmapped peer
        if any(st is not None for st in status):
            try:
                ctx.exchange_destroy()
            except Exception:  # noqa: BLE001
                pass
            raise RuntimeError("peer-memory exchange unavailable: " +
                               "; ".join(f"rank {r}: {st}" for r, st in enumerate(status) if st is not None))
        self._open = True

    def allreduce(self, slot: int, mode: int = 0) -> None:
        self.ctx.exchange_allreduce(slot, mode)

    def check(self) -> None:
        """Blocking: raises if a peer failed to reach a barrier of an earlier call."""
        self.ctx.exchange_status()

    def close(self) -> None:
        import torch.distributed as dist
        if self._open:
            self._open = False
            self.ctx.sync()
            dist.barrier(group=self.group)
            self.ctx.exchange_destroy()
