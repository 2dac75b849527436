"""world_size-2 check (gloo, CPU) of the multi-GPU plumbing: each rank holds the histogram of its own
contiguous ray-id range, one all-reduce(sum, int64) gives every rank the complete histogram, and the result
equals the unsharded trace bit for bit.  The per-rank histograms come from the oracle here (no GPU in this
container); on the GPU box the same helpers drive the CUDA path (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realisticaudioraytracing2d_b200 import scenes
from realisticaudioraytracing2d_b200.host.sharding import (allreduce_histogram, dispatched_threads, gather_handles,
                                                           interleaved_chunks, shard_range)
from tests.common import oracle_params, oracle_walls, trace_kwargs


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc, ray_count=3000)
    lo, hi = shard_range(dispatched_threads(3000), rank, world)
    part = O.trace(oracle_walls(O, sc.walls), oracle_params(O, dict(kw, ray_begin=lo, ray_end=hi)), n_threads=2).hist
    t = torch.from_numpy(part.copy())
    allreduce_histogram(t)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), t.numpy())
    # the block-cyclic shard bench.py uses at N > 1 (rar_trace_interleaved): chunks rank, rank + world, ... of 2^8 ids
    part = np.zeros_like(part)
    for lo, hi in interleaved_chunks(dispatched_threads(3000), rank, world, 8):
        part += O.trace(oracle_walls(O, sc.walls), oracle_params(O, dict(kw, ray_begin=lo, ray_end=hi)), n_threads=2).hist
    t = torch.from_numpy(part.copy())
    allreduce_histogram(t)
    np.save(os.path.join(out_dir, f"cyclic{rank}.npy"), t.numpy())
    # the transport of the peer-memory exchange's handles: rank order, every rank sees all of them
    handles = gather_handles(bytes([rank]) * 80)
    assert handles == [bytes([r]) * 80 for r in range(world)]
    dist.destroy_process_group()


def test_two_rank_allreduce_is_bit_identical_to_one_rank(tmp_path, oracle):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sc = scenes.smoll_room()
    whole = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, trace_kwargs(sc, ray_count=3000))).hist
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"rank{r}.npy"), whole)
        assert np.array_equal(np.load(tmp_path / f"cyclic{r}.npy"), whole)


class _StubContext:
    """Stands in for _capi.Context in the set-up protocol of PeerExchange (no GPU here)."""

    def __init__(self, rank, fail_connect_on):
        self.rank, self.fail_on, self.destroyed, self.connected = rank, fail_connect_on, False, None

    def exchange_create(self, capacity_words):
        return bytes([self.rank]) * 80

    def exchange_connect(self, rank, world, handles):
        if rank == self.fail_on:
            raise RuntimeError("cudaIpcOpenMemHandle: simulated failure")
        self.connected = list(handles)

    def exchange_destroy(self):
        self.destroyed = True

    def sync(self):
        pass


def _exchange_worker(rank, world, port, out_dir, fail_on):
    from realisticaudioraytracing2d_b200.host.sharding import PeerExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _StubContext(rank, fail_on)
    try:
        ex = PeerExchange(ctx, 1000)
        outcome = "connected" if ctx.connected == [bytes([r]) * 80 for r in range(world)] else "bad handles"
        ex.close()
        outcome += "+closed" if ctx.destroyed else ""
    except RuntimeError as e:
        outcome = f"raised destroyed={ctx.destroyed}: {e}"
    with open(os.path.join(out_dir, f"outcome{rank}.txt"), "w") as f:
        f.write(outcome)
    dist.destroy_process_group()


def test_peer_exchange_setup_is_all_or_nothing(tmp_path):
    """Either every rank ends up connected or every rank raises (so a caller can fall back consistently)."""
    for fail_on, sub in ((-1, "ok"), (1, "fail")):
        d = tmp_path / sub
        d.mkdir()
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        mp.spawn(_exchange_worker, args=(2, port, str(d), fail_on), nprocs=2, join=True)
        got = [(d / f"outcome{r}.txt").read_text() for r in range(2)]
        if fail_on < 0:
            assert got == ["connected+closed"] * 2
        else:
            assert all(g.startswith("raised destroyed=True") and "rank 1: connect" in g for g in got), got


def test_allreduce_is_a_noop_without_a_process_group():
    t = torch.arange(8, dtype=torch.int64)
    allreduce_histogram(t)
    assert t.tolist() == list(range(8))
