"""One process per GPU: the all-reduce of the ray-range sharding done by the library's own kernel over CUDA-IPC
peer memory (rar_exchange_*), checked bit for bit against the unsharded oracle trace.  The process group (gloo)
only carries the 80-byte handles.  With >= 2 GPUs (`gpurun --gpus 2`) every process has its own device; on a 1-GPU box
the two processes share the device: their kernels are then time-sliced instead of concurrent, which the flag barrier
tolerates (a rank spins until the peer's kernel gets its slice), so the protocol and the kernel are still checked."""
import os
import socket

import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from realisticaudioraytracing2d_b200.host.sharding import dispatched_threads, shard_range
from tests.common import capi_params, oracle_params, oracle_walls, trace_kwargs

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_exchange_state_machine_on_one_rank(ctx):
    ctx.ir_clear(0, 1000, 1)
    with pytest.raises(_capi.RarError) as e:
        ctx.exchange_allreduce(0)
    assert e.value.code == -3                                   # RAR_ERR_STATE: not connected
    with pytest.raises(_capi.RarError):
        ctx.exchange_status()                                   # nothing created yet
    h = ctx.exchange_create(1000)
    assert len(h) == _capi.RAR_EXCHANGE_HANDLE_BYTES
    with pytest.raises(ValueError):
        ctx.exchange_connect(0, 2, [h])                         # one handle per rank
    ctx.exchange_connect(0, 1, [h])
    with pytest.raises(_capi.RarError):
        ctx.exchange_connect(0, 1, [h])                         # already connected
    ir = np.zeros(1000, np.float32)
    ir[7] = 0.25
    ctx.ir_write(0, ir)
    ctx.exchange_allreduce(0)                                   # world 1: leaves the slot alone
    ctx.exchange_status()
    assert np.array_equal(ctx.ir_read(0, 1000), ir)
    ctx.ir_clear(1, 2000, 1)
    with pytest.raises(_capi.RarError):
        ctx.exchange_allreduce(1)                               # larger than the capacity
    ctx.exchange_destroy()
    ctx.exchange_destroy()                                      # idempotent


def _odd_payload(rank, words):
    """Small integers times 2^-10: exactly representable, so the quantised words are known on the host."""
    return (np.random.default_rng(100 + rank).integers(-1000, 1000, words) / 1024.0).astype(np.float32)


def _worker(rank, world, port, out_dir, bands):
    import torch
    import torch.distributed as dist
    from realisticaudioraytracing2d_b200.host.sharding import PeerExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    device = rank % torch.cuda.device_count()
    torch.cuda.set_device(device)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.maze(n_segments=600, ray_count=40_000, max_bounces=10, bands=8)
    kw = trace_kwargs(sc, bands=bands)
    n = kw["impulse_length"]
    total = dispatched_threads(kw["ray_count"])
    lo, hi = shard_range(total, rank, world)
    ctx = _capi.Context(device)
    try:
        ctx.set_walls(sc.walls)
        ctx.set_wall_band_absorption(sc.band_absorption)
        ex = PeerExchange(ctx, n * bands)
        # several calls back to back: both modes, both parities of the double-buffered staging, no host sync between
        for it, mode in enumerate([_capi.RAR_EXCHANGE_ONE_SHOT, _capi.RAR_EXCHANGE_TWO_SHOT, _capi.RAR_EXCHANGE_TWO_SHOT,
                                   _capi.RAR_EXCHANGE_ONE_SHOT, _capi.RAR_EXCHANGE_AUTO]):
            ctx.ir_clear(it, n, bands)
            ctx.trace(capi_params(_capi, dict(kw, ray_begin=lo, ray_end=hi, rng_state_offset=1 + it % 2)), it)
            ex.allreduce(it, mode)
        # an odd word count (the last 16-byte vector is half padding) and a single word, in both modes
        odd = []
        for k, (words, mode) in enumerate([(4801, _capi.RAR_EXCHANGE_ONE_SHOT), (4801, _capi.RAR_EXCHANGE_TWO_SHOT),
                                           (1, _capi.RAR_EXCHANGE_ONE_SHOT), (1, _capi.RAR_EXCHANGE_TWO_SHOT)]):
            ctx.ir_write(10 + k, _odd_payload(rank, words))
            ex.allreduce(10 + k, mode)
            odd.append(ctx.ir_read_fixed(10 + k, words))
        ex.check()
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack([ctx.ir_read_fixed(it, n * bands) for it in range(5)]))
        np.save(os.path.join(out_dir, f"odd{rank}.npy"), np.concatenate(odd))
        ex.close()
    finally:
        ctx.destroy()
        dist.destroy_process_group()


@pytest.mark.parametrize("bands", [1, 8])
def test_peer_exchange_matches_unsharded_trace(tmp_path, oracle, bands):
    import torch.multiprocessing as mp
    world = min(max(_device_count(), 2), 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), bands), nprocs=world, join=True)
    sc = scenes.maze(n_segments=600, ray_count=40_000, max_bounces=10, bands=8)
    kw = trace_kwargs(sc, bands=bands)
    want = [oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, dict(kw, rng_state_offset=f)),
                         band_abs=sc.band_absorption if bands > 1 else None).hist for f in (1, 2)]
    assert want[0].any() and not np.array_equal(want[0], want[1])
    odd_want = np.concatenate([sum((_odd_payload(r, words).astype(np.float64) * 2.0 ** 40).astype(np.int64) for r in range(world))
                               for words in (4801, 4801, 1, 1)])
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        for it in range(5):
            assert np.array_equal(got[it], want[it % 2]), (r, it)
        assert np.array_equal(np.load(tmp_path / f"odd{r}.npy"), odd_want), r
