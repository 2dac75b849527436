"""Single-process multi-GPU path (what a one-process host such as Unity uses): one context per device, disjoint ray
ranges, rar_allreduce_slots over NVLink peer memory.  On a 1-GPU box the same code runs with two contexts on the one
device (the reduce kernel then reads its peer's histogram from local memory), so the sharding logic and the kernel are
exercised there too; `gpurun --gpus 2` and up give the real peer-memory path."""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from realisticaudioraytracing2d_b200.host.sharding import dispatched_threads, shard_range
from tests.common import capi_params, oracle_params, oracle_walls, trace_kwargs

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_single_context_allreduce_is_a_noop(ctx):
    ctx.ir_clear(0, 128, 1)
    _capi.allreduce_slots([ctx], 0)
    assert not ctx.ir_read_fixed(0, 128).any()


@pytest.mark.parametrize("bands", [1, 8])
def test_peer_memory_allreduce_matches_unsharded_trace(oracle, bands):
    real = _device_count()
    n_dev = min(max(real, 2), 8)                      # at least two contexts, wrapped onto the devices there are
    sc = scenes.maze(n_segments=600, ray_count=40_000, max_bounces=10, bands=8)
    kw = trace_kwargs(sc, bands=bands)
    n = kw["impulse_length"]
    ctxs = [_capi.Context(d % real) for d in range(n_dev)]
    try:
        total = dispatched_threads(kw["ray_count"])
        for r, c in enumerate(ctxs):
            c.set_walls(sc.walls)
            c.set_wall_band_absorption(sc.band_absorption)
            c.ir_clear(0, n, bands)
            lo, hi = shard_range(total, r, n_dev)
            c.trace(capi_params(_capi, dict(kw, ray_begin=lo, ray_end=hi)), 0)
        _capi.allreduce_slots(ctxs, 0)
        want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), band_abs=sc.band_absorption if bands > 1 else None).hist
        for c in ctxs:
            assert np.array_equal(c.ir_read_fixed(0, n * bands), want)
    finally:
        for c in ctxs:
            c.destroy()
