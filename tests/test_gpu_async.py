"""Frame-rate entry points do not stall the caller, and the device-built uniform grid
equals the host builder's.

The reference issues all of these from Unity's main thread once per frame or per FixedUpdate
(RayTraceManager.cs:50-53,64-89,246-250), where a blocked call is a dropped frame."""
import time

import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from tests import emulation
from tests.common import capi_params, oracle_params, oracle_walls, trace_kwargs
from tests.test_gpu_trace import _soup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["smoll", "big", "maze2000", "maze10000", "soup700", "soup3000", "shoebox", "degenerate"])
def test_device_built_grid_equals_the_host_builder(ctx, case):
    if case == "smoll":
        walls = scenes.smoll_room().walls
    elif case == "big":
        walls = scenes.big_room().walls
    elif case.startswith("maze"):
        walls = scenes.maze(n_segments=int(case[4:]), ray_count=64, max_bounces=2, bands=8).walls
    elif case.startswith("soup"):
        walls = _soup(7, int(case[4:]))
    elif case == "shoebox":
        walls = scenes.shoebox().walls
    else:
        walls = scenes.smoll_room().walls.copy()
        walls["end"][3] = walls["start"][3]                  # zero-length wall
        walls = np.concatenate([walls, walls[:5]])           # duplicates
    ctx.set_walls(walls)
    got, want = ctx.debug_grid(), emulation.grid_digest(walls)
    assert got == want and got["nx"] > 0 and got["n_items"] >= len(walls)


def _busy(ctx, seconds=0.25):
    """Enqueues a brute-force trace that keeps the device busy for roughly `seconds`; returns a ticket that completes
    when it has finished."""
    sc = scenes.maze(n_segments=10000, ray_count=int(148 * 1024 * 8 * seconds / 0.25), max_bounces=16, bands=8)
    n = sc.impulse_length
    ctx.set_walls(sc.walls)
    ctx.ir_clear(30, n, 1)
    ctx.sync()
    ctx.trace(capi_params(_capi, trace_kwargs(sc)), 30)
    return ctx.ir_read_begin(30, 16), sc


def test_frame_rate_entry_points_return_while_the_device_is_busy(ctx, oracle):
    """set_walls, set_wall_band_absorption, a grid trace on NEW walls (device grid build), trace_listeners and the
    batched response load are all enqueued behind a long-running kernel and return at once."""
    small = scenes.maze(n_segments=3000, ray_count=4096, max_bounces=6, bands=8, seed=4)
    n = small.impulse_length
    listeners = np.stack([np.linspace(20, 60, 16), np.linspace(25, 55, 16)], 1).astype(np.float32)
    cv = _capi.Convolver(ctx, 8, 256, 48000)
    irs = np.random.default_rng(0).standard_normal((8, 48000)).astype(np.float32) * np.float32(0.01)
    # everything below has run once: buffers are at their final sizes (growth is allowed to wait)
    ctx.set_walls(small.walls)
    ctx.set_wall_band_absorption(small.band_absorption)
    for l in range(16):
        ctx.ir_clear(40 + l, n, 1)
    ctx.ir_clear(31, n, 8)
    ctx.trace_listeners(capi_params(_capi, trace_kwargs(small)), listeners, 40)
    ctx.trace(capi_params(_capi, trace_kwargs(small, flags=_capi.RAR_FLAG_USE_GRID)), 40)
    ctx.trace(capi_params(_capi, trace_kwargs(small, bands=8)), 31)   # (a kernel's very first launch loads its code, which waits for the device)
    cv.set_irs(0, irs)
    ctx.sync()
    try:
        ticket, _ = _busy(ctx, 0.6)
        spent = {}

        def timed(name, fn, *a):
            t = time.perf_counter()
            fn(*a)
            spent[name] = spent.get(name, 0.0) + (time.perf_counter() - t) * 1e3

        timed("set_walls", ctx.set_walls, small.walls)                   # UpdateGeometry (RayTraceManager.cs:246-250)
        timed("set_wall_band_absorption", ctx.set_wall_band_absorption, small.band_absorption)
        for l in range(16):
            timed("ir_clear", ctx.ir_clear, 40 + l, n, 1)
        timed("ir_clear", ctx.ir_clear, 31, n, 8)
        timed("trace (grid built on the device)", ctx.trace, capi_params(_capi, trace_kwargs(small, flags=_capi.RAR_FLAG_USE_GRID)), 40)
        timed("trace (8 bands)", ctx.trace, capi_params(_capi, trace_kwargs(small, bands=8)), 31)
        timed("trace_listeners", ctx.trace_listeners, capi_params(_capi, trace_kwargs(small)), listeners, 40)
        timed("conv.set_irs", cv.set_irs, 0, irs)
        dt = sum(spent.values()) * 1e-3
        still_running = not ctx.poll(ticket)
        ctx.ir_read_end(ticket, 16)
        report = ", ".join(f"{k} {v:.2f} ms" for k, v in spent.items())
        assert dt < 0.05, f"the calls took {dt * 1e3:.1f} ms of host time behind a busy device: {report}"
        assert still_running, f"the long kernel finished before the calls returned ({report}): nothing was demonstrated"
        # ... and what they enqueued is right
        want = oracle.trace(oracle_walls(oracle, small.walls), oracle_params(oracle, dict(trace_kwargs(small), listener=(float(listeners[3, 0]), float(listeners[3, 1]))))).hist
        assert np.array_equal(ctx.ir_read_fixed(43, n), want)
        want8 = oracle.trace(oracle_walls(oracle, small.walls), oracle_params(oracle, trace_kwargs(small, bands=8)), band_abs=small.band_absorption).hist
        assert np.array_equal(ctx.ir_read_fixed(31, n * 8), want8)
    finally:
        cv.destroy()


def test_band_table_update_is_ordered_behind_a_trace_in_flight(ctx, oracle):
    """the band-absorption upload used a blocking copy on the legacy stream, unordered against the
    context's non-blocking stream.  A banded trace in flight must see the table it was launched with."""
    sc = scenes.maze(n_segments=2000, ray_count=300_000, max_bounces=16, bands=8, seed=6)
    n = sc.impulse_length
    kw = trace_kwargs(sc, bands=8)
    other = (sc.band_absorption[::-1] * np.float32(0.5)).copy()
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(32, n, 8)
    ctx.ir_clear(33, n, 8)
    ctx.sync()
    ctx.trace(capi_params(_capi, kw), 32)                 # in flight with the first table ...
    ctx.set_wall_band_absorption(other)                   # ... while the second one is uploaded
    ctx.trace(capi_params(_capi, kw), 33)
    a, b = ctx.ir_read_fixed(32, n * 8), ctx.ir_read_fixed(33, n * 8)
    O = oracle
    assert np.array_equal(a, O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw), band_abs=sc.band_absorption).hist)
    assert np.array_equal(b, O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw), band_abs=other).hist)


def test_async_reads_are_ordered_against_later_writes_of_the_same_slot(ctx, oracle):
    """rar_ir_read_begin converts and copies on the context's read stream so that the next frame (other slot) does not
    queue behind it.  A read in flight must still see the histogram as it was when the read was requested, whatever is
    enqueued on the slot afterwards (clear, another trace), and reads of both slots may be in flight at once."""
    sc = scenes.maze(n_segments=2000, ray_count=200_000, max_bounces=12, bands=1, seed=9)
    n = sc.impulse_length
    ctx.set_walls(sc.walls)
    O = oracle
    want = [O.trace(oracle_walls(O, sc.walls), oracle_params(O, trace_kwargs(sc, rng_state_offset=f))).hist for f in (1, 2, 3)]
    for rep in range(3):
        tickets = []
        for k, f in enumerate((1, 2, 3)):
            s = k & 1
            ctx.ir_clear(s, n, 1)                            # (for k = 2: slot 0 again, while its first read may be in flight)
            ctx.trace(capi_params(_capi, trace_kwargs(sc, rng_state_offset=f)), s)
            tickets.append(ctx.ir_read_begin(s, n))
        ctx.ir_clear(0, n, 1)                                # ... and cleared once more behind the last read
        ctx.ir_clear(1, n, 1)
        for k, t in enumerate(tickets):
            got = ctx.ir_read_end(t, n)
            assert np.array_equal(got, O.ir_to_float(want[k])), (rep, k)
    assert not ctx.ir_read_fixed(0, n).any() and not ctx.ir_read_fixed(1, n).any()
