"""GPU parity of the trace kernel against the CPU oracle, through the C-ABI.

Bars (BASELINE.json north_star): integer histogram bins bit-exact; hit distance <= 1e-5 relative.
Because the kernel and the oracle obey the same binary32 arithmetic contract the hit lists are in
fact expected to be bit-identical; the test asserts that and reports the relative error otherwise.
"""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from tests.common import capi_params, oracle_params, oracle_walls, sort_hits, trace_kwargs

pytestmark = pytest.mark.gpu


def _run_pair(ctx, O, sc, kw, frames=(1,)):
    bands = kw["bands"]
    n_words = kw["impulse_length"] * bands
    ctx.set_walls(sc.walls)
    if bands > 1:
        ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(0, kw["impulse_length"], bands)
    ctx.get_counters(reset=True)
    hist = np.zeros(n_words, dtype=np.int64)
    ctr = None
    for f in frames:
        k = dict(kw, rng_state_offset=f)
        ctx.trace(capi_params(_capi, dict(k, flags=k["flags"] | _capi.RAR_FLAG_COUNT_TESTS)), 0)
        r = O.trace(oracle_walls(O, sc.walls), oracle_params(O, k), band_abs=sc.band_absorption if bands > 1 else None, hist=hist)
        ctr = r.counters if ctr is None else {a: ctr[a] + r.counters[a] for a in ctr}
    got = ctx.ir_read_fixed(0, n_words)
    return got, hist, ctx.get_counters(), ctr


@pytest.mark.parametrize("name,frames", [("smoll", (1, 2, 3)), ("big", (1,)), ("smoll1000", (7,))])
def test_bundled_rooms_histogram_bit_exact(ctx, oracle, name, frames):
    sc = scenes.big_room() if name == "big" else scenes.smoll_room()
    kw = trace_kwargs(sc, ray_count=1000 if name == "smoll1000" else sc.ray_count)
    got, want, gc, oc = _run_pair(ctx, oracle, sc, kw, frames)
    assert np.count_nonzero(want) > 500
    assert np.array_equal(got, want)
    assert gc == oc


def test_shoebox_histogram_bit_exact(ctx, oracle):
    sc = scenes.shoebox(ray_count=200_000, max_bounces=32)
    got, want, gc, oc = _run_pair(ctx, oracle, sc, trace_kwargs(sc))
    assert np.array_equal(got, want)
    assert gc == oc


def test_shoebox_scattering_transmission_bit_exact(ctx, oracle):
    sc = scenes.shoebox(ray_count=100_000, max_bounces=16, scattering=0.35, transmission=0.2, ior=1.3)
    got, want, gc, oc = _run_pair(ctx, oracle, sc, trace_kwargs(sc))
    assert np.array_equal(got, want)
    assert gc == oc


@pytest.mark.parametrize("bands", [1, 8])
def test_maze_histogram_bit_exact(ctx, oracle, bands):
    sc = scenes.maze(n_segments=2000, ray_count=20_000, max_bounces=24, bands=8)
    got, want, gc, oc = _run_pair(ctx, oracle, sc, trace_kwargs(sc, bands=bands))
    assert np.count_nonzero(want) > 100
    assert np.array_equal(got, want)
    assert gc == oc


@pytest.mark.parametrize("n_segments", [5000, 10000, 16000])
def test_large_scene_staging_modes_bit_exact(ctx, oracle, n_segments):
    # 5000 walls: endpoint plane in shared memory, materials global; 10000: 160 KB plane, one CTA per SM;
    # 16000: does not fit in shared memory, read-only global path.
    sc = scenes.maze(n_segments=n_segments, ray_count=4096, max_bounces=8, bands=8)
    got, want, gc, oc = _run_pair(ctx, oracle, sc, trace_kwargs(sc, bands=1))
    assert np.array_equal(got, want)
    assert gc == oc


def test_hit_list_matches_oracle(ctx, oracle):
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc)
    ctx.set_walls(sc.walls)
    hits, keys, n = ctx.trace_hits(capi_params(_capi, kw), capacity=sc.ray_count * sc.max_bounces * 2 + 1024)
    r = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), want_hist=False, want_hits=True)
    assert n == r.n_hits == len(hits)
    hits, keys = sort_hits(hits, keys)
    assert np.array_equal(keys["ray"], r.hits["ray"]) and np.array_equal(keys["bounce"], r.hits["bounce"])
    assert np.array_equal(keys["kind"], r.hits["kind"])
    # tolerance of the north star: arrival time (distance / c) within 1e-5 relative
    rel = np.abs(hits["time_delay"].astype(np.float64) - r.hits["time_delay"]) / r.hits["time_delay"]
    assert rel.max() <= 1e-5
    # and, stronger, bit identity of every field
    assert np.array_equal(hits["time_delay"].view(np.uint32), r.hits["time_delay"].view(np.uint32))
    assert np.array_equal(hits["energy"].view(np.uint32), r.hits["energy"].view(np.uint32))
    assert np.array_equal(hits["hit_point"][:, 0].view(np.uint32), r.hits["hit_x"].view(np.uint32))
    assert np.array_equal(hits["hit_point"][:, 1].view(np.uint32), r.hits["hit_y"].view(np.uint32))


def test_ray_range_sharding_is_bit_identical(ctx, oracle):
    """Rays sharded by contiguous id range (the multi-GPU partition) sum to the unsharded histogram."""
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n, 1)
    ctx.trace(capi_params(_capi, kw), 0)
    whole = ctx.ir_read_fixed(0, n)
    ctx.ir_clear(1, n, 1)
    total = 15040
    for lo, hi in [(0, 3000), (3000, 3001), (3001, 9999), (9999, total)]:
        ctx.trace(capi_params(_capi, dict(kw, ray_begin=lo, ray_end=hi)), 1)
    assert np.array_equal(ctx.ir_read_fixed(1, n), whole)


@pytest.mark.parametrize("world,chunk_log2", [(2, 10), (3, 5), (8, 12), (8, 14)])
def test_block_cyclic_sharding_is_bit_identical(ctx, oracle, world, chunk_log2):
    """rar_trace_interleaved: every rank's share (chunks rank, rank + world, ... of 2^k thread ids) accumulated into one
    slot is the unsharded histogram; a dispatch that is not a multiple of the chunk, and more ranks than chunks."""
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc, ray_count=15000 if chunk_log2 < 14 else 5000)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n, 1)
    for r in range(world):
        ctx.trace_interleaved(capi_params(_capi, kw), 0, r, world, chunk_log2)
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw)).hist
    assert np.array_equal(ctx.ir_read_fixed(0, n), want)
    with pytest.raises(_capi.RarError):
        ctx.trace_interleaved(capi_params(_capi, dict(kw, ray_begin=0, ray_end=64)), 0, 0, world, chunk_log2)
    with pytest.raises(_capi.RarError):
        ctx.trace_interleaved(capi_params(_capi, kw), 0, world, world, chunk_log2)


def test_exact_ray_count_flag_and_dispatch_rounding(ctx, oracle):
    sc = scenes.smoll_room()
    for flags in (0, _capi.RAR_FLAG_EXACT_RAY_COUNT):
        kw = trace_kwargs(sc, ray_count=1001, flags=flags)
        got, want, gc, oc = _run_pair(ctx, oracle, sc, kw)
        assert np.array_equal(got, want)
        assert gc == oc


def test_edge_cases(ctx, oracle):
    sc = scenes.smoll_room()
    n = 4800
    # no walls at all: only direct listener crossings, rays end after one iteration
    empty = sc.walls[:0]
    ctx.set_walls(empty)
    ctx.ir_clear(0, n, 1)
    kw = trace_kwargs(sc, ray_count=4096, impulse_length=n, source=(0.0, 0.0), listener=(3.0, 0.0))
    ctx.trace(capi_params(_capi, kw), 0)
    r = oracle.trace(empty.view(oracle.SEGMENT_DTYPE), oracle_params(oracle, kw))
    assert np.array_equal(ctx.ir_read_fixed(0, n), r.hist) and np.count_nonzero(r.hist) > 0
    # zero bounces: nothing is deposited
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n, 1)
    ctx.trace(capi_params(_capi, trace_kwargs(sc, max_bounce_count=0, impulse_length=n)), 0)
    assert not ctx.ir_read_fixed(0, n).any()
    # impulse response shorter than the arrivals: late hits are dropped, not wrapped
    kw = trace_kwargs(sc, impulse_length=3000)
    got, want, _, _ = _run_pair(ctx, oracle, sc, kw)
    assert np.array_equal(got, want)
    # a single ray
    kw = trace_kwargs(sc, ray_count=1, flags=_capi.RAR_FLAG_EXACT_RAY_COUNT)
    got, want, _, _ = _run_pair(ctx, oracle, sc, kw)
    assert np.array_equal(got, want)


def test_error_behaviour(ctx):
    sc = scenes.smoll_room()
    ctx.set_walls(sc.walls)
    with pytest.raises(_capi.RarError):
        ctx.trace(capi_params(_capi, trace_kwargs(sc)), 9)          # slot never configured
    ctx.ir_clear(0, 100, 1)
    with pytest.raises(_capi.RarError):
        ctx.trace(capi_params(_capi, trace_kwargs(sc)), 0)          # impulse_length mismatch
    with pytest.raises(_capi.RarError):
        ctx.trace(capi_params(_capi, trace_kwargs(sc, bands=3, impulse_length=100)), 0)
    assert not ctx.ir_read_fixed(5, 16).any()                        # unconfigured slot reads as zeros


def test_debug_rays(ctx):
    sc = scenes.smoll_room()
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, sc.impulse_length, 1)
    ctx.trace(capi_params(_capi, trace_kwargs(sc, debug_ray_count=100)), 0)
    rays = ctx.get_debug_rays(100 * (sc.max_bounces + 1)).reshape(100, sc.max_bounces + 1, 4)
    assert np.allclose(rays[:, 0, 0], sc.source[0]) and np.allclose(rays[:, 0, 1], sc.source[1])
    assert np.all(rays[:, 0, 2] == np.float32(sc.input_gain))
    assert np.all(rays[:, 1, 2] == np.float32(sc.input_gain))       # energy before the first absorption


def test_debug_rays_hold_no_stale_vertices(ctx):
    """The kernel re-initialises the debug row of every ray it traces (no separate clear per frame): the buffer after a
    frame does not depend on the frames before it, and rows without a ray read as zero."""
    sc = scenes.smoll_room()
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, sc.impulse_length, 1)
    n4 = 100 * (sc.max_bounces + 1)

    def frame(f, **over):
        ctx.trace(capi_params(_capi, trace_kwargs(sc, debug_ray_count=100, rng_state_offset=f, **over)), 0)
        return ctx.get_debug_rays(n4)
    first = frame(5)
    other = frame(6)
    again = frame(5)
    assert np.array_equal(first, again) and not np.array_equal(first, other)
    few = frame(5, ray_count=64).reshape(100, sc.max_bounces + 1, 4)      # 64 rays, 100 rows: rows 64.. have no ray
    assert few[:64].any() and not few[64:].any()
    shard = frame(5, ray_begin=32, ray_end=64).reshape(100, sc.max_bounces + 1, 4)   # a shard clears the whole buffer first
    assert shard[32:64].any() and not shard[:32].any() and not shard[64:].any()


def test_full_size_shoebox_properties(ctx, oracle):
    """BASELINE config 2 at full size (1M rays x 32 bounces): size-independent checks.
    (a) linearity in input_gain by a power of two is exact in fixed point up to the truncation of each deposit;
    (b) the histogram is independent of how the ray range is split;
    (c) the first arrival is the direct path at (|S-L| - r)/c."""
    sc = scenes.shoebox()
    kw = trace_kwargs(sc)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n, 1)
    ctx.trace(capi_params(_capi, kw), 0)
    whole = ctx.ir_read_fixed(0, n)
    ctx.ir_clear(1, n, 1)
    cuts = [0, 1 << 18, 3 << 18, (1 << 20) - 7, 1 << 20]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ctx.trace(capi_params(_capi, dict(kw, ray_begin=lo, ray_end=hi)), 1)
    assert np.array_equal(ctx.ir_read_fixed(1, n), whole)
    first = int(np.flatnonzero(whole)[0])
    d = np.hypot(sc.listener[0] - sc.source[0], sc.listener[1] - sc.source[1]) - sc.listener_radius
    assert abs(first - d / sc.speed_of_sound * sc.sample_rate) <= 1.5
    # oracle on a 1/64 sample of the same dispatch
    sub = dict(kw, ray_begin=123_456, ray_end=123_456 + 16_384)
    ctx.ir_clear(1, n, 1)
    ctx.trace(capi_params(_capi, sub), 1)
    r = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, sub))
    assert np.array_equal(ctx.ir_read_fixed(1, n), r.hist)


def test_batched_listeners_config4_shape(ctx, oracle):
    """BASELINE config 4 at reduced size: one source, a grid of listeners, one IR slot per listener; every slot
    must equal the oracle's histogram for that listener."""
    sc = scenes.maze(n_segments=400, ray_count=4096, max_bounces=5, bands=8)
    kw = trace_kwargs(sc, bands=1, impulse_length=24000)
    gx, gy = np.meshgrid(np.linspace(20, 60, 3), np.linspace(25, 55, 2))
    listeners = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)
    ctx.set_walls(sc.walls)
    for l in range(len(listeners)):
        ctx.ir_clear(10 + l, 24000, 1)
    ctx.trace_listeners(capi_params(_capi, kw), listeners, 10)
    for l, (lx, ly) in enumerate(listeners):
        want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, dict(kw, listener=(float(lx), float(ly))))).hist
        assert np.array_equal(ctx.ir_read_fixed(10 + l, 24000), want), l


def test_executed_test_count_is_a_subset_and_result_is_unchanged(ctx, oracle):
    """RAR_FLAG_COUNT_EXECUTED counts what the production kernel evaluates: never more shadow tests than the
    reference performs, the same nearest tests, and the same histogram."""
    sc = scenes.shoebox(ray_count=65536, max_bounces=32)
    kw = trace_kwargs(sc)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    res = {}
    for name, flags in (("ref", _capi.RAR_FLAG_COUNT_TESTS), ("exe", _capi.RAR_FLAG_COUNT_TESTS | _capi.RAR_FLAG_COUNT_EXECUTED), ("prod", 0)):
        ctx.ir_clear(0, n, 1)
        ctx.get_counters(reset=True)
        ctx.trace(capi_params(_capi, dict(kw, flags=flags)), 0)
        res[name] = (ctx.ir_read_fixed(0, n), ctx.get_counters())
    assert np.array_equal(res["ref"][0], res["exe"][0]) and np.array_equal(res["ref"][0], res["prod"][0])
    r, e = res["ref"][1], res["exe"][1]
    assert e["nearest_tests"] == r["nearest_tests"] and e["nee_hits"] == r["nee_hits"] and e["direct_hits"] == r["direct_hits"]
    assert 0 < e["shadow_tests"] < r["shadow_tests"]
    assert res["prod"][1]["nearest_tests"] == 0                     # no counting without the flag


def test_fused_listener_kernel_equals_single_listener_traces(ctx, oracle):
    """The fused batched-listener kernel against one rar_trace per listener (which the oracle tests pin):
    bundled room with transmission and scattering (wall_depth > 0 paths), and a 3000-wall scene
    (warp-cooperative shadow rays), 40 listeners each."""
    rng = np.random.default_rng(4)
    for sc, kw, box in ((scenes.smoll_room(), None, (-19, 19, -4, 9)),
                        (scenes.maze(n_segments=3000, ray_count=2048, max_bounces=6, bands=8), None, (5, 95, 5, 95))):
        kw = trace_kwargs(sc, ray_count=min(sc.ray_count, 4096), impulse_length=24000)
        listeners = np.stack([rng.uniform(box[0], box[1], 40), rng.uniform(box[2], box[3], 40)], 1).astype(np.float32)
        ctx.set_walls(sc.walls)
        for l in range(40):
            ctx.ir_clear(100 + l, 24000, 1)
            ctx.ir_clear(200 + l, 24000, 1)
        ctx.trace_listeners(capi_params(_capi, kw), listeners, 100)
        for l, (lx, ly) in enumerate(listeners):
            ctx.trace(capi_params(_capi, dict(kw, listener=(float(lx), float(ly)))), 200 + l)
        total = 0
        for l in range(40):
            fused, single = ctx.ir_read_fixed(100 + l, 24000), ctx.ir_read_fixed(200 + l, 24000)
            assert np.array_equal(fused, single), l
            total += np.count_nonzero(single)
        assert total > 1000
        # oracle spot check on one listener
        lx, ly = map(float, listeners[7])
        want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, dict(kw, listener=(lx, ly)))).hist
        assert np.array_equal(ctx.ir_read_fixed(107, 24000), want)


@pytest.mark.parametrize("bands", [1, 8])
def test_grid_mode_is_bit_identical_to_brute_force(ctx, oracle, bands):
    """RAR_FLAG_USE_GRID changes which walls are looked at, never the result: same histogram as the brute-force
    kernel and as the oracle, for the broadband and the banded kernel."""
    sc = scenes.maze(n_segments=10000, ray_count=30_000, max_bounces=24, bands=8)
    kw = trace_kwargs(sc, bands=bands)
    n = kw["impulse_length"] * bands
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(0, kw["impulse_length"], bands)
    ctx.trace(capi_params(_capi, kw), 0)
    brute = ctx.ir_read_fixed(0, n)
    ctx.ir_clear(1, kw["impulse_length"], bands)
    ctx.get_counters(reset=True)
    ctx.trace(capi_params(_capi, dict(kw, flags=_capi.RAR_FLAG_USE_GRID | _capi.RAR_FLAG_COUNT_TESTS)), 1)
    grid = ctx.ir_read_fixed(1, n)
    c = ctx.get_counters()
    assert np.array_equal(grid, brute) and np.count_nonzero(brute) > 100
    assert 0 < c["nearest_tests"] < c["ray_bounces"] * 200          # a few walls per query instead of 10 000
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), band_abs=sc.band_absorption if bands > 1 else None).hist
    assert np.array_equal(grid, want)


def test_grid_mode_bundled_room_listeners_and_geometry_update(ctx, oracle):
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc, flags=_capi.RAR_FLAG_USE_GRID)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n, 1)
    ctx.trace(capi_params(_capi, kw), 0)
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, trace_kwargs(sc))).hist
    assert np.array_equal(ctx.ir_read_fixed(0, n), want)
    # new geometry invalidates the cached grid
    sc2 = scenes.big_room()
    ctx.set_walls(sc2.walls)
    kw2 = trace_kwargs(sc2, flags=_capi.RAR_FLAG_USE_GRID)
    ctx.ir_clear(0, n, 1)
    ctx.trace(capi_params(_capi, kw2), 0)
    want = oracle.trace(oracle_walls(oracle, sc2.walls), oracle_params(oracle, trace_kwargs(sc2))).hist
    assert np.array_equal(ctx.ir_read_fixed(0, n), want)
    # fused listeners through the grid
    rng = np.random.default_rng(3)
    listeners = np.stack([rng.uniform(-150, 150, 12), rng.uniform(-40, 80, 12)], 1).astype(np.float32)
    for l in range(12):
        ctx.ir_clear(20 + l, n, 1)
    ctx.trace_listeners(capi_params(_capi, kw2), listeners, 20)
    for l in (0, 5, 11):
        lx, ly = map(float, listeners[l])
        want = oracle.trace(oracle_walls(oracle, sc2.walls), oracle_params(oracle, dict(trace_kwargs(sc2), listener=(lx, ly)))).hist
        assert np.array_equal(ctx.ir_read_fixed(20 + l, n), want), l


def _soup(seed, n, extent=50.0):
    from realisticaudioraytracing2d_b200.host.scene_helper import SEGMENT_DTYPE
    rng = np.random.default_rng(seed)
    a = rng.uniform(0, extent, (n, 2)).astype(np.float32)
    length = np.where(rng.random(n) < 0.1, rng.uniform(5, 40, n), rng.uniform(0.01, 2.0, n))
    ang = rng.uniform(0, 2 * np.pi, n)
    b = (a + np.stack([np.cos(ang), np.sin(ang)], 1) * length[:, None]).astype(np.float32)
    walls = np.zeros(n, dtype=SEGMENT_DTYPE)
    walls["start"], walls["end"] = a, b
    d = b - a
    walls["normal"] = (np.stack([d[:, 1], -d[:, 0]], 1) / np.maximum(np.hypot(d[:, 0], d[:, 1]), 1e-9)[:, None]).astype(np.float32)
    walls["absorption"] = rng.uniform(0.02, 0.4, n)
    walls["scattering"] = rng.uniform(0, 1, n) * (rng.random(n) < 0.7)
    walls["transmission"] = rng.uniform(0, 0.6, n) * (rng.random(n) < 0.5)
    walls["ior"] = rng.uniform(0.3, 2.0, n)
    return walls


@pytest.mark.parametrize("seed,n", [(1, 37), (2, 700), (3, 3000)])
def test_random_segment_soup_brute_and_grid(ctx, oracle, seed, n):
    """Crossing, overlapping, tiny and long walls with per-wall random materials (refraction, jitter, diffuse
    reflection, nested media): brute force, the grid and the oracle must agree bit for bit."""
    sc = scenes.smoll_room()
    sc.walls, sc.source, sc.listener = _soup(seed, n), (25.0, 25.0), (25.9, 25.4)
    kw = trace_kwargs(sc, ray_count=20_000, max_bounce_count=16, impulse_length=48000)
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw))
    ctx.set_walls(sc.walls)
    for flags in (_capi.RAR_FLAG_COUNT_TESTS, 0, _capi.RAR_FLAG_USE_GRID):
        ctx.ir_clear(0, 48000, 1)
        ctx.get_counters(reset=True)
        ctx.trace(capi_params(_capi, dict(kw, flags=flags)), 0)
        assert np.array_equal(ctx.ir_read_fixed(0, 48000), want.hist), flags
        if flags == _capi.RAR_FLAG_COUNT_TESTS:
            assert ctx.get_counters() == want.counters
    assert np.count_nonzero(want.hist) >= 1 and want.counters["ray_bounces"] > 20_000


def test_degenerate_inputs_match_the_oracle(ctx, oracle):
    """Zero-length and duplicated walls, the source on a wall endpoint, a listener inside a wall, a tiny listener,
    huge gain (fixed-point saturation path), slow sound (late arrivals dropped)."""
    sc = scenes.smoll_room()
    w = np.concatenate([sc.walls, sc.walls[:3], sc.walls[5:6]])
    w["end"][20] = w["start"][20]                                  # zero-length wall
    cases = [
        dict(source=tuple(map(float, sc.walls["start"][0])), listener=sc.listener),       # source on an endpoint
        dict(source=sc.source, listener=(0.0, 10.0)),                                     # listener inside the top wall
        dict(source=sc.source, listener=sc.listener, listener_radius=1e-3),
        dict(source=sc.source, listener=sc.listener, input_gain=1e9),
        dict(source=sc.source, listener=sc.listener, speed_of_sound=3.0),
        dict(source=(5.0, 3.0), listener=(5.0, 3.0)),                                     # listener at the source
    ]
    ctx.set_walls(w)
    for i, over in enumerate(cases):
        kw = trace_kwargs(sc, ray_count=4096, **over)
        want = oracle.trace(oracle_walls(oracle, w), oracle_params(oracle, kw)).hist
        for flags in (0, _capi.RAR_FLAG_USE_GRID):
            ctx.ir_clear(0, kw["impulse_length"], 1)
            ctx.trace(capi_params(_capi, dict(kw, flags=flags)), 0)
            assert np.array_equal(ctx.ir_read_fixed(0, kw["impulse_length"]), want), (i, flags)


def test_banded_layout_with_time_divisor(ctx, oracle):
    """The banded deposit of RaytraceOcclusion2D.compute:241-248: bin = (int)(t*SampleRate/WindowSize), slot layout
    IR[bin*WindowSize + band], with WindowSize = 8 bands."""
    sc = scenes.maze(n_segments=800, ray_count=20_000, max_bounces=12, bands=8)
    kw = trace_kwargs(sc, bands=8, time_divisor=8.0, impulse_length=6000)
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(0, 6000, 8)
    ctx.trace(capi_params(_capi, kw), 0)
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), band_abs=sc.band_absorption).hist
    got = ctx.ir_read_fixed(0, 48000)
    assert np.array_equal(got, want) and np.count_nonzero(want) > 1000
    # band b of a bin never exceeds the broadband-equivalent ordering: lower absorption bands carry more energy
    e = got.reshape(6000, 8).sum(0)
    assert (e > 0).all()


def test_trace_frames_equals_successive_frames(ctx, oracle):
    """rar_trace_frames: n frames of the dispatch in one launch == n launches with successive rngStateOffset."""
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc, rng_state_offset=40, debug_ray_count=100)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n, 1)
    ctx.trace_frames(capi_params(_capi, kw), 0, 7)
    hist = np.zeros(n, np.int64)
    for f in range(40, 47):
        oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, dict(kw, rng_state_offset=f)), hist=hist)
    assert np.array_equal(ctx.ir_read_fixed(0, n), hist)
    ctx.ir_clear(0, n, 1)
    ctx.trace_frames(capi_params(_capi, kw), 0, 0)                  # zero frames: nothing
    assert not ctx.ir_read_fixed(0, n).any()
    with pytest.raises(_capi.RarError):
        ctx.trace_frames(capi_params(_capi, kw), 0, -1)


@pytest.mark.parametrize("bands", [2, 3, 20, 128])
def test_arbitrary_band_counts_are_traced_in_chunks_of_eight(ctx, oracle, bands):
    """Slots of any band count up to 128 (the experimental variant's WindowSize default): chunks of 8 bands, the
    same rays per chunk; must equal the oracle's single pass."""
    sc = scenes.maze(n_segments=300, ray_count=6000, max_bounces=10, bands=bands, seed=5)
    kw = trace_kwargs(sc, bands=bands, impulse_length=6000, flags=_capi.RAR_FLAG_COUNT_TESTS)
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(0, 6000, bands)
    ctx.get_counters(reset=True)
    ctx.trace(capi_params(_capi, kw), 0)
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), band_abs=sc.band_absorption)
    assert np.array_equal(ctx.ir_read_fixed(0, 6000 * bands), want.hist) and np.count_nonzero(want.hist) > 100
    assert ctx.get_counters() == want.counters                      # counted once, not once per chunk
    ctx.ir_clear(0, 6000, bands)
    ctx.trace(capi_params(_capi, dict(kw, flags=_capi.RAR_FLAG_USE_GRID)), 0)
    assert np.array_equal(ctx.ir_read_fixed(0, 6000 * bands), want.hist)


@pytest.mark.parametrize("case", ["shoebox_diffuse", "maze600", "maze600_b8", "maze600_grid", "maze120", "maze120_b8",
                                  "shoebox_b8"])
def test_production_kernels_for_opaque_scenes(ctx, oracle, case):
    """Scenes without any transmitting wall run kernels compiled without the transmit/refract branch (production
    mode only: no counters).  They must reproduce the oracle, which always evaluates `rngVal < transmission`.
    shoebox*: the four-wall, range-checked-once instantiation (FAST 3); maze120*: range-checked-once, general wall
    count (FAST 1); maze600*: cooperative / grid kernels."""
    if case == "shoebox_diffuse":
        sc, bands, flags = scenes.shoebox(ray_count=100_000, max_bounces=24, scattering=0.4), 1, 0
    elif case == "shoebox_b8":
        sc, bands, flags = scenes.shoebox(ray_count=50_000, max_bounces=24, scattering=0.2), 8, 0
        sc.band_absorption = np.random.default_rng(5).uniform(0.02, 0.3, size=(4, 8)).astype(np.float32)
    elif case.startswith("maze120"):
        sc = scenes.maze(n_segments=120, ray_count=30_000, max_bounces=24, bands=8, seed=3)
        bands, flags = (8 if case.endswith("_b8") else 1), 0
    else:
        sc = scenes.maze(n_segments=600, ray_count=30_000, max_bounces=16, bands=8, seed=11)
        bands = 8 if case == "maze600_b8" else 1
        flags = _capi.RAR_FLAG_USE_GRID if case == "maze600_grid" else 0
    assert not (sc.walls["transmission"] > 0).any()
    kw = trace_kwargs(sc, bands=bands, flags=flags)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    if bands > 1:
        ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(0, n, bands)
    ctx.trace(capi_params(_capi, kw), 0)
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, dict(kw, flags=0)),
                        band_abs=sc.band_absorption if bands > 1 else None).hist
    assert np.array_equal(ctx.ir_read_fixed(0, n * bands), want) and np.count_nonzero(want) > 500


def test_arithmetic_selftest(ctx):
    """The range-checked-once forms of 1/x, sqrt(x), a/b (csrc/rar_math.cuh) equal the correctly rounded intrinsics
    bit for bit over the admitted operand ranges (2^27 operand sets per seed)."""
    for seed in (1, 2):
        assert ctx.selftest_arithmetic(1 << 27, seed) == [0, 0, 0, 0, 0]


def test_fast_and_guarded_kernels_agree(ctx, oracle, monkeypatch):
    """RAR_NO_FAST=1 selects the kernels with a guard around every division / square root; same histogram."""
    for sc in (scenes.shoebox(ray_count=300_000, max_bounces=32), scenes.maze(n_segments=90, ray_count=50_000, max_bounces=20, bands=1, seed=9)):
        kw = trace_kwargs(sc)
        n = kw["impulse_length"]
        ctx.set_walls(sc.walls)
        out = []
        for no_fast in ("0", "1"):
            monkeypatch.setenv("RAR_NO_FAST", no_fast)
            ctx.ir_clear(0, n, 1)
            ctx.trace(capi_params(_capi, kw), 0)
            out.append(ctx.ir_read_fixed(0, n))
        monkeypatch.delenv("RAR_NO_FAST")
        assert np.array_equal(out[0], out[1]) and np.count_nonzero(out[0]) > 1000


def test_degenerate_inputs_opaque_scene(ctx, oracle):
    """The degenerate placements of test_degenerate_inputs_match_the_oracle in an OPAQUE four-wall room, i.e. through
    the range-checked-once kernels: a zero distance to the listener, a source on a corner, out-of-range speed and
    coordinates (which must select the guarded kernels), a huge gain."""
    base = scenes.shoebox(ray_count=4096, max_bounces=16, scattering=0.3)
    far = scenes.shoebox(width=3e9, height=2e9, ray_count=4096, max_bounces=8)     # coordinates beyond 2^30
    cases = [
        (base, dict(listener=(10.0, 3.0))),                   # listener centre ON a wall: a hit point can coincide with it
        (base, dict(source=(0.0, 0.0))),                      # source on a corner
        (base, dict(source=(5.0, 3.0), listener=(5.0, 3.0))),
        (base, dict(listener_radius=1e-3)),
        (base, dict(input_gain=1e9)),
        (base, dict(speed_of_sound=1e-7)),                    # below the range the reciprocal trick admits
        (base, dict(speed_of_sound=3e6)),
        (far, dict(source=(1e9, 1e9), listener=(2e9, 1.5e9), speed_of_sound=3e8)),
    ]
    for i, (sc, over) in enumerate(cases):
        kw = trace_kwargs(sc, **over)
        want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw)).hist
        ctx.set_walls(sc.walls)
        ctx.ir_clear(0, kw["impulse_length"], 1)
        ctx.trace(capi_params(_capi, kw), 0)
        assert np.array_equal(ctx.ir_read_fixed(0, kw["impulse_length"]), want), i


def test_asynchronous_ir_readback(ctx, oracle):
    """rar_ir_read_begin/_end (AsyncGPUReadback analogue): same values as the blocking read, several in flight."""
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc, ray_count=2000)
    n = kw["impulse_length"]
    ctx.set_walls(sc.walls)
    tickets = []
    for s in (0, 1):
        ctx.ir_clear(s, n, 1)
        ctx.trace(capi_params(_capi, dict(kw, rng_state_offset=1 + s)), s)
        tickets.append(ctx.ir_read_begin(s, n + 5))         # longer than the slot: zero tail
    assert tickets[0] != tickets[1]
    got = [ctx.ir_read_end(t, n + 5) for t in tickets]
    for s in (0, 1):
        assert np.array_equal(got[s][:n], ctx.ir_read(s, n)) and not got[s][n:].any() and got[s].any()
    assert not np.array_equal(got[0], got[1])
    t = ctx.ir_read_begin(77, 16)                           # never configured: reads as zeros
    while not ctx.poll(t):
        pass
    assert not ctx.ir_read_end(t, 16).any()
    with pytest.raises(_capi.RarError):
        ctx.ir_read_end(t, 16)                              # the ticket was released


def _adversarial_walls(n, seed):
    """A maze plus walls chosen to hit the zeros of the filter: axis-parallel walls through the source's coordinates
    (num1 = 0 or num2 = 0 for whole fans of rays, dotP = 0 for the axis-parallel rays), zero-length walls, duplicates,
    and an odd total (the last pair record holds one wall and a NaN)."""
    sc = scenes.maze(n_segments=n, ray_count=60_000, max_bounces=12, bands=8, seed=seed)
    walls = sc.walls.copy()
    sx, sy = sc.source
    k = 0
    for (a, b) in (((sx - 7.0, sy), (sx - 3.0, sy)), ((sx + 2.0, sy), (sx + 9.0, sy)),      # on the source's horizontal line
                   ((sx, sy + 1.5), (sx, sy + 6.0)), ((sx, sy - 8.0), (sx, sy - 2.5)),      # ... vertical line
                   ((sx + 4.0, sy + 4.0), (sx + 4.0, sy + 4.0)),                            # zero length
                   ((sx - 5.0, sy - 5.0), (sx - 1.0, sy - 1.0))):                           # on a diagonal through the source
        w = walls[k]
        w["start"], w["end"] = a, b
        d = np.array(b, np.float32) - np.array(a, np.float32)
        ln = float(np.hypot(*d))
        w["normal"] = (-d[1] / ln, d[0] / ln) if ln > 0 else (0.0, 1.0)
        k += 37
    walls[5] = walls[4]                                                                      # a duplicate
    return sc, walls


@pytest.mark.parametrize("n,bands", [(301, 1), (1000, 8), (4001, 1)])
def test_packed_and_scalar_wall_scans_agree(ctx, oracle, monkeypatch, n, bands):
    """FAST bit 2 (packed FP32, two walls per instruction) against RAR_NO_PACKED=1 (scalar) and the oracle, on geometry
    that sits on the zeros of the filter quantities -- where the packed forms carry some zeros with the other sign."""
    sc, walls = _adversarial_walls(n, n)
    kw = trace_kwargs(sc, bands=bands)
    m = kw["impulse_length"]
    ctx.set_walls(walls)
    if bands > 1:
        ctx.set_wall_band_absorption(sc.band_absorption)
    out = []
    for no_packed in ("0", "1"):
        monkeypatch.setenv("RAR_NO_PACKED", no_packed)
        ctx.ir_clear(0, m, bands)
        ctx.trace(capi_params(_capi, kw), 0)
        out.append(ctx.ir_read_fixed(0, m * bands))
    monkeypatch.delenv("RAR_NO_PACKED")
    want = oracle.trace(oracle_walls(oracle, walls), oracle_params(oracle, kw), band_abs=sc.band_absorption if bands > 1 else None).hist
    assert np.array_equal(out[0], out[1])
    assert np.array_equal(out[0], want) and np.count_nonzero(want) > 500


def test_packed_four_wall_kernel_on_the_zeros_of_the_filter(ctx, oracle, monkeypatch):
    """The FAST = 7 kernel (config 2's) with the source on the room's axes of symmetry and on a wall's line: rays
    parallel to walls (dotP = 0), num1 = 0 and num2 = 0 for the first bounce of whole fans of rays."""
    for src, lst in (((5.0, 3.0), (7.5, 3.0)), ((5.0, 0.0), (2.0, 3.0)), ((0.0, 3.0), (9.0, 5.0)), ((2.5, 1.5), (2.5, 4.5))):
        sc = scenes.shoebox(ray_count=65536, max_bounces=24)
        kw = trace_kwargs(sc, source=src, listener=lst)
        m = kw["impulse_length"]
        ctx.set_walls(sc.walls)
        out = []
        for no_packed in ("0", "1"):
            monkeypatch.setenv("RAR_NO_PACKED", no_packed)
            ctx.ir_clear(0, m, 1)
            ctx.trace(capi_params(_capi, kw), 0)
            out.append(ctx.ir_read_fixed(0, m))
        monkeypatch.delenv("RAR_NO_PACKED")
        want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw)).hist
        assert np.array_equal(out[0], out[1]), (src, lst)
        assert np.array_equal(out[0], want), (src, lst)
