"""GPU parity of the banded model's convolution: a banded slot -- IR[bin*bands + band],
RaytraceOcclusion2D.compute:241-248 -- is convolved through its filter-bank synthesis.  The GPU does the synthesis
as overlap-add in the frequency domain; the oracle restates it in direct form (double accumulation).
Bar (north star): convolved audio <= 1e-4 relative L2; the synthesised response is held to the same bar."""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from tests.common import capi_params, oracle_params, oracle_walls, rel_l2, trace_kwargs

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _trace_banded(ctx, O, sc, slot, bands, **over):
    kw = trace_kwargs(sc, bands=bands, **over)
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption[:, :bands] if sc.band_absorption.shape[1] != bands else sc.band_absorption)
    ctx.ir_clear(slot, kw["impulse_length"], bands)
    ctx.trace(capi_params(_capi, kw), slot)
    hist = ctx.ir_read_fixed(slot, kw["impulse_length"] * bands)
    return kw, hist


def test_config3_geometry_eight_bands_synthesis_and_convolution(ctx, oracle):
    """BASELINE config 3 geometry (10 000-wall maze, 8 absorption bands) at a reduced ray count."""
    sc = scenes.maze(n_segments=10000, ray_count=60_000, max_bounces=24, bands=8)
    kw, hist = _trace_banded(ctx, oracle, sc, 0, 8)
    n = kw["impulse_length"]
    assert np.count_nonzero(hist) > 2000
    want_ir = oracle.synthesize_ir(hist, n, 8)
    got_ir = ctx.synthesize_ir(0, n)
    assert rel_l2(got_ir, want_ir) <= TOL
    clip = scenes.synthetic_clip(9600)
    got = ctx.convolve(0, clip, 3, n)
    want = oracle.convolve(clip, want_ir, 3)
    assert len(got) == len(clip) + n and got[-1] == 0.0
    assert rel_l2(got, want) <= TOL
    # the bands do differ: the synthesis is not just one band's response
    assert rel_l2(oracle.ir_to_float(hist.reshape(n, 8)[:, 0].copy()), want_ir) > 0.005


def test_equal_bands_synthesise_to_their_common_response(ctx, oracle):
    """The band filters sum to a unit impulse: a slot whose bands all hold h synthesises to h."""
    rng = np.random.default_rng(3)
    n, bands = 5000, 8
    h = (rng.random(n) * (rng.random(n) < 0.2) * 1e-3).astype(np.float32)
    ctx.ir_write(0, np.repeat(h[:, None], bands, 1).ravel(), bands=bands)
    got = ctx.synthesize_ir(0, n)
    want = ctx.ir_read(0, n * bands).reshape(n, bands)[:, 0]
    assert rel_l2(got, want) <= 2e-6


@pytest.mark.parametrize("n,bands", [(100, 8), (256, 4), (257, 8), (2047, 12), (2048, 8), (2049, 4), (5000, 20)])
def test_register_transform_synthesis_shapes(ctx, oracle, n, bands):
    """band_synth.cu (bands a multiple of 4): segment tails, responses shorter than a segment, partially filled
    CTAs (eight segments each) and chunk counts other than two."""
    rng = np.random.default_rng(n + bands)
    ir = (rng.random((n, bands)) * (rng.random((n, bands)) < 0.3) * 2e-3).astype(np.float32)
    ir[0] = 1e-3
    ir[-1] = 3e-3
    ctx.ir_write(58, ir.ravel(), bands=bands)
    hist = ctx.ir_read_fixed(58, n * bands)
    want = oracle.synthesize_ir(hist, n, bands)
    assert rel_l2(ctx.synthesize_ir(58, n), want) <= 3e-6


@pytest.mark.parametrize("bands,edges_hz", [(3, None), (8, [0, 88, 177, 354, 707, 1414, 2828, 5657, 24000]), (128, None)])
def test_band_counts_and_custom_edges(ctx, oracle, bands, edges_hz):
    """Other band counts (WindowSize = 128 is the experimental variant's default, RayTraceManagerComplex.cs:27) and
    octave-style band edges instead of the default equal-width bands."""
    rng = np.random.default_rng(bands)
    n = 3000
    ir = (rng.random((n, bands)) * (rng.random((n, bands)) < 0.05) * 2e-3).astype(np.float32)
    ctx.ir_write(1, ir.ravel(), bands=bands)
    hist = ctx.ir_read_fixed(1, n * bands)
    ctx.set_band_edges(edges_hz, 48000)
    try:
        got = ctx.synthesize_ir(1, n)
        edges = None if edges_hz is None else np.asarray(edges_hz, np.float32) / np.float32(24000)
        want = oracle.synthesize_ir(hist, n, bands, 1, edges)
        assert rel_l2(got, want) <= TOL
        x = scenes.synthetic_clip(1500, seed=bands)
        assert rel_l2(ctx.convolve(1, x, 2, n), oracle.convolve(x, want, 2)) <= TOL
    finally:
        ctx.set_band_edges(None, 48000)


def test_coarse_time_bins_of_the_reference_layout(ctx, oracle):
    """time_divisor = W > 1 (RaytraceOcclusion2D.compute:241-243: bin = (int)(t * SampleRate / WindowSize)): bin k stands
    on sample k*W, the response has impulse_length * W samples."""
    W, bands = 4, 4
    sc = scenes.maze(n_segments=300, ray_count=20_000, max_bounces=12, bands=8, seed=5)
    kw, hist = _trace_banded(ctx, oracle, sc, 2, bands, time_divisor=float(W), impulse_length=12000)
    r = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), band_abs=sc.band_absorption[:, :bands])
    assert np.array_equal(hist, r.hist) and np.count_nonzero(hist) > 500
    n = 12000 * W
    want_ir = oracle.synthesize_ir(hist, 12000, bands, W)
    assert rel_l2(ctx.synthesize_ir(2, n), want_ir) <= TOL
    x = scenes.synthetic_clip(2000, seed=1)
    got = ctx.convolve(2, x, 1, n)                       # ir_len = impulse_length * W
    assert len(got) == 2000 + n
    assert rel_l2(got, oracle.convolve(x, want_ir, 1)) <= TOL
    # a non-integer divisor leaves no sample grid to synthesise on
    kw2, _ = _trace_banded(ctx, oracle, sc, 3, bands, time_divisor=2.5, impulse_length=12000)
    with pytest.raises(_capi.RarError) as e:
        ctx.convolve(3, x, 1, 12000)
    assert e.value.code == -5


def test_streaming_convolver_takes_a_banded_slot(ctx, oracle):
    sc = scenes.maze(n_segments=500, ray_count=30_000, max_bounces=12, bands=8, seed=2)
    kw, hist = _trace_banded(ctx, oracle, sc, 4, 8, impulse_length=6000)
    n = 6000
    want_ir = oracle.synthesize_ir(hist, n, 8) / np.float32(5)
    cv = _capi.Convolver(ctx, 2, 256, n)
    try:
        cv.set_ir_from_slot(1, 4, 5)
        cv.set_ir(0, np.zeros(n, np.float32))
        rng = np.random.default_rng(0)
        x = rng.uniform(-1, 1, (2, 256 * 30)).astype(np.float32)
        y = np.concatenate([cv.process(x[:, k * 256:(k + 1) * 256]) for k in range(30)], axis=1)
        want = oracle.convolve(x[1], want_ir, 1)[: 256 * 30]
        assert rel_l2(y[1], want) <= TOL and not y[0].any()
    finally:
        cv.destroy()


def test_batched_responses_from_slots(ctx, oracle):
    """rar_conv_set_irs_from_slots: 40 streams (more than one launch group of 32) from banded and from broadband slots
    in single calls; every stream answers a unit impulse with the first block of its slot's response."""
    rng = np.random.default_rng(8)
    n, S = 3000, 40
    for bands in (8, 1):
        for k in range(S):
            ir = (rng.random((n, bands)) * (rng.random((n, bands)) < 0.05) * 1e-3).astype(np.float32)
            ctx.ir_write(400 + k, ir.ravel(), bands=bands)
        cv = _capi.Convolver(ctx, S + 2, 256, n)
        try:
            accum = np.arange(1, S + 1, dtype=np.int32)
            cv.set_irs_from_slots(1, np.arange(400, 400 + S), accum)
            x = np.zeros((S + 2, 256), np.float32)
            x[:, 0] = 1.0
            y = cv.process(x)
            assert not y[0].any() and not y[S + 1].any()          # streams outside the range keep their (zero) response
            for k in range(S):
                want = ctx.synthesize_ir(400 + k, 256) / np.float32(accum[k])
                assert rel_l2(y[1 + k], want) <= 2e-5, (bands, k)
        finally:
            cv.destroy()


@pytest.mark.parametrize("case", ["maze300_b8", "maze10000_b8", "shoebox_b8", "maze_b20", "grid"])
def test_air_absorption_bit_exact(ctx, oracle, case):
    """rar_set_air_absorption: every band arrival scaled by exp(-alpha_b * path length), a fixed polynomial kernel of
    the arithmetic contract: histograms bit-exact against the oracle in the per-thread, cooperative, four-wall
    (range-checked-once), chunked (20 bands) and grid kernels; switching it off restores the plain banded trace."""
    bands, flags = 8, 0
    if case == "maze300_b8":
        sc = scenes.maze(n_segments=300, ray_count=20_000, max_bounces=12, bands=8, seed=5)
    elif case == "maze10000_b8":
        sc = scenes.maze(n_segments=10000, ray_count=20_000, max_bounces=16, bands=8)
    elif case == "shoebox_b8":
        sc = scenes.shoebox(ray_count=50_000, max_bounces=24, scattering=0.2)
        sc.band_absorption = np.random.default_rng(5).uniform(0.02, 0.3, size=(4, 8)).astype(np.float32)
    elif case == "maze_b20":
        sc, bands = scenes.maze(n_segments=600, ray_count=10_000, max_bounces=10, bands=20, seed=3), 20
    else:
        sc, flags = scenes.maze(n_segments=2000, ray_count=20_000, max_bounces=12, bands=8, seed=2), _capi.RAR_FLAG_USE_GRID
    air = np.geomspace(1e-4, 0.08, bands).astype(np.float32)
    air[0] = 0.0
    kw = trace_kwargs(sc, bands=bands, impulse_length=24000, flags=flags)
    n = 24000 * bands
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption)
    O = oracle
    try:
        ctx.set_air_absorption(air)
        ctx.ir_clear(6, 24000, bands)
        ctx.trace(capi_params(_capi, kw), 6)
        want = O.trace(oracle_walls(O, sc.walls), oracle_params(O, dict(kw, flags=0)), band_abs=sc.band_absorption, air=air).hist
        assert np.array_equal(ctx.ir_read_fixed(6, n), want) and np.count_nonzero(want) > 1000
        with pytest.raises(_capi.RarError):                       # the table is for `bands` bands
            ctx.set_air_absorption(air[:-1])
            ctx.ir_clear(7, 24000, bands)
            ctx.trace(capi_params(_capi, kw), 7)
    finally:
        ctx.set_air_absorption(None)
    ctx.ir_clear(6, 24000, bands)
    ctx.trace(capi_params(_capi, kw), 6)
    plain = O.trace(oracle_walls(O, sc.walls), oracle_params(O, dict(kw, flags=0)), band_abs=sc.band_absorption).hist
    assert np.array_equal(ctx.ir_read_fixed(6, n), plain) and not np.array_equal(plain, want)


def test_band_edge_errors(ctx):
    with pytest.raises(_capi.RarError):
        ctx.set_band_edges([0, 5000, 4000, 24000], 48000)          # not ascending
    with pytest.raises(_capi.RarError):
        ctx.set_band_edges([100, 5000, 24000], 48000)              # does not start at 0
    ctx.set_band_edges([0, 1000, 24000], 48000)                    # two bands ...
    try:
        ctx.ir_write(57, np.ones(3 * 100, np.float32) * 1e-3, bands=3)
        with pytest.raises(_capi.RarError) as e:                   # ... but the slot has three
            ctx.synthesize_ir(57, 100)
        assert e.value.code == -3
    finally:
        ctx.set_band_edges(None, 48000)
