"""CPU checks of the product's device headers compiled for the host (tests/host_emulation.cpp) against
the independent oracle: the conservative intersection filter, the batched inner loops, the lazy /
specular shortcuts and the fixed-point deposit must reproduce the oracle bit for bit; the hand-written
FFT and the partitioned overlap-save pipeline must stay within the 1e-4 relative-L2 bar.

These run without a GPU and are NOT a product path: they exist so that logic errors are caught here
before GPU time is spent.  The GPU parity tests proper are tests/test_gpu_*.py.
"""
import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import scenes
from tests import emulation
from tests.common import oracle_params, oracle_walls, rel_l2, trace_kwargs


def _check(O, sc, kw):
    P = oracle_params(O, kw)
    ba = sc.band_absorption if kw["bands"] > 1 else None
    r = O.trace(oracle_walls(O, sc.walls), P, band_abs=ba, want_hits=True)
    hist, hits, ctr = emulation.trace(O, sc.walls, P, ba)
    assert np.array_equal(hist, r.hist)
    assert len(hits) == len(r.hits) and hits.tobytes() == r.hits.tobytes()
    assert ctr == r.counters
    # the production instantiation (no counters: shadow rays below the deposit threshold are skipped)
    hist2, hits2, _ = emulation.trace(O, sc.walls, P, ba, counting=False)
    assert np.array_equal(hist2, r.hist) and hits2.tobytes() == r.hits.tobytes()
    return r


@pytest.mark.parametrize("frame", [1, 2, 77])
def test_smoll_room(oracle, frame):
    sc = scenes.smoll_room()
    r = _check(oracle, sc, trace_kwargs(sc, rng_state_offset=frame))
    assert r.counters["ray_bounces"] > 75000


def test_big_room_ten_bounces(oracle):
    sc = scenes.big_room()
    _check(oracle, sc, trace_kwargs(sc, max_bounce_count=10))


@pytest.mark.parametrize("scattering,transmission,ior", [(0.0, 0.0, 1.0), (0.3, 0.0, 1.0), (0.0, 0.4, 1.4), (0.7, 0.5, 0.6)])
def test_shoebox_materials(oracle, scattering, transmission, ior):
    sc = scenes.shoebox(ray_count=20000, max_bounces=32, scattering=scattering, transmission=transmission, ior=ior)
    _check(oracle, sc, trace_kwargs(sc))


def test_shoebox_axis_aligned_rays_hit_zero_components(oracle):
    # 4 rays: directions are (almost) axis aligned; specular directions get exact-zero components, the case
    # the specular shortcut must hand to the general path.
    sc = scenes.shoebox(ray_count=4, max_bounces=16)
    sc.source = (5.0, 3.0)
    for frame in range(1, 40):
        _check(oracle, sc, trace_kwargs(sc, ray_count=4, rng_state_offset=frame, flags=1))


@pytest.mark.parametrize("bands", [1, 8])
@pytest.mark.parametrize("n_segments", [1000, 1003])
def test_maze(oracle, bands, n_segments):
    sc = scenes.maze(n_segments=n_segments, ray_count=3000, max_bounces=16, bands=8)
    _check(oracle, sc, trace_kwargs(sc, bands=bands))


@pytest.mark.parametrize("n_walls", [0, 1, 2, 3, 5, 7])
def test_wall_counts_around_the_batch_size(oracle, n_walls):
    sc = scenes.smoll_room()
    sc.walls = sc.walls[[0, 4, 8, 12, 16, 17, 18][:n_walls]] if n_walls else sc.walls[:0]
    _check(oracle, sc, trace_kwargs(sc, ray_count=2000))


def test_fft_against_numpy():
    rng = np.random.default_rng(0)
    for _ in range(4):
        w = rng.standard_normal(512).astype(np.float32)
        got = emulation.rfft512(w).astype(np.complex128)
        ref = np.fft.rfft(w.astype(np.float64))
        assert abs(got[0].real - ref[0].real) <= 1e-4 and abs(got[0].imag - ref[256].real) <= 1e-4   # packed DC / Nyquist
        assert np.abs(got[1:] - ref[1:256]).max() <= 2e-6 * np.abs(ref).max() * 10
        assert np.abs(emulation.irfft512(got.astype(np.complex64)) - w).max() <= 2e-6


@pytest.mark.parametrize("n_in,n_ir", [(4800, 72000), (1000, 300), (256, 256), (257, 255), (5000, 1), (3, 7)])
def test_partitioned_convolution_against_direct_form(oracle, n_in, n_ir):
    rng = np.random.default_rng(n_in + n_ir)
    x = rng.uniform(-0.5, 0.5, n_in).astype(np.float32)
    x[::7] *= np.float32(1e-4)
    h = scenes.decaying_noise_ir(n_ir, 3, decay_s=0.4)
    got = emulation.convolve(x, h, 3)
    want = oracle.convolve(x, h, 3)
    assert rel_l2(got, want) <= 1e-4


def test_filtered_intersection_equals_literal_formula(oracle):
    """The division-free filter + exact path against the oracle's literal `intersect`, on random and
    adversarial inputs: axis-aligned rays and walls, rays through wall endpoints, near-parallel pairs,
    tiny and huge scales."""
    rng = np.random.default_rng(11)
    n = 200_000
    o = rng.uniform(-50, 50, (n, 2))
    ang = rng.uniform(0, 2 * np.pi, n)
    d = np.stack([np.cos(ang), np.sin(ang)], 1)
    a = rng.uniform(-50, 50, (n, 2))
    b = a + rng.uniform(-30, 30, (n, 2))
    k = n // 8
    d[:k] = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]])[rng.integers(0, 4, k)]        # axis-aligned rays
    b[k:2 * k, 0] = a[k:2 * k, 0]                                                      # vertical walls
    b[2 * k:3 * k, 1] = a[2 * k:3 * k, 1]                                              # horizontal walls
    t = rng.uniform(0.5, 40, k)[:, None]
    a[3 * k:4 * k] = o[3 * k:4 * k] + d[3 * k:4 * k] * t                               # ray through endpoint a
    b[4 * k:5 * k] = o[4 * k:5 * k] + d[4 * k:5 * k] * t                               # ray through endpoint b
    b[5 * k:6 * k] = a[5 * k:6 * k] + d[5 * k:6 * k] * t + rng.normal(0, 1e-5, (k, 2))  # nearly parallel
    scale = np.ones(n)
    scale[6 * k:7 * k] = 1e-3
    scale[7 * k:] = 1e4
    o, a, b = o * scale[:, None], a * scale[:, None], b * scale[:, None]
    rays = np.concatenate([o, d], 1).astype(np.float32)
    segs = np.concatenate([a, b], 1).astype(np.float32)
    closest = np.where(rng.random(n) < 0.5, 1e8, rng.uniform(0.1, 60, n) * scale).astype(np.float32)
    got = emulation.intersect_many(rays, segs, closest)
    L = oracle.lib()
    want = np.array([L.orc_intersect(*map(float, rays[i]), *map(float, segs[i])) for i in range(n)], np.float32)
    want = np.where(want < closest, want, np.float32(1e8))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (want < 1e8).sum() > n // 20


# ---- uniform-grid mode (RAR_FLAG_USE_GRID): identical results to the brute-force oracle ------------------------

def _check_grid(O, sc, kw):
    P = oracle_params(O, dict(kw, flags=kw["flags"] | 8))
    r = O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw), want_hits=True)
    for counting in (True, False):
        hist, hits, _ = emulation.trace(O, sc.walls, P, None, counting=counting)
        assert np.array_equal(hist, r.hist)
        assert len(hits) == len(r.hits) and hits.tobytes() == r.hits.tobytes()
    return r


@pytest.mark.parametrize("frame", [1, 5])
def test_grid_mode_bundled_rooms(oracle, frame):
    for sc in (scenes.smoll_room(), scenes.big_room()):      # long rotated walls spanning many cells
        _check_grid(oracle, sc, trace_kwargs(sc, ray_count=4000, rng_state_offset=frame, max_bounce_count=8))


@pytest.mark.parametrize("n_segments,seed", [(300, 1), (2000, 2), (6000, 3)])
def test_grid_mode_mazes(oracle, n_segments, seed):
    sc = scenes.maze(n_segments=n_segments, ray_count=3000, max_bounces=20, bands=8, seed=seed)
    r = _check_grid(oracle, sc, trace_kwargs(sc))
    assert r.n_hits > 50


def test_grid_mode_shoebox_and_axis_aligned_rays(oracle):
    sc = scenes.shoebox(ray_count=20000, max_bounces=32, scattering=0.2, transmission=0.3, ior=1.2)
    _check_grid(oracle, sc, trace_kwargs(sc))
    sc = scenes.shoebox(ray_count=4, max_bounces=16)
    sc.source = (5.0, 3.0)                                   # rays along cell boundaries of the grid
    for frame in range(1, 20):
        _check_grid(oracle, sc, trace_kwargs(sc, ray_count=4, rng_state_offset=frame, flags=1))


def test_grid_mode_rays_leaving_an_open_scene(oracle):
    sc = scenes.smoll_room()
    sc.walls = sc.walls[[0, 1, 2, 3, 16, 17]]                # open on several sides: rays escape the grid
    _check_grid(oracle, sc, trace_kwargs(sc, ray_count=5000))
    sc.source = (-80.0, 30.0)                                # source outside the walls' bounding box
    _check_grid(oracle, sc, trace_kwargs(sc, ray_count=5000))


def test_grid_mode_random_soup_of_segments(oracle):
    """Random overlapping, crossing, tiny and long segments with shared endpoints: nothing grid-friendly."""
    from realisticaudioraytracing2d_b200.host.scene_helper import SEGMENT_DTYPE
    rng = np.random.default_rng(9)
    n = 1500
    a = rng.uniform(0, 50, (n, 2)).astype(np.float32)
    length = np.where(rng.random(n) < 0.1, rng.uniform(5, 40, n), rng.uniform(0.01, 2.0, n))
    ang = rng.uniform(0, 2 * np.pi, n)
    b = (a + np.stack([np.cos(ang), np.sin(ang)], 1) * length[:, None]).astype(np.float32)
    b[1::7] = a[0:-1:7][: len(b[1::7])]                      # shared endpoints
    walls = np.zeros(n, dtype=SEGMENT_DTYPE)
    walls["start"], walls["end"] = a, b
    d = b - a
    nrm = np.stack([d[:, 1], -d[:, 0]], 1) / np.maximum(np.hypot(d[:, 0], d[:, 1]), 1e-9)[:, None]
    walls["normal"] = nrm.astype(np.float32)
    walls["absorption"], walls["scattering"], walls["transmission"], walls["ior"] = 0.1, 0.3, 0.2, 1.1
    sc = scenes.smoll_room()
    sc.walls, sc.source, sc.listener = walls, (25.0, 25.0), (30.0, 22.0)
    r = _check_grid(oracle, sc, trace_kwargs(sc, ray_count=4000, max_bounce_count=12))
    assert r.counters["ray_bounces"] > 20000


def test_air_attenuation_host_logic_matches_the_oracle(oracle):
    """rar_ray.cuh's band arrivals with air absorption (counting and production instantiations) against the oracle."""
    sc = scenes.maze(n_segments=300, ray_count=6000, max_bounces=12, bands=8, seed=5)
    air = np.array([0.0, 1e-4, 5e-4, 1e-3, 3e-3, 1e-2, 3e-2, 0.1], np.float32)
    P = oracle_params(oracle, trace_kwargs(sc, bands=8, impulse_length=12000))
    want = oracle.trace(oracle_walls(oracle, sc.walls), P, band_abs=sc.band_absorption, air=air).hist
    assert np.count_nonzero(want) > 3000
    for counting in (True, False):
        got, _, _ = emulation.trace(oracle, oracle_walls(oracle, sc.walls), P, band_abs=sc.band_absorption, counting=counting, air=air)
        assert np.array_equal(got, want), counting


def test_fft16_register_transform():
    """rar_synth16.cuh: the 16-point transform of the synthesis kernel (forward / inverse, full / upper half zero)."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex64)
    assert np.abs(emulation.fft16(x) - np.fft.fft(x.astype(np.complex128))).max() < 2e-6 * np.abs(x).sum()
    assert np.abs(emulation.fft16(x, inverse=True) - 16 * np.fft.ifft(x.astype(np.complex128))).max() < 2e-6 * np.abs(x).sum()
    h = x.copy()
    h[8:] = 0
    junk = x.copy()                                                     # the upper half is not read
    assert np.abs(emulation.fft16(junk, half=True) - np.fft.fft(h.astype(np.complex128))).max() < 2e-6 * np.abs(x).sum()
    assert np.abs(emulation.fft16(junk, inverse=True, half=True) - 16 * np.fft.ifft(h.astype(np.complex128))).max() < 2e-6 * np.abs(x).sum()


@pytest.mark.parametrize("bins,bands", [(900, 8), (256, 4), (257, 8), (1300, 12), (100, 8)])
def test_band_synthesis_with_register_transforms(oracle, bins, bands):
    """The arithmetic of csrc/band_synth.cu (16 x 16 register transforms, real zero-phase weights, one partner exchange
    per segment) against the oracle's direct-form filter bank: 1e-4 relative L2 is the contract, ~1e-6 is observed."""
    rng = np.random.default_rng(bins + bands)
    ir = (rng.random((bins, bands)) * (rng.random((bins, bands)) < 0.3) * 1e-3).astype(np.float32)
    ir[-1] = 1e-3                                                       # the last bin carries energy (tail handling)
    ir[0] = 2e-3
    hist = np.array([oracle.lib().orc_quantize(float(v)) for v in ir.ravel()], np.int64)
    want = oracle.synthesize_ir(hist, bins, bands)
    taps = np.zeros((bands, 256), np.float32)
    for b in range(bands):
        taps[b, :255] = oracle.band_filter_taps(b / bands, (b + 1) / bands)
    got = emulation.band_synth16(hist, bands, 1.0, taps, bins)
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < 2e-6, err
    half = emulation.band_synth16(hist, bands, 0.5, taps, bins)
    assert np.linalg.norm(half - 0.5 * want) / np.linalg.norm(want) < 2e-6


@pytest.mark.parametrize("n", [4, 7, 20])
def test_pair_planes_layout(n):
    """The planes the packed wall scans read: record p = walls 2p and 2p+1 component by component (start x, start y,
    edge x, edge y); a missing second wall is NaN, which no comparison of the filter accepts."""
    walls = scenes.maze(n_segments=max(n, 8), ray_count=64, max_bounces=2, bands=8).walls[:n].copy()
    a, b = emulation.pair_planes(walls)
    assert a.shape == ((n + 1) // 2, 4)
    ex = (walls["end"][:, 0] - walls["start"][:, 0]).astype(np.float32)
    ey = (walls["end"][:, 1] - walls["start"][:, 1]).astype(np.float32)
    for w in range(n):
        p, h = divmod(w, 2)
        assert a[p, h] == walls["start"][w, 0] and a[p, 2 + h] == walls["start"][w, 1]
        assert b[p, h] == ex[w] and b[p, 2 + h] == ey[w]
    if n % 2:
        assert np.isnan(a[-1, 1]) and np.isnan(a[-1, 3]) and np.isnan(b[-1, 1]) and np.isnan(b[-1, 3])
