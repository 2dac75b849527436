"""Pins the CPU oracle (oracle/rar_oracle.c) against everything the reference and first principles
offer.  The reference ships no tests or golden vectors, so the pins are:
  * PRNG known-answer values derived by hand from Common.hlsl:8-12;
  * literal cases of intersect / intersectCircle / Refract (Common.hlsl:14-43);
  * analytic acoustics: direct-path arrival time and 1/d^2 energy (Raytrace2D.compute:74-84), image-source
    arrival times in a shoebox with specular walls;
  * AudioConvolve.compute:13-31 properties: impulse, delayed impulse, epsilon gate, N+M outputs;
  * the scene fixtures of Appendix B.1-B.2 and the workload statistics derived there.
"""
import ctypes as C

import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import scenes
from tests.common import oracle_params, oracle_walls, trace_kwargs

# (seed, state after 1 draw, float32 bits of the first value)  -- SURVEY.md Appendix B.3
PRNG_KAT = [
    (0, 2891336453, 0x3CF765FC), (1, 3639132858, 0x3F28BEEA), (2, 91961967, 0x3EF4FD99),
    (63, 2757869712, 0x3F259979), (14999, 635037896, 0x3E0212DD),
    (719393, 556825434, 0x3ED86A85),      # frame 1, id 0
    (5035753, 930254018, 0x3F7AA0C3),     # frame 7, id 2
]


@pytest.mark.parametrize("seed,state,bits", PRNG_KAT)
def test_prng_known_answers(oracle, seed, state, bits):
    vals, states = oracle.random_sequence(seed, 1)
    assert states[0] == state
    assert int(vals.view(np.uint32)[0]) == bits


def test_prng_sequence_and_upper_edge(oracle):
    vals, states = oracle.random_sequence(0, 3)
    assert states == [2891336453, 1192405134, 568162667]
    assert [int(v) for v in vals.view(np.uint32)] == [0x3CF765FC, 0x3E0ADADB, 0x3E6FDB83]
    # (float)uint rounds to nearest: res >= 4294967168 gives exactly 1.0 (Appendix B.3 "Edge")
    assert np.float32(4294967167) * np.float32(2.0 ** -32) == np.float32(0.99999994)
    assert np.float32(4294967168) * np.float32(2.0 ** -32) == np.float32(1.0)
    big = oracle.random_sequence(12345, 200000)[0]
    assert big.min() >= 0.0 and big.max() <= 1.0 and abs(float(big.mean()) - 0.5) < 5e-3


def test_intersect_literal_cases(oracle):
    I = oracle.lib().orc_intersect
    inf = np.float32(1e8)
    assert I(0, 0, 1, 0, 5, -1, 5, 1) == 5.0                       # perpendicular hit at t1 = 5, t2 = 0.5
    assert I(0, 0, 1, 0, 5, 0, 5, 2) == 5.0                        # t2 = 0 (endpoint a) is inside
    assert I(0, 0, 1, 0, 5, -2, 5, 0) == 5.0                       # t2 = 1 (endpoint b) is inside
    assert I(0, 0, 1, 0, 5, 0.001, 5, 2) == inf                    # just past the endpoint
    assert I(0, 0, -1, 0, 5, -1, 5, 1) == inf                      # behind the ray
    assert I(0, 0, 1, 0, 0, 1, 10, 1) == inf                       # parallel: |dotP| < eps
    assert I(0, 0, 1, 0, 5e-5, -1, 5e-5, 1) == inf                 # t1 < eps
    assert I(0, 0, 1, 0, 1e-4, -1, 1e-4, 1) == np.float32(1e-4)    # t1 == eps is a hit (>=)
    # the parallel test is scale dependent (v2 is not normalised): a short wall at a grazing angle is skipped
    assert I(0, 0, 1, 0, 5, 0, 5 + 1e-3, 5e-5) == inf


def test_intersect_circle_literal_cases(oracle):
    Ic = oracle.lib().orc_intersect_circle
    inf = np.float32(1e8)
    assert Ic(0, 0, 1, 0, 10, 0, 0.5) == 9.5                       # nearest root
    assert Ic(0, 0, -1, 0, 10, 0, 0.5) == inf                      # behind (tca < 0)
    assert Ic(0, 0, 1, 0, 10, 0.6, 0.5) == inf                     # misses (d2 > r2)
    assert Ic(10, 0, 1, 0, 10, 0, 0.5) == 0.5                      # from inside: far root
    assert abs(Ic(0, 0, 1, 0, 10, 0.5, 0.5) - 10.0) < 1e-3         # tangent


def test_refract_literal_cases(oracle):
    ok, tx, ty = oracle.refract(0.0, -1.0, 0.0, 1.0, 0.5)          # normal incidence passes straight through
    assert ok == 1 and tx == 0.0 and ty == np.float32(-1.0)
    s = np.float32(np.sqrt(0.5))
    ok, tx, ty = oracle.refract(float(s), float(-s), 0.0, 1.0, 2.0)  # eta*sin(i) > 1: total internal reflection
    assert ok == 0 and tx == 0.0 and ty == 0.0
    ok, tx, ty = oracle.refract(float(s), float(-s), 0.0, 1.0, 0.5)  # Snell: sin(t) = eta sin(i)
    assert ok == 1 and abs(tx / np.hypot(tx, ty) - 0.5 * s) < 1e-6


def test_transcendental_kernels_accuracy(oracle):
    xs = np.linspace(-7.0, 7.0, 5001).astype(np.float32)
    err = 0.0
    for x in xs:
        s, c = oracle.sincos(float(x))
        err = max(err, abs(s - np.sin(np.float64(x))), abs(c - np.cos(np.float64(x))))
    assert err < 2e-7
    ys = np.linspace(-1.0, 1.0, 4001).astype(np.float32)
    ea = max(abs(oracle.lib().orc_asinf(float(y)) - np.arcsin(np.float64(y))) for y in ys)
    assert ea < 3e-7


def test_quantiser_and_time_bin(oracle):
    L = oracle.lib()
    assert L.orc_quantize(1.0) == 1 << 40 and L.orc_quantize(-0.5) == -(1 << 39)
    assert L.orc_quantize(1e-13) == 0 and L.orc_quantize(float("nan")) == 0        # truncation toward zero
    assert L.orc_quantize(1e30) == 1 << 62 and L.orc_quantize(-1e30) == -(1 << 62)  # saturation at +-2^22
    assert L.orc_time_bin(0.5, 48000, 1.0, 48000) == 24000
    assert L.orc_time_bin(1.0, 48000, 1.0, 48000) == -1                             # index < ImpulseLength
    assert L.orc_time_bin(-1e-6, 48000, 1.0, 48000) == 0                            # (int) truncates -0.048 to 0
    assert L.orc_time_bin(-1e-3, 48000, 1.0, 48000) == -1
    assert L.orc_time_bin(0.5, 48000, 128.0, 48000) == 187                          # banded variant: /WindowSize


def test_direct_path_arrival_time_and_energy(oracle):
    """No walls: each ray either crosses the listener circle or not; a crossing at distance d arrives at d/c
    with energy gain / max(1, d^2) (Raytrace2D.compute:78-81)."""
    sc = scenes.smoll_room()
    kw = trace_kwargs(sc, source=(0.0, 0.0), listener=(6.0, 0.0), ray_count=4096, input_gain=2.0)
    r = oracle.trace(sc.walls[:0].view(oracle.SEGMENT_DTYPE), oracle_params(oracle, kw), want_hits=True)
    h = r.hits
    assert len(h) > 0 and np.all(h["kind"] == 0) and np.all(h["bounce"] == 0)
    d = np.hypot(h["hit_x"], h["hit_y"]).astype(np.float64)
    assert np.allclose(h["time_delay"], d / 343.0, rtol=1e-6)
    assert np.allclose(h["energy"], 2.0 / np.maximum(1.0, d * d), rtol=1e-6)
    assert abs(d.min() - 5.5) < 1e-3                                 # (|S-L| - r)
    # the fraction of rays that hit is the angular size of the listener: 2*asin(r/|S-L|) / 2pi
    assert abs(len(h) / 4096 - 2 * np.arcsin(0.5 / 6.0) / (2 * np.pi)) < 2e-3
    assert r.counters["nearest_tests"] == 0 and r.counters["ray_bounces"] == 4096


def test_shoebox_image_source_arrival_times(oracle):
    """Specular shoebox: every listener crossing after k reflections lies on a straight line from an image
    source, so its arrival time is (|image - crossing point unfolded|)/c, i.e. path length / c."""
    sc = scenes.shoebox(ray_count=20000, max_bounces=6)
    kw = trace_kwargs(sc)
    r = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), want_hits=True)
    direct = r.hits[r.hits["kind"] == 0]
    assert len(direct) > 100
    W, H = 10.0, 6.0
    sx, sy = sc.source
    checked = 0
    for h in direct[:: max(1, len(direct) // 400)]:
        best = 1e9
        path = float(h["time_delay"]) * 343.0
        for ix in range(-7, 8):
            for iy in range(-7, 8):
                if abs(ix) + abs(iy) != int(h["bounce"]):
                    continue
                px = ix * W + (sx if ix % 2 == 0 else W - sx)
                py = iy * H + (sy if iy % 2 == 0 else H - sy)
                # unfold the crossing point into the same image cell
                qx = ix * W + (h["hit_x"] if ix % 2 == 0 else W - h["hit_x"])
                qy = iy * H + (h["hit_y"] if iy % 2 == 0 else H - h["hit_y"])
                best = min(best, abs(np.hypot(qx - sx, qy - sy) - path), abs(np.hypot(px - h["hit_x"], py - h["hit_y"]) - path))
        assert best < 5e-3, (h, best)       # eps offsets at each bounce move the path by ~1e-4 per bounce
        checked += 1
    assert checked > 100


def test_smoll_room_workload_statistics(oracle):
    """Appendix B 'derived workload statistics' (numpy emulation, statistical): 15 040 threads, 75 185
    ray-bounce iterations, 1 503 700 nearest tests, ~1.07 M shadow tests, ~12.3 k hits, 0.0616-0.2866 s."""
    sc = scenes.smoll_room()
    r = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, trace_kwargs(sc)), want_hits=True)
    c = r.counters
    assert c["ray_bounces"] == 75185 and c["nearest_tests"] == 1503700
    assert abs(c["shadow_tests"] - 1069134) < 2000
    assert abs(r.n_hits - 12332) < 50 and abs(c["direct_hits"] - 699) < 10
    assert abs(r.hits["time_delay"].min() - 0.0616) < 2e-4 and abs(r.hits["time_delay"].max() - 0.2866) < 2e-4
    assert abs(r.hist.sum() / 2.0 ** 40 - 0.79) < 0.01
    assert r.hits["energy"].min() > 1e-5 and abs(r.hits["energy"].max() - 1.57e-3) < 2e-5


def test_scene_fixtures_match_appendix_b(oracle):
    sc = scenes.smoll_room()
    w = sc.walls
    assert len(w) == 20 and w.dtype.itemsize == 40
    assert np.allclose(w["start"][0], (-50, 9.5)) and np.allclose(w["end"][0], (50, 9.5)) and np.allclose(w["normal"][0], (0, -1))
    assert np.allclose(w["start"][4], (-49.99, -5.5)) and np.allclose(w["normal"][6], (0, 1))
    assert np.allclose(w["start"][8], (-19.5, -10), atol=1e-5) and np.allclose(w["normal"][8], (1, 0), atol=1e-6)
    assert np.allclose(w["start"][16], (-38.5389, -35.0726), atol=1e-3) and np.allclose(w["end"][16], (15.7785, 48.8894), atol=1e-3)
    assert np.allclose(w["normal"][16], (0.83962, -0.543175), atol=1e-5)
    assert np.allclose(w["absorption"][:16], 0.507) and np.allclose(w["ior"][16:], 0.6)
    big = scenes.big_room().walls
    assert np.allclose(big["start"][16], (-386.189, -350.726), atol=2e-2) and np.allclose(big["end"][17], (148.589, 494.326), atol=2e-2)
    # the host mirror of AddLoopToSegments and the oracle's restatement agree bit for bit
    unit = [(-0.5, -0.5), (0.5, -0.5), (0.5, 0.5), (-0.5, 0.5)]
    o = oracle.add_loop(unit, (-11.8, 7.18), 0.47792548, 0.8784004, (100.0, 1.0), (0.148, 1.0, 1.0, 0.6))
    assert o.tobytes() == np.ascontiguousarray(w[16:20]).tobytes()
    o = oracle.add_loop(unit, (20.0, 0.0), 0.7071068, 0.7071068, (20.0, 1.0), (0.507, 0.5, 0.271, 0.01))
    assert o.tobytes() == np.ascontiguousarray(w[12:16]).tobytes()


def test_convolution_properties(oracle):
    rng = np.random.default_rng(1)
    ir = rng.standard_normal(300).astype(np.float32)
    x = np.zeros(50, np.float32)
    x[0] = 1.0
    y = oracle.convolve(x, ir, 4)
    assert y.shape == (350,) and y[-1] == 0.0                      # N+M outputs, the last always 0
    assert np.array_equal(y[:300], ir / np.float32(4))             # impulse reproduces ir / accumCount
    x = np.zeros(50, np.float32)
    x[17] = 1.0
    y = oracle.convolve(x, ir, 1)
    assert np.array_equal(y[17:317], ir) and not y[:17].any()      # delayed impulse shifts it
    x = rng.uniform(-1, 1, 400).astype(np.float32)
    x[::2] = rng.uniform(-1e-4, 1e-4, 200).astype(np.float32)      # |x| <= eps contributes nothing
    xz = x.copy()
    xz[np.abs(xz) <= np.float32(1e-4)] = 0
    assert np.array_equal(oracle.convolve(x, ir, 2), oracle.convolve(xz, ir, 2))
    assert not oracle.convolve(x, ir, 0).any()                     # accumCount <= 0 -> zeros
    from scipy.signal import fftconvolve
    ref = fftconvolve(xz.astype(np.float64), ir.astype(np.float64)) / 2
    assert np.linalg.norm(oracle.convolve(x, ir, 2)[:-1] - ref) / np.linalg.norm(ref) < 1e-6


def test_openmp_result_is_thread_count_independent(oracle):
    sc = scenes.smoll_room()
    P = oracle_params(oracle, trace_kwargs(sc, ray_count=3000))
    a = oracle.trace(oracle_walls(oracle, sc.walls), P, n_threads=1).hist
    b = oracle.trace(oracle_walls(oracle, sc.walls), P, n_threads=5).hist
    assert np.array_equal(a, b)


def test_band_filter_bank_properties(oracle):
    """The banded model's filter bank (this build's definition, include/rar2d.h): the band filters sum to a unit
    impulse at the filter delay; each is linear-phase; synthesis is linear, shift-invariant on the sample grid, and
    equals scipy's convolution of the taps with the band responses."""
    from scipy.signal import fftconvolve
    for bands in (2, 8, 128):
        g = np.stack([oracle.band_filter_taps(b / bands, (b + 1) / bands) for b in range(bands)]).astype(np.float64)
        delta = np.zeros(255)
        delta[127] = 1.0
        assert np.abs(g.sum(0) - delta).max() < 1e-6
        assert np.abs(g - g[:, ::-1]).max() < 1e-7                      # symmetric about tap 127
    rng = np.random.default_rng(1)
    bins, bands = 900, 8
    ir = (rng.random((bins, bands)) * (rng.random((bins, bands)) < 0.1) * 1e-3).astype(np.float32)
    hist = np.array([oracle.lib().orc_quantize(float(v)) for v in ir.ravel()], np.int64)
    out = oracle.synthesize_ir(hist, bins, bands)
    hf = oracle.ir_to_float(hist).reshape(bins, bands).astype(np.float64)
    want = np.zeros(bins + 254)
    for b in range(bands):
        want += fftconvolve(hf[:, b], oracle.band_filter_taps(b / bands, (b + 1) / bands).astype(np.float64))
    assert np.abs(out - want[127:127 + bins]).max() <= 1e-9
    # a band that carries everything alone is band-limited: energy outside [lo, hi] is small
    one = np.zeros((bins, bands), np.int64)
    one[100, 3] = 1 << 38
    spec = np.abs(np.fft.rfft(oracle.synthesize_ir(one.ravel(), bins, bands), 4096))
    f = np.arange(len(spec)) / (len(spec) - 1)
    inside = (f > 3 / 8 - 0.02) & (f < 4 / 8 + 0.02)
    assert (spec[~inside] ** 2).sum() < 2e-3 * (spec ** 2).sum()
    # coarse time bins: bin k stands on sample k * stride
    s4 = oracle.synthesize_ir(hist, bins, bands, 4)
    up = np.zeros((bins * 4, bands), np.int64)
    up[::4] = hist.reshape(bins, bands)
    assert np.array_equal(s4, oracle.synthesize_ir(up.ravel(), bins * 4, bands, 1))


# SURVEY.md Appendix B.2, every row: the four corners of each box in segment order (segment k runs corner k -> k+1)
# and the four outward normals, rounded to 6 significant digits there.
_B2 = {
    "smoll": [
        ([(-50, 9.5), (50, 9.5), (50, 10.5), (-50, 10.5)], [(0, -1), (1, 0), (0, 1), (-1, 0)]),
        ([(-49.99, -5.5), (50.01, -5.5), (50.01, -4.5), (-49.99, -4.5)], [(0, -1), (1, 0), (0, 1), (-1, 0)]),
        ([(-19.5, -10), (-19.5, 10), (-20.5, 10), (-20.5, -10)], [(1, 0), (0, 1), (-1, 0), (0, -1)]),
        ([(20.5, -10), (20.5, 10), (19.5, 10), (19.5, -10)], [(1, 0), (0, 1), (-1, 0), (0, -1)]),
        ([(-38.5389, -35.0726), (15.7785, 48.8894), (14.9389, 49.4326), (-39.3785, -34.5294)],
         [(.83962, -.543175), (.543175, .83962), (-.83962, .543175), (-.543175, -.83962)]),
    ],
    "big": [
        ([(-500, 99.5), (500, 99.5), (500, 100.5), (-500, 100.5)], [(0, -1), (1, 0), (0, 1), (-1, 0)]),
        ([(-499.99, -50.5), (500.01, -50.5), (500.01, -49.5), (-499.99, -49.5)], [(0, -1), (1, 0), (0, 1), (-1, 0)]),
        ([(-199.5, -100), (-199.5, 100), (-200.5, 100), (-200.5, -100)], [(1, 0), (0, 1), (-1, 0), (0, -1)]),
        ([(200.5, -100), (200.5, 100), (199.5, 100), (199.5, -100)], [(1, 0), (0, 1), (-1, 0), (0, -1)]),
        ([(-386.189, -350.726), (156.985, 488.894), (148.589, 494.326), (-394.585, -345.294)],
         [(.83962, -.543175), (.543175, .83962), (-.83962, .543175), (-.543175, -.83962)]),
    ],
}


@pytest.mark.parametrize("room", ["smoll", "big"])
def test_every_segment_of_appendix_b2_to_six_significant_digits(room):
    """All 20 segments of both bundled rooms -- start, end, normal and material -- against the table the survey derived
    from the scene transforms (Helpers/SceneHelper.cs:49-58,78-97), to the six significant digits it prints."""
    w = (scenes.smoll_room() if room == "smoll" else scenes.big_room()).walls

    def six(a, b):        # equal after rounding to 6 significant digits (absolute 5e-7 near zero: sin/cos of 90 degrees)
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return np.all(np.abs(a - b) <= np.maximum(5.1e-6 * np.abs(b), 5e-7))
    border, mat = (0.507, 0.5, 0.271, 0.01), (0.148, 1.0, 1.0, 0.6)
    for box, (corners, normals) in enumerate(_B2[room]):
        for k in range(4):
            seg = w[4 * box + k]
            assert six(seg["start"], corners[k]), (box, k, seg["start"], corners[k])
            assert six(seg["end"], corners[(k + 1) % 4]), (box, k, seg["end"])
            assert six(seg["normal"], normals[k]), (box, k, seg["normal"])
            m = mat if box == 4 else border
            assert (float(seg["absorption"]), float(seg["scattering"]), float(seg["transmission"]), float(seg["ior"])) == \
                tuple(float(np.float32(v)) for v in m)


def test_air_attenuation_kernel_and_trace(oracle):
    """The banded model's air absorption (this build's extension): exp(-alpha d) as a fixed polynomial kernel --
    accurate to 4e-6 relative down to the smallest normal number, exactly 0 below and exactly 1 at d = 0 -- scales the
    BAND energies of every arrival and nothing else (band with alpha = 0, counters and ray paths unchanged)."""
    xs = np.concatenate([np.linspace(0, 87, 5001), [0.0, 1e-8, 87.3]]).astype(np.float32)
    got = np.array([oracle.lib().orc_exp_neg(1.0, float(x)) for x in xs])
    want = np.exp(-xs.astype(np.float64))
    assert np.max(np.abs(got - want) / want) < 4e-6
    assert oracle.lib().orc_exp_neg(0.5, 0.0) == 1.0 and oracle.lib().orc_exp_neg(1.0, 200.0) == 0.0
    assert oracle.lib().orc_exp_neg(2.0, 3.0) == oracle.lib().orc_exp_neg(3.0, 2.0)
    sc = scenes.maze(n_segments=300, ray_count=8000, max_bounces=12, bands=8, seed=5)
    air = np.array([0.0, 1e-4, 5e-4, 1e-3, 3e-3, 1e-2, 3e-2, 0.1], np.float32)
    P = oracle_params(oracle, trace_kwargs(sc, bands=8, impulse_length=12000))
    r0 = oracle.trace(oracle_walls(oracle, sc.walls), P, band_abs=sc.band_absorption)
    r1 = oracle.trace(oracle_walls(oracle, sc.walls), P, band_abs=sc.band_absorption, air=air)
    h0, h1 = r0.hist.reshape(-1, 8), r1.hist.reshape(-1, 8)
    assert r0.counters == r1.counters and np.count_nonzero(h0) > 5000
    assert np.array_equal(h0[:, 0], h1[:, 0])                                  # alpha = 0
    ratio = h1.sum(0) / h0.sum(0)
    assert np.all(np.diff(ratio) < 0) and ratio[-1] < 0.1 < ratio[4]            # more absorption, less energy
    assert np.all(h1 <= h0)
