"""Golden fixtures (tests/golden/*.npz, minted from the oracle by tests/golden/make_golden.py -- the
reference has none): the oracle must keep reproducing them (CPU), and the CUDA path must reproduce them
without the oracle in the loop (GPU)."""
import hashlib
import os

import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import scenes
from tests.common import capi_params, oracle_params, oracle_walls, rel_l2, sort_hits, trace_kwargs
from tests.golden.make_golden import CASES

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CTR = ("ray_bounces", "nearest_tests", "shadow_tests", "direct_hits", "nee_hits")


def _load(name):
    g = np.load(os.path.join(HERE, name + ".npz"))
    hist = np.zeros(int(g["n_words"]), np.int64)
    hist[g["idx"]] = g["val"]
    return g, hist


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    sc, over = CASES[name]()
    kw = trace_kwargs(sc, **over)
    g, hist = _load(name)
    r = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw),
                     band_abs=sc.band_absorption if kw["bands"] > 1 else None, want_hits=True)
    assert np.array_equal(r.hist, hist)
    assert r.n_hits == int(g["n_hits"])
    assert hashlib.sha256(r.hits.tobytes()).hexdigest().encode() == bytes(g["hits_sha256"])
    assert [r.counters[k] for k in CTR] == list(g["counters"])


def test_oracle_reproduces_golden_convolution(oracle):
    g = np.load(os.path.join(HERE, "conv_1000x300.npz"))
    assert np.array_equal(oracle.convolve(g["x"], g["ir"], int(g["accum"])), g["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_golden(ctx, name):
    from realisticaudioraytracing2d_b200 import _capi
    sc, over = CASES[name]()
    kw = trace_kwargs(sc, **over)
    g, hist = _load(name)
    ctx.set_walls(sc.walls)
    if kw["bands"] > 1:
        ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(0, kw["impulse_length"], kw["bands"])
    ctx.get_counters(reset=True)
    ctx.trace(capi_params(_capi, dict(kw, flags=_capi.RAR_FLAG_COUNT_TESTS)), 0)
    assert np.array_equal(ctx.ir_read_fixed(0, hist.size), hist)
    c = ctx.get_counters()
    assert [c[k] for k in CTR] == list(g["counters"])
    if kw["bands"] == 1:
        hits, keys, n = ctx.trace_hits(capi_params(_capi, kw), capacity=int(g["n_hits"]) + 16)
        assert n == int(g["n_hits"])
        hits, keys = sort_hits(hits, keys)
        packed = np.zeros(n, dtype=[("time_delay", "<f4"), ("energy", "<f4"), ("hit_x", "<f4"), ("hit_y", "<f4"),
                                    ("ray", "<u4"), ("bounce", "<u2"), ("kind", "<u2")])
        packed["time_delay"], packed["energy"] = hits["time_delay"], hits["energy"]
        packed["hit_x"], packed["hit_y"] = hits["hit_point"][:, 0], hits["hit_point"][:, 1]
        packed["ray"], packed["bounce"], packed["kind"] = keys["ray"], keys["bounce"], keys["kind"]
        assert hashlib.sha256(packed.tobytes()).hexdigest().encode() == bytes(g["hits_sha256"])


@pytest.mark.gpu
def test_gpu_reproduces_golden_convolution(ctx):
    g = np.load(os.path.join(HERE, "conv_1000x300.npz"))
    ctx.ir_write(0, g["ir"])
    assert np.array_equal(ctx.ir_read(0, 300), g["ir"])            # already on the Q23.40 grid
    got = ctx.convolve(0, g["x"], int(g["accum"]), 300)
    assert rel_l2(got, g["out"]) <= 1e-4
