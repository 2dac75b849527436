"""Bisects a histogram mismatch: per-ray-range bounce counters, hit lists and histograms of the GPU kernels against the
CPU oracle on the config-2 shoebox (written while chasing a ptxas miscompile of a predicate-derived shared address)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from realisticaudioraytracing2d_b200 import _capi, scenes
from tests.common import capi_params, oracle_params, oracle_walls, sort_hits, trace_kwargs
O.build()
ctx = _capi.Context(0)
sc = scenes.shoebox(ray_count=1 << 20, max_bounces=32)
ctx.set_walls(sc.walls)
n = sc.impulse_length
# per-range bounce counters from the COUNT kernel
for lo, hi in [(0, 32), (32, 64), (0, 256), (256, 512), (0, 4096), (4096, 8192), (151552, 151552 + 4096), (1 << 19, (1 << 19) + 4096), (0, 1 << 20)]:
    kw = trace_kwargs(sc, ray_begin=lo, ray_end=hi, flags=_capi.RAR_FLAG_COUNT_TESTS)
    ctx.ir_clear(0, n, 1)
    ctx.get_counters(reset=True)
    ctx.trace(capi_params(_capi, kw), 0)
    c = ctx.get_counters()
    r = O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw))
    same = np.array_equal(ctx.ir_read_fixed(0, n), r.hist)
    print(lo, hi, "gpu bounces", c["ray_bounces"], "oracle", r.counters["ray_bounces"], "hist equal", same, flush=True)
# hit lists on a small range
kw = trace_kwargs(sc, ray_begin=0, ray_end=4096)
hits, keys, cnt = ctx.trace_hits(capi_params(_capi, kw), capacity=4096 * 64 + 1024)
r = O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw), want_hist=False, want_hits=True)
print("hits gpu", cnt, "oracle", r.n_hits)
hits, keys = sort_hits(hits, keys)
m = min(len(hits), len(r.hits))
bad = np.flatnonzero((keys["ray"][:m] != r.hits["ray"][:m]) | (keys["bounce"][:m] != r.hits["bounce"][:m]) | (hits["time_delay"][:m] != r.hits["time_delay"][:m]))
print("first differing hit index", bad[:5], keys[bad[:3]] if len(bad) else None, r.hits[bad[:3]] if len(bad) else None)
# production kernel on the small range
ctx.ir_clear(0, n, 1)
ctx.trace(capi_params(_capi, kw), 0)
want = O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw)).hist
got = ctx.ir_read_fixed(0, n)
print("prod small range equal", np.array_equal(got, want), np.count_nonzero(got), np.count_nonzero(want), int(got.sum()), int(want.sum()))
