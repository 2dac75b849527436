"""Fused listener kernel against single-listener traces at growing ray counts (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from realisticaudioraytracing2d_b200 import _capi, scenes

ctx = _capi.Context(0)
walls, bounces = 2000, 5
for rays, nl in ((1 << 16, 8), (1 << 20, 8), (1 << 22, 8), (1 << 22, 128)):
    sc = scenes.maze(n_segments=walls, ray_count=rays, max_bounces=bounces, bands=8)
    n = sc.impulse_length
    gx, gy = np.meshgrid(np.linspace(8, 92, 32), np.linspace(8, 92, 32))
    grid = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)[:nl]
    ctx.set_walls(sc.walls)
    def prm(listener=(0.0, 0.0), flags=0):
        return _capi.make_trace_params(sc.source, listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain, bounces, 1,
                                       rays, 0, sc.sample_rate, n, 1, 1.0, flags, 0, 0)
    for l in range(nl):
        ctx.ir_clear(100 + l, n, 1)
    ctx.trace_listeners(prm(), grid, 100)
    bad = 0
    for l in range(min(nl, 8)):
        ctx.ir_clear(400, n, 1)
        ctx.trace(prm(listener=(float(grid[l, 0]), float(grid[l, 1]))), 400)
        a, b = ctx.ir_read_fixed(100 + l, n), ctx.ir_read_fixed(400, n)
        d = np.flatnonzero(a != b)
        if len(d):
            bad += 1
            print(f"rays {rays} listeners {nl}: listener {l} differs in {len(d)} bins; sums {int(a.sum())} vs {int(b.sum())}; nonzero {np.count_nonzero(a)} vs {np.count_nonzero(b)}; first {d[:4]} {a[d[:4]]} {b[d[:4]]}")
    print(f"rays {rays} listeners {nl}: {bad} of {min(nl, 8)} listeners differ", flush=True)
