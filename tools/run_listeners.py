#!/usr/bin/env python
"""One fused-listener trace of the config-4 share at a reduced ray count (the command line ncu wraps).

    python tools/run_listeners.py [log2_rays=18] [listeners=128]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realisticaudioraytracing2d_b200 import _capi, scenes  # noqa: E402


def main():
    rays = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    ctx = _capi.Context(0)
    sc = scenes.maze(n_segments=2000, ray_count=rays, max_bounces=5, bands=8)
    n = sc.impulse_length
    gx, gy = np.meshgrid(np.linspace(8, 92, 32), np.linspace(8, 92, 32))
    mine = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)[256:256 + per]
    ctx.set_walls(sc.walls)
    for l in range(per):
        ctx.ir_clear(100 + l, n, 1)
    prm = _capi.make_trace_params(sc.source, (0.0, 0.0), sc.listener_radius, sc.speed_of_sound, sc.input_gain, 5, 1, rays, 0,
                                  sc.sample_rate, n, 1, 1.0, 0, 0, 0)
    for _ in range(2):
        ctx.trace_listeners(prm, mine, 100)
    ctx.sync()
    ctx.destroy()


if __name__ == "__main__":
    main()
