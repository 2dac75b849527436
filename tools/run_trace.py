#!/usr/bin/env python
"""Runs the trace kernel (or the streaming convolver) a few times; the short command line that ncu wraps.

    python tools/run_trace.py c2|c1|maze|maze8|wallsN|conv|clip [reps]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realisticaudioraytracing2d_b200 import _capi, scenes  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "c2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ctx = _capi.Context(0)
    if what == "conv":
        S, B, n_ir = 256, 256, 480000
        cv = _capi.Convolver(ctx, S, B, n_ir)
        ir = scenes.decaying_noise_ir(n_ir, 1, 3.0)
        for s in range(S):
            cv.set_ir(s, np.roll(ir, s))
        x = np.random.default_rng(0).uniform(-1, 1, (S, B)).astype(np.float32)
        for _ in range(reps):
            t0 = time.perf_counter()
            cv.process(x)
            print(f"conv block {1e3 * (time.perf_counter() - t0):.3f} ms (host copies included)")
        return
    if what == "clip":                      # bench.py's clip_prep leg: 256 stereo 10 s clips, 44.1 -> 48 kHz
        import torch
        n_clips, samples, ch, freq, rate = 256, 441000, 2, 44100, 48000
        n_out = _capi.prepared_length(samples, freq, rate)
        raw = torch.rand((n_clips, samples, ch), device="cuda") * 2 - 1
        out = torch.empty((n_clips, n_out), device="cuda")
        torch.cuda.synchronize()
        for _ in range(reps):
            t0 = time.perf_counter()
            ctx.prepare_clips_device(raw.data_ptr(), samples, ch, freq, rate, n_clips, out.data_ptr(), n_out)
            ctx.sync()
            print(f"clip prep {1e3 * (time.perf_counter() - t0):.3f} ms")
        return
    if what == "c2":
        sc, bands = scenes.shoebox(), 1
    elif what == "c1":
        sc, bands = scenes.smoll_room(), 1
    elif what.startswith("walls"):          # e.g. walls16000: staging-mode comparison at a given wall count
        sc, bands = scenes.maze(n_segments=int(what[5:]), ray_count=148 * 1024 * 2, max_bounces=8, bands=8), 1
    else:
        sc, bands = scenes.maze(n_segments=10000, ray_count=148 * 1024 * 2, max_bounces=16, bands=8), (8 if what == "maze8" else 1)
    n = sc.impulse_length
    ctx.set_walls(sc.walls)
    if sc.band_absorption is not None:
        ctx.set_wall_band_absorption(sc.band_absorption)
    p = _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain, sc.max_bounces,
                                1, sc.ray_count, 100 if what == "c2" else 0, sc.sample_rate, n, bands)
    if os.environ.get("RAR_GRID"):
        p.flags = _capi.RAR_FLAG_USE_GRID
    for r in range(reps):
        ctx.ir_clear(0, n, bands)
        ctx.sync()
        t0 = time.perf_counter()
        ctx.trace(p, 0)
        ctx.sync()
        print(f"{what} trace {1e3 * (time.perf_counter() - t0):.3f} ms")
    ctx.get_counters(reset=True)
    p.flags |= _capi.RAR_FLAG_COUNT_TESTS
    ctx.ir_clear(0, n, bands)
    ctx.trace(p, 0)
    c = ctx.get_counters()
    print("tests per launch", c["nearest_tests"] + c["shadow_tests"], c)
    print("nonzero bins", int(np.count_nonzero(ctx.ir_read_fixed(0, n * bands))))


if __name__ == "__main__":
    main()
