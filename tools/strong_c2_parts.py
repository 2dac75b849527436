"""strong.c2 on ONE GPU, piece by piece: one rank's block-cyclic share of the 1 Mi-ray dispatch at world = 1, 2, 4, 8
(clear + trace, device time, mean of 20), to separate the trace share from the exchange in the N-GPU `ir_build_ms`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realisticaudioraytracing2d_b200 import _capi, scenes

ctx = _capi.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
sc = scenes.shoebox(ray_count=1 << 20, max_bounces=32)
n = sc.impulse_length
ctx.set_walls(sc.walls); ctx.ir_clear(0, n, 1)
def prm():
    return _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain, sc.max_bounces, 1, sc.ray_count, 0, sc.sample_rate, n, 1, 1.0, 0, 0, 0)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("rays", sc.ray_count, "bounces", sc.max_bounces, "walls", len(sc.walls))
for world in (1, 2, 4, 8):
    def step():
        ctx.ir_clear(0, n, 1)
        if world == 1: ctx.trace(prm(), 0)
        else: ctx.trace_interleaved(prm(), 0, 0, world, 14)
    def trace_only():
        if world == 1: ctx.trace(prm(), 0)
        else: ctx.trace_interleaved(prm(), 0, 0, world, 14)
    print(f"world {world}: clear + trace {timed(step) * 1e3:.1f} us   trace only {timed(trace_only) * 1e3:.1f} us   ideal {timed(lambda: ctx.trace(prm(), 0)) * 1e3 / world:.1f} us", flush=True)
