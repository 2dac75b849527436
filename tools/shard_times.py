"""One rank's share of the strong.c3 leg measured alone on ONE GPU: contiguous range against block-cyclic chunks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realisticaudioraytracing2d_b200 import _capi, scenes
from realisticaudioraytracing2d_b200.host.sharding import shard_range

world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rays = 148 * 1024 * 32
ctx = _capi.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
sc = scenes.maze(n_segments=10000, ray_count=rays, max_bounces=64, bands=8)
n = sc.impulse_length
ctx.set_walls(sc.walls); ctx.set_wall_band_absorption(sc.band_absorption); ctx.ir_clear(0, n, 8)
def prm(b=0, e=0):
    return _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain, 64, 1, rays, 0, sc.sample_rate, n, 8, 1.0, 0, b, e)
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for r in (0, world // 2, world - 1):
    lo, hi = shard_range(rays, r, world)
    print(f"world {world} rank {r}: contiguous {timed(lambda: ctx.trace(prm(lo, hi), 0)):.1f} ms   block-cyclic 2^14 {timed(lambda: ctx.trace_interleaved(prm(), 0, r, world, 14)):.1f} ms"
          f"   block-cyclic 2^17 {timed(lambda: ctx.trace_interleaved(prm(), 0, r, world, 17)):.1f} ms", flush=True)
