"""Per-rank times of the weak-scaled config-2 leg measured on ONE GPU: rank r of N traces ray ids [r, r+1) x 1 Mi of a
dispatch of N Mi rays, i.e. one angular sector; the step time of the N-GPU run is the maximum."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from realisticaudioraytracing2d_b200 import _capi, scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = _capi.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
sc = scenes.shoebox(ray_count=(1 << 20) * N, max_bounces=32)
n = sc.impulse_length
ctx.set_walls(sc.walls)
ctx.ir_clear(0, n, 1)
for r in range(N):
    lo, hi = r << 20, (r + 1) << 20
    p = _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain, 32, 1, sc.ray_count, 100,
                                sc.sample_rate, n, 1, 1.0, 0, lo, hi)
    ctx.trace(p, 0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        ctx.ir_clear(0, n, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.trace(p, 0); e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    p.flags = _capi.RAR_FLAG_COUNT_TESTS
    ctx.get_counters(reset=True); ctx.ir_clear(0, n, 1); ctx.trace(p, 0)
    c = ctx.get_counters()
    print(f"rank {r}/{N}: {min(ts):.4f} ms  direct {c['direct_hits']} nee {c['nee_hits']} shadow {c['shadow_tests']}", flush=True)
