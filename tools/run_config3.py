#!/usr/bin/env python
"""BASELINE config 3 at (or near) full size: 10 000-wall maze, 8 frequency bands, N rays x B bounces, rays
sharded across the ranks by contiguous id range (STRONG scaling: the dispatch is fixed), one int64 all-reduce.

    python tools/run_config3.py [--rays 67108864] [--bounces 64] [--walls 10000] [--bands 8] [--count]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/run_config3.py ...

Prints one JSON line with the IR-build time (max over ranks, CUDA events), tests/s when --count is given, and the
SHA-256 of the complete histogram, which must be identical for every GPU count.
"""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from realisticaudioraytracing2d_b200 import _capi, scenes  # noqa: E402
from realisticaudioraytracing2d_b200.host import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 26)
    ap.add_argument("--bounces", type=int, default=64)
    ap.add_argument("--walls", type=int, default=10000)
    ap.add_argument("--bands", type=int, default=8)
    ap.add_argument("--count", action="store_true")
    ap.add_argument("--grid", action="store_true", help="RAR_FLAG_USE_GRID: same histogram, far fewer tests")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = _capi.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    sc = scenes.maze(n_segments=a.walls, ray_count=a.rays, max_bounces=a.bounces, bands=8)
    n = sc.impulse_length
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption)
    ctx.ir_clear(0, n, a.bands)
    hist = sharding.DeviceHistogram(ctx, 0, dev).tensor
    lo, hi = sharding.shard_range(sharding.dispatched_threads(a.rays), rank, world)

    base_flags = _capi.RAR_FLAG_USE_GRID if a.grid else 0

    def prm(flags=0):
        flags |= base_flags
        return _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain,
                                       a.bounces, 1, a.rays, 0, sc.sample_rate, n, a.bands, 1.0, flags, lo, hi)
    # warm-up on a sliver of the range
    warm = _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain,
                                   2, 1, a.rays, 0, sc.sample_rate, n, a.bands, 1.0, 0, lo, min(hi, lo + 4096))
    ctx.trace(warm, 0)
    sharding.allreduce_histogram(hist)
    ctx.ir_clear(0, n, a.bands)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.trace(prm(), 0)
    sharding.allreduce_histogram(hist)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    digest = hashlib.sha256(hist.cpu().numpy().tobytes()).hexdigest()
    nz = int((hist != 0).sum())
    tests = None
    if a.count:
        ctx.ir_clear(1, n, a.bands)
        ctx.get_counters(reset=True)
        ctx.trace(prm(_capi.RAR_FLAG_COUNT_TESTS), 1)
        c = ctx.get_counters()
        t = torch.tensor([c["nearest_tests"] + c["shadow_tests"], c["ray_bounces"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tests = float(t[0])
    if rank == 0:
        print(json.dumps({"config": f"config3: {a.walls}-wall maze, {a.rays} rays x {a.bounces} bounces, {a.bands} bands, 48000 bins",
                          "n_gpus": world, "scaling": "strong", "grid": bool(a.grid), "ir_build_ms": float(ms[0]), "tests": tests,
                          "tests_per_s": tests / (float(ms[0]) * 1e-3) if tests else None,
                          "histogram_sha256": digest, "nonzero_words": nz}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.destroy()


if __name__ == "__main__":
    main()
