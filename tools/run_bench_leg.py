#!/usr/bin/env python
"""Runs single legs of bench.py on one GPU (the short command line that ncu wraps, and a debugging aid).

    python tools/run_bench_leg.py strong|listeners|banded|maze|config1 [...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from realisticaudioraytracing2d_b200 import _capi, scenes  # noqa: E402
from realisticaudioraytracing2d_b200.host import sharding  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = _capi.Context(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks, kind = bench._peaks()
    env = dict(ctx=ctx, capi=_capi, scenes=scenes, torch=torch, dist=None, sharding=sharding, stream=stream, flush=flush, dev=dev,
               rank=0, world=1, ex=None, use_nccl=False, barrier=torch.cuda.synchronize, fp32_peak=ctx.measure_fp32_peak())
    for leg in sys.argv[1:] or ["strong"]:
        if leg == "strong":
            out = bench.bench_strong(env)
        elif leg == "listeners":
            out = bench.bench_listeners(env)
        elif leg == "banded":
            out = bench.bench_banded(env, peaks, kind)
        elif leg == "maze":
            out = bench.bench_maze(ctx, _capi, scenes, torch, stream, flush, env["fp32_peak"])
        elif leg == "config1":
            out = bench.bench_config1(ctx, _capi, scenes, torch, stream)
        else:
            raise SystemExit(f"unknown leg {leg}")
        print(json.dumps({leg: out}), flush=True)
    ctx.destroy()


if __name__ == "__main__":
    main()
