#!/usr/bin/env python
"""BASELINE config 4 (batched auralisation): one source, a grid of listeners in a 2 000-wall scene, R rays each,
listeners sharded across the ranks by contiguous range (no collective).  BASELINE leaves the bounce depth
unspecified; the reference default (5) is used unless --bounces is given.

    python tools/run_config4.py [--listeners 1024] [--rays 4194304] [--bounces 5] [--compare 8]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/run_config4.py ...

--compare K additionally times K listeners of this rank's share as K separate single-listener traces and checks
that the fused kernel's histograms are identical.
"""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from realisticaudioraytracing2d_b200 import _capi, scenes  # noqa: E402
from realisticaudioraytracing2d_b200.host import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--listeners", type=int, default=1024)
    ap.add_argument("--rays", type=int, default=1 << 22)
    ap.add_argument("--bounces", type=int, default=5)
    ap.add_argument("--walls", type=int, default=2000)
    ap.add_argument("--compare", type=int, default=0)
    ap.add_argument("--count", action="store_true")
    ap.add_argument("--grid", action="store_true", help="RAR_FLAG_USE_GRID: same histograms, far fewer tests")
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
    g = int(round(a.listeners ** 0.5))
    gx, gy = np.meshgrid(np.linspace(8, 92, g), np.linspace(8, 92, max(1, a.listeners // g)))
    all_listeners = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)[: a.listeners]
    lo, hi = sharding.shard_range(len(all_listeners), rank, world)
    mine = all_listeners[lo:hi]
    ctx.set_walls(sc.walls)
    for l in range(len(mine)):
        ctx.ir_clear(l, n, 1)

    base_flags = _capi.RAR_FLAG_USE_GRID if a.grid else 0

    def prm(listener=(0.0, 0.0), flags=0, rays=a.rays):
        flags |= base_flags
        return _capi.make_trace_params(sc.source, listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain,
                                       a.bounces, 1, rays, 0, sc.sample_rate, n, 1, 1.0, flags, 0, 0)
    ctx.trace_listeners(prm(rays=4096), mine[: min(2, len(mine))], 0)      # warm-up
    for l in range(min(2, len(mine))):
        ctx.ir_clear(l, n, 1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.trace_listeners(prm(), mine, 0)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    h = hashlib.sha256()
    for l in range(len(mine)):
        h.update(ctx.ir_read_fixed(l, n).tobytes())
    out = {"config": f"config4: {a.walls}-wall scene, {len(all_listeners)} listeners, {a.rays} rays x {a.bounces} bounces each, 48000 bins",
           "n_gpus": world, "grid": bool(a.grid), "listeners_per_gpu": len(mine), "fused_ms": float(ms[0]), "rank0_histograms_sha256": h.hexdigest()}
    if a.count:
        k = min(len(mine), 16)
        base = len(mine) + 8
        for l in range(k):
            ctx.ir_clear(base + l, n, 1)
        ctx.get_counters(reset=True)
        ctx.trace_listeners(prm(flags=_capi.RAR_FLAG_COUNT_TESTS | _capi.RAR_FLAG_COUNT_EXECUTED), mine[:k], base)
        c = ctx.get_counters()
        out["executed_tests_for_%d_listeners" % k] = c["nearest_tests"] + c["shadow_tests"]
        out["executed_tests_estimate_rank0"] = c["nearest_tests"] + c["shadow_tests"] * len(mine) / k
        out["executed_tests_per_s_rank0_estimate"] = out["executed_tests_estimate_rank0"] / (float(ms[0]) * 1e-3)
    if a.compare > 0:
        k = min(a.compare, len(mine))
        base = len(mine) + 32
        for l in range(k):
            ctx.ir_clear(base + l, n, 1)
        torch.cuda.synchronize()
        e0.record(stream)
        for l in range(k):
            ctx.trace(prm(listener=(float(mine[l, 0]), float(mine[l, 1]))), base + l)
        e1.record(stream)
        torch.cuda.synchronize()
        same = all(np.array_equal(ctx.ir_read_fixed(l, n), ctx.ir_read_fixed(base + l, n)) for l in range(k))
        out["unfused_ms_per_listener"] = e0.elapsed_time(e1) / k
        out["fused_ms_per_listener"] = float(ms[0]) / len(mine)
        out["fused_equals_unfused"] = bool(same)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.destroy()


if __name__ == "__main__":
    main()
