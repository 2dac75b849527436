"""Latency of the histogram all-reduce, peer-memory exchange kernel vs ncclAllReduce, under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_exchange.py [--iters 200]

Every rank fills its slot with rank-dependent words, runs `iters` back-to-back all-reduces per method between
two CUDA events, and rank 0 prints one JSON line per histogram size (us per call, max over ranks) after checking
that all methods agree bit for bit."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realisticaudioraytracing2d_b200 import _capi  # noqa: E402
from realisticaudioraytracing2d_b200.host import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = _capi.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    sizes = [48_000, 96_000 * 8, 96_000 * 64]
    ex = sharding.PeerExchange(ctx, max(sizes))
    rng = np.random.default_rng(1234 + rank)
    for words in sizes:
        ctx.ir_clear(0, words, 1)
        hist = sharding.DeviceHistogram(ctx, 0, dev).tensor
        mine = torch.from_numpy(rng.integers(-2**40, 2**40, words, dtype=np.int64)).to(dev)
        results, timing = {}, {}
        methods = {"nccl": lambda: dist.all_reduce(hist), "one_shot": lambda: ex.allreduce(0, _capi.RAR_EXCHANGE_ONE_SHOT),
                   "two_shot": lambda: ex.allreduce(0, _capi.RAR_EXCHANGE_TWO_SHOT)}
        for name, fn in methods.items():
            hist.copy_(mine)
            fn()
            torch.cuda.synchronize()
            results[name] = hist.clone()
            for _ in range(10):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.iters):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / args.iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            timing[name] = round(float(t), 2)
        ex.check()
        same = all(torch.equal(results["nccl"], results[k]) for k in ("one_shot", "two_shot"))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"world": world, "words": words, "bytes": words * 8, "us_per_call": timing,
                              "bit_identical": bool(int(flag))}), flush=True)
    ex.close()
    ctx.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
