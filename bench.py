#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (BASELINE.json: ray-segment tests/s, IR-build ms,
convolved samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one IR build of BASELINE config 2 (synthetic shoebox,
4 walls, 1 Mi rays x 32 bounces, 1 s IR at 48 kHz): clear the histogram slot, trace + deposit, and
-- for N > 1 -- one all-reduce of the int64 histogram (the library's peer-memory kernel).  Each rank traces its share
of a dispatch of N x 1 Mi rays (weak scaling): chunks rank, rank + N, ... of 2^14 contiguous ray ids.  The same line carries the convolution stage
(config 5: 256 streams x 10 s IRs per GPU, block 256) under "conv", a large-scene trace (config 3
geometry, reduced ray count) under "maze", the roofline of the dominant kernel and the CPU baseline.

Legs added for the multi-GPU record (N >= 1, same launch): `strong.c3` -- BASELINE config 3 geometry (10 000 walls,
8 bands, 64 bounces, brute force) with a FIXED total of 4 849 664 rays split over the N ranks in block-cyclic chunks of
contiguous ray ids, two-shot all-reduce of the 3.07 MB histogram -- and `strong.c2` -- config 2 with a fixed total of 1 Mi rays
(the latency regime).  Both compare the SHA-256 of the all-reduced histogram with the hash the CPU oracle minted for
the unsharded dispatch (tests/golden/strong_scaling.json): `parity_ok`.  `unsharded_equal` does the same for the weak
leg against a single-GPU trace of the whole dispatch.

`--impl reference` times the CPU oracle (the only CPU implementation of this path that exists: the
reference itself is HLSL compute run by Unity) on the host cores, on bounded samples of the same
workload.  It is the one place besides cpu_baseline where bench.py executes oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray_segment_tests_per_sec"
UNIT = "tests/s"
WORKLOAD = "config2: synthetic shoebox 10x6 m (4 walls), 1048576 rays x 32 bounces, 1 s IR @ 48 kHz, per GPU"
RAYS_PER_GPU = 1 << 20
BOUNCES = 32
FLOPS_PER_TEST = 20.0  # ~20 fp32 lane-ops per ray-segment test
CHUNK_LOG2 = 14        # N > 1: a rank traces chunks rank, rank + N, ... of 2^14 contiguous ray ids (rar_trace_interleaved)
IR_BINS = 48000


L2_NOTE = "flushed between timed steps (256 MiB write); per-step CUDA events summed"
EXCHANGE_PEER = ("all-reduce(sum,int64) of the histogram by the library's peer-memory kernel "
                 "(CUDA IPC over NVLink, one launch per rank)")
EXCHANGE_NCCL = "ncclAllReduce(sum,int64) of the histogram"


def _config(n_gpus: int, use_nccl: bool = False) -> dict:
    """The `config` object BOTH arms print, key for key and value for value (the driver compares them); what is
    specific to the CPU arm is said in its `config_note`."""
    return {"workload": WORKLOAD, "rays_total": RAYS_PER_GPU * n_gpus, "bounces": BOUNCES, "walls": 4, "ir_bins": IR_BINS,
            "l2": L2_NOTE, "exchange": "none" if n_gpus == 1 else (EXCHANGE_NCCL if use_nccl else EXCHANGE_PEER)}


def _strong_golden():
    with open(os.path.join(ROOT, "tests", "golden", "strong_scaling.json")) as f:
        return json.load(f)


def _traffic(kernel: str, key: str = "bytes"):
    """A per-launch figure of a kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[kernel][key]
    except Exception:
        return None


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _host_threads() -> int:
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1; ignore that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def _oracle_step_sample(O, scenes, n_gpus: int, rays: int, frame: int, threads: int = 0):
    """One bounded sample of the workload on the CPU oracle; returns (tests, seconds)."""
    threads = threads or _host_threads()
    sc = scenes.shoebox(ray_count=RAYS_PER_GPU * n_gpus, max_bounces=BOUNCES)
    P = O.make_params(source_x=sc.source[0], source_y=sc.source[1], listener_x=sc.listener[0], listener_y=sc.listener[1],
                      listener_radius=sc.listener_radius, speed_of_sound=sc.speed_of_sound, input_gain=sc.input_gain,
                      max_bounce_count=BOUNCES, rng_state_offset=frame, ray_count=sc.ray_count,
                      sample_rate=sc.sample_rate, impulse_length=sc.impulse_length, ray_begin=0, ray_end=rays)
    walls = np.ascontiguousarray(sc.walls).view(O.SEGMENT_DTYPE)
    t0 = time.perf_counter()
    r = O.trace(walls, P, n_threads=threads)
    dt = time.perf_counter() - t0
    return r.counters["nearest_tests"] + r.counters["shadow_tests"], dt


def run_reference(args):
    """The reference arm: the CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from realisticaudioraytracing2d_b200 import scenes
    O.build()
    cores = _host_threads()
    rays = RAYS_PER_GPU  # bounded sample: one GPU's share of the dispatch per step (a fraction of a second)
    for w in range(args.warmup):
        _oracle_step_sample(O, scenes, args.gpus, rays, 1000 + w)
    tests = secs = 0.0
    for k in range(args.steps):
        t, dt = _oracle_step_sample(O, scenes, args.gpus, rays, 1 + k)
        tests += t
        secs += dt
    value = tests / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.gpus, os.environ.get("RAR_BENCH_EXCHANGE", "peer") == "nccl"),
        "config_note": "CPU arm of the same workload: one host, no GPU, no exchange; each step is the bounded sample named in "
                       "cpu_baseline.sample (the `l2` and `exchange` entries of config describe the GPU arm it is compared with)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"rays [0,{rays}) of the dispatch x {BOUNCES} bounces per step, OpenMP over rays"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from realisticaudioraytracing2d_b200 import _capi, scenes
    from realisticaudioraytracing2d_b200.host import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    peaks, peaks_kind = _peaks()
    ctx = _capi.Context(local)
    # One explicit (non-default) stream carries both the library's kernels and torch's collectives, so
    # that CUDA events recorded on it bracket exactly the step's work.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    info = ctx.device_info()

    # ---- workload ------------------------------------------------------------------------------
    sc = scenes.shoebox(ray_count=RAYS_PER_GPU * world, max_bounces=BOUNCES)
    n_bins = sc.impulse_length
    ctx.set_walls(sc.walls)
    ctx.ir_clear(0, n_bins, 1)
    hist_t = sharding.DeviceHistogram(ctx, 0, dev).tensor

    def params(frame, flags=0):
        return _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain,
                                       BOUNCES, frame, sc.ray_count, 100, sc.sample_rate, n_bins, 1, 1.0, flags, 0, 0)

    def trace(frame, slot, flags=0):
        """This rank's share of the dispatch of N x 1 Mi rays: block-cyclic chunks of contiguous ray ids (one contiguous
        range per rank leaves the rank whose angular sector faces the listener 13 % behind the others)."""
        if world == 1:
            ctx.trace(params(frame, flags), slot)
        else:
            ctx.trace_interleaved(params(frame, flags), slot, rank, world, CHUNK_LOG2)

    # The all-reduce of the ray-range sharding: the library's own kernel over CUDA-IPC peer memory (the process
    # group only delivers the handles); RAR_BENCH_EXCHANGE=nccl selects ncclAllReduce on the same buffer instead.
    use_nccl = os.environ.get("RAR_BENCH_EXCHANGE", "peer") == "nccl"
    ex = None
    if world > 1 and not use_nccl:
        try:
            ex = sharding.PeerExchange(ctx, n_bins * 8)          # room for the 8-band histogram of the strong.c3 leg
        except RuntimeError as e:   # raised on every rank or on none: CUDA IPC is unavailable between these processes
            sys.stderr.write(f"[bench] {e}; using ncclAllReduce for the exchange\n")
            use_nccl = True

    def exchange(nccl=use_nccl):
        if world == 1:
            return
        if nccl:
            sharding.allreduce_histogram(hist_t)
        else:
            ex.allreduce(0)

    def step(frame, nccl=use_nccl):
        ctx.ir_clear(0, n_bins, 1)
        trace(frame, 0)
        exchange(nccl)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # tests per frame, counted by the kernel's own counters (checked against the oracle in tests/)
    frames = list(range(1, args.steps + 1))
    tests_total = 0
    ctx.get_counters(reset=True)
    for f in frames:
        ctx.ir_clear(0, n_bins, 1)
        trace(f, 0, _capi.RAR_FLAG_COUNT_TESTS)
    c = ctx.get_counters(reset=True)
    tests_total = c["nearest_tests"] + c["shadow_tests"]
    # ... and the tests the production kernel actually evaluates (shadow rays whose estimate cannot clear the
    # deposit threshold are skipped): the honest numerator of the roofline's "achieved".
    for f in frames:
        ctx.ir_clear(0, n_bins, 1)
        trace(f, 0, _capi.RAR_FLAG_COUNT_TESTS | _capi.RAR_FLAG_COUNT_EXECUTED)
    c = ctx.get_counters(reset=True)
    tests_executed = c["nearest_tests"] + c["shadow_tests"]

    for w in range(args.warmup):
        step(1000 + w)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for f in frames:
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step(f)
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms = ms  # the step is clear + one trace kernel (+ all-reduce)

    # the same steps with the other all-reduce, for comparison (and a bit-for-bit check of the two)
    exchange_compare = None
    if world > 1 and ex is not None:
        ex.check()
        other = []
        for f in frames:
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(f, not use_nccl)
            e1.record(stream)
            other.append((e0, e1))
        barrier()
        other_ms = sum(a.elapsed_time(b) for a, b in other)
        step(frames[0], True)
        torch.cuda.synchronize()
        want = hist_t.clone()
        step(frames[0], False)
        torch.cuda.synchronize()
        same = bool(torch.equal(want, hist_t))
        peer_ms, nccl_ms = (other_ms, ms) if use_nccl else (ms, other_ms)
        exchange_compare = [peer_ms, nccl_ms, float(same)]

    # ---- end-to-end through the C-ABI with host buffers --------------------------------------------
    # Every step uploads the scene from host memory and reads the finished IR back into the caller's host buffer.
    # Steps alternate between the ping and pong slots (RayTraceManager.cs:36) and the readback is asynchronous
    # (rar_ir_read_begin/_end, the analogue of AsyncGPUReadback), so step k+1's upload and trace are enqueued
    # while step k's result is still on its way; every step's result is collected inside the timed region.
    ir_host = [torch.empty(n_bins, dtype=torch.float32).pin_memory() for _ in range(2)]   # the caller's result buffers
    walls_host = np.ascontiguousarray(sc.walls)
    ctx.ir_clear(1, n_bins, 1)
    hist_ts = [hist_t, sharding.DeviceHistogram(ctx, 1, dev).tensor]

    def e2e_run(fr):
        pending = None
        for k, frame in enumerate(fr):
            s = k & 1
            ctx.set_walls(walls_host)                      # H2D: 40 B per wall
            ctx.ir_clear(s, n_bins, 1)
            trace(frame, s)
            if world > 1:
                if use_nccl:
                    sharding.allreduce_histogram(hist_ts[s])
                else:
                    ex.allreduce(s)
            t = ctx.ir_read_begin(s, n_bins)               # D2H: the float IR
            if pending is not None:
                ctx.ir_read_end(pending[0], n_bins, ir_host[pending[1]].data_ptr())
            pending = (t, s)
        ctx.ir_read_end(pending[0], n_bins, ir_host[pending[1]].data_ptr())

    e2e_run([2000 + w for w in range(max(2, min(args.warmup, 3)))])
    barrier()
    t0 = time.perf_counter()
    e2e_run(frames)
    barrier()
    e2e_s = time.perf_counter() - t0

    # max over ranks
    tt = torch.tensor([ms, e2e_s * 1e3, float(tests_total), float(tests_executed)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        if exchange_compare is not None:
            cmp_t = torch.tensor([exchange_compare[0], exchange_compare[1], -exchange_compare[2]], dtype=torch.float64, device=dev)
            dist.all_reduce(cmp_t, op=dist.ReduceOp.MAX)
            exchange_compare = {"peer_kernel_ms_per_step": float(cmp_t[0]) / len(frames),
                                "nccl_ms_per_step": float(cmp_t[1]) / len(frames),
                                "bit_identical": bool(float(cmp_t[2]) == -1.0)}
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, tests_all, tests_exec_all = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
    else:
        e2e_ms, tests_all, tests_exec_all = e2e_s * 1e3, float(tests_total), float(tests_executed)

    value = tests_all / (ms * 1e-3)
    e2e_value = tests_all / (e2e_ms * 1e-3)

    # ---- secondary measurements (reported by rank 0) ----------------------------------------------------
    extra = {}
    fp32_peak = ctx.measure_fp32_peak()
    env = dict(ctx=ctx, capi=_capi, scenes=scenes, torch=torch, dist=dist, sharding=sharding, stream=stream, flush=flush, dev=dev,
               rank=rank, world=world, ex=ex, use_nccl=use_nccl, barrier=barrier, fp32_peak=fp32_peak)
    try:
        extra["unsharded_equal"] = check_unsharded(env, sc, params, step, hist_t, n_bins)
    except Exception as e:  # keep the headline even if a secondary leg fails
        extra["unsharded_equal"] = {"error": str(e)}
    try:
        extra["strong"] = bench_strong(env)
    except Exception as e:
        extra["strong"] = {"error": str(e)}
    try:
        extra["listeners"] = bench_listeners(env)
    except Exception as e:
        extra["listeners"] = {"error": str(e)}
    try:
        extra["maze"] = bench_maze(ctx, _capi, scenes, torch, stream, flush, fp32_peak)
    except Exception as ex:  # keep the headline even if a secondary leg fails
        extra["maze"] = {"error": str(ex)}
    try:
        extra["config1"] = bench_config1(ctx, _capi, scenes, torch, stream)
    except Exception as ex:
        extra["config1"] = {"error": str(ex)}
    try:
        extra["conv"] = bench_conv(ctx, _capi, scenes, torch, stream, dev, peaks, peaks_kind, args, world, dist)
    except Exception as ex:
        extra["conv"] = {"error": str(ex)}
    try:
        extra["banded"] = bench_banded(env, peaks, peaks_kind)
    except Exception as ex:
        extra["banded"] = {"error": str(ex)}
    try:
        extra["clip_prep"] = bench_clip_prep(ctx, _capi, torch, stream, dev, peaks, peaks_kind)
    except Exception as ex:
        extra["clip_prep"] = {"error": str(ex)}

    clocks = sampler.stop() if rank == 0 else None   # sampled across the timed region and the secondary legs

    cpu = None
    if rank == 0 and world == 1:
        from oracle import oracle as O
        O.build()
        t_acc = n_acc = 0.0
        rays = RAYS_PER_GPU
        k = 0
        while t_acc < 10.0 and k < 200:
            t, dt = _oracle_step_sample(O, scenes, 1, rays, 1 + k)
            n_acc += t
            t_acc += dt
            k += 1
        cpu = {"value": n_acc / t_acc, "unit": UNIT, "cores": _host_threads(), "kind": "port",
               "sample": f"{k} x rays [0,{rays}) of the same dispatch x {BOUNCES} bounces ({t_acc:.1f} s of CPU work)"}

    if rank == 0:
        per_launch_tests = tests_exec_all / world / len(frames)
        ach = per_launch_tests * FLOPS_PER_TEST / (kernel_ms / len(frames) * 1e-3) / 1e12
        peak = fp32_peak / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / len(frames), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": _config(world, use_nccl),
            "value_executed": tests_exec_all / (ms * 1e-3),
            "ir_build_ms": ms / len(frames),
            "tests_per_step": tests_all / len(frames),
            "tests_executed_per_step": tests_exec_all / len(frames),
            "counting_rule": "value counts the intersect() evaluations the reference algorithm performs for this input "
                             "(SURVEY 8d: nearest-hit tests + early-exit-aware shadow tests, counted by the kernel and checked "
                             "against the oracle); the kernel evaluates fewer (tests_executed_per_step) because it skips shadow "
                             "rays whose estimate cannot pass the deposit threshold; value_executed and roofline.achieved use the "
                             "executed count -- quote value_executed as the absolute rate of work done",
            "wall_ms_per_step_incl_flush": t_wall / len(frames) * 1e3,
            "roofline": {"bound": "fp32-issue", "achieved": ach, "peak": peak, "unit": "Tlaneop/s", "frac": ach / peak,
                         "traffic": _traffic("trace_deposit_kernel"), "kernel": "trace_deposit_kernel",
                         "ncu_issue_slot_utilisation_pct": _traffic("trace_deposit_kernel", "sm_inst_issued_pct_of_peak"),
                         "note": f"{FLOPS_PER_TEST:g} fp32 lane-ops per ray-segment test (SURVEY 8d) x tests EXECUTED per launch "
                                 "(value_executed is the same count per second, whole job) / CUDA-event duration; peak = FFMA issue rate measured in this run (rar_measure_fp32_peak); "
                                 "HBM traffic is negligible for this kernel (scene 160 B, histogram 384 KB, L2 resident)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(walls_host.nbytes + 88),
                    "d2h_bytes_per_step": int(n_bins * 4), "ms_per_step": e2e_ms / len(frames)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "device": info,
            "peaks": {"kind": peaks_kind, "hbm_gbs": peaks.get("hbm_gbs"), "fp32_laneops_per_s_measured": fp32_peak},
        }
        if exchange_compare:
            line["exchange_compare"] = exchange_compare
        line.update(extra)
        _emit(line)
    if world > 1:
        if ex is not None:
            ex.close()
        dist.barrier()
        dist.destroy_process_group()
    ctx.destroy()


def _sha(a) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _all_true(env, flag: bool) -> bool:
    """Logical AND of `flag` over the ranks."""
    if env["world"] == 1:
        return bool(flag)
    t = env["torch"].tensor([1.0 if flag else 0.0], dtype=env["torch"].float64, device=env["dev"])
    env["dist"].all_reduce(t, op=env["dist"].ReduceOp.MIN)
    return bool(float(t[0]) == 1.0)


def _max_over_ranks(env, values):
    if env["world"] == 1:
        return [float(v) for v in values]
    t = env["torch"].tensor([float(v) for v in values], dtype=env["torch"].float64, device=env["dev"])
    env["dist"].all_reduce(t, op=env["dist"].ReduceOp.MAX)
    return [float(v) for v in t]


def _sum_over_ranks(env, values):
    if env["world"] == 1:
        return [float(v) for v in values]
    t = env["torch"].tensor([float(v) for v in values], dtype=env["torch"].float64, device=env["dev"])
    env["dist"].all_reduce(t, op=env["dist"].ReduceOp.SUM)
    return [float(v) for v in t]


def check_unsharded(env, sc, params, step, hist_t, n_bins):
    """Multi-GPU parity of the weak-scaled headline leg: one more step on a fixed frame (every rank traces its own
    ray range, then the all-reduce), after which rank 0 ALONE traces all N ranges into a second slot and compares
    the two histograms; the ranks' copies of the all-reduced histogram are compared through their hashes."""
    ctx, capi, torch, dist, world, rank = env["ctx"], env["capi"], env["torch"], env["dist"], env["world"], env["rank"]
    frame = 777
    step(frame)
    torch.cuda.synchronize()
    mine = ctx.ir_read_fixed(0, n_bins)
    equal = True
    if rank == 0:
        ctx.ir_clear(1, n_bins, 1)
        ctx.trace(params(frame), 1)                       # the whole dispatch of N x 1 Mi rays, unsharded, on this GPU alone
        equal = bool(np.array_equal(ctx.ir_read_fixed(1, n_bins), mine)) and bool(mine.any())
    digests = [_sha(mine)]
    if world > 1:
        digests = [None] * world
        dist.all_gather_object(digests, _sha(mine))
    return {"equal": _all_true(env, equal), "ranks_identical": len(set(digests)) == 1, "hist_sha256": digests[0],
            "what": f"all-reduced histogram of {world} x {RAYS_PER_GPU} rays (frame {frame}, block-cyclic shards) == rank 0 tracing the whole dispatch alone"}


STRONG_LEGS = {   # must equal tests/golden/make_strong_golden.py LEGS
    "c3": dict(kind="maze", walls=10000, rays=148 * 1024 * 32, bounces=64, bands=8, frame=1, reps=3, slot=4),
    "c2": dict(kind="shoebox", walls=4, rays=1 << 20, bounces=32, bands=1, frame=1, reps=10, slot=5),
}


def bench_strong(env):
    """STRONG scaling: a fixed dispatch split over the N ranks in block-cyclic chunks of ray ids, one all-reduce of the
    histogram per step.  c3 = BASELINE config 3 geometry (10 000-wall maze, 8 bands, 64 bounces, brute force) with
    4 849 664 rays in total; c2 = config 2 with 1 Mi rays in total (latency regime).  The SHA-256 of the all-reduced
    histogram is compared with the one the CPU oracle produced for the unsharded dispatch."""
    ctx, capi, scenes, torch, dist = env["ctx"], env["capi"], env["scenes"], env["torch"], env["dist"]
    world, rank, stream, flush, ex, use_nccl = env["world"], env["rank"], env["stream"], env["flush"], env["ex"], env["use_nccl"]
    golden = _strong_golden()
    out = {}
    for name, leg in STRONG_LEGS.items():
        g = golden.get(name)
        if g is None or any(g[k] != leg[k] for k in ("kind", "walls", "rays", "bounces", "bands", "frame")):
            out[name] = {"error": "tests/golden/strong_scaling.json does not describe this dispatch"}
            continue
        if leg["kind"] == "maze":
            sc = scenes.maze(n_segments=leg["walls"], ray_count=leg["rays"], max_bounces=leg["bounces"], bands=8)
        else:
            sc = scenes.shoebox(ray_count=leg["rays"], max_bounces=leg["bounces"])
        n, bands, slot = sc.impulse_length, leg["bands"], leg["slot"]
        ctx.set_walls(sc.walls)
        if bands > 1:
            ctx.set_wall_band_absorption(sc.band_absorption)
        ctx.ir_clear(slot, n, bands)
        hist_t = env["sharding"].DeviceHistogram(ctx, slot, env["dev"]).tensor if (world > 1 and use_nccl) else None

        def prm(flags=0, b=0, e=0, bounces=leg["bounces"]):
            return capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain,
                                          bounces, leg["frame"], leg["rays"], 0, sc.sample_rate, n, bands, 1.0, flags, b, e)

        def trace_share(flags=0):
            if world == 1:
                ctx.trace(prm(flags), slot)
            else:
                ctx.trace_interleaved(prm(flags), slot, rank, world, CHUNK_LOG2)

        def one_step():
            ctx.ir_clear(slot, n, bands)
            trace_share()
            if world > 1:
                if use_nccl:
                    env["sharding"].allreduce_histogram(hist_t)
                else:
                    ex.allreduce(slot)

        ctx.trace(prm(b=0, e=4096, bounces=2), slot)                      # first launch of this kernel variant
        one_step()                                                        # warm-up at full size
        env["barrier"]()
        evs = []
        for _ in range(leg["reps"]):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one_step()
            e1.record(stream)
            evs.append((e0, e1))
        env["barrier"]()
        ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        digest = _sha(ctx.ir_read_fixed(slot, n * bands))
        # the tests the production kernel evaluates on this rank's share (roofline numerator)
        ctx.get_counters(reset=True)
        ctx.ir_clear(slot, n, bands)
        trace_share(capi.RAR_FLAG_COUNT_TESTS | capi.RAR_FLAG_COUNT_EXECUTED)
        c = ctx.get_counters(reset=True)
        (ms,) = _max_over_ranks(env, [ms])
        (executed,) = _sum_over_ranks(env, [c["nearest_tests"] + c["shadow_tests"]])
        tests = g["counters"]["nearest_tests"] + g["counters"]["shadow_tests"]
        parity = _all_true(env, digest == g["hist_sha256"])
        res = {"workload": f"{sc.name}: {leg['walls']} walls, {leg['rays']} rays in total x {leg['bounces']} bounces, {bands} band(s), "
                           f"brute force; ray ids split over {world} GPU(s) in block-cyclic chunks of {1 << CHUNK_LOG2}, "
                           f"all-reduce of {n * bands * 8} bytes",
               "n_gpus": world, "ir_build_ms": ms, "steps": leg["reps"], "tests": tests, "tests_executed": executed,
               "tests_per_s": tests / (ms * 1e-3), "tests_executed_per_s": executed / (ms * 1e-3),
               "hist_sha256": digest, "golden_sha256": g["hist_sha256"], "parity_ok": parity,
               "golden": "tests/golden/strong_scaling.json (minted by the CPU oracle, unsharded)"}
        if env["fp32_peak"]:
            ach = executed / world * FLOPS_PER_TEST / (ms * 1e-3) / 1e12
            res["roofline"] = {"bound": "fp32-issue", "achieved": ach, "peak": env["fp32_peak"] / 1e12, "unit": "Tlaneop/s",
                               "frac": ach / (env["fp32_peak"] / 1e12), "traffic": None, "kernel": "trace_deposit_kernel",
                               "note": "per GPU: tests executed / N x 20 nominal lane-ops / step time (exchange included); the packed-FP32 "
                                       "wall scans need ~10 issue slots per test, so the fraction can exceed 1 (see maze.roofline.note)"}
        out[name] = res
    return out


def bench_listeners(env):
    """BASELINE config 4 (batched auralisation), one GPU's share: 128 listener positions (this rank's contiguous
    range of the 32 x 32 grid) in the 2 000-wall scene, 4 Mi rays x 5 bounces each, fused listener kernel, one IR
    slot per listener, no collective (listeners shard across GPUs)."""
    ctx, capi, scenes, torch, stream = env["ctx"], env["capi"], env["scenes"], env["torch"], env["stream"]
    world, rank = env["world"], env["rank"]
    per_gpu, rays, bounces, walls, first = 128, 1 << 22, 5, 2000, 100
    sc = scenes.maze(n_segments=walls, ray_count=rays, max_bounces=bounces, bands=8)
    n = sc.impulse_length
    gx, gy = np.meshgrid(np.linspace(8, 92, 32), np.linspace(8, 92, 32))
    grid = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)
    start = (256 + per_gpu * rank) % 1024          # rank 0 takes grid rows 8-11, the ones around the source
    mine = grid[start:start + per_gpu]
    ctx.set_walls(sc.walls)

    def clear():
        for l in range(per_gpu):
            ctx.ir_clear(first + l, n, 1)

    def prm(flags=0, r=rays, listener=(0.0, 0.0)):
        return capi.make_trace_params(sc.source, listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain, bounces, 1,
                                      r, 0, sc.sample_rate, n, 1, 1.0, flags, 0, 0)
    clear()
    ctx.trace_listeners(prm(r=4096), mine, first)               # warm-up
    clear()
    env["barrier"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.trace_listeners(prm(), mine, first)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # the busiest listener again as a plain single-listener trace: identical histogram (most listeners of the maze
    # hear nothing within 5 bounces)
    heads = [ctx.ir_read_fixed(first + l, n) for l in range(per_gpu)]
    k = int(np.argmax([np.count_nonzero(h) for h in heads]))
    reached = int(sum(1 for h in heads if h.any()))
    ctx.ir_clear(first + per_gpu, n, 1)
    ctx.trace(prm(listener=(float(mine[k, 0]), float(mine[k, 1]))), first + per_gpu)
    single = ctx.ir_read_fixed(first + per_gpu, n)
    same = bool(np.array_equal(single, heads[k]))        # (a rank whose listeners all hear nothing compares two empty histograms)
    if not same:
        sys.stderr.write(f"[bench] listeners: listener {k} fused nonzero {np.count_nonzero(heads[k])} sum {int(heads[k].sum())}; "
                         f"single nonzero {np.count_nonzero(single)} sum {int(single.sum())}\n")
    clear()
    ctx.get_counters(reset=True)
    ctx.trace_listeners(prm(capi.RAR_FLAG_COUNT_TESTS | capi.RAR_FLAG_COUNT_EXECUTED), mine, first)
    c = ctx.get_counters(reset=True)
    executed = c["nearest_tests"] + c["shadow_tests"]
    (ms,) = _max_over_ranks(env, [ms])
    executed_all, reached_all = _sum_over_ranks(env, [executed, reached])
    res = {"workload": f"config4 share: {per_gpu} listeners per GPU x {rays} rays x {bounces} bounces, {walls}-wall scene, "
                       f"{n} bins per listener, fused listener kernel (each ray traced once per launch)",
           "n_gpus": world, "listeners_total": per_gpu * world, "ms": ms, "ms_per_listener": ms / per_gpu,
           "tests_executed": executed_all, "tests_executed_per_s": executed_all / (ms * 1e-3),
           "listeners_with_arrivals": int(reached_all), "fused_equals_single_listener_trace": _all_true(env, same)}
    if env["fp32_peak"]:
        ach = executed * FLOPS_PER_TEST / (ms * 1e-3) / 1e12
        res["roofline"] = {"bound": "fp32-issue", "achieved": ach, "peak": env["fp32_peak"] / 1e12, "unit": "Tlaneop/s",
                           "frac": ach / (env["fp32_peak"] / 1e12), "traffic": None, "kernel": "trace_listeners_kernel"}
    return res


def bench_maze(ctx, _capi, scenes, torch, stream, flush, fp32_peak=None):
    """Config 3 geometry (10 000 walls, 8 bands) with a reduced ray count: the regime where the inner loop
    over walls dominates and the shared-memory staging matters."""
    sc = scenes.maze(n_segments=10000, ray_count=148 * 1024 * 3, max_bounces=16, bands=8)  # 3 full waves of 1024-thread CTAs
    n = sc.impulse_length
    out = {}
    ctx.set_walls(sc.walls)
    ctx.set_wall_band_absorption(sc.band_absorption)
    for bands in (1, 8):
        def prm(flags=0):
            return _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain,
                                           sc.max_bounces, 1, sc.ray_count, 0, sc.sample_rate, n, bands, 1.0, flags, 0, 0)
        ctx.ir_clear(2, n, bands)
        ctx.get_counters(reset=True)
        ctx.trace(prm(_capi.RAR_FLAG_COUNT_TESTS), 2)
        c = ctx.get_counters(reset=True)
        tests = c["nearest_tests"] + c["shadow_tests"]
        ctx.ir_clear(2, n, bands)
        ctx.trace(prm(_capi.RAR_FLAG_COUNT_TESTS | _capi.RAR_FLAG_COUNT_EXECUTED), 2)
        c = ctx.get_counters(reset=True)
        tests_exec = c["nearest_tests"] + c["shadow_tests"]   # what the production kernel evaluates (skipped shadow rays)
        def timed(flags, reps=5):
            """Mean launch time over `reps` launches (after one untimed launch), L2 flushed before each."""
            ctx.ir_clear(2, n, bands)
            ctx.trace(prm(flags), 2)
            tot = 0.0
            for _ in range(reps):
                flush.zero_()
                ctx.ir_clear(2, n, bands)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.trace(prm(flags), 2)
                e1.record(stream)
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / reps
        best = timed(0)
        out[f"bands{bands}"] = {"tests": tests, "ms": best, "launches_averaged": 5, "tests_per_s": tests / (best * 1e-3),
                                "tests_executed_per_s": tests_exec / (best * 1e-3)}
        if fp32_peak:
            # the wall-dominated regime: nearly every instruction of the kernel belongs to a wall test, so the
            # 20-op rule of SURVEY 8(d) describes the kernel; achieved uses the tests EXECUTED
            ach = tests_exec * FLOPS_PER_TEST / (best * 1e-3) / 1e12
            out[f"bands{bands}"]["tests_executed"] = tests_exec
            out[f"bands{bands}"]["roofline"] = {"bound": "fp32-issue", "achieved": ach, "peak": fp32_peak / 1e12,
                                               "unit": "Tlaneop/s", "frac": ach / (fp32_peak / 1e12), "traffic": None,
                                               "kernel": "trace_deposit_kernel (1024-thread cooperative variant, packed FP32 wall scans)",
                                               "note": "20 nominal lane-ops per executed test (SURVEY 8d) against the scalar FFMA issue rate; "
                                                       "since the wall scans use packed FP32 (FFMA2: two results per issue slot, same results/s) "
                                                       "the kernel needs ~10 issue slots and 13 FP32 results per test, so this fraction can exceed 1",
                                               "fp32_results_frac": tests_exec * 13.0 / (best * 1e-3) / fp32_peak}
        # the same IR through the optional uniform grid (identical histogram, far fewer tests evaluated)
        ref = ctx.ir_read_fixed(2, n * bands)
        gbest = timed(_capi.RAR_FLAG_USE_GRID)
        out[f"bands{bands}"]["grid_ms"] = gbest
        out[f"bands{bands}"]["grid_identical"] = bool(np.array_equal(ctx.ir_read_fixed(2, n * bands), ref))
    out["workload"] = f"config3 geometry: 10000-wall maze, {sc.ray_count} rays x 16 bounces (reduced ray count)"
    return out


def bench_config1(ctx, _capi, scenes, torch, stream):
    """Config 1, the reference's own real-time case: bundled SmollRoom, 15 000 rays x 5 bounces per frame,
    a 0.1 s chunk (4800 samples) convolved with the 1.5 s IR (72 000 taps) through the host-buffer API, and the
    whole 42 624-sample clip in one call (BakeAudio)."""
    sc = scenes.smoll_room()
    n = sc.impulse_length
    ctx.set_walls(sc.walls)
    ctx.ir_clear(3, n, 1)

    def prm(frame):
        return _capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain,
                                       sc.max_bounces, frame, sc.ray_count, 100, sc.sample_rate, n)
    for f in range(3):
        ctx.trace(prm(f + 1), 3)
    torch.cuda.synchronize()
    reps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for f in range(reps):
        ctx.trace(prm(10 + f), 3)
    e1.record(stream)
    torch.cuda.synchronize()
    frame_ms = e0.elapsed_time(e1) / reps
    ctx.trace_frames(prm(100), 3, 10)                     # ten frames in one launch (rar_trace_frames)
    torch.cuda.synchronize()
    e0.record(stream)
    for f in range(10):
        ctx.trace_frames(prm(200 + 10 * f), 3, 10)
    e1.record(stream)
    torch.cuda.synchronize()
    batched_ms = e0.elapsed_time(e1) / 100
    ctx.get_counters(reset=True)
    ctx.ir_clear(4, n, 1)
    ctx.trace(_capi.make_trace_params(sc.source, sc.listener, sc.listener_radius, sc.speed_of_sound, sc.input_gain, sc.max_bounces,
                                      10, sc.ray_count, 100, sc.sample_rate, n, 1, 1.0, _capi.RAR_FLAG_COUNT_TESTS), 4)
    c = ctx.get_counters(reset=True)
    tests_per_frame = c["nearest_tests"] + c["shadow_tests"]      # frame 10, counted by the kernel
    clip = scenes.synthetic_clip()
    chunk = clip[:4800]
    # one whole frame of the reference's real-time loop through host buffers (RayTraceManager.cs:50-53,64-123):
    # wall upload -> clear -> trace -> 4800-sample chunk convolved with the fresh IR -> output back on the host
    walls_host = np.ascontiguousarray(sc.walls)

    def frame(f):
        ctx.set_walls(walls_host)
        ctx.ir_clear(3, n, 1)
        ctx.trace(prm(f), 3)
        return ctx.convolve(3, chunk, 1, n)
    for f in range(3):
        frame(300 + f)
    t0 = time.perf_counter()
    for f in range(30):
        frame(400 + f)
    frame_e2e_ms = (time.perf_counter() - t0) / 30 * 1e3
    for f in range(reps + 3):
        ctx.trace(prm(500 + f), 3)
    ctx.convolve(3, chunk, reps + 3, n)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.convolve(3, chunk, reps + 3, n)
    chunk_ms = (time.perf_counter() - t0) / 20 * 1e3
    ctx.convolve(3, clip, reps + 3, n)                      # first call grows the ticket's pinned/device buffers
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.convolve(3, clip, reps + 3, n)
    clip_ms = (time.perf_counter() - t0) / 10 * 1e3
    return {"workload": "config1: SmollRoom 20 walls, 15000 rays x 5 bounces per frame; 1.5 s IR; 4800-sample chunk / 42624-sample clip",
            "ir_build_ms_per_frame": frame_ms, "ir_build_ms_per_frame_batched_x10": batched_ms, "tests_per_frame": tests_per_frame,
            "tests_per_s": tests_per_frame / (frame_ms * 1e-3),
            "frame_e2e_ms": frame_e2e_ms, "frame_e2e": "host walls -> clear -> trace -> 4800-sample chunk convolved with the fresh 72000-tap IR -> "
                                                        "76800 output samples on the host; wall clock per frame, blocking",
            "chunk_convolve_e2e_ms": chunk_ms, "chunk_realtime_factor": 100.0 / chunk_ms,
            "clip_convolve_e2e_ms": clip_ms, "clip_samples_per_s": (len(clip) + n) / (clip_ms * 1e-3)}


def bench_banded(env, peaks, peaks_kind):
    """banded responses into the streaming convolver.  16 streams, each taking its response from its own
    banded slot of 480 000 bins x 8 bands (a 10 s IR per band, 30.7 MB of Q23.40 words per slot, 491 MB in all, > L2):
    filter-bank synthesis (band_synth16_kernel) followed by the partition spectra (ir_spectra16_kernel), per stream."""
    ctx, capi, scenes, torch, stream = env["ctx"], env["capi"], env["scenes"], env["torch"], env["stream"]
    S, n, bands, first = 16, 480000, 8, 300
    cv = capi.Convolver(ctx, S, 256, n)
    t = np.arange(n, dtype=np.float32) / np.float32(48000)
    for st in range(S):
        base = np.abs(scenes.decaying_noise_ir(n, seed=40 + st, decay_s=3.0))
        ir = base[:, None] * np.exp(-t[:, None] * np.arange(bands, dtype=np.float32)[None, :] * np.float32(0.4))
        ctx.ir_write(first + st, ir.astype(np.float32).ravel(), bands=bands)
    slots, ones = np.arange(first, first + S, dtype=np.int32), np.ones(S, dtype=np.int32)
    cv.set_irs_from_slots(0, slots, ones)                 # builds the band filters, first launches
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        cv.set_irs_from_slots(0, slots, ones)             # one synthesis launch + one spectra launch for the 16 streams
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # a unit impulse into stream 3 must come back as the first block of that slot's synthesised response
    x = np.zeros((S, 256), np.float32)
    x[3, 0] = 1.0
    y = cv.process(x)
    want = ctx.synthesize_ir(first + 3, 256)
    err = float(np.linalg.norm(y[3] - want) / max(np.linalg.norm(want), 1e-30))
    cv.destroy()
    n_part = (n + 255) // 256
    nbytes = S * (n * bands * 8 + 2 * n * 4 + n_part * 2048)      # histogram read, response write + read, spectra write
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"workload": f"{S} banded slots x {n} bins x {bands} bands -> filter-bank synthesis -> partition spectra of {S} streams",
            "ms": ms, "ms_per_stream": ms / S, "responses_per_s": S / (ms * 1e-3), "impulse_check_rel_l2": err,
            "gpu_launches_per_step": 2,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": _traffic("band_synth16_kernel"), "kernel": "band_synth16_kernel", "peak_kind": peaks_kind,
                         "algorithmic_bytes_per_launch": S * (n * bands * 8 + n * 4),
                         "note": "bytes of the whole step (response clear, synthesis, spectra) over its time; the synthesis "
                                 "kernel reads the 8-byte histogram words once and adds into the 4-byte response"}}


def bench_clip_prep(ctx, _capi, torch, stream, dev, peaks, peaks_kind):
    """LoadSample (mono mix + linear resample) for a batch of clips resident in HBM.
    256 stereo clips of 10 s at 44.1 kHz -> mono 48 kHz; working set 1.39 GB (> L2)."""
    n_clips, samples, ch, freq, rate = 256, 441000, 2, 44100, 48000
    n_out = _capi.prepared_length(samples, freq, rate)
    g = torch.Generator(device=dev).manual_seed(3)
    raw = torch.rand((n_clips, samples, ch), generator=g, device=dev) * 2 - 1
    out = torch.empty((n_clips, n_out), device=dev)
    for _ in range(2):
        ctx.prepare_clips_device(raw.data_ptr(), samples, ch, freq, rate, n_clips, out.data_ptr(), n_out)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.prepare_clips_device(raw.data_ptr(), samples, ch, freq, rate, n_clips, out.data_ptr(), n_out)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = raw.numel() * 4 + out.numel() * 4
    peak = peaks.get("hbm_gbs") or 6550.0
    ach = nbytes / (ms * 1e-3) / 1e9
    res = {"workload": f"{n_clips} stereo clips x {samples} samples @ {freq} Hz -> mono {n_out} samples @ {rate} Hz, resident in HBM",
           "ms": ms, "output_samples_per_s": n_clips * n_out / (ms * 1e-3),
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": _traffic("prepare_clips_kernel"),
                        "kernel": "prepare_clips_kernel", "peak_kind": peaks_kind, "algorithmic_bytes_per_launch": nbytes}}
    del raw, out
    return res


def bench_conv(ctx, _capi, scenes, torch, stream, dev, peaks, peaks_kind, args, world, dist):
    """Config 5: 256 concurrent 48 kHz streams per GPU x 10 s IRs, uniformly partitioned overlap-save, block 256."""
    S, B, n_ir = 256, 256, 480000
    cv = _capi.Convolver(ctx, S, B, n_ir)
    base = [scenes.decaying_noise_ir(n_ir, seed=100 + k, decay_s=3.0) for k in range(8)]
    irs = np.stack([np.roll(base[s % 8], s // 8) for s in range(S)])      # 256 distinct IRs from 8 seeded ones
    t0 = time.perf_counter()
    cv.set_irs(0, irs)                                   # one call: pinned staging in halves, no stream synchronisation
    load_call_ms = (time.perf_counter() - t0) * 1e3
    ctx.sync()
    load_ms = (time.perf_counter() - t0) * 1e3
    del irs
    g = torch.Generator(device="cpu").manual_seed(1)
    x_host = (torch.rand((S, B), generator=g) * 2 - 1).pin_memory()
    y_host = torch.empty((S, B)).pin_memory()
    x_dev = x_host.to(dev)
    y_dev = torch.empty_like(x_dev)
    steps, warm = max(args.steps, 20), max(args.warmup, 3)
    for _ in range(warm):
        cv.process_device(x_dev.data_ptr(), y_dev.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        cv.process_device(x_dev.data_ptr(), y_dev.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        cv.process_host_ptr(x_host.data_ptr(), y_host.data_ptr())
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    bytes_per_block = cv.bytes_per_block()
    cv.destroy()
    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    sps = S * B * world / (ms * 1e-3)
    gbs = bytes_per_block / (ms * 1e-3) / 1e9
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return {
        "workload": f"config5: {S} streams per GPU x {n_ir}-tap IRs, block {B} (1875 partitions), working set "
                    f"{bytes_per_block / 1e9:.2f} GB per step per GPU (> L2, no flush needed)",
        "samples_per_s": sps, "ms_per_block": ms, "realtime_48k_streams": sps / 48000.0,
        "e2e_samples_per_s": S * B * world / (e2e_ms * 1e-3), "e2e_ms_per_block": e2e_ms,
        "h2d_bytes_per_step": S * B * 4, "d2h_bytes_per_step": S * B * 4, "gpu_launches_per_step": 3,
        "load_256_responses_ms": load_ms, "load_256_responses_call_returns_after_ms": load_call_ms,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "traffic": _traffic("stream_cmac_kernel"), "kernel": "stream_cmac_kernel", "peak_kind": peaks_kind,
                     "algorithmic_bytes_per_launch": bytes_per_block},
    }


_JSON_FD = None


def _claim_stdout():
    """Libraries (NCCL: "NCCL version ...") write to fd 1 from C.  The contract is ONE JSON line on stdout, so
    everything else is sent to stderr and the JSON line goes to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
