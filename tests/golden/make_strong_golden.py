#!/usr/bin/env python
"""Mints tests/golden/strong_scaling.json: the SHA-256 of the complete (unsharded) histogram of the two
strong-scaling legs of bench.py, computed by the CPU ORACLE (oracle/rar_oracle.c), plus its test counters.

`bench.py --gpus N` traces the same dispatches split over N GPUs by contiguous ray-id range, all-reduces the
int64 histograms and compares the hash of the result with the one committed here (`parity_ok`): multi-GPU parity
against the oracle in the driver-run record, without executing the oracle on the GPU box.

    python tests/golden/make_strong_golden.py [--threads 7] [--only c2|c3]

The config-3 leg is ~3.8e12 literal intersect() evaluations (about an hour on 8 host cores), so it is traced in
ray-range pieces with the partial histograms kept under tests/golden/_partial/ (git-ignored): the run can be
interrupted and resumed.  Integer histograms add exactly, so the pieces sum to the unsharded result.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from realisticaudioraytracing2d_b200 import scenes  # noqa: E402
from tests.common import oracle_params, oracle_walls, trace_kwargs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "strong_scaling.json")
PARTIAL = os.path.join(HERE, "_partial")

# The two dispatches.  bench.py builds the same scenes from these numbers (STRONG_LEGS there must stay equal).
LEGS = {
    # BASELINE config 3 geometry with a FIXED total of 148 x 1024 x 32 = 4 849 664 rays: four full waves of
    # 1024-thread CTAs per GPU at N = 8, thirty-two at N = 1
    "c3": dict(kind="maze", walls=10000, rays=148 * 1024 * 32, bounces=64, bands=8, frame=1, pieces=256),
    # BASELINE config 2 with a fixed total of 1 Mi rays: the latency regime (the whole dispatch takes ~0.5 ms on one GPU)
    "c2": dict(kind="shoebox", walls=4, rays=1 << 20, bounces=32, bands=1, frame=1, pieces=1),
}


def build_scene(leg):
    if leg["kind"] == "maze":
        return scenes.maze(n_segments=leg["walls"], ray_count=leg["rays"], max_bounces=leg["bounces"], bands=8)
    return scenes.shoebox(ray_count=leg["rays"], max_bounces=leg["bounces"])


def mint(name, leg, threads):
    sc = build_scene(leg)
    n_words = sc.impulse_length * leg["bands"]
    total = np.zeros(n_words, dtype=np.int64)
    counters = {}
    os.makedirs(PARTIAL, exist_ok=True)
    rays, pieces = leg["rays"], leg["pieces"]
    assert rays % 64 == 0, "the dispatch must not round up (Raytrace2D.compute:49-52)"
    t_all = time.time()
    for k in range(pieces):
        lo, hi = rays * k // pieces, rays * (k + 1) // pieces
        path = os.path.join(PARTIAL, f"{name}_{rays}_{leg['bounces']}_{k:03d}_of_{pieces}.npz")
        if os.path.exists(path):
            z = np.load(path)
            hist, ctr = z["hist"], {k2: int(v) for k2, v in zip(z["names"], z["values"])}
        else:
            kw = trace_kwargs(sc, bands=leg["bands"], rng_state_offset=leg["frame"], max_bounce_count=leg["bounces"],
                              ray_begin=lo, ray_end=hi)
            t0 = time.time()
            r = O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw),
                        band_abs=sc.band_absorption if leg["bands"] > 1 else None, n_threads=threads)
            hist, ctr = r.hist, {k2: int(v) for k2, v in r.counters.items()}
            np.savez_compressed(path, hist=hist, names=np.array(list(ctr.keys())), values=np.array(list(ctr.values()), np.int64))
            print(f"{name} piece {k + 1}/{pieces}: rays [{lo},{hi}) {time.time() - t0:.1f} s", flush=True)
        total += hist
        for k2, v in ctr.items():
            counters[k2] = counters.get(k2, 0) + v
    return {
        "scene": sc.name, **{k: leg[k] for k in ("kind", "walls", "rays", "bounces", "bands", "frame")},
        "impulse_length": sc.impulse_length, "n_words": int(n_words),
        "hist_sha256": hashlib.sha256(total.tobytes()).hexdigest(),
        "hist_sum": int(total.sum()), "hist_nonzero": int(np.count_nonzero(total)),
        "counters": counters,
        "minted_by": "oracle/rar_oracle.c (CPU restatement), unsharded sum of ray-range pieces",
        "oracle_seconds": round(time.time() - t_all, 1),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    O.build()
    out = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            out = json.load(f)
    for name, leg in LEGS.items():
        if a.only and a.only != name:
            continue
        out[name] = mint(name, leg, a.threads)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
            f.write("\n")
        print(name, out[name]["hist_sha256"][:16], out[name]["counters"], flush=True)


if __name__ == "__main__":
    main()
