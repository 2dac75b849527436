#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/.

The reference (HLSL compute + Unity C#) cannot run in this image and ships no golden vectors, so these
are MINTED FROM THE ORACLE RESTATEMENT (oracle/rar_oracle.c), not inherited from the reference: they pin
the oracle against regressions and give the GPU tests a check that does not execute the oracle.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from realisticaudioraytracing2d_b200 import scenes  # noqa: E402
from tests.common import oracle_params, oracle_walls, trace_kwargs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "smoll_frame1": lambda: (scenes.smoll_room(), {}),
    "big_room_frame3": lambda: (scenes.big_room(), {"rng_state_offset": 3}),
    "shoebox_4096x32": lambda: (scenes.shoebox(ray_count=4096, max_bounces=32), {}),
    "shoebox_scatter_transmit": lambda: (scenes.shoebox(ray_count=4096, max_bounces=16, scattering=0.35, transmission=0.2, ior=1.3), {}),
    "maze500_b8": lambda: (scenes.maze(n_segments=500, ray_count=2048, max_bounces=12, bands=8), {"bands": 8}),
}


def sparse(hist):
    idx = np.flatnonzero(hist).astype(np.int32)
    return idx, hist[idx].astype(np.int64)


def main():
    for name, make in CASES.items():
        sc, over = make()
        kw = trace_kwargs(sc, **over)
        r = O.trace(oracle_walls(O, sc.walls), oracle_params(O, kw), band_abs=sc.band_absorption if kw["bands"] > 1 else None,
                    want_hits=True)
        idx, val = sparse(r.hist)
        digest = hashlib.sha256(r.hits.tobytes()).hexdigest()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), idx=idx, val=val, n_words=np.int64(r.hist.size),
                            n_hits=np.int64(r.n_hits), hits_sha256=np.bytes_(digest.encode()),
                            counters=np.array([r.counters[k] for k in ("ray_bounces", "nearest_tests", "shadow_tests", "direct_hits", "nee_hits")], np.int64))
        print(name, len(idx), "bins", r.n_hits, "hits", digest[:12])
    rng = np.random.default_rng(2024)
    x = rng.uniform(-0.5, 0.5, 1000).astype(np.float32)
    x[::9] *= np.float32(1e-4)
    ir = O.ir_to_float(np.array([O.lib().orc_quantize(float(v)) for v in scenes.decaying_noise_ir(300, 5, decay_s=0.004)], np.int64))
    out = O.convolve(x, ir, 3)
    np.savez_compressed(os.path.join(HERE, "conv_1000x300.npz"), x=x, ir=ir, accum=np.int32(3), out=out)
    print("conv_1000x300", float(np.abs(out).max()))


if __name__ == "__main__":
    main()
