"""Shared helpers of the parity tests: one parameter dict drives both the oracle and the C-ABI."""
from __future__ import annotations

import numpy as np


def trace_kwargs(scene, **over):
    d = dict(source=scene.source, listener=scene.listener, listener_radius=scene.listener_radius,
             speed_of_sound=scene.speed_of_sound, input_gain=scene.input_gain, max_bounce_count=scene.max_bounces,
             rng_state_offset=1, ray_count=scene.ray_count, debug_ray_count=0, sample_rate=scene.sample_rate,
             impulse_length=scene.impulse_length, bands=1, time_divisor=1.0, flags=0, ray_begin=0, ray_end=0)
    unknown = set(over) - set(d)
    assert not unknown, unknown
    d.update(over)
    return d


def oracle_params(O, kw):
    k = dict(kw)
    src, lis = k.pop("source"), k.pop("listener")
    return O.make_params(source_x=src[0], source_y=src[1], listener_x=lis[0], listener_y=lis[1], **k)


def capi_params(capi, kw):
    return capi.make_trace_params(**kw)


def oracle_walls(O, walls):
    return np.ascontiguousarray(walls).view(O.SEGMENT_DTYPE)


def sort_hits(hits, keys):
    order = np.lexsort((keys["kind"], keys["bounce"], keys["ray"]))
    return hits[order], keys[order]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a))
