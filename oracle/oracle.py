"""ctypes view of the CPU oracle (oracle/rar_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.  Parity status of the
oracle itself: *parity unpinned* (the reference has no tests, golden vectors or CPU path); see the
header of rar_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librar_oracle.so")

SEGMENT_DTYPE = np.dtype(
    [("ax", "<f4"), ("ay", "<f4"), ("bx", "<f4"), ("by", "<f4"), ("nx", "<f4"), ("ny", "<f4"),
     ("absorption", "<f4"), ("scattering", "<f4"), ("transmission", "<f4"), ("ior", "<f4")]
)
HIT_DTYPE = np.dtype(
    [("time_delay", "<f4"), ("energy", "<f4"), ("hit_x", "<f4"), ("hit_y", "<f4"),
     ("ray", "<u4"), ("bounce", "<u2"), ("kind", "<u2")]
)
assert SEGMENT_DTYPE.itemsize == 40 and HIT_DTYPE.itemsize == 24


class TraceParams(C.Structure):
    _fields_ = [
        ("source_x", C.c_float), ("source_y", C.c_float), ("listener_x", C.c_float), ("listener_y", C.c_float),
        ("listener_radius", C.c_float), ("speed_of_sound", C.c_float), ("input_gain", C.c_float),
        ("max_bounce_count", C.c_int32), ("rng_state_offset", C.c_uint32), ("ray_count", C.c_int32),
        ("debug_ray_count", C.c_int32), ("sample_rate", C.c_int32), ("impulse_length", C.c_int32),
        ("bands", C.c_int32), ("time_divisor", C.c_float), ("flags", C.c_uint32),
        ("ray_begin", C.c_int64), ("ray_end", C.c_int64),
    ]


class Counters(C.Structure):
    _fields_ = [("ray_bounces", C.c_uint64), ("nearest_tests", C.c_uint64), ("shadow_tests", C.c_uint64),
                ("direct_hits", C.c_uint64), ("nee_hits", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rar_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        f32, i32, i64, u32p = C.c_float, C.c_int32, C.c_int64, C.POINTER(C.c_uint32)
        L.orc_random.restype = f32
        L.orc_random.argtypes = [u32p]
        L.orc_intersect.restype = f32
        L.orc_intersect.argtypes = [f32] * 8
        L.orc_intersect_circle.restype = f32
        L.orc_intersect_circle.argtypes = [f32] * 7
        L.orc_refract.restype = C.c_int
        L.orc_refract.argtypes = [f32] * 5 + [C.POINTER(f32)] * 2
        L.orc_sincosf.restype = None
        L.orc_sincosf.argtypes = [f32, C.POINTER(f32), C.POINTER(f32)]
        L.orc_asinf.restype = f32
        L.orc_asinf.argtypes = [f32]
        L.orc_quantize.restype = i64
        L.orc_quantize.argtypes = [f32]
        L.orc_time_bin.restype = i32
        L.orc_time_bin.argtypes = [f32, i32, f32, i32]
        L.orc_dispatch_threads.restype = i64
        L.orc_dispatch_threads.argtypes = [C.POINTER(TraceParams)]
        L.orc_trace.restype = C.c_int
        L.orc_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(TraceParams), C.c_void_p, C.c_void_p,
                                i64, C.POINTER(i64), C.POINTER(Counters), C.c_int]
        L.orc_trace_air.restype = C.c_int
        L.orc_trace_air.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(TraceParams), C.c_void_p, C.c_void_p,
                                    i64, C.POINTER(i64), C.POINTER(Counters), C.c_int]
        L.orc_exp_neg.restype = f32
        L.orc_exp_neg.argtypes = [f32, f32]
        L.orc_ir_to_float.restype = None
        L.orc_ir_to_float.argtypes = [C.c_void_p, i64, C.c_void_p]
        L.orc_convolve.restype = None
        L.orc_convolve.argtypes = [C.c_void_p, i32, C.c_void_p, i32, i32, C.c_void_p, C.c_int]
        L.orc_convolve_f64.restype = None
        L.orc_convolve_f64.argtypes = [C.c_void_p, i32, C.c_void_p, i32, i32, C.c_void_p, C.c_int]
        L.orc_band_filter_taps.restype = None
        L.orc_band_filter_taps.argtypes = [C.c_double, C.c_double, C.c_void_p]
        L.orc_synthesize_ir.restype = None
        L.orc_synthesize_ir.argtypes = [C.c_void_p, i32, i32, i32, C.c_void_p, C.c_void_p]
        L.orc_load_sample.restype = i64
        L.orc_load_sample.argtypes = [C.c_void_p, i64, i32, i32, i32, C.c_void_p]
        L.orc_add_loop.restype = C.c_int
        L.orc_add_loop.argtypes = [C.c_void_p, C.c_int] + [f32] * 10 + [C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def random_sequence(seed: int, n: int):
    """n successive values of Common.hlsl:8-12 `random` from the given state; returns (floats, states)."""
    st = C.c_uint32(seed & 0xFFFFFFFF)
    vals, states = [], []
    for _ in range(n):
        vals.append(lib().orc_random(C.byref(st)))
        states.append(st.value)
    return np.array(vals, dtype=np.float32), states


def sincos(x: float):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincosf(x, C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


def refract(ix, iy, nx, ny, eta):
    tx, ty = C.c_float(), C.c_float()
    ok = lib().orc_refract(ix, iy, nx, ny, eta, C.byref(tx), C.byref(ty))
    return ok, np.float32(tx.value), np.float32(ty.value)


def make_params(**kw) -> TraceParams:
    d = dict(source_x=0.0, source_y=0.0, listener_x=1.0, listener_y=0.0, listener_radius=0.5,
             speed_of_sound=343.0, input_gain=1.0, max_bounce_count=5, rng_state_offset=1, ray_count=1000,
             debug_ray_count=100, sample_rate=48000, impulse_length=96000, bands=1, time_divisor=1.0,
             flags=0, ray_begin=0, ray_end=0)
    unknown = set(kw) - set(d)
    if unknown:
        raise TypeError(f"unknown trace parameter(s): {sorted(unknown)}")
    d.update(kw)
    return TraceParams(**d)


@dataclass
class TraceResult:
    hist: np.ndarray | None
    hits: np.ndarray | None
    counters: dict
    n_hits: int


def trace(walls: np.ndarray, params: TraceParams, band_abs: np.ndarray | None = None, hist: np.ndarray | None = None,
          want_hist: bool = True, want_hits: bool = False, hit_cap: int | None = None, n_threads: int = 0,
          air: np.ndarray | None = None) -> TraceResult:
    walls = np.ascontiguousarray(walls, dtype=SEGMENT_DTYPE)
    bands = max(1, params.bands)
    if want_hist and hist is None:
        hist = np.zeros(params.impulse_length * bands, dtype=np.int64)
    if band_abs is not None:
        band_abs = np.ascontiguousarray(band_abs, dtype=np.float32)
        assert band_abs.shape == (len(walls), bands)
    hits = None
    if want_hits:
        if hit_cap is None:
            n = params.ray_end - params.ray_begin or lib().orc_dispatch_threads(C.byref(params))
            hit_cap = int(n) * params.max_bounce_count * 2
        hits = np.zeros(hit_cap, dtype=HIT_DTYPE)
    cnt = C.c_int64(0)
    ctr = Counters()
    if air is not None:
        air = np.ascontiguousarray(air, dtype=np.float32)
        assert air.shape == (bands,)
    rc = lib().orc_trace_air(walls.ctypes.data, len(walls), band_abs.ctypes.data if band_abs is not None else None,
                         air.ctypes.data if air is not None else None,
                         C.byref(params), hist.ctypes.data if want_hist else None,
                         hits.ctypes.data if hits is not None else None, hit_cap or 0, C.byref(cnt), C.byref(ctr),
                         n_threads)
    if rc != 0:
        raise RuntimeError(f"orc_trace failed: {rc}")
    if hits is not None:
        hits = hits[: min(cnt.value, len(hits))]
        hits = hits[np.lexsort((hits["kind"], hits["bounce"], hits["ray"]))]
    counters = {k: getattr(ctr, k) for k, _ in Counters._fields_}
    return TraceResult(hist if want_hist else None, hits, counters, cnt.value)


def ir_to_float(hist: np.ndarray) -> np.ndarray:
    hist = np.ascontiguousarray(hist, dtype=np.int64)
    out = np.empty(hist.shape, dtype=np.float32)
    lib().orc_ir_to_float(hist.ctypes.data, hist.size, out.ctypes.data)
    return out


def convolve(inp: np.ndarray, ir: np.ndarray, accum_count: int, n_threads: int = 0, f64: bool = False) -> np.ndarray:
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    ir = np.ascontiguousarray(ir, dtype=np.float32)
    out = np.empty(len(inp) + len(ir), dtype=np.float64 if f64 else np.float32)
    fn = lib().orc_convolve_f64 if f64 else lib().orc_convolve
    fn(inp.ctypes.data, len(inp), ir.ctypes.data, len(ir), accum_count, out.ctypes.data, n_threads)
    return out


def band_filter_taps(lo: float, hi: float) -> np.ndarray:
    """The 255 taps of the band-pass filter of a band with edges lo, hi (fractions of Nyquist)."""
    g = np.zeros(255, dtype=np.float32)
    lib().orc_band_filter_taps(lo, hi, g.ctypes.data)
    return g


def synthesize_ir(hist: np.ndarray, bins: int, bands: int, stride: int = 1, edges=None) -> np.ndarray:
    """Filter-bank synthesis of a banded histogram [bins][bands] in direct form (this build's banded model)."""
    hist = np.ascontiguousarray(hist, dtype=np.int64)
    assert hist.size == bins * bands
    e = None if edges is None else np.ascontiguousarray(edges, dtype=np.float32)
    assert e is None or len(e) == bands + 1
    out = np.zeros(bins * stride, dtype=np.float32)
    lib().orc_synthesize_ir(hist.ctypes.data, bins, bands, stride, e.ctypes.data if e is not None else None, out.ctypes.data)
    return out


def add_loop(local_xy, pos, qz, qw, scale, mat) -> np.ndarray:
    """Helpers/SceneHelper.cs:78-98 for one closed loop of local points under a 2-D transform."""
    pts = np.ascontiguousarray(local_xy, dtype=np.float32).reshape(-1, 2)
    out = np.zeros(len(pts), dtype=SEGMENT_DTYPE)
    lib().orc_add_loop(pts.ctypes.data, len(pts), pos[0], pos[1], qz, qw, scale[0], scale[1],
                       mat[0], mat[1], mat[2], mat[3], out.ctypes.data)
    return out


def load_sample(raw, samples: int, channels: int, clip_frequency: int, sample_rate: int) -> np.ndarray:
    """RayTraceManager.cs:135-167 LoadSample of one interleaved clip."""
    a = np.ascontiguousarray(raw, dtype=np.float32)
    assert a.size >= samples * channels
    n = lib().orc_load_sample(a.ctypes.data, samples, channels, clip_frequency, sample_rate, None)
    out = np.zeros(n, dtype=np.float32)
    if n:
        lib().orc_load_sample(a.ctypes.data, samples, channels, clip_frequency, sample_rate, out.ctypes.data)
    return out


def num_threads() -> int:
    return lib().orc_num_threads()
