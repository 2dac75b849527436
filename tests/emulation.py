"""Builds and binds tests/host_emulation.cpp (TEST TOOLING: the product's device headers compiled for
the host so their logic can be compared with the oracle without a GPU; not a product path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "_build", "libhost_emulation.so")
_SRC = os.path.join(_HERE, "host_emulation.cpp")
_DEPS = [_SRC] + [os.path.join(_ROOT, "realisticaudioraytracing2d_b200", "csrc", h)
                  for h in ("rar_math.cuh", "rar_ray.cuh", "rar_fft.cuh", "rar_synth16.cuh", "rar_layout.h")]


def build(force: bool = False) -> str:
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in _DEPS)
    if stale:
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off",
                               "-fno-fast-math", "-mfma", "-mavx2", "-Wall", "-Wno-unknown-pragmas", "-o", _SO, _SRC])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def trace(O, walls, params, band_abs=None, counting=True, air=None):
    """Runs the product's ray logic on the host. `params` is an oracle TraceParams (same layout as
    rar_trace_params).  Returns (hist, hits sorted by ray/bounce/kind, counters dict)."""
    L = lib()
    bands = max(1, params.bands)
    hist = np.zeros(params.impulse_length * bands, np.int64)
    n = O.lib().orc_dispatch_threads(C.byref(params)) if params.ray_end == 0 else params.ray_end - params.ray_begin
    cap = int(n) * max(1, params.max_bounce_count) * 2
    hits = np.zeros(cap, O.HIT_DTYPE)
    cnt = C.c_int64(0)
    ctr = O.Counters()
    walls = np.ascontiguousarray(walls)
    ba = np.ascontiguousarray(band_abs, dtype=np.float32) if band_abs is not None else None
    a = np.ascontiguousarray(air, dtype=np.float32) if air is not None else None
    L.emu_set_air(C.c_void_p(a.ctypes.data) if a is not None else None, len(a) if a is not None else 0)
    fn = L.emu_trace if counting else L.emu_trace_nocount
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                   C.POINTER(C.c_int64), C.POINTER(O.Counters)]
    rc = fn(walls.ctypes.data if len(walls) else None, len(walls), ba.ctypes.data if ba is not None else None,
                     C.addressof(params), hist.ctypes.data, hits.ctypes.data, cap, C.byref(cnt), C.byref(ctr))
    L.emu_set_air(None, 0)
    assert rc == 0, rc
    hits = hits[: cnt.value]
    hits = hits[np.lexsort((hits["kind"], hits["bounce"], hits["ray"]))]
    return hist, hits, {k: getattr(ctr, k) for k, _ in O.Counters._fields_}


def rfft512(w):
    w = np.ascontiguousarray(w, np.float32)
    P = np.zeros(512, np.float32)
    lib().emu_rfft512(C.c_void_p(w.ctypes.data), C.c_void_p(P.ctypes.data))
    return P.view(np.complex64)


def irfft512(P):
    P = np.ascontiguousarray(P, np.complex64)
    w = np.zeros(512, np.float32)
    lib().emu_irfft512(C.c_void_p(P.ctypes.data), C.c_void_p(w.ctypes.data))
    return w / np.float32(256)


def convolve(x, ir, accum):
    x = np.ascontiguousarray(x, np.float32)
    ir = np.ascontiguousarray(ir, np.float32)
    out = np.zeros(len(x) + len(ir), np.float32)
    lib().emu_convolve(C.c_void_p(x.ctypes.data), len(x), C.c_void_p(ir.ctypes.data), len(ir), int(accum), C.c_void_p(out.ctypes.data))
    return out


def fft16(x, inverse=False, half=False):
    v = np.ascontiguousarray(x, np.complex64).copy()
    lib().emu_fft16(C.c_void_p(v.ctypes.data), int(inverse), int(half))
    return v


def band_synth16(hist, bands, scale, taps, out_len):
    """Filter-bank synthesis with the arithmetic of csrc/band_synth.cu (rar_synth16.cuh) on the host."""
    hist = np.ascontiguousarray(hist, np.int64)
    taps = np.ascontiguousarray(taps, np.float32)
    assert taps.shape == (bands, 256)
    out = np.zeros(out_len, np.float32)
    lib().emu_band_synth16(C.c_void_p(hist.ctypes.data), len(hist) // bands, int(bands), C.c_float(scale), C.c_void_p(taps.ctypes.data),
                           C.c_void_p(out.ctypes.data), int(out_len))
    return out


def intersect_many(rays, segs, closest):
    rays = np.ascontiguousarray(rays, np.float32)
    segs = np.ascontiguousarray(segs, np.float32)
    closest = np.ascontiguousarray(closest, np.float32)
    out = np.zeros(len(rays), np.float32)
    lib().emu_intersect_many(C.c_void_p(rays.ctypes.data), C.c_void_p(segs.ctypes.data), C.c_void_p(closest.ctypes.data),
                             len(rays), C.c_void_p(out.ctypes.data))
    return out


def pair_planes(walls):
    """rar_layout.h pair_planes: (pair_a, pair_b), each [(n + 1) // 2][4]."""
    walls = np.ascontiguousarray(walls)
    n = len(walls)
    a = np.zeros(((n + 1) // 2, 4), np.float32)
    b = np.zeros_like(a)
    lib().emu_pair_planes(C.c_void_p(walls.ctypes.data), n, C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data))
    return a, b


def grid_digest(walls):
    """Dimensions, list length and digest of the HOST-built uniform grid (rar_layout.h build_grid)."""
    walls = np.ascontiguousarray(walls)
    nx, ny, n, d = C.c_int(), C.c_int(), C.c_longlong(), C.c_ulonglong()
    lib().emu_grid_digest(C.c_void_p(walls.ctypes.data if len(walls) else None), len(walls), C.byref(nx), C.byref(ny), C.byref(n), C.byref(d))
    return {"nx": nx.value, "ny": ny.value, "n_items": n.value, "digest": d.value}
