"""ctypes binding of the C-ABI in include/rar2d.h (librar2d.so, built in-tree by build.py).

This is the only way the Python host code reaches the compute path.  There is no fallback: if the
shared library is missing, or no B200-class CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librar2d.so")

RAR_OK = 0
RAR_FLAG_EXACT_RAY_COUNT = 1
RAR_FLAG_COUNT_TESTS = 2
RAR_FLAG_COUNT_EXECUTED = 4
RAR_FLAG_USE_GRID = 8
RAR_EXCHANGE_HANDLE_BYTES = 80
RAR_EXCHANGE_MAX_RANKS = 16
RAR_EXCHANGE_AUTO, RAR_EXCHANGE_ONE_SHOT, RAR_EXCHANGE_TWO_SHOT = 0, 1, 2

# include/rar2d.h rar_segment / rar_ray_info / rar_hit_key
SEGMENT_DTYPE = np.dtype(
    [("start", "<f4", (2,)), ("end", "<f4", (2,)), ("normal", "<f4", (2,)),
     ("absorption", "<f4"), ("scattering", "<f4"), ("transmission", "<f4"), ("ior", "<f4")]
)
RAY_INFO_DTYPE = np.dtype([("time_delay", "<f4"), ("energy", "<f4"), ("hit_point", "<f4", (2,))])
HIT_KEY_DTYPE = np.dtype([("ray", "<u4"), ("bounce", "<u2"), ("kind", "<u2")])
assert SEGMENT_DTYPE.itemsize == 40 and RAY_INFO_DTYPE.itemsize == 16 and HIT_KEY_DTYPE.itemsize == 8


class TraceParams(C.Structure):
    """include/rar2d.h rar_trace_params."""
    _fields_ = [
        ("source_pos", C.c_float * 2), ("listener_pos", C.c_float * 2),
        ("listener_radius", C.c_float), ("speed_of_sound", C.c_float), ("input_gain", C.c_float),
        ("max_bounce_count", C.c_int32), ("rng_state_offset", C.c_uint32), ("ray_count", C.c_int32),
        ("debug_ray_count", C.c_int32), ("sample_rate", C.c_int32), ("impulse_length", C.c_int32),
        ("bands", C.c_int32), ("time_divisor", C.c_float), ("flags", C.c_uint32),
        ("ray_begin", C.c_int64), ("ray_end", C.c_int64),
    ]


class Counters(C.Structure):
    _fields_ = [("ray_bounces", C.c_uint64), ("nearest_tests", C.c_uint64), ("shadow_tests", C.c_uint64),
                ("direct_hits", C.c_uint64), ("nee_hits", C.c_uint64)]

    def as_dict(self) -> dict:
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# name -> (restype, argtypes).  Every symbol include/rar2d.h declares.
_p, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "rar_version": (C.c_int, []),
    "rar_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "rar_destroy": (C.c_int, [_p]),
    "rar_last_error": (C.c_char_p, [_p]),
    "rar_set_stream": (C.c_int, [_p, _p]),
    "rar_sync": (C.c_int, [_p]),
    "rar_set_walls": (C.c_int, [_p, _p, _i32]),
    "rar_set_wall_band_absorption": (C.c_int, [_p, _p, _i32, _i32]),
    "rar_set_air_absorption": (C.c_int, [_p, _p, _i32]),
    "rar_ir_clear": (C.c_int, [_p, _i32, _i32, _i32]),
    "rar_ir_read": (C.c_int, [_p, _i32, _p, _i64]),
    "rar_ir_read_begin": (C.c_int, [_p, _i32, _i64, C.POINTER(_i32)]),
    "rar_ir_read_end": (C.c_int, [_p, _i32, _p, _i64]),
    "rar_ir_read_fixed": (C.c_int, [_p, _i32, _p, _i64]),
    "rar_ir_write": (C.c_int, [_p, _i32, _p, _i32, _i32]),
    "rar_ir_device_ptr": (C.c_int, [_p, _i32, C.POINTER(_p), C.POINTER(_i64)]),
    "rar_allreduce_slots": (C.c_int, [C.POINTER(_p), _i32, _i32]),
    "rar_exchange_create": (C.c_int, [_p, _i64, _p]),
    "rar_exchange_connect": (C.c_int, [_p, _i32, _i32, _p]),
    "rar_exchange_allreduce": (C.c_int, [_p, _i32, _i32]),
    "rar_exchange_status": (C.c_int, [_p]),
    "rar_exchange_destroy": (C.c_int, [_p]),
    "rar_prepared_length": (_i64, [_i64, _i32, _i32]),
    "rar_prepare_clips": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _p, _i64]),
    "rar_prepare_clips_device": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _p, _i64]),
    "rar_trace": (C.c_int, [_p, C.POINTER(TraceParams), _i32]),
    "rar_trace_interleaved": (C.c_int, [_p, C.POINTER(TraceParams), _i32, _i32, _i32, _i32]),
    "rar_trace_frames": (C.c_int, [_p, C.POINTER(TraceParams), _i32, _i32]),
    "rar_trace_listeners": (C.c_int, [_p, C.POINTER(TraceParams), _p, _i32, _i32]),
    "rar_trace_hits": (C.c_int, [_p, C.POINTER(TraceParams), _p, _p, _i64, C.POINTER(_i64)]),
    "rar_get_counters": (C.c_int, [_p, C.POINTER(Counters), _i32]),
    "rar_get_debug_rays": (C.c_int, [_p, _p, _i64]),
    "rar_set_band_edges": (C.c_int, [_p, _p, _i32, _i32]),
    "rar_synthesize_ir": (C.c_int, [_p, _i32, _p, _i64]),
    "rar_convolve": (C.c_int, [_p, _i32, _p, _i32, _i32, _p, _i32]),
    "rar_convolve_begin": (C.c_int, [_p, _i32, _p, _i32, _i32, C.POINTER(_i32)]),
    "rar_poll": (C.c_int, [_p, _i32]),
    "rar_convolve_end": (C.c_int, [_p, _i32, _p, _i32]),
    "rar_conv_create": (C.c_int, [_p, _i32, _i32, _i32, C.POINTER(_p)]),
    "rar_conv_destroy": (C.c_int, [_p]),
    "rar_conv_set_ir": (C.c_int, [_p, _i32, _p, _i32, _f32]),
    "rar_conv_set_ir_from_slot": (C.c_int, [_p, _i32, _i32, _i32]),
    "rar_conv_set_irs": (C.c_int, [_p, _i32, _i32, _p, _i32, _i64, _f32]),
    "rar_conv_set_irs_from_slots": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "rar_conv_update_ir": (C.c_int, [_p, _i32, _p, _i32, _f32]),
    "rar_conv_update_ir_from_slot": (C.c_int, [_p, _i32, _i32, _i32]),
    "rar_conv_reset": (C.c_int, [_p]),
    "rar_conv_process": (C.c_int, [_p, _p, _p]),
    "rar_conv_process_device": (C.c_int, [_p, _p, _p]),
    "rar_conv_bytes_per_block": (_i64, [_p]),
    "rar_ring_create": (C.c_int, [_i32, _f32, C.POINTER(_p)]),
    "rar_ring_destroy": (C.c_int, [_p]),
    "rar_ring_reset": (C.c_int, [_p]),
    "rar_ring_stop": (C.c_int, [_p]),
    "rar_ring_size": (_i32, [_p]),
    "rar_ring_is_pinned": (_i32, [_p]),
    "rar_ring_frames_drained": (_i64, [_p]),
    "rar_ring_push": (C.c_int, [_p, _p, _i32, _i64]),
    "rar_ring_drain": (C.c_int, [_p, _p, _i32, _i32]),
    "rar_conv_process_to_ring": (C.c_int, [_p, _p, _p, _i64]),
    "rar_device_info": (C.c_int, [_p, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "rar_measure_fp32_peak": (C.c_int, [_p, C.POINTER(C.c_double)]),
    "rar_selftest_arithmetic": (C.c_int, [_p, _i64, C.c_uint32, C.POINTER(C.c_uint64)]),
    "rar_debug_grid": (C.c_int, [_p, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(C.c_uint64)]),
    "rar_launch_count": (_i64, [_p]),
}

_lib = None


class RarError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rar2d error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Loads librar2d.so and declares every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RarError(-2, f"{LIB_PATH} is missing: build it with `python -m realisticaudioraytracing2d_b200.build` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(ctx, rc: int) -> int:
    if rc < 0:
        msg = load().rar_last_error(ctx)
        raise RarError(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc


def make_trace_params(source, listener, listener_radius=0.5, speed_of_sound=343.0, input_gain=1.0,
                      max_bounce_count=5, rng_state_offset=1, ray_count=1000, debug_ray_count=0,
                      sample_rate=48000, impulse_length=96000, bands=1, time_divisor=1.0, flags=0,
                      ray_begin=0, ray_end=0) -> TraceParams:
    p = TraceParams()
    p.source_pos[0], p.source_pos[1] = float(source[0]), float(source[1])
    p.listener_pos[0], p.listener_pos[1] = float(listener[0]), float(listener[1])
    p.listener_radius, p.speed_of_sound, p.input_gain = listener_radius, speed_of_sound, input_gain
    p.max_bounce_count, p.rng_state_offset, p.ray_count = max_bounce_count, rng_state_offset & 0xFFFFFFFF, ray_count
    p.debug_ray_count, p.sample_rate, p.impulse_length = debug_ray_count, sample_rate, impulse_length
    p.bands, p.time_divisor, p.flags = bands, time_divisor, flags
    p.ray_begin, p.ray_end = ray_begin, ray_end
    return p


class Context:
    """Thin object wrapper over a rar_context*; methods map 1:1 onto the C entry points."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = _p()
        rc = self._lib.rar_create(device, C.byref(h))
        if rc < 0:
            msg = self._lib.rar_last_error(None)
            raise RarError(rc, msg.decode() if msg else "")
        self._h = h
        self.device = device

    # lifetime ---------------------------------------------------------------------------------
    def destroy(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rar_destroy(self._h)
            self._h = None

    close = destroy

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.destroy()

    def _ck(self, rc):
        return check(self._h, rc)

    def set_stream(self, cuda_stream_ptr) -> None:
        self._ck(self._lib.rar_set_stream(self._h, _p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def sync(self) -> None:
        self._ck(self._lib.rar_sync(self._h))

    # geometry ---------------------------------------------------------------------------------
    def set_walls(self, segments: np.ndarray) -> None:
        seg = np.ascontiguousarray(segments)
        if seg.dtype.itemsize != 40:
            raise ValueError("walls must be 40-byte Segment records")
        self._ck(self._lib.rar_set_walls(self._h, seg.ctypes.data if len(seg) else None, len(seg)))

    def set_wall_band_absorption(self, table: np.ndarray) -> None:
        t = np.ascontiguousarray(table, dtype=np.float32)
        self._ck(self._lib.rar_set_wall_band_absorption(self._h, t.ctypes.data if t.size else None, t.shape[0], t.shape[1]))

    def set_air_absorption(self, alpha_per_m) -> None:
        """rar_set_air_absorption: per-band air absorption in 1/m; None switches it off."""
        if alpha_per_m is None:
            self._ck(self._lib.rar_set_air_absorption(self._h, None, 0))
            return
        a = np.ascontiguousarray(alpha_per_m, dtype=np.float32)
        self._ck(self._lib.rar_set_air_absorption(self._h, a.ctypes.data, len(a)))

    # IR slots ---------------------------------------------------------------------------------
    def ir_clear(self, slot: int, impulse_length: int, bands: int = 1) -> None:
        self._ck(self._lib.rar_ir_clear(self._h, slot, impulse_length, bands))

    def ir_read(self, slot: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float32)
        self._ck(self._lib.rar_ir_read(self._h, slot, out.ctypes.data, n))
        return out

    def ir_read_into(self, slot: int, host_ptr: int, n: int) -> None:
        """rar_ir_read into caller-owned host memory (e.g. a pinned buffer), n floats."""
        self._ck(self._lib.rar_ir_read(self._h, slot, _p(host_ptr), n))

    def ir_read_begin(self, slot: int, n: int) -> int:
        """rar_ir_read_begin: enqueue the readback of n floats, returns a ticket (poll() / ir_read_end())."""
        t = _i32(-1)
        self._ck(self._lib.rar_ir_read_begin(self._h, slot, n, C.byref(t)))
        return t.value

    def ir_read_end(self, ticket: int, n: int, host_ptr: int = 0) -> np.ndarray:
        """rar_ir_read_end into a new array, or into caller-owned memory when host_ptr is given."""
        if host_ptr:
            self._ck(self._lib.rar_ir_read_end(self._h, ticket, _p(host_ptr), n))
            return None
        out = np.empty(n, dtype=np.float32)
        self._ck(self._lib.rar_ir_read_end(self._h, ticket, out.ctypes.data, n))
        return out

    def ir_read_fixed(self, slot: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int64)
        self._ck(self._lib.rar_ir_read_fixed(self._h, slot, out.ctypes.data, n))
        return out

    def ir_write(self, slot: int, ir: np.ndarray, bands: int = 1) -> None:
        a = np.ascontiguousarray(ir, dtype=np.float32)
        self._ck(self._lib.rar_ir_write(self._h, slot, a.ctypes.data if a.size else None, a.size // bands, bands))

    def ir_device_ptr(self, slot: int):
        ptr, n = _p(), _i64()
        self._ck(self._lib.rar_ir_device_ptr(self._h, slot, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    # multi-process peer-memory all-reduce -------------------------------------------------------
    def exchange_create(self, capacity_words: int) -> bytes:
        buf = C.create_string_buffer(RAR_EXCHANGE_HANDLE_BYTES)
        self._ck(self._lib.rar_exchange_create(self._h, capacity_words, buf))
        return buf.raw

    def exchange_connect(self, rank: int, world: int, handles) -> None:
        blob = b"".join(handles)
        if len(blob) != world * RAR_EXCHANGE_HANDLE_BYTES:
            raise ValueError("need one %d-byte handle per rank" % RAR_EXCHANGE_HANDLE_BYTES)
        self._ck(self._lib.rar_exchange_connect(self._h, rank, world, blob))

    def exchange_allreduce(self, slot: int, mode: int = RAR_EXCHANGE_AUTO) -> None:
        self._ck(self._lib.rar_exchange_allreduce(self._h, slot, mode))

    def exchange_status(self) -> None:
        self._ck(self._lib.rar_exchange_status(self._h))

    def exchange_destroy(self) -> None:
        self._ck(self._lib.rar_exchange_destroy(self._h))

    # clip preparation ---------------------------------------------------------------------------
    def prepare_clips(self, raw: np.ndarray, samples: int, channels: int, clip_frequency: int, sample_rate: int,
                      n_clips: int = 1) -> np.ndarray:
        """rar_prepare_clips: LoadSample of n_clips interleaved clips [n_clips][samples][channels] -> [n_clips][newLength]."""
        a = np.ascontiguousarray(raw, dtype=np.float32)
        if a.size != n_clips * samples * channels:
            raise ValueError("raw must hold n_clips * samples * channels values")
        n = prepared_length(samples, clip_frequency, sample_rate)
        out = np.zeros((n_clips, n), dtype=np.float32)
        self._ck(self._lib.rar_prepare_clips(self._h, a.ctypes.data if a.size else None, samples, channels, clip_frequency,
                                             sample_rate, n_clips, out.ctypes.data if out.size else None, n))
        return out

    def prepare_clips_device(self, d_raw: int, samples: int, channels: int, clip_frequency: int, sample_rate: int,
                             n_clips: int, d_out: int, out_stride: int) -> None:
        self._ck(self._lib.rar_prepare_clips_device(self._h, _p(d_raw), samples, channels, clip_frequency, sample_rate,
                                                    n_clips, _p(d_out), out_stride))

    # trace ------------------------------------------------------------------------------------
    def trace(self, params: TraceParams, slot: int) -> None:
        self._ck(self._lib.rar_trace(self._h, C.byref(params), slot))

    def trace_interleaved(self, params: TraceParams, slot: int, rank: int, world: int, chunk_log2: int = 14) -> None:
        """rar_trace_interleaved: this rank's block-cyclic share (chunks of 2^chunk_log2 thread ids) of the dispatch."""
        self._ck(self._lib.rar_trace_interleaved(self._h, C.byref(params), slot, rank, world, chunk_log2))

    def trace_frames(self, params: TraceParams, slot: int, n_frames: int) -> None:
        self._ck(self._lib.rar_trace_frames(self._h, C.byref(params), slot, n_frames))

    def trace_listeners(self, params: TraceParams, listeners_xy: np.ndarray, first_slot: int) -> None:
        xy = np.ascontiguousarray(listeners_xy, dtype=np.float32).reshape(-1, 2)
        self._ck(self._lib.rar_trace_listeners(self._h, C.byref(params), xy.ctypes.data if len(xy) else None, len(xy), first_slot))

    def trace_hits(self, params: TraceParams, capacity: int):
        hits = np.zeros(capacity, dtype=RAY_INFO_DTYPE)
        keys = np.zeros(capacity, dtype=HIT_KEY_DTYPE)
        cnt = _i64()
        self._ck(self._lib.rar_trace_hits(self._h, C.byref(params), hits.ctypes.data, keys.ctypes.data, capacity, C.byref(cnt)))
        n = min(cnt.value, capacity)
        return hits[:n], keys[:n], cnt.value

    def get_counters(self, reset: bool = True) -> dict:
        c = Counters()
        self._ck(self._lib.rar_get_counters(self._h, C.byref(c), 1 if reset else 0))
        return c.as_dict()

    def get_debug_rays(self, n_float4: int) -> np.ndarray:
        out = np.zeros((n_float4, 4), dtype=np.float32)
        self._ck(self._lib.rar_get_debug_rays(self._h, out.ctypes.data, n_float4))
        return out

    # banded model ------------------------------------------------------------------------------
    def set_band_edges(self, edges_hz, sample_rate: int) -> None:
        """rar_set_band_edges; edges_hz=None restores equal-width bands."""
        if edges_hz is None:
            self._ck(self._lib.rar_set_band_edges(self._h, None, 0, 0))
            return
        e = np.ascontiguousarray(edges_hz, dtype=np.float32)
        self._ck(self._lib.rar_set_band_edges(self._h, e.ctypes.data, len(e) - 1, sample_rate))

    def synthesize_ir(self, slot: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float32)
        self._ck(self._lib.rar_synthesize_ir(self._h, slot, out.ctypes.data, n))
        return out

    # convolution ------------------------------------------------------------------------------
    def convolve(self, slot: int, samples: np.ndarray, accum_count: int, ir_len: int) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float32)
        out = np.empty(len(x) + ir_len, dtype=np.float32)
        self._ck(self._lib.rar_convolve(self._h, slot, x.ctypes.data if len(x) else None, len(x), accum_count,
                                        out.ctypes.data, len(out)))
        return out

    def convolve_begin(self, slot: int, samples: np.ndarray, accum_count: int) -> int:
        x = np.ascontiguousarray(samples, dtype=np.float32)
        t = _i32(-1)
        self._ck(self._lib.rar_convolve_begin(self._h, slot, x.ctypes.data if len(x) else None, len(x), accum_count, C.byref(t)))
        return t.value

    def poll(self, ticket: int) -> bool:
        return self._ck(self._lib.rar_poll(self._h, ticket)) == 1

    def convolve_end(self, ticket: int, out_len: int) -> np.ndarray:
        out = np.empty(out_len, dtype=np.float32)
        self._ck(self._lib.rar_convolve_end(self._h, ticket, out.ctypes.data, out_len))
        return out

    # measurement ------------------------------------------------------------------------------
    def device_info(self) -> dict:
        a, b, c = _i32(), _i32(), _i32()
        self._ck(self._lib.rar_device_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"sm_count": a.value, "sm_clock_khz": b.value, "smem_optin_bytes": c.value}

    def measure_fp32_peak(self) -> float:
        v = C.c_double()
        self._ck(self._lib.rar_measure_fp32_peak(self._h, C.byref(v)))
        return v.value

    def selftest_arithmetic(self, n_samples: int, seed: int = 1) -> list:
        """rar_selftest_arithmetic: mismatch counts [rcp, sqrt, div, div-by-constant, range test]; all zero = pass."""
        m = (C.c_uint64 * 5)()
        self._ck(self._lib.rar_selftest_arithmetic(self._h, n_samples, seed & 0xFFFFFFFF, m))
        return [int(v) for v in m]

    def debug_grid(self) -> dict:
        """rar_debug_grid: dimensions, list length and digest of the (device-built) uniform grid of the current walls."""
        nx, ny, n, d = _i32(), _i32(), _i64(), C.c_uint64()
        self._ck(self._lib.rar_debug_grid(self._h, C.byref(nx), C.byref(ny), C.byref(n), C.byref(d)))
        return {"nx": nx.value, "ny": ny.value, "n_items": n.value, "digest": d.value}

    def launch_count(self) -> int:
        return int(self._lib.rar_launch_count(self._h))


class Ring:
    """rar_ring*: AudioManager's playback ring, lock-free for one producer and one consumer thread.  Needs no GPU."""

    def __init__(self, output_sample_rate: int = 48000, reverb_duration: float = 1.5):
        self._lib = load()
        h = _p()
        rc = self._lib.rar_ring_create(output_sample_rate, reverb_duration, C.byref(h))
        if rc < 0:
            raise RarError(rc, "rar_ring_create failed")
        self._h = h

    def destroy(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rar_ring_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def size(self) -> int:
        return int(self._lib.rar_ring_size(self._h))

    @property
    def pinned(self) -> bool:
        return bool(self._lib.rar_ring_is_pinned(self._h))

    @property
    def frames_drained(self) -> int:
        return int(self._lib.rar_ring_frames_drained(self._h))

    def reset(self) -> None:
        self._lib.rar_ring_reset(self._h)

    def stop(self) -> None:
        self._lib.rar_ring_stop(self._h)

    def push(self, samples: np.ndarray, sample_offset: int) -> None:
        a = np.ascontiguousarray(samples, dtype=np.float32)
        rc = self._lib.rar_ring_push(self._h, a.ctypes.data if a.size else None, a.size, sample_offset)
        if rc < 0:
            raise RarError(rc, "rar_ring_push: bad arguments")

    def drain(self, data: np.ndarray, channels: int = 1) -> None:
        """OnAudioFilterRead(data, channels): fills data in place (float32, contiguous)."""
        if data.dtype != np.float32 or not data.flags.c_contiguous:
            raise ValueError("data must be a contiguous float32 array")
        rc = self._lib.rar_ring_drain(self._h, data.ctypes.data if data.size else None, data.size, channels)
        if rc < 0:
            raise RarError(rc, "rar_ring_drain: bad arguments")


def prepared_length(samples: int, clip_frequency: int, sample_rate: int) -> int:
    """rar_prepared_length: the length LoadSample produces (RayTraceManager.cs:150-153); host-only, needs no device."""
    return int(load().rar_prepared_length(samples, clip_frequency, sample_rate))


def allreduce_slots(contexts, slot: int) -> None:
    """rar_allreduce_slots: sum `slot` over the contexts of this process (one per device), total everywhere."""
    arr = (_p * len(contexts))(*[c._h for c in contexts])
    check(contexts[0]._h, load().rar_allreduce_slots(arr, len(contexts), slot))


class Convolver:
    """rar_convolver*: the batched streaming partitioned convolver (BASELINE config 5)."""

    def __init__(self, ctx: Context, n_streams: int, block: int, max_ir_len: int):
        self._ctx = ctx
        self._lib = ctx._lib
        h = _p()
        ctx._ck(self._lib.rar_conv_create(ctx._h, n_streams, block, max_ir_len, C.byref(h)))
        self._h = h
        self.n_streams, self.block, self.max_ir_len = n_streams, block, max_ir_len

    def destroy(self) -> None:
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            self._lib.rar_conv_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def set_ir(self, stream: int, ir: np.ndarray, scale: float = 1.0) -> None:
        a = np.ascontiguousarray(ir, dtype=np.float32)
        self._ctx._ck(self._lib.rar_conv_set_ir(self._h, stream, a.ctypes.data if a.size else None, a.size, scale))

    def set_ir_from_slot(self, stream: int, slot: int, accum_count: int) -> None:
        self._ctx._ck(self._lib.rar_conv_set_ir_from_slot(self._h, stream, slot, accum_count))

    def set_irs(self, first_stream: int, irs: np.ndarray, scale: float = 1.0) -> None:
        """rar_conv_set_irs: irs[k] becomes the response of stream first_stream + k; no stream synchronisation."""
        a = np.ascontiguousarray(irs, dtype=np.float32)
        if a.ndim != 2:
            raise ValueError("irs must be [n_streams][ir_len]")
        self._ctx._ck(self._lib.rar_conv_set_irs(self._h, first_stream, a.shape[0], a.ctypes.data if a.size else None, a.shape[1],
                                                 a.shape[1], scale))

    def set_irs_from_slots(self, first_stream: int, slots, accum_counts) -> None:
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        ac = np.ascontiguousarray(accum_counts, dtype=np.int32)
        if sl.shape != ac.shape:
            raise ValueError("one accum_count per slot")
        self._ctx._ck(self._lib.rar_conv_set_irs_from_slots(self._h, first_stream, len(sl), sl.ctypes.data, ac.ctypes.data))

    def update_ir(self, stream: int, ir: np.ndarray, scale: float = 1.0) -> None:
        """rar_conv_update_ir: new response, cross-faded in over the next processed block."""
        a = np.ascontiguousarray(ir, dtype=np.float32)
        self._ctx._ck(self._lib.rar_conv_update_ir(self._h, stream, a.ctypes.data if a.size else None, a.size, scale))

    def update_ir_from_slot(self, stream: int, slot: int, accum_count: int) -> None:
        self._ctx._ck(self._lib.rar_conv_update_ir_from_slot(self._h, stream, slot, accum_count))

    def reset(self) -> None:
        self._ctx._ck(self._lib.rar_conv_reset(self._h))

    def process(self, block_in: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(block_in, dtype=np.float32)
        if x.shape != (self.n_streams, self.block):
            raise ValueError(f"expected input of shape {(self.n_streams, self.block)}")
        out = np.empty_like(x)
        self._ctx._ck(self._lib.rar_conv_process(self._h, x.ctypes.data, out.ctypes.data))
        return out

    def process_host_ptr(self, in_ptr: int, out_ptr: int) -> None:
        self._ctx._ck(self._lib.rar_conv_process(self._h, _p(in_ptr), _p(out_ptr)))

    def process_device(self, d_in_ptr: int, d_out_ptr: int) -> None:
        self._ctx._ck(self._lib.rar_conv_process_device(self._h, _p(d_in_ptr), _p(d_out_ptr)))

    def process_to_rings(self, block_in: np.ndarray, rings, sample_offset: int) -> None:
        """rar_conv_process_to_ring: one block; stream s goes straight into rings[s] (None: dropped)."""
        x = np.ascontiguousarray(block_in, dtype=np.float32)
        if x.shape != (self.n_streams, self.block) or len(rings) != self.n_streams:
            raise ValueError("one input row and one ring (or None) per stream")
        arr = (_p * self.n_streams)(*[(r._h if r is not None else None) for r in rings])
        self._ctx._ck(self._lib.rar_conv_process_to_ring(self._h, x.ctypes.data, arr, sample_offset))

    def bytes_per_block(self) -> int:
        return int(self._lib.rar_conv_bytes_per_block(self._h))
