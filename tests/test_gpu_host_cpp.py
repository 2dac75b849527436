"""The compiled C++ host mirror (host_cpp/rar2d_host.hpp: SceneToData2D, RayTraceManager, AudioManager, BakeAudio)
driven like the reference's Unity loop by tests/host_cpp_driver.cpp, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import _capi, scenes
from tests.common import rel_l2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "host_cpp_driver")


def _build():
    src = os.path.join(ROOT, "tests", "host_cpp_driver.cpp")
    hdr = os.path.join(ROOT, "host_cpp", "rar2d_host.hpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-ffp-contract=off", "-o", EXE, src,
                               "-L" + os.path.dirname(_capi.LIB_PATH), "-l:librar2d.so",
                               "-Wl,-rpath," + os.path.dirname(_capi.LIB_PATH)])


def test_cpp_host_mirror_against_oracle(tmp_path, oracle):
    _build()
    out = tmp_path / "dump.bin"
    r = subprocess.run([EXE, str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "segments 20 frames 3 accum 3" in r.stdout and "streaming pending 0" in r.stdout
    raw = out.read_bytes()
    sc = scenes.smoll_room()
    segs = np.frombuffer(raw[: 20 * 40], dtype=sc.walls.dtype)
    assert segs.tobytes() == sc.walls.tobytes()                    # C++ SceneToData2D == Python mirror == oracle add_loop
    q = np.frombuffer(raw[800: 800 + 72000 * 8], dtype=np.int64)
    hist = np.zeros(72000, np.int64)
    for f in (1, 2, 3):
        P = oracle.make_params(source_x=-18.0, source_y=9.0, listener_x=0.0, listener_y=-3.68, ray_count=15000,
                               max_bounce_count=5, rng_state_offset=f, impulse_length=72000, debug_ray_count=100)
        oracle.trace(sc.walls.view(oracle.SEGMENT_DTYPE), P, hist=hist)
    assert np.array_equal(q, hist)
    baked = np.frombuffer(raw[800 + 72000 * 8:], dtype=np.float32)
    s, clip = 12345, np.zeros(6000, np.float32)
    for i in range(6000):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        clip[i] = np.float32(((s >> 8) & 0xFFFF) / 65536.0 - 0.5)
    want = oracle.convolve(clip, oracle.ir_to_float(hist), 3)
    want = want / np.abs(want).max()
    assert baked.shape == want.shape and rel_l2(baked, want) <= 1e-4
    energy = float(r.stdout.split("ring_energy")[1])
    assert energy > 0
