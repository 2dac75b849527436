"""SURVEY 5 / VERDICT r1 item 6(ii): the CPU oracle and the host-compiled product headers (tests/host_emulation.cpp)
under AddressSanitizer + UndefinedBehaviorSanitizer.  (compute-sanitizer is closed on the GPU pool, so the device
side has no such run; the logic both sides share -- rar_ray.cuh, rar_layout.h -- is what gets sanitised here.)"""
import os
import struct
import subprocess

import numpy as np
import pytest

from realisticaudioraytracing2d_b200 import scenes
from tests.common import oracle_params, oracle_walls, trace_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
SAN = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-g", "-O1"]
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1", OMP_NUM_THREADS="2")


def _scene_file(tmp_path, O, sc, kw, name):
    bands = kw["bands"]
    P = oracle_params(O, kw)
    path = tmp_path / name
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", len(sc.walls), bands))
        f.write(bytes(P))
        f.write(np.ascontiguousarray(sc.walls).tobytes())
        if bands > 1:
            f.write(np.ascontiguousarray(sc.band_absorption, dtype=np.float32).tobytes())
    return str(path)


CASES = {
    "smoll": lambda: (scenes.smoll_room(), dict(ray_count=3000)),
    "shoebox": lambda: (scenes.shoebox(ray_count=2048, max_bounces=16, scattering=0.3, transmission=0.2, ior=1.3), {}),
    "maze_b8": lambda: (scenes.maze(n_segments=300, ray_count=1024, max_bounces=10, bands=8), dict(bands=8, impulse_length=6000)),
    "no_walls": lambda: (scenes.smoll_room(), dict(ray_count=256)),
}


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_under_asan_ubsan(oracle, tmp_path, case):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "sanitize"])
    sc, over = CASES[case]()
    if case == "no_walls":
        sc.walls = sc.walls[:0]
    kw = trace_kwargs(sc, **over)
    path = _scene_file(tmp_path, oracle, sc, kw, case + ".bin")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_build", "sanitize_oracle"), path], capture_output=True, text=True, env=ENV, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    # same counters as the ordinary build
    want = oracle.trace(oracle_walls(oracle, sc.walls), oracle_params(oracle, kw), band_abs=sc.band_absorption if kw["bands"] > 1 else None)
    line = r.stdout.splitlines()[0].split()
    assert int(line[5]) == want.counters["ray_bounces"] and int(line[7]) == want.counters["nearest_tests"]
    assert int(line[9]) == want.counters["shadow_tests"] and int(line[3]) == want.counters["direct_hits"] + want.counters["nee_hits"]


def test_host_emulation_under_asan_ubsan(oracle, tmp_path):
    """The product's per-ray logic, grid builder and FFT index logic compiled for the host, with a driver that calls
    the same entry points the emulation tests use."""
    drv = tmp_path / "drv.cpp"
    drv.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../realisticaudioraytracing2d_b200/csrc/rar_layout.h"
extern "C" int emu_trace(const rar_segment *, int, const float *, const rar_trace_params *, long long *, void *, long long, long long *, rar_counters *);
extern "C" int emu_trace_nocount(const rar_segment *, int, const float *, const rar_trace_params *, long long *, void *, long long, long long *, rar_counters *);
extern "C" int emu_grid_digest(const rar_segment *, int, int *, int *, long long *, unsigned long long *);
extern "C" void emu_convolve(const float *, int, const float *, int, int, float *);
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb");
    int n = 0, bands = 0;
    rar_trace_params p;
    if (!f || fread(&n, 4, 1, f) != 1 || fread(&bands, 4, 1, f) != 1 || fread(&p, sizeof p, 1, f) != 1) return 2;
    std::vector<rar_segment> w(n);
    if (n && fread(w.data(), sizeof(rar_segment), n, f) != (size_t)n) return 2;
    std::vector<float> ba((size_t)n * (bands > 1 ? bands : 0));
    if (!ba.empty() && fread(ba.data(), 4, ba.size(), f) != ba.size()) return 2;
    fclose(f);
    const long long words = (long long)p.impulse_length * (bands > 1 ? bands : 1);
    unsigned long long sums[3] = {0, 0, 0};
    for (int mode = 0; mode < 3; mode++) {                      // counting, production, production through the grid
        if (mode == 2 && bands > 1) continue;
        std::vector<long long> hist(words, 0);
        std::vector<char> hits(24 * 64);
        long long cnt = 0;
        rar_counters c;
        rar_trace_params q = p;
        if (mode == 2) q.flags |= RAR_FLAG_USE_GRID;
        int rc = (mode == 0 ? emu_trace : emu_trace_nocount)(w.data(), n, ba.empty() ? nullptr : ba.data(), &q, hist.data(), hits.data(), 64, &cnt, &c);
        if (rc != 0 && !(mode == 2 && rc == -6)) return 3;
        for (long long v : hist) sums[mode] = sums[mode] * 1099511628211ull + (unsigned long long)v;
    }
    int nx, ny; long long items; unsigned long long dig;
    emu_grid_digest(w.data(), n, &nx, &ny, &items, &dig);
    std::vector<float> x(700, 0.25f), ir(1000, 0.001f), y(1700);
    emu_convolve(x.data(), 700, ir.data(), 1000, 2, y.data());
    printf("%llu %llu %llu grid %d %d %lld\n", sums[0], sums[1], sums[2], nx, ny, items);
    return (sums[0] == sums[1] && (bands > 1 || n == 0 || sums[2] == sums[0] || nx == 0)) ? 0 : 4;
}
''')
    exe = tmp_path / "drv"
    subprocess.check_call([CXX, "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-mavx2", "-Wno-unknown-pragmas", *SAN,
                           "-I", os.path.join(ROOT, "tests"), "-o", str(exe), str(drv), os.path.join(ROOT, "tests", "host_emulation.cpp")],
                          cwd=os.path.join(ROOT, "tests"))
    for case in ("smoll", "shoebox", "maze_b8"):
        sc, over = CASES[case]()
        kw = trace_kwargs(sc, **over)
        path = _scene_file(tmp_path, oracle, sc, kw, case + ".bin")
        r = subprocess.run([str(exe), path], capture_output=True, text=True, env=ENV, timeout=300)
        assert r.returncode == 0, (case, r.stdout, r.stderr[-3000:])
        assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
