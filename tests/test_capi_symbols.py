"""The C-ABI library loads without a GPU and exports exactly what include/rar2d.h declares; the ctypes
layouts match the C structs; and the compute entry points FAIL LOUDLY when no device is present (there is
no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from realisticaudioraytracing2d_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rar2d.h")


def _declared():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"RAR_API\s+[\w\s\*]+?\b(rar_\w+)\s*\(", txt)))


def test_header_and_binding_declare_the_same_symbols():
    assert _declared() == sorted(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    for name in _declared():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rar_\w+)", out))
    assert exported == set(_declared())          # nothing else leaks out of the library


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rar2d.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(rar_segment),sizeof(rar_ray_info),sizeof(rar_hit_key),sizeof(rar_trace_params),sizeof(rar_counters),"
                   "offsetof(rar_trace_params,ray_begin),offsetof(rar_trace_params,flags));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = list(map(int, subprocess.check_output([str(exe)]).split()))
    assert sizes[0] == 40 == _capi.SEGMENT_DTYPE.itemsize          # Helpers/SceneHelper.cs:15-22
    assert sizes[1] == 16 == _capi.RAY_INFO_DTYPE.itemsize         # RayTraceManager.cs:43
    assert sizes[2] == 8 == _capi.HIT_KEY_DTYPE.itemsize
    assert sizes[3] == C.sizeof(_capi.TraceParams)
    assert sizes[4] == C.sizeof(_capi.Counters)
    assert sizes[5] == _capi.TraceParams.ray_begin.offset and sizes[6] == _capi.TraceParams.flags.offset


def test_version_and_null_handling():
    lib = _capi.load()
    assert lib.rar_version() == 100
    assert lib.rar_destroy(None) == 0                               # `buffer?.Release()`
    assert lib.rar_conv_destroy(None) == 0
    assert lib.rar_sync(None) < 0 and lib.rar_launch_count(None) == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this check is for the CPU-only container")
    with pytest.raises(_capi.RarError) as e:
        _capi.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_csharp_binding_declares_every_symbol():
    """csharp/RarNative.cs (the P/Invoke side a maintainer drops into Assets/Script/, not compilable here) names
    exactly the entry points of include/rar2d.h."""
    cs = open(os.path.join(ROOT, "csharp", "RarNative.cs")).read()
    assert sorted(set(re.findall(r"extern\s+\w+\s+(rar_\w+)\s*\(", cs))) == _declared()
