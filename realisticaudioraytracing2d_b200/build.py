"""Builds csrc/ into the in-tree shared library `librar2d.so` for sm_100a with nvcc.

    python -m realisticaudioraytracing2d_b200.build [--force] [--verbose]

The library is built in-tree (next to this file) so that it travels with the repository snapshot to
the GPU box; nothing is installed into site-packages.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librar2d.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "-Xptxas", "-v"]

# The ray stage has a bit-exact arithmetic contract: no implicit FMA contraction, IEEE division/sqrt,
# no flush-to-zero.  The convolution stage has a tolerance contract and may contract freely.
UNITS = [
    ("trace_kernel.cu", ["--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]),
    ("conv_kernels.cu", []),
    ("band_synth.cu", []),
    ("exchange_kernel.cu", []),
    ("clip_kernel.cu", ["--fmad=false", "-prec-div=true", "-ftz=false"]),
    ("ring.cu", []),
    ("grid_kernel.cu", ["--fmad=false"]),
    ("rar2d_api.cu", []),
]
HEADERS = ["rar_math.cuh", "rar_ray.cuh", "rar_fft.cuh", "rar_synth16.cuh", "rar_layout.h", "rar_internal.h", "../../include/rar2d.h"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = r.stdout + r.stderr
            with open(o + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if verbose or r.returncode != 0:
                sys.stderr.write(log)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src} (see {o}.log)")
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
