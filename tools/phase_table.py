#!/usr/bin/env python
"""Executed warp instructions of a trace_deposit_kernel capture by PHASE of the bounce.

    python tools/phase_table.py gpurun_out/r02_c2.ncu-rep [warp_bounces]

ncu's source page counts an instruction under one line only, which says little for code that is inlined five deep.
This joins the per-SASS-instruction counts of the report (--page source --print-source sass) with the inline chains
nvdisasm prints for the same kernel of the library in the tree (-gi: "line N inlined at line M ..."), and names the
phase after the call that appears in the chain (bounce_nearest, listener_direct, ... in csrc/rar_ray.cuh and
csrc/trace_kernel.cu).  The library must be the build the capture was taken from (same SASS: checked by opcode).
With `warp_bounces` (rays / 32 x bounces of the run) the counts are also printed per warp-bounce: DESIGN.md 4.1's table.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "realisticaudioraytracing2d_b200", "librar2d.so")
PHASES = [  # (text that appears on a line of the inline chain, phase)
    ("bounce_nearest<", "nearest hit"),
    ("listener_direct<", "listener circle test, direct arrival"),
    ("deposit_direct<", "direct arrival: deposit"),
    ("bounce_advance(", "advance, material fetch"),
    ("listener_nee<", "next-event estimate"),
    ("check_vis(", "shadow ray"),
    ("coop_shadow<", "shadow ray (cooperative)"),
    ("nee_arrival<", "NEE arrival (time bin, energy)"),
    ("bounce_scatter<", "absorb, RNG, reflect / transmit"),
    ("deposit_hist<", "NEE deposit"),
    ("ray_init(", "ray set-up"),
    ("stage_scene<", "scene staging"),
]


def sass_counts(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hdr = rows[1]
    ia, isrc, ic = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    base = int(rows[2][ia], 16)
    return name, [(int(r[ia], 16) - base, r[isrc].split()[0 if not r[isrc].strip().startswith("@") else 1], int(r[ic])) for r in rows[2:] if len(r) > ic]


def mangled_args(name):
    """'trace_deposit_kernel<(int)1, (bool)0, ...>' -> 'ILi1ELb0E...' as it appears in the mangled symbol."""
    args = re.search(r"trace_deposit_kernel<(.*)>\(", name).group(1).split(", ")
    return "I" + "".join(("Li" if a.startswith("(int)") else "Lb") + a.split(")")[1] + "E" for a in args) + "E"


def chains(symbol_part):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "trace_kernel", LIB], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and "trace_deposit_kernel" + symbol_part in l)
    res, chain, fresh = [], [], True
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            fresh = True
            ops = m.group(2).split()
            res.append((int(m.group(1), 16), ops[1] if ops[0].startswith("@") else ops[0], list(chain)))
    return res


_src = {}


def text(f, ln):
    if f not in _src:
        try:
            _src[f] = open(f).read().split("\n")
        except OSError:
            _src[f] = []
    return _src[f][ln - 1] if 0 < ln <= len(_src[f]) else ""


def main():
    rep = sys.argv[1]
    wb = float(sys.argv[2]) if len(sys.argv) > 2 else None
    name, counts = sass_counts(rep)
    ch = {o: (op, c) for o, op, c in chains(mangled_args(name))}
    tot, table, bad = 0, {}, 0
    for off, op, n in counts:
        if off not in ch or ch[off][0].split(".")[0] != op.split(".")[0]:
            bad += 1
            continue
        tot += n
        lines = [text(f, ln) for f, ln in ch[off][1]]
        phase = next((p for key, p in PHASES if any(key in t for t in lines)), None)
        if phase is None:
            outer = ch[off][1][-1] if ch[off][1] else ("?", 0)
            t = text(*outer).strip()
            phase = "loop head (any-ray-alive vote, per-bounce set-up)" if ("__any_sync" in t or "i < max_b" in t or "else if (alive)" in t or "if (alive) {" in t or "want_shadow" in t) \
                else "tile loop, claim, debug rows, tail"
        if phase in ("nearest hit", "shadow ray") and any("intersect_exact" in t or "_EXACT(" in t or "_EXACT2(" in t for t in lines):
            phase += ": exact evaluation of survivors"
        elif phase in ("nearest hit", "shadow ray"):
            phase += ": wall filters"
        table[phase] = table.get(phase, 0) + n
    if bad:
        print(f"# WARNING: {bad} SASS rows did not match the library's disassembly (different build?)")
    print(f"# {name}\n# {tot} warp instructions" + (f", {tot / wb:.1f} per warp-bounce ({wb:.0f} warp-bounces)" if wb else ""))
    for phase, n in sorted(table.items(), key=lambda kv: -kv[1]):
        print(f"{n:12d} {n / tot * 100:5.1f}% " + (f"{n / wb:7.1f}  " if wb else " ") + phase)


if __name__ == "__main__":
    main()
