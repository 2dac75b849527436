#!/usr/bin/env python
"""Counts the SASS mnemonics that evidence the design per kernel of the built objects (no GPU needed):
UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), REDUX (warp reduce), MATCH (match_any), ATOMG/REDG/RED (global atomics),
FFMA2 / FMUL2 / FADD2 (packed FP32), UBLKPF (bulk L2 prefetch), MUFU, LDS, and the instruction count.

    python tools/sass_grep.py > profiles/r02_sass_grep.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "realisticaudioraytracing2d_b200", "_obj")
KEYS = ["UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "REDUX", "MATCH", "ATOMG", "REDG", "FFMA2", "FMUL2", "FADD2", "MUFU", "LDS", "LDG", "FCHK", "BSSY"]


def main():
    print("# cuobjdump -sass of the objects linked into librar2d.so (sm_100a); counts of instructions whose mnemonic starts with the key")
    print("# kernel | instructions | " + " ".join(KEYS))
    for obj in sorted(os.listdir(OBJ)):
        if not obj.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        arch = re.search(r"arch = (sm_\w+)", out)
        print(f"## {obj} ({arch.group(1) if arch else '?'})")
        for chunk in out.split("\t\tFunction : ")[1:]:
            name = chunk.split("\n", 1)[0].strip()
            ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[0-9T] )?([A-Z0-9_]+)", chunk)
            c = collections.Counter()
            for o in ops:
                for k in KEYS:
                    if o.startswith(k):
                        c[k] += 1
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
            demangled = re.sub(r"rar::\(anonymous namespace\)::", "", demangled)
            print(f"{demangled[:110]:112s} {len(ops):6d}  " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))


if __name__ == "__main__":
    main()
