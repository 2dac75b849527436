#!/usr/bin/env python
"""Per-source-line summary of an ncu report captured with --import-source on (-lineinfo builds).

    python tools/ncu_lines.py gpurun_out/prof_trace.ncu-rep [top_n]

Prints, for the first kernel in the report, the source lines ranked by executed warp instructions with
their share and stall samples.  Used to write the summaries under profiles/.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, func, seen_funcs = None, None, []
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1]
            if func not in seen_funcs:
                seen_funcs.append(func)
        elif r[0] == "Line No":
            hdr = r
        elif r[0].isdigit() and func == seen_funcs[0]:
            ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
            try:
                lines.append((int(r[ci]), int(r[cs]), cur_file, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    tot = sum(l[0] for l in lines)
    tsm = sum(l[1] for l in lines) or 1
    print(f"# {seen_funcs[0]}")
    print(f"# total warp instructions {tot}, stall samples {tsm}")
    for inst, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{inst:12d} {inst / tot * 100:5.1f}%  smp {smp / tsm * 100:5.1f}%  {f}:{ln:<4d} {src}")


if __name__ == "__main__":
    main()
