#!/usr/bin/env python
"""Turns the ncu artefacts in gpurun_out/ into the text summaries committed under profiles/.

    python tools/summarize_profiles.py <tag>      # r02 (default) or r01_final
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:3]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for m in METRICS:
            if m in hdr:
                lines.append(f"{m:85s} {r[hdr.index(m)]:>20s} {units[hdr.index(m)]}")
    return "\n".join(lines)


def launches(csv_path, title):
    rows = csv.DictReader([l for l in open(csv_path) if not l.startswith("==")])
    agg = collections.OrderedDict()
    for row in rows:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        agg.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = [f"# {title}", "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes"]
    for k, v in agg.items():
        out.append(f"{k[:118]:120s} n={len(v):4d} mean_us={sum(v) / len(v):10.1f} total_us={sum(v):11.1f} share={sum(v) / tot * 100:5.1f}%")
    return "\n".join(out)


JOBS = {
    "r01_final": [("prof_c2_final", "ncu_trace_c2", "trace_deposit_kernel on BASELINE config 2 (tools/run_trace.py c2)"),
                  ("prof_maze_final", "ncu_trace_maze", "trace_deposit_kernel on the 10 000-wall maze, brute force (tools/run_trace.py maze)"),
                  ("prof_grid_final", "ncu_trace_grid", "trace_deposit_kernel on the 10 000-wall maze, RAR_FLAG_USE_GRID (RAR_GRID=1 tools/run_trace.py maze)"),
                  ("prof_cmac_final", "ncu_cmac", "stream_cmac_kernel on BASELINE config 5 (tools/run_trace.py conv)")],
    # tools/profile_round2.sh
    "r02": [("r02_c2", "ncu_trace_c2", "trace_deposit_kernel<..., FAST=7> on BASELINE config 2 (tools/run_trace.py c2): four-wall, range-checked-once variant with packed FP32 wall tests"),
            ("r02_c2_guarded", "ncu_trace_c2_guarded", "the same dispatch with RAR_NO_FAST=1: the kernel with a guard around every division / square root"),
            ("r02_maze8", "ncu_trace_maze8", "trace_deposit_kernel on the 10 000-wall maze, 8 bands, brute force (tools/run_trace.py maze8)"),
            ("r02_c1", "ncu_trace_c1", "trace_deposit_kernel on BASELINE config 1, one 15 000-ray frame of SmollRoom (tools/run_trace.py c1)"),
            ("r02_band_synth", "ncu_band_synth", "band_synth_kernel (the first, shared-memory-FFT synthesis kernel; RAR_NO_FAST_SYNTH=1 now): 16 banded slots x 480 000 bins x 8 bands in one launch (tools/run_bench_leg.py banded)"),
            ("r02_listeners", "ncu_listeners", "trace_listeners_kernel (fused listeners, packed cooperative shadow scans): 128 listeners x 2^18 rays x 5 bounces, 2 000 walls (tools/run_listeners.py)"),
            ("r02_band_synth16", "ncu_band_synth16", "band_synth16_kernel (production synthesis kernel, register transforms): 16 banded slots x 480 000 bins x 8 bands in one launch (tools/run_bench_leg.py banded)")],
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    if tag.startswith("-") or "/" in tag:      # `--help` once became a file-name prefix under profiles/
        raise SystemExit(__doc__)
    for rep, name, title in JOBS.get(tag, []):
        path = os.path.join(OUT, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), path, "30"], capture_output=True, text=True).stdout
        with open(os.path.join(PROF, f"{tag}_{name}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on: {title}\n{raw(path)}\n\n"
                    "# per-source-line warp instructions (with inlining one SASS instruction is attributed to every line of its\n"
                    "# inline chain, so the per-line counts overlap; the kernel total is smsp__inst_executed.sum above)\n" + lines)
    lc = os.path.join(OUT, "launches_final.csv" if tag == "r01_final" else f"{tag}_launches.csv")
    if os.path.exists(lc):
        with open(os.path.join(PROF, f"{tag}_launches_summary.txt"), "w") as f:
            f.write(launches(lc, "ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 2 --warmup 3") + "\n")
        with open(lc) as src, open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as dst:
            dst.write(src.read())


if __name__ == "__main__":
    main()
