#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun -- in two calls, tools/profile_round2b.sh holds the c2 captures
# and the launch list: gpurun brings back at most 64 MiB per call and six full captures exceed that; artefacts land in gpurun_out/, summaries are made from
# them by tools/summarize_profiles.py r02).  Every capture follows a plain run of the same command that exited 0.
set -u
out=gpurun_out
py=python
ncu_full="ncu --set full --import-source on --clock-control none"
run() { echo "== $*" >&2; "$@"; }

$py tools/run_trace.py c2 3 > $out/r02_plain_c2.log 2>&1 || exit 1
run $ncu_full -k regex:trace_deposit -c 1 -s 2 -o $out/r02_c2 -f $py tools/run_trace.py c2 3 > $out/r02_ncu_c2.log 2>&1
RAR_NO_FAST=1 run $ncu_full -k regex:trace_deposit -c 1 -s 2 -o $out/r02_c2_guarded -f $py tools/run_trace.py c2 3 > $out/r02_ncu_c2g.log 2>&1
$py tools/run_trace.py maze8 2 > $out/r02_plain_maze8.log 2>&1 || exit 1
run $ncu_full -k regex:trace_deposit -c 1 -s 1 -o $out/r02_maze8 -f $py tools/run_trace.py maze8 2 > $out/r02_ncu_maze8.log 2>&1
$py tools/run_trace.py c1 3 > $out/r02_plain_c1.log 2>&1 || exit 1
run $ncu_full -k regex:trace_deposit -c 1 -s 2 -o $out/r02_c1 -f $py tools/run_trace.py c1 3 > $out/r02_ncu_c1.log 2>&1
$py tools/run_bench_leg.py banded > $out/r02_plain_banded.log 2>&1 || exit 1
run $ncu_full -k regex:band_synth16 -c 1 -s 2 -o $out/r02_band_synth16 -f $py tools/run_bench_leg.py banded > $out/r02_ncu_banded.log 2>&1
RAR_NO_FAST_SYNTH=1 run $ncu_full -k regex:band_synth_kernel -c 1 -s 2 -o $out/r02_band_synth -f $py tools/run_bench_leg.py banded > $out/r02_ncu_banded_old.log 2>&1
$py bench.py --steps 2 --warmup 3 > $out/r02_bench_for_launches.json 2> $out/r02_bench_for_launches.err || exit 1
run ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/r02_launches.csv $py bench.py --steps 2 --warmup 3 > $out/r02_ncu_bench.log 2>&1
ls -la $out/r02_* >&2
