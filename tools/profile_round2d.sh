#!/bin/bash
# launch list of the final bench + one capture of the fused-listener kernel (packed shadow scans); see profile_round2.sh
set -u
out=gpurun_out
python bench.py --steps 2 --warmup 3 > $out/r02_bench_for_launches.json 2> $out/r02_bench_for_launches.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/r02_launches.csv python bench.py --steps 2 --warmup 3 > $out/r02_ncu_bench.log 2>&1
python tools/run_listeners.py > $out/r02_plain_listeners.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:trace_listeners -c 1 -s 1 -o $out/r02_listeners -f python tools/run_listeners.py > $out/r02_ncu_listeners.log 2>&1
ls -la $out/r02_listeners.ncu-rep $out/r02_launches.csv
