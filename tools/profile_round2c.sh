#!/bin/bash
# the final captures of the packed trace kernels (config 2 and the 8-band maze); see profile_round2.sh
set -u
out=gpurun_out
ncu_full="ncu --set full --import-source on --clock-control none"
python tools/run_trace.py c2 3 > $out/r02_plain_c2.log 2>&1 || exit 1
$ncu_full -k regex:trace_deposit -c 1 -s 2 -o $out/r02_c2 -f python tools/run_trace.py c2 3 > $out/r02_ncu_c2.log 2>&1
python tools/run_trace.py maze8 2 > $out/r02_plain_maze8.log 2>&1 || exit 1
$ncu_full -k regex:trace_deposit -c 1 -s 1 -o $out/r02_maze8 -f python tools/run_trace.py maze8 2 > $out/r02_ncu_maze8.log 2>&1
ls -la $out/r02_c2.ncu-rep $out/r02_maze8.ncu-rep
