set -u
out=gpurun_out
ncu_full="ncu --set full --import-source on --clock-control none"
python tools/run_trace.py c2 3 > $out/r02_plain_c2.log 2>&1 || exit 1
$ncu_full -k regex:trace_deposit -c 1 -s 2 -o $out/r02_c2 -f python tools/run_trace.py c2 3 > $out/r02_ncu_c2.log 2>&1
RAR_NO_FAST=1 RAR_NO_PACKED=1 $ncu_full -k regex:trace_deposit -c 1 -s 2 -o $out/r02_c2_guarded -f python tools/run_trace.py c2 3 > $out/r02_ncu_c2g.log 2>&1
python bench.py --steps 2 --warmup 3 > $out/r02_bench_for_launches.json 2> $out/r02_bench_for_launches.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/r02_launches.csv python bench.py --steps 2 --warmup 3 > $out/r02_ncu_bench.log 2>&1
ls -la $out | head -20
