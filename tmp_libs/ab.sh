#!/bin/bash
# A/B: time config 2 and the maze with two builds of the library
cd /root/repo
for lib in tmp_libs/librar2d_before.so new; do
  if [ "$lib" != new ]; then cp realisticaudioraytracing2d_b200/librar2d.so /tmp/new.so; cp $lib realisticaudioraytracing2d_b200/librar2d.so; fi
  echo "== $lib"
  python tools/run_trace.py c2 6 | tail -4 | head -3
  python tools/run_trace.py maze 3 | tail -3 | head -2
  python tools/run_trace.py c1 6 | tail -3 | head -2
  if [ "$lib" != new ]; then cp /tmp/new.so realisticaudioraytracing2d_b200/librar2d.so; fi
done
