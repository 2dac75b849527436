#!/bin/bash
# gpurun_retry.sh <log> <gpurun args...>: repeats a gpurun call while the pod answers "transient"/busy (nothing charged).
log=$1; shift
for try in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    if grep -q "status=transient\|status=busy\|retry in" "$log"; then sleep 150; continue; fi
    break
done
echo "tries: $try" >> "$log"
