#!/bin/bash
# retry a gpurun call while the pod answers "busy" (nothing is charged for those); usage: gpu_retry.sh <tries> <timeout_s> '<command>'
tries=$1; limit=$2; shift 2
for i in $(seq 1 "$tries"); do
    /usr/local/graft/bin/gpurun --timeout "$limit" -- "$@" > gpurun_out/retry.log 2>&1
    if ! grep -q "status=transient" gpurun_out/retry.log; then break; fi
    sleep 45
done
tail -30 gpurun_out/retry.log
