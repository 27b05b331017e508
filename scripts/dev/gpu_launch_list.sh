#!/bin/bash
TAG=r1v4
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
python scripts/profile_step.py cifar 256 3 > gpurun_out/events_${TAG}.txt 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2>gpurun_out/bench_${TAG}.err; tail -c 200 gpurun_out/bench_${TAG}.json
du -sh gpurun_out
