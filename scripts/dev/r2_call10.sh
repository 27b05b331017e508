#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/phase_times.py > gpurun_out/r2_phases.txt 2>&1
tail -5 gpurun_out/r2_phases.txt
