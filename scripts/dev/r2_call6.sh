#!/bin/bash
mkdir -p gpurun_out
MMDGAN_PROF=1 MMDGAN_DEBUG=8 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2f_prof_ts.txt 2>&1
grep PROF gpurun_out/r2f_prof_ts.txt | tail -27 | head -14 | cut -c1-300
