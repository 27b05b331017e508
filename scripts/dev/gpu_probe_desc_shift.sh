#!/bin/bash
# Round-2 experiment: shifted-start K-major descriptors (scripts/dev/probe_desc_shift.cu).  Run under gpurun, 1 GPU.
mkdir -p gpurun_out
nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mmd-gan_b200/csrc scripts/dev/probe_desc_shift.cu \
     -o /tmp/probe_desc_shift || exit 1
timeout 120 /tmp/probe_desc_shift 2>&1 | tee gpurun_out/probe_desc_shift.txt
