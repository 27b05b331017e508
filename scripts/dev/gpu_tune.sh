#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1','ms/step',round(d['ms_per_step'],4))"; }
MMDGAN_PAIR_N64=0 run base
MMDGAN_PAIR_N64=1 run pairn64
MMDGAN_PAIR_N64=0 run base
MMDGAN_PAIR_N64=1 run pairn64
python scripts/profile_step.py cifar 256 3 2>&1 | awk '$8==64 && $6>1'
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -q -m gpu -x 2>&1 | tail -2
