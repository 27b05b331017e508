#!/bin/bash
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1','ms/step',round(d['ms_per_step'],4))"; }
run default222
MMDGAN_WGRAD_CTAS=185 run ctas185
MMDGAN_WGRAD_CTAS=259 run ctas259
run default222
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_golden_api.py -q -m gpu -x --timeout 600 2>&1 | tail -2
