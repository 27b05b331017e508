#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1','ms/step',round(d['ms_per_step'],4))"; }
run base
MMDGAN_BN256_AUX=1 run bn256aux
MMDGAN_PAIR_MIN_TILES=128 run pair128
MMDGAN_PAIR_MIN_TILES=512 run pair512
MMDGAN_WGRAD_BN=128 run wgrad128
