#!/bin/bash
mkdir -p gpurun_out
one() { env "${@:2}" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $1 2>/dev/null | tail -1 > gpurun_out/r2_bench_$1${3:+_alt}.json; }
one cifar A=0 
one stl A=0
one celeba A=0
one lsun A=0
MMDGAN_DIRECT_CONV=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 > gpurun_out/r2_bench_cifar_nodirect.json
bash scripts/ncu_step.sh r2
for f in cifar stl celeba lsun cifar_nodirect; do python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(l['value']), round(l['ms_per_step'],3), round(l['e2e']['ms_per_step'],3), (l.get('roofline') or {}).get('frac'))
"; done
