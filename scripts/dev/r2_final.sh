#!/bin/bash
# Round-2 final single-GPU artefacts: full GPU test suite, smoke, ncu captures, bench lines of the four workloads, reference arm.
mkdir -p gpurun_out
rm -f gpurun_out/prof_r2_*
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2_final_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 bash scripts/ncu_step.sh r2 > gpurun_out/ncu_step.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_cifar.json 2> gpurun_out/r2_bench_cifar.err; tail -c 300 gpurun_out/r2_bench_cifar.json
for w in stl celeba lsun; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_$w.json').read().strip().splitlines()[-1]);print('$w','ms/step',d['ms_per_step'],'img/s',d['value'],'frac',d['roofline']['frac'])" || tail -3 gpurun_out/r2_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r2_bench_reference.json
timeout 200 python scripts/phase_times.py > gpurun_out/r2_phases.txt 2>&1; tail -2 gpurun_out/r2_phases.txt
timeout 200 python scripts/mmd_sweep.py > gpurun_out/r2_mmd_sweep.txt 2>&1
du -sh gpurun_out
