#!/bin/bash
# 2 (or N) GPU validation: NCCL data-parallel path (overlapped all-reduce), bench with dp_equals_single + strong point, then the
# NVSwitch-multicast path.  Every step is bounded by `timeout`.
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== NCCL multi_gpu_check" | tee gpurun_out/r2_multi_$N.txt
run 29711 scripts/multi_gpu_check.py 2>&1 | tail -8 | tee -a gpurun_out/r2_multi_$N.txt
echo "== bench NCCL" | tee -a gpurun_out/r2_multi_$N.txt
run 29713 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r2_bench_nccl_$N.err | tail -1 > gpurun_out/r2_bench_nccl_$N.json
tail -c 1500 gpurun_out/r2_bench_nccl_$N.json | tee -a gpurun_out/r2_multi_$N.txt; tail -5 gpurun_out/r2_bench_nccl_$N.err | tee -a gpurun_out/r2_multi_$N.txt
echo "== NVLS multi_gpu_check" | tee -a gpurun_out/r2_multi_$N.txt
MMDGAN_NVLS_ADAM=1 run 29712 scripts/multi_gpu_check.py 2>&1 | tail -25 | tee -a gpurun_out/r2_multi_$N.txt
echo "== NVLS + SYNC_BN multi_gpu_check" | tee -a gpurun_out/r2_multi_$N.txt
MMDGAN_NVLS_ADAM=1 MMDGAN_SYNC_BN=1 run 29715 scripts/multi_gpu_check.py 2>&1 | tail -25 | tee -a gpurun_out/r2_multi_$N.txt
echo "== bench NVLS" | tee -a gpurun_out/r2_multi_$N.txt
MMDGAN_NVLS_ADAM=1 run 29714 bench.py --gpus $N --steps 20 --warmup 5 --no-dp-check --no-strong 2> gpurun_out/r2_bench_nvls_$N.err | tail -1 > gpurun_out/r2_bench_nvls_$N.json
tail -c 600 gpurun_out/r2_bench_nvls_$N.json | tee -a gpurun_out/r2_multi_$N.txt; tail -5 gpurun_out/r2_bench_nvls_$N.err | tee -a gpurun_out/r2_multi_$N.txt
