#!/bin/bash
# First hardware run of the opt-in NVSwitch-multicast data-parallel path (csrc/nvls.cu).  Run under `gpurun --gpus 2` (or 8):
# correctness (2-rank step == single-process step, replicas bit-identical) with and without MMDGAN_NVLS_ADAM=1, then the
# bench line of both paths.  Every step is bounded by `timeout` so that a cross-rank barrier that never completes cannot hold the box.
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== NCCL path" | tee gpurun_out/nvls_check.txt
run 29711 scripts/multi_gpu_check.py 2>&1 | tail -5 | tee -a gpurun_out/nvls_check.txt
echo "== NVLS path" | tee -a gpurun_out/nvls_check.txt
MMDGAN_NVLS_ADAM=1 run 29712 scripts/multi_gpu_check.py 2>&1 | tail -25 | tee -a gpurun_out/nvls_check.txt
echo "== NVLS path + global-batch batch norm" | tee -a gpurun_out/nvls_check.txt
MMDGAN_NVLS_ADAM=1 MMDGAN_SYNC_BN=1 run 29715 scripts/multi_gpu_check.py 2>&1 | tail -25 | tee -a gpurun_out/nvls_check.txt
echo "== bench, NCCL path" | tee -a gpurun_out/nvls_check.txt
run 29713 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_nccl_$N.json
echo "== bench, NVLS path" | tee -a gpurun_out/nvls_check.txt
MMDGAN_NVLS_ADAM=1 run 29714 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_nvls_$N.json
