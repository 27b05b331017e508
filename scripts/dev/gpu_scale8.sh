#!/bin/bash
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --steps 20 --warmup 5 --workload $2 --no-roofline 2>/dev/null | tail -1 > gpurun_out/scale_$2_$1.json; python -c "
import json;d=json.loads(open('gpurun_out/scale_$2_$1.json').read().strip().splitlines()[-1]);print('$2 N=$1','ms/step',round(d['ms_per_step'],3),'img/s',round(d['value']),'e2e',round(d['e2e']['value']))" || tail -c 300 gpurun_out/scale_$2_$1.json; }
tr 8 cifar
tr 4 cifar
tr 8 lsun
tr 4 celeba
# strong scaling of BASELINE's CIFAR configuration (global batch 256 split over the ranks), NCCL and NVSwitch-multicast paths
ts() { $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1)) bench.py --gpus $1 --steps 20 --warmup 5 --scaling strong --no-roofline 2>/dev/null | tail -1 > gpurun_out/strong_$2_$1.json; python -c "
import json;d=json.loads(open('gpurun_out/strong_$2_$1.json').read().strip().splitlines()[-1]);print('strong $2 N=$1','ms/step',round(d['ms_per_step'],3),'img/s',round(d['value']))" || tail -c 300 gpurun_out/strong_$2_$1.json; }
ts 8 nccl env
ts 8 nvls "env MMDGAN_NVLS_ADAM=1"
