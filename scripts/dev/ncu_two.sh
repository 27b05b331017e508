#!/bin/bash
mkdir -p gpurun_out
cap() {
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/prof_r1v3_$1 -f python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
}
cap conv_fwd 'conv_gemm_pair_kernel<[^0-9]*256[^0-9]+3[^0-9]+(1|true)' 2 1
cap conv_dgrad 'conv_gemm_pair_kernel<[^0-9]*256[^0-9]+3[^0-9]+(1|true)' 5 1
du -sh gpurun_out
