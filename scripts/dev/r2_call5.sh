#!/bin/bash
mkdir -p gpurun_out
cap() {
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/prof_r2e_$1 -f python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
}
cap conv_fwd 'conv_gemm_pair_kernel<[^0-9]*256[^0-9]+3[^0-9]+(1|true)' 2 1
cap conv_dgrad_n64 'conv_gemm_kernel<[^0-9]*64[^0-9]+3[^0-9]+(1|true)' 0 1
ls -la gpurun_out/*.ncu-rep
